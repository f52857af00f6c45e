"""MODEL.PLN.DISTANCE_TYPE = L1 / L2 (and COS with two representatives per class), pinned to the reference's OWN code:
``tests/golden/golden_ref_pln_dist_v1.npz`` was produced by executing the unmodified reference ``PLN.loss`` /
``PLN.inference`` (``tests/golden/make_golden_pln_dist.py``; nothing in it comes from ``oracle/`` or the product).
The oracle must reproduce it on the CPU, the CUDA kernels (through the C ABI) on the GPU.

Tolerances (fp32): loss rtol 2e-5; gradients rtol 1e-3 / atol 2e-6 of entries that are O(1e-3) - the L1 gradient is a sum
of +-1 signs pushed through two normalisations, the L2 one divides by the distance: summation order and one rounding of
the distance are the only differences.  Rows whose distance is within 1e-5 of a hinge threshold (or whose two nearest
prototypes are within 1e-5) are excluded: a flipped hinge is a discontinuity of the gradient, not an error."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import pln as opln

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_ref_pln_dist_v1.npz"))
CASES = [str(c) for c in G["cases"]]
K = 20


def t(case, name, dev="cpu"):
    return torch.from_numpy(G[f"{case}_{name}"]).to(dev)


def params(case):
    rpc, alpha, beta, unk_thr, w, iou_thr = (float(v) for v in G[f"{case}_params"])
    return dict(reps_per_class=int(rpc), alpha=alpha, beta=beta, loss_weight=w, iou_threshold=iou_thr,
                distance_type=str(G[f"{case}_dist"])), unk_thr


def safe_rows(case):
    """Rows of the training batch that are NOT within 1e-5 of a decision boundary (computed from the reference's emb)."""
    kw, _ = params(case)
    eh, rh = F.normalize(t(case, "emb")), F.normalize(t(case, "reps"))
    d = opln.pln_distance(eh, rh, kw["distance_type"])
    rpc = kw["reps_per_class"]
    d3 = d.reshape(-1, K, rpc)
    clear = torch.ones(d.shape[0], dtype=torch.bool)
    if rpc > 1:
        top2 = torch.topk(d3, 2, dim=2, largest=False).values
        clear = ((top2[:, :, 1] - top2[:, :, 0]) > 1e-5).all(dim=1)
    dmin = d3.min(dim=2)[0]
    cls, ious = t(case, "gt_classes"), t(case, "ious")
    fg = (cls >= 0) & (cls < K) & (ious > kw["iou_threshold"])
    y = cls.clamp(0, K - 1)
    intra = dmin.gather(1, y[:, None])[:, 0]
    dm = dmin.clone()
    dm.scatter_(1, y[:, None], 1000.0)
    top2 = torch.topk(dm, 2, dim=1, largest=False).values
    safe = ((intra - kw["alpha"]).abs() > 1e-5) & ((kw["beta"] - top2[:, 0]).abs() > 1e-5) & ((top2[:, 1] - top2[:, 0]) > 1e-5)
    return (safe & clear) | ~fg


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_reference_pln_distance_types(case):
    kw, unk_thr = params(case)
    x = t(case, "x").requires_grad_(True)
    enc_w = t(case, "enc_w").requires_grad_(True)
    reps = t(case, "reps").requires_grad_(True)
    emb, rec, loss = opln.pln_loss(x, enc_w, t(case, "enc_b"), t(case, "dec_w"), t(case, "dec_b"), reps,
                                   t(case, "gt_classes"), t(case, "ious"), num_known_classes=K, **kw)
    emb.retain_grad()
    loss.backward()
    torch.testing.assert_close(emb.detach(), t(case, "emb"), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(rec.detach(), t(case, "rec"), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(loss.detach(), t(case, "loss"), rtol=1e-6, atol=0)
    torch.testing.assert_close(emb.grad, t(case, "grad_emb"), rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(reps.grad, t(case, "grad_reps"), rtol=1e-5, atol=1e-8)
    torch.testing.assert_close(enc_w.grad, t(case, "grad_enc_w"), rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(x.grad, t(case, "grad_x"), rtol=1e-4, atol=1e-7)
    for n in range(2):
        rec_n, pred = opln.pln_inference(t(case, f"inf_feats{n}"), t(case, "enc_w"), t(case, "enc_b"), t(case, "dec_w"),
                                         t(case, "dec_b"), t(case, "reps"), num_known_classes=K,
                                         reps_per_class=kw["reps_per_class"], unk_thr=unk_thr,
                                         distance_type=kw["distance_type"], unknown_id=80)
        assert torch.equal(pred, t(case, f"inf_pred{n}"))
        torch.testing.assert_close(rec_n, t(case, f"inf_rec{n}"), rtol=1e-6, atol=1e-6)


def test_fixture_covers_every_distance_type_and_hinge():
    assert {str(G[f"{c}_dist"]) for c in CASES} == {"COS", "L1", "L2"}
    for case in CASES:
        assert float(t(case, "grad_reps").abs().max()) > 0 and float(t(case, "grad_emb").abs().max()) > 0
        pred = torch.cat([t(case, "inf_pred0"), t(case, "inf_pred1")])
        assert 0 < int((pred == 80).sum()) < pred.numel()


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_pln_distance_types_match_reference(case):
    from osr_b200.pln import PLN, pln_loss_from_emb, pln_loss_fwd_bwd
    from osr_b200.structures import Instances
    kw, unk_thr = params(case)
    dev = "cuda:0"
    pln = PLN(num_classes=81, num_known_classes=K, feature_dim=64, embedding_dim=256, unk_thr=unk_thr,
              opendet_benchmark=True, **{k: v for k, v in kw.items() if k not in ("iou_threshold",)},
              iou_threshold=kw["iou_threshold"])
    with torch.no_grad():
        pln.encoder.weight.copy_(t(case, "enc_w", dev)); pln.encoder.bias.copy_(t(case, "enc_b", dev))
        pln.decoder.weight.copy_(t(case, "dec_w", dev)); pln.decoder.bias.copy_(t(case, "dec_b", dev))
        pln.representatives.copy_(t(case, "reps", dev))
    cls, ious = t(case, "gt_classes", dev), t(case, "ious", dev)
    n0 = int(G[f"{case}_split"][0])
    props = []
    for lo, hi in ((0, n0), (n0, cls.numel())):
        p = Instances((100, 100))
        p.set("gt_classes", cls[lo:hi]); p.set("ious", ious[lo:hi])
        props.append(p)
    # (1) the module, through autograd: PLN.loss(roi_features, proposals)
    x = t(case, "x", dev).requires_grad_(True)
    emb, rec, loss = pln.loss(x, props)
    emb.retain_grad()
    loss.backward()
    torch.testing.assert_close(loss.detach().cpu(), t(case, "loss"), rtol=2e-5, atol=0)
    torch.testing.assert_close(emb.detach().cpu(), t(case, "emb"), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(rec.detach().cpu(), t(case, "rec"), rtol=1e-5, atol=1e-5)
    safe = safe_rows(case)
    assert safe.float().mean() > 0.9
    torch.testing.assert_close(emb.grad.cpu()[safe], t(case, "grad_emb")[safe], rtol=1e-3, atol=2e-6)
    # (2) the functional forms on the reference's own embedding, unsafe rows taken out of the foreground on both sides
    if bool(safe.all()):
        torch.testing.assert_close(pln.representatives.grad.cpu(), t(case, "grad_reps"), rtol=1e-3, atol=2e-6)
        torch.testing.assert_close(pln.encoder.weight.grad.cpu(), t(case, "grad_enc_w"), rtol=2e-3, atol=2e-6)
        torch.testing.assert_close(x.grad.cpu(), t(case, "grad_x"), rtol=2e-3, atol=2e-6)
    emb_ref = t(case, "emb", dev)
    kwf = dict(num_known_classes=K, **kw)
    l2, g_emb, g_reps = pln_loss_fwd_bwd(emb_ref, t(case, "reps", dev), cls, ious, **kwf)
    torch.testing.assert_close(l2.cpu(), t(case, "loss"), rtol=2e-5, atol=0)
    torch.testing.assert_close(g_emb.cpu()[safe], t(case, "grad_emb")[safe], rtol=1e-3, atol=2e-6)
    # fused forward+backward == autograd pair, bit for bit
    ea = emb_ref.clone().requires_grad_(True); ra = t(case, "reps", dev).requires_grad_(True)
    pln_loss_from_emb(ea, ra, cls, ious, **kwf).backward()
    assert torch.equal(ea.grad, g_emb) and torch.equal(ra.grad, g_reps)
    # prototype gradient against the oracle's autograd with the unsafe rows removed on both sides
    ious_s = torch.where(safe.to(dev), ious, torch.zeros_like(ious))
    _, _, g_reps_s = pln_loss_fwd_bwd(emb_ref, t(case, "reps", dev), cls, ious_s, **kwf)
    rb = t(case, "reps").requires_grad_(True)
    opln.pln_loss_from_emb(t(case, "emb"), rb, t(case, "gt_classes"), ious_s.cpu(), num_known_classes=K, **kw).backward()
    torch.testing.assert_close(g_reps_s.cpu(), rb.grad, rtol=1e-3, atol=2e-6)
    # (3) inference
    fg = []
    for n in range(2):
        q = Instances((100, 100))
        q.set("features", t(case, f"inf_feats{n}", dev))
        fg.append(q)
    with torch.no_grad():
        out = pln.inference(fg)
    for n, r in enumerate(out):
        e = F.normalize(F.linear(t(case, f"inf_feats{n}"), t(case, "enc_w"), t(case, "enc_b")))
        d = opln.pln_distance(e, F.normalize(t(case, "reps")), kw["distance_type"])
        dmin = d.reshape(-1, K, kw["reps_per_class"]).min(dim=2)[0]
        top2 = torch.topk(dmin, 2, dim=1, largest=False).values
        ok = ((top2[:, 0] - unk_thr).abs() > 1e-5) & ((top2[:, 1] - top2[:, 0]) > 1e-5)
        assert ok.float().mean() > 0.9
        assert torch.equal(r.get("pred_classes").cpu()[ok], t(case, f"inf_pred{n}")[ok])
        torch.testing.assert_close(r.get("features").cpu(), t(case, f"inf_rec{n}"), rtol=1e-5, atol=1e-5)
