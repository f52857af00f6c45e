"""GPU parity: PLN loss fwd/bwd + inference kernels vs the oracle (torch ops, fp32).
Tolerances: loss rtol 1e-5 (fp32 dot products in a different summation order); gradients compared on all rows
with atol 1e-7 (entries are O(1e-4)) excluding rows whose hinge margin is within 1e-5 of a threshold (a flipped
hinge is a discontinuity of the gradient, not an error)."""
import pytest
import torch

from oracle import pln as opln

pytestmark = pytest.mark.gpu


def _kw(K=20, alpha=0.1, beta=0.9, w=0.5):
    return dict(num_known_classes=K, alpha=alpha, beta=beta, loss_weight=w, iou_threshold=0.5)


def _margin_mask(emb, reps, labels, ious, K, alpha, beta, rpc=1):
    eh = torch.nn.functional.normalize(emb); rh = torch.nn.functional.normalize(reps)
    d = 1 - eh @ rh.t()
    if rpc > 1:   # class distance = min over the class's representatives; also require a clear arg-min inside each class
        d3 = d.reshape(-1, K, rpc)
        top2 = torch.topk(d3, 2, dim=2, largest=False).values
        clear = ((top2[:, :, 1] - top2[:, :, 0]) > 1e-5).all(dim=1)
        d = d3.min(dim=2)[0]
    else:
        clear = torch.ones(d.shape[0], dtype=torch.bool, device=d.device)
    fg = (labels >= 0) & (labels < K) & (ious > 0.5)
    y = labels.clamp(0, K - 1)
    intra = d.gather(1, y[:, None])[:, 0]
    dm = d.clone(); dm.scatter_(1, y[:, None], 1000.0)
    inter = dm.min(1)[0]
    dm2 = torch.topk(dm, 2, dim=1, largest=False).values
    safe = ((intra - alpha).abs() > 1e-5) & ((beta - inter).abs() > 1e-5) & ((dm2[:, 1] - dm2[:, 0]) > 1e-5) & clear
    return safe | ~fg


@pytest.mark.parametrize("R,K,alpha,beta,w", [(8192, 20, 0.1, 0.9, 0.5), (4096, 28, 0.05, 0.95, 2.0), (37, 20, 0.1, 0.9, 0.5)])
def test_loss_and_gradients_match_oracle(R, K, alpha, beta, w):
    from osr_b200 import synth
    from osr_b200.pln import pln_loss_from_emb
    pi = synth.make_pln_inputs(R, num_known=K, num_classes=81 if K == 20 else 88, seed=R, device="cuda:0")
    emb0 = (pi.roi_features @ pi.enc_w.t()).detach()
    kw = _kw(K, alpha, beta, w)
    emb_a = emb0.clone().requires_grad_(True); reps_a = pi.reps.clone().requires_grad_(True)
    la = pln_loss_from_emb(emb_a, reps_a, pi.gt_classes, pi.ious, **kw)
    (la * 1.7).backward()
    emb_b = emb0.clone().requires_grad_(True); reps_b = pi.reps.clone().requires_grad_(True)
    lb = opln.pln_loss_from_emb(emb_b, reps_b, pi.gt_classes, pi.ious, **kw)
    (lb * 1.7).backward()
    torch.testing.assert_close(la, lb, rtol=1e-5, atol=1e-7)
    safe = _margin_mask(emb0, pi.reps, pi.gt_classes, pi.ious, K, alpha, beta)
    assert safe.float().mean() > 0.99
    torch.testing.assert_close(emb_a.grad[safe], emb_b.grad[safe], rtol=1e-4, atol=1e-7)
    # prototype gradient: ALWAYS checked - rows within 1e-5 of a hinge threshold (where a flipped hinge is a discontinuity,
    # not an error) are taken out of the foreground in BOTH implementations by zeroing their IoU
    ious2 = torch.where(safe, pi.ious, torch.zeros_like(pi.ious))
    ra = pi.reps.clone().requires_grad_(True); rb = pi.reps.clone().requires_grad_(True)
    pln_loss_from_emb(emb0, ra, pi.gt_classes, ious2, **kw).backward()
    opln.pln_loss_from_emb(emb0, rb, pi.gt_classes, ious2, **kw).backward()
    torch.testing.assert_close(ra.grad, rb.grad, rtol=1e-4, atol=1e-6)
    # CPU oracle too
    lc = opln.pln_loss_from_emb(emb0.cpu(), pi.reps.cpu(), pi.gt_classes.cpu(), pi.ious.cpu(), **kw)
    torch.testing.assert_close(la.cpu(), lc, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("rpc,D", [(2, 256), (5, 256), (1, 64), (2, 128), (5, 64)])
def test_reps_per_class_and_small_embedding_dims(rpc, D):
    """REPS_PER_CLASS > 1 (min over a class's representatives, `prototype_learning_network.py:164`) and EMD_DIM < 256:
    loss and both gradients against the autograd oracle (same neutralisation of near-threshold rows as above)."""
    from osr_b200 import synth
    from osr_b200.pln import pln_loss_from_emb
    K, R = 20, 1500
    pi = synth.make_pln_inputs(R, emb_dim=D, num_known=K, seed=rpc * 100 + D, device="cuda:0")
    g = torch.Generator("cuda:0").manual_seed(rpc + D)
    reps0 = torch.randn(K * rpc, D, device="cuda:0", generator=g)
    emb0 = (pi.roi_features @ pi.enc_w.t()).detach()
    kw = dict(_kw(K), reps_per_class=rpc)
    safe = _margin_mask(emb0, reps0, pi.gt_classes, pi.ious, K, 0.1, 0.9, rpc)
    assert safe.float().mean() > 0.98
    ious2 = torch.where(safe, pi.ious, torch.zeros_like(pi.ious))
    ea = emb0.clone().requires_grad_(True); ra = reps0.clone().requires_grad_(True)
    eb = emb0.clone().requires_grad_(True); rb = reps0.clone().requires_grad_(True)
    la = pln_loss_from_emb(ea, ra, pi.gt_classes, ious2, **kw)
    lb = opln.pln_loss_from_emb(eb, rb, pi.gt_classes, ious2, **kw)
    la.backward(); lb.backward()
    torch.testing.assert_close(la, lb, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(ea.grad, eb.grad, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(ra.grad, rb.grad, rtol=1e-4, atol=1e-6)


def test_no_foreground_rows_only_center_term():
    from osr_b200 import synth
    from osr_b200.pln import pln_loss_from_emb
    pi = synth.make_pln_inputs(64, seed=2, device="cuda:0")
    labels = torch.full((64,), 81, dtype=torch.int64, device="cuda:0")
    emb = (pi.roi_features @ pi.enc_w.t()).requires_grad_(True)
    reps = pi.reps.clone().requires_grad_(True)
    la = pln_loss_from_emb(emb, reps, labels, pi.ious, **_kw())
    la.backward()
    lb = opln.pln_loss_from_emb(emb.detach(), pi.reps, labels, pi.ious, **_kw())
    torch.testing.assert_close(la, lb, rtol=1e-5, atol=1e-8)
    assert float(emb.grad.abs().max()) == 0.0


def test_deterministic():
    from osr_b200 import synth
    from osr_b200.pln import pln_loss_from_emb
    pi = synth.make_pln_inputs(4096, seed=6, device="cuda:0")
    emb0 = (pi.roi_features @ pi.enc_w.t()).detach()
    outs = []
    for _ in range(2):
        e = emb0.clone().requires_grad_(True); r = pi.reps.clone().requires_grad_(True)
        l = pln_loss_from_emb(e, r, pi.gt_classes, pi.ious, **_kw())
        l.backward()
        outs.append((l.detach().clone(), e.grad.clone(), r.grad.clone()))
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


def test_module_loss_matches_reference_formulation():
    from osr_b200 import synth
    from osr_b200.pln import PLN
    from osr_b200.structures import Instances
    torch.manual_seed(0)
    m = PLN(81, 20, 1024, 256, "COS", 1, 0.1, 0.9, 0.5, opendet_benchmark=True)
    pi = synth.make_pln_inputs(1024, seed=8, device="cuda:0")
    props = []
    for i in range(2):
        inst = Instances((800, 1333))
        inst.gt_classes = pi.gt_classes[i * 512:(i + 1) * 512]
        inst.ious = pi.ious[i * 512:(i + 1) * 512]
        props.append(inst)
    emb, rec, loss = m.loss(pi.roi_features, props)
    e2, r2, l2 = opln.pln_loss(pi.roi_features, m.encoder.weight, m.encoder.bias, m.decoder.weight, m.decoder.bias,
                               m.representatives, pi.gt_classes, pi.ious, **_kw())
    torch.testing.assert_close(emb, e2); torch.testing.assert_close(rec, r2)
    torch.testing.assert_close(loss, l2, rtol=1e-5, atol=1e-7)
    loss.backward()
    assert m.encoder.weight.grad is not None and m.representatives.grad is not None
    assert set(dict(m.named_parameters())) == {"encoder.weight", "encoder.bias", "decoder.weight", "decoder.bias",
                                               "representatives"}


def test_inference_matches_oracle():
    from osr_b200 import synth
    from osr_b200.pln import PLN
    from osr_b200.structures import Instances
    torch.manual_seed(1)
    m = PLN(81, 20, 1024, 256, "COS", 1, 0.1, 0.9, 0.5, unk_thr=0.93, opendet_benchmark=True)
    pi = synth.make_pln_inputs(300, seed=9, device="cuda:0")
    insts = []
    for a, b in ((0, 100), (100, 100), (100, 300)):
        inst = Instances((800, 1333)); inst.features = pi.roi_features[a:b]; insts.append(inst)
    with torch.no_grad():
        out = m.inference(insts)
        for inst, (a, b) in zip(out, ((0, 100), (100, 100), (100, 300))):
            rec, pred = opln.pln_inference(pi.roi_features[a:b], m.encoder.weight, m.encoder.bias, m.decoder.weight,
                                           m.decoder.bias, m.representatives, num_known_classes=20, unk_thr=0.93,
                                           unknown_id=80)
            assert torch.equal(inst.pred_classes, pred)
            torch.testing.assert_close(inst.features, rec)
        assert any((o.pred_classes == 80).any() for o in out if len(o)) and any((o.pred_classes < 20).any() for o in out if len(o))


def _tf32(t):
    """Operand rounding of tcgen05 kind::tf32: the tensor core reads the top 19 bits of an fp32 (sign, 8 exponent, 10
    mantissa bits) - truncation of the low 13 mantissa bits."""
    return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("variant", [0, 1], ids=["tf32_direct", "bf16_cast"])
@pytest.mark.parametrize("R", [8192, 1000, 128, 77])
def test_tcgen05_encoder_matches_reference(R, variant):
    """tcgen05 encoder GEMM, fp32 accumulate in TMEM.  Shipped form (OSR_TUNE_PLN_VARIANT 0): fp32 x and W are loaded as
    they are and multiplied as TF32 (kind::tf32, no cast pass); variant 1: bf16 copies (kind::f16).  Each is compared (a)
    against the same REDUCED-precision operands multiplied in fp64 by torch - what remains is the fp32 accumulation -
    and (b) against the fp32 nn.Linear with the operand-rounding tolerance (tf32: 2^-10, bf16: 2^-8 relative per operand,
    K = 1024 products of magnitude ~6e-3)."""
    from osr_b200 import _lib, synth
    from osr_b200.pln import pln_encode_tc
    pi = synth.make_pln_inputs(R, seed=R, device="cuda:0")
    bias = torch.randn(256, device="cuda:0", generator=torch.Generator("cuda:0").manual_seed(1)) * 0.01
    prev = _lib.set_tuning("pln", variant)
    try:
        got = pln_encode_tc(pi.roi_features, pi.enc_w, bias)
    finally:
        _lib.set_tuning("pln", prev)
    ref_fp32 = torch.nn.functional.linear(pi.roi_features, pi.enc_w, bias)
    if variant == 0:
        red = (_tf32(pi.roi_features).double() @ _tf32(pi.enc_w).double().t() + bias.double()).float()
        torch.testing.assert_close(got, red, rtol=1e-4, atol=2e-5)
        torch.testing.assert_close(got, ref_fp32, rtol=5e-3, atol=1e-3)
    else:
        red = pi.roi_features.bfloat16().float() @ pi.enc_w.bfloat16().float().t() + bias
        torch.testing.assert_close(got, red, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(got, ref_fp32, rtol=2e-2, atol=3e-3)


def test_tcgen05_encoder_in_module_loss_and_grads():
    from osr_b200 import synth
    from osr_b200.pln import PLN
    torch.manual_seed(0)
    m = PLN(81, 20, 1024, 256, "COS", 1, 0.1, 0.9, 0.5, opendet_benchmark=True, encoder_impl="tcgen05")
    m32 = PLN(81, 20, 1024, 256, "COS", 1, 0.1, 0.9, 0.5, opendet_benchmark=True)
    m32.load_state_dict(m.state_dict())
    pi = synth.make_pln_inputs(2048, seed=12, device="cuda:0")
    _, _, l_tc = m.loss_from_tensors(pi.roi_features, pi.gt_classes, pi.ious)
    _, _, l_32 = m32.loss_from_tensors(pi.roi_features, pi.gt_classes, pi.ious)
    torch.testing.assert_close(l_tc, l_32, rtol=2e-2, atol=1e-4)   # bf16 operands: stated tolerance of the north star
    l_tc.backward()
    assert m.encoder.weight.grad is not None and torch.isfinite(m.encoder.weight.grad).all()


def test_zero_rows_gives_separation_term_only():
    """R == 0 (an image batch with no sampled RoIs): the loss is the prototype-separation term alone."""
    from osr_b200 import synth
    from osr_b200.pln import pln_loss_from_emb
    pi = synth.make_pln_inputs(8, seed=3, device="cuda:0")
    emb = torch.zeros(0, pi.enc_w.shape[0], device="cuda:0", requires_grad=True)
    labels = torch.zeros(0, dtype=torch.int64, device="cuda:0")
    ious = torch.zeros(0, device="cuda:0")
    reps = pi.reps.clone().requires_grad_(True)
    la = pln_loss_from_emb(emb, reps, labels, ious, **_kw())
    la.backward()
    lb = opln.pln_loss_from_emb(emb.detach().cpu(), pi.reps.cpu(), labels.cpu(), ious.cpu(), **_kw())
    torch.testing.assert_close(la.cpu(), lb, rtol=1e-5, atol=1e-8)
    assert torch.isfinite(reps.grad).all()


@pytest.mark.parametrize("dist,rpc,R,K", [("L1", 1, 4096, 20), ("L2", 1, 4096, 20), ("L2", 2, 1500, 28), ("L1", 5, 333, 28)])
def test_l1_l2_distances_loss_and_gradients_match_oracle(dist, rpc, R, K):
    """MODEL.PLN.DISTANCE_TYPE L1 / L2 (prototype_learning_network.py:156-161,171-176) at training sizes against the
    oracle's torch.cdist + autograd; alpha / beta sit at the medians of the intra / inter distances so that both hinges are
    half active.  Rows within 1e-5 of a decision boundary are taken out of the foreground on both sides."""
    from osr_b200 import synth
    from osr_b200.pln import pln_loss_from_emb, pln_loss_fwd_bwd, pln_nearest
    pi = synth.make_pln_inputs(R, num_known=K, num_classes=81 if K == 20 else 88, seed=R + rpc, device="cuda:0")
    emb0 = (pi.roi_features @ pi.enc_w.t()).detach()
    g = torch.Generator(device="cuda:0").manual_seed(R)
    reps = torch.randn(K * rpc, emb0.shape[1], device="cuda:0", generator=g)
    eh, rh = torch.nn.functional.normalize(emb0), torch.nn.functional.normalize(reps)
    d3 = opln.pln_distance(eh, rh, dist).reshape(R, K, rpc)
    clear = torch.ones(R, dtype=torch.bool, device="cuda:0")
    if rpc > 1:
        t2 = torch.topk(d3, 2, dim=2, largest=False).values
        clear = ((t2[:, :, 1] - t2[:, :, 0]) > 1e-5).all(dim=1)
    dmin = d3.min(dim=2)[0]
    fg = (pi.gt_classes >= 0) & (pi.gt_classes < K) & (pi.ious > 0.5)
    y = pi.gt_classes.clamp(0, K - 1)
    intra = dmin.gather(1, y[:, None])[:, 0]
    dm = dmin.clone(); dm.scatter_(1, y[:, None], 1000.0)
    top2 = torch.topk(dm, 2, dim=1, largest=False).values
    alpha, beta = float(intra[fg].median()), float(top2[fg, 0].median())
    safe = ((intra - alpha).abs() > 1e-5) & ((beta - top2[:, 0]).abs() > 1e-5) & ((top2[:, 1] - top2[:, 0]) > 1e-5) & clear
    ious = torch.where(safe | ~fg, pi.ious, torch.zeros_like(pi.ious))
    kw = dict(num_known_classes=K, reps_per_class=rpc, alpha=alpha, beta=beta, loss_weight=0.5, iou_threshold=0.5,
              distance_type=dist)
    ea = emb0.clone().requires_grad_(True); ra = reps.clone().requires_grad_(True)
    la = pln_loss_from_emb(ea, ra, pi.gt_classes, ious, **kw)
    (la * 1.3).backward()
    eb = emb0.clone().requires_grad_(True); rb = reps.clone().requires_grad_(True)
    lb = opln.pln_loss_from_emb(eb, rb, pi.gt_classes, ious, **kw)
    (lb * 1.3).backward()
    torch.testing.assert_close(la, lb, rtol=2e-5, atol=1e-7)
    torch.testing.assert_close(ea.grad, eb.grad, rtol=1e-3, atol=1e-7)
    torch.testing.assert_close(ra.grad, rb.grad, rtol=1e-3, atol=1e-6)
    assert float(ea.grad.abs().max()) > 0 and float(ra.grad.abs().max()) > 0
    # fused forward + backward: bit-identical to the autograd pair
    l2, ge, gr = pln_loss_fwd_bwd(emb0, reps, pi.gt_classes, ious, grad_loss=torch.full((1,), 1.3, device="cuda:0"), **kw)
    assert torch.equal(l2, la.detach()) and torch.equal(ge, ea.grad) and torch.equal(gr, ra.grad)
    # nearest prototype under the same distance
    thr = float(dmin.min(dim=1)[0].median())
    pred, md = pln_nearest(emb0, reps, num_known_classes=K, reps_per_class=rpc, unk_thr=thr, unknown_id=80, distance_type=dist)
    best, arg = dmin.min(dim=1)
    exp = torch.where(best > thr, torch.full_like(arg, 80), arg)
    t2 = torch.topk(dmin, 2, dim=1, largest=False).values
    ok = ((best - thr).abs() > 1e-5) & ((t2[:, 1] - t2[:, 0]) > 1e-5)
    assert torch.equal(pred[ok], exp[ok])
    torch.testing.assert_close(md, best, rtol=1e-5, atol=1e-6)


def test_unknown_distance_type_is_a_config_error():
    from osr_b200.pln import PLN
    with pytest.raises(ValueError, match="DISTANCE_TYPE"):
        PLN(num_classes=81, num_known_classes=20, feature_dim=64, embedding_dim=256, distance_type="L3", reps_per_class=1,
            alpha=0.1, beta=0.9, loss_weight=0.5)


@pytest.mark.parametrize("dist", ["COS", "L2"])
def test_split_forward_backward_equals_one_call(dist):
    """osr_pln_loss_fwd_bwd_phase 1 (rows) then 2 (prototype gradient + loss, here on a side stream) == osr_pln_loss_fwd_bwd."""
    from osr_b200 import synth
    from osr_b200.pln import pln_loss_fwd_bwd
    pi = synth.make_pln_inputs(4096, num_known=20, num_classes=81, seed=77, device="cuda:0")
    emb = (pi.roi_features @ pi.enc_w.t()).detach()
    kw = dict(num_known_classes=20, alpha=0.1 if dist == "COS" else 1.40, beta=0.9 if dist == "COS" else 1.33, loss_weight=0.5,
              iou_threshold=0.5, distance_type=dist)
    l0, ge0, gr0 = pln_loss_fwd_bwd(emb, pi.reps, pi.gt_classes, pi.ious, **kw)
    finish, ge1 = pln_loss_fwd_bwd(emb, pi.reps, pi.gt_classes, pi.ious, split=True, **kw)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        l1, gr1 = finish()
    torch.cuda.current_stream().wait_stream(side)
    assert torch.equal(l0, l1) and torch.equal(ge0, ge1) and torch.equal(gr0, gr1)
    assert float(gr1.abs().max()) > 0
