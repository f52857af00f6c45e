"""Full-size parity at the BASELINE.json configurations (cfg 2-5 shapes), on the GPU.

At these sizes the CPU oracle would take minutes, so the checker is the oracle's code run ON THE GPU (the reference's
own ops there: torch.topk, torchvision CUDA roi_align / nms, autograd) plus size-independent properties
(adjoint identity, run-to-run bit-identity).  Bars: proposals / keep indices / level ids bit-exact; ROIAlign features
rtol 1e-5 atol 2e-5; gradients rtol 1e-4 atol 1e-4 * max|g|; PLN loss rtol 1e-5 (fp32 encoder).
"""
import pytest
import torch

from oracle import pln as opln
from oracle import roi_align as ora
from oracle import rpn as orpn
from oracle.structures import Boxes as OBoxes

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _proposals_pair(ho, pre_k, training, mode="as_shipped", post_k=None, nms_thresh=0.7):
    from osr_b200 import proposals as P
    kw = dict(pre_nms_topk=pre_k, post_nms_topk=post_k or pre_k, training=training, mode=mode, nms_thresh=nms_thresh)
    ours = P.predict_proposals([a.to(DEV) for a in ho.anchors], [d.to(DEV) for d in ho.deltas],
                               [c.to(DEV) for c in ho.centerness], ho.image_sizes, **kw)
    ref = orpn.predict_proposals([OBoxes(a.to(DEV)) for a in ho.anchors], [d.to(DEV) for d in ho.deltas],
                                 [c.to(DEV) for c in ho.centerness], ho.image_sizes, topk_impl="torch", **kw)
    for o, r in zip(ours, ref):
        assert len(o) == len(r)
        assert torch.equal(o.proposal_boxes.tensor, r.proposal_boxes.tensor)
        assert torch.equal(o.objectness_logits, r.objectness_logits)
    return ours


def _poolers():
    from osr_b200.poolers import ROIPooler
    from osr_b200 import synth
    return ROIPooler(7, synth.POOL_SCALES, 0, "ROIAlignV2"), ora.ROIPooler(7, synth.POOL_SCALES, 0, "ROIAlignV2")


def _check_pool(feats, boxes, *, backward):
    ours, ref = _poolers()
    fa = [f.detach().clone().requires_grad_(backward) for f in feats]
    fb = [f.detach().contiguous().clone().requires_grad_(backward) for f in feats]
    ob = [OBoxes(b) for b in boxes]
    out, lvl = ours.forward_with_levels(fa, ob)
    exp = ref.forward(fb, ob)
    assert torch.equal(lvl.long(), ref.level_assignments(ob))
    torch.testing.assert_close(out, exp, rtol=1e-5, atol=2e-5)
    if not backward:
        return
    gout = torch.randn(out.shape, device=DEV, generator=torch.Generator(DEV).manual_seed(3))
    ga = torch.autograd.grad(out, fa, gout)
    gb = torch.autograd.grad(exp, fb, gout)
    for a, b in zip(ga, gb):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-4 * max(1.0, float(b.abs().max())))
    lhs = (out.double() * gout.double()).sum()
    rhs = sum((f.double() * g.double()).sum() for f, g in zip(fa, ga))
    assert abs(lhs - rhs) <= 1e-5 * max(abs(lhs), 1.0) + 1e-2


def test_cfg2_training_step_full_size():
    """cfg 2: 16 images 800x1333, k = 2000 (7 323 proposals/img as shipped), 512 RoIs/img, K = 20."""
    from osr_b200 import synth
    from osr_b200.pln import pln_loss_from_emb
    ho = synth.make_head_outputs(16, (800, 1333), seed=1234)
    props = _proposals_pair(ho, 2000, training=True)
    assert all(len(p) <= 7323 for p in props) and sum(len(p) for p in props) > 16 * 6000
    g = torch.Generator().manual_seed(1)
    boxes = [p.proposal_boxes.tensor[torch.randperm(len(p), generator=g)[:512].to(DEV)] for p in props]
    feats = synth.make_features(16, (800, 1333), 256, seed=2, device=DEV, channels_last=True)
    _check_pool(feats, boxes, backward=True)
    pi = synth.make_pln_inputs(16 * 512, seed=3, device=DEV)
    emb = (pi.roi_features @ pi.enc_w.t() + pi.enc_b).requires_grad_(True)
    kw = dict(num_known_classes=20, alpha=0.1, beta=0.9, loss_weight=0.5, iou_threshold=0.5)
    la = pln_loss_from_emb(emb, pi.reps.clone().requires_grad_(True), pi.gt_classes, pi.ious, **kw)
    lb = opln.pln_loss_from_emb(emb.detach(), pi.reps, pi.gt_classes, pi.ious, **kw)
    torch.testing.assert_close(la, lb, rtol=1e-5, atol=1e-7)


def test_cfg3_graspnet_shapes():
    """cfg 3 per-GPU shard: 8 images 750x1333 (padded 768x1344), K = 28 known classes."""
    from osr_b200 import synth
    from osr_b200.pln import pln_loss_from_emb
    ho = synth.make_head_outputs(8, (750, 1333), seed=77)
    props = _proposals_pair(ho, 2000, training=True)
    g = torch.Generator().manual_seed(2)
    boxes = [p.proposal_boxes.tensor[torch.randperm(len(p), generator=g)[:512].to(DEV)] for p in props]
    feats = synth.make_features(8, (750, 1333), 256, seed=4, device=DEV, channels_last=True)
    _check_pool(feats, boxes, backward=True)
    pi = synth.make_pln_inputs(8 * 512, num_known=28, seed=5, device=DEV)
    emb = (pi.roi_features @ pi.enc_w.t()).requires_grad_(True)
    kw = dict(num_known_classes=28, alpha=0.1, beta=0.9, loss_weight=1.0, iou_threshold=0.5)
    la = pln_loss_from_emb(emb, pi.reps.clone().requires_grad_(True), pi.gt_classes, pi.ious, **kw)
    lb = opln.pln_loss_from_emb(emb.detach(), pi.reps, pi.gt_classes, pi.ious, **kw)
    torch.testing.assert_close(la, lb, rtol=1e-5, atol=1e-7)


def test_cfg4_inference_batch32_nominal_nms():
    """cfg 4: 32 images eval, k = 1000, NMS 0.7 -> <= 1000 proposals/img (nominal mode), ROIAlign forward of all."""
    from osr_b200 import synth
    ho = synth.make_head_outputs(32, (800, 1333), seed=404)
    props = _proposals_pair(ho, 1000, training=False, mode="nominal", post_k=1000)
    assert all(len(p) <= 1000 for p in props)
    boxes = [p.proposal_boxes.tensor for p in props]
    for cl in (True, False):
        feats = synth.make_features(32, (800, 1333), 256, seed=6, device=DEV, channels_last=cl)
        _check_pool(feats, boxes, backward=False)
        del feats
        torch.cuda.empty_cache()


def test_cfg5_stress_1333_square_k4000_1024_rois():
    """cfg 5 shapes: 1333x1333 (padded 1344^2), 4000 pre-NMS proposals per level, 1024 RoIs/img (4 images here)."""
    from osr_b200 import synth
    ho = synth.make_head_outputs(4, (1333, 1333), seed=55)
    props = _proposals_pair(ho, 4000, training=True)
    assert max(len(p) for p in props) > 10000
    g = torch.Generator().manual_seed(3)
    boxes = [p.proposal_boxes.tensor[torch.randperm(len(p), generator=g)[:1024].to(DEV)] for p in props]
    feats = synth.make_features(4, (1333, 1333), 256, seed=7, device=DEV, channels_last=True)
    _check_pool(feats, boxes, backward=True)
    # backward is run-to-run bit-identical at this size too
    ours, _ = _poolers()
    ob = [OBoxes(b) for b in boxes]
    gout = torch.randn(4 * 1024, 256, 7, 7, device=DEV)
    grads = []
    for _ in range(2):
        fa = [f.detach().clone().requires_grad_(True) for f in feats]
        grads.append(torch.autograd.grad(ours.forward(fa, ob), fa, gout))
    for a, b in zip(*grads):
        assert torch.equal(a, b)
