"""Committed fixtures (tests/golden/golden_v1.npz, produced by tests/golden/make_golden.py from the real torchvision /
ATen CPU binaries): the oracle must reproduce them on the CPU, the CUDA kernels must match them on the GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import nms as onms, pln as opln, roi_align as ora, rpn as orpn
from oracle.structures import Boxes
from osr_b200 import synth

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))
HW_ROI, HW_RPN = (128, 160), (96, 128)
KW = dict(num_known_classes=20, alpha=0.1, beta=0.9, loss_weight=0.5, iou_threshold=0.5)


def t(name):
    return torch.from_numpy(G[name])


# ------------------------------------------------------------------------------------------------ CPU: oracle
def test_oracle_roi_align_reproduces_golden():
    feats = [t(f"roi_feat{l}").requires_grad_(True) for l in range(4)]
    boxes = [Boxes(t("roi_rois0")), Boxes(t("roi_rois1"))]
    p = ora.ROIPooler(7, synth.POOL_SCALES, 0)
    out = p.forward(feats, boxes)
    assert torch.equal(out.detach(), t("roi_pooled"))
    assert torch.equal(p.level_assignments(boxes), t("roi_levels"))
    grads = torch.autograd.grad(out, feats, t("roi_gout"))
    for l in range(4):
        torch.testing.assert_close(grads[l], t(f"roi_grad{l}"), rtol=1e-6, atol=1e-6)
    # loop-level restatement of the torchvision kernel agrees with the binary's golden output (per level)
    fmt = ora.convert_boxes_to_pooler_format(boxes)
    lv = t("roi_levels")
    for l in range(4):
        m = lv == l
        if int(m.sum()) == 0:
            continue
        loops = ora.roi_align_loops(feats[l].detach(), fmt[m], 7, synth.POOL_SCALES[l], 0)
        torch.testing.assert_close(loops, t("roi_pooled")[m], rtol=1e-5, atol=1e-5)


def test_oracle_nms_reproduces_golden():
    b, s, i = t("nms_boxes"), t("nms_scores"), t("nms_idxs")
    for thr in (0.5, 0.7, 1.0):
        assert torch.equal(onms.nms(b, s, thr), t(f"nms_keep_{thr}"))
        assert torch.equal(onms.nms_loops(b, s, thr, "cpu"), t(f"nms_keep_{thr}"))
    assert torch.equal(onms.batched_nms(b, s, i, 0.5), t("nms_batched_keep_0.5"))


def test_oracle_rpn_reproduces_golden():
    ho = synth.make_head_outputs(2, HW_RPN, seed=21, mixed_sizes=True)
    anchors = [Boxes(a) for a in ho.anchors]
    props = orpn.predict_proposals(anchors, ho.deltas, ho.centerness, ho.image_sizes, pre_nms_topk=60, post_nms_topk=60,
                                   training=False, topk_impl="stable")
    for n, p in enumerate(props):
        assert torch.equal(p.proposal_boxes.tensor, t(f"rpn_boxes{n}"))
        assert torch.equal(p.objectness_logits, t(f"rpn_scores{n}"))
        assert torch.equal(p.level_ids, t(f"rpn_levels{n}"))


def test_oracle_pln_reproduces_golden():
    emb = t("pln_emb").requires_grad_(True)
    reps = t("pln_reps").requires_grad_(True)
    loss = opln.pln_loss_from_emb(emb, reps, t("pln_labels"), t("pln_ious"), **KW)
    torch.testing.assert_close(loss.detach(), t("pln_loss"), rtol=1e-6, atol=1e-8)
    ge, gr = torch.autograd.grad(loss, [emb, reps])
    torch.testing.assert_close(ge, t("pln_grad_emb"), rtol=1e-5, atol=1e-9)
    ge2, gr2 = opln.pln_loss_grad_closed_form(t("pln_emb").double(), t("pln_reps").double(), t("pln_labels"),
                                              t("pln_ious").double(), **KW)
    torch.testing.assert_close(ge2.float(), t("pln_grad_emb"), rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(gr2.float(), t("pln_grad_reps"), rtol=1e-4, atol=1e-7)


# ------------------------------------------------------------------------------------------------ GPU: kernels
@pytest.mark.gpu
@pytest.mark.parametrize("channels_last", [False, True])
def test_gpu_roi_align_matches_golden(channels_last):
    from osr_b200.poolers import ROIPooler
    fmt = torch.channels_last if channels_last else torch.contiguous_format
    feats = [t(f"roi_feat{l}").cuda().contiguous(memory_format=fmt).requires_grad_(True) for l in range(4)]
    boxes = [Boxes(t("roi_rois0").cuda()), Boxes(t("roi_rois1").cuda())]
    p = ROIPooler(7, synth.POOL_SCALES, 0, "ROIAlignV2")
    out, lvl = p.forward_with_levels(feats, boxes)
    assert torch.equal(lvl.cpu().long(), t("roi_levels"))
    torch.testing.assert_close(out.detach().cpu(), t("roi_pooled"), rtol=1e-5, atol=1e-4)
    grads = torch.autograd.grad(out, feats, t("roi_gout").cuda())
    for l in range(4):
        ref = t(f"roi_grad{l}")
        torch.testing.assert_close(grads[l].cpu(), ref, rtol=1e-4, atol=1e-4 * max(1.0, float(ref.abs().max())))


@pytest.mark.gpu
def test_gpu_nms_matches_golden():
    from osr_b200.nms import batched_nms, nms
    b, s, i = t("nms_boxes").cuda(), t("nms_scores").cuda(), t("nms_idxs").cuda()
    for thr in (0.5, 0.7, 1.0):
        assert torch.equal(nms(b, s, thr).cpu(), t(f"nms_keep_{thr}"))
    assert torch.equal(batched_nms(b, s, i, 0.5).cpu(), t("nms_batched_keep_0.5"))


@pytest.mark.gpu
def test_gpu_rpn_matches_golden():
    from osr_b200 import proposals as P
    ho = synth.make_head_outputs(2, HW_RPN, seed=21, mixed_sizes=True)
    kw = dict(pre_nms_topk=60, post_nms_topk=60, training=False)
    props = P.predict_proposals([a.cuda() for a in ho.anchors], [d.cuda() for d in ho.deltas],
                                [c.cuda() for c in ho.centerness], ho.image_sizes, **kw)
    for n, p in enumerate(props):
        assert torch.equal(p.proposal_boxes.tensor.cpu(), t(f"rpn_boxes{n}"))
        assert torch.equal(p.objectness_logits.cpu(), t(f"rpn_scores{n}"))
    nom = P.predict_proposals([a.cuda() for a in ho.anchors], [d.cuda() for d in ho.deltas],
                              [c.cuda() for c in ho.centerness], ho.image_sizes, nms_thresh=0.7, pre_nms_topk=60,
                              post_nms_topk=40, training=False, mode="nominal")
    for n, p in enumerate(nom):
        assert torch.equal(p.proposal_boxes.tensor.cpu(), t(f"rpn_nominal_boxes{n}"))
        assert torch.equal(p.objectness_logits.cpu(), t(f"rpn_nominal_scores{n}"))


@pytest.mark.gpu
def test_gpu_pln_matches_golden():
    from osr_b200.pln import pln_loss_from_emb
    emb = t("pln_emb").cuda().requires_grad_(True)
    reps = t("pln_reps").cuda().requires_grad_(True)
    loss = pln_loss_from_emb(emb, reps, t("pln_labels").cuda(), t("pln_ious").cuda(), **KW)
    torch.testing.assert_close(loss.detach().cpu(), t("pln_loss"), rtol=1e-5, atol=1e-7)
    ge, gr = torch.autograd.grad(loss, [emb, reps])
    torch.testing.assert_close(ge.cpu(), t("pln_grad_emb"), rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(gr.cpu(), t("pln_grad_reps"), rtol=1e-4, atol=1e-6)
