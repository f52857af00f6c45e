"""CPU tests of the RPN oracle: restatement vs the real ATen ops and closed-form expectations."""
import torch

from oracle import rpn as orpn
from oracle.structures import Boxes
from osr_b200 import synth


def test_anchor_generators_agree_and_are_exact():
    grids = synth.fpn_grid_sizes(800, 1333)
    assert grids == [(200, 336), (100, 168), (50, 84), (25, 42), (13, 21)]
    a1 = orpn.generate_anchors(grids, synth.RPN_STRIDES, synth.RPN_SIZES)
    a2 = synth.make_anchors(grids)
    for x, y in zip(a1, a2):
        assert torch.equal(x.tensor, y)
    # cell (y=3, x=5) of p3: [5*8-32, 3*8-32, 5*8+32, 3*8+32]
    assert a2[1][3 * 168 + 5].tolist() == [8.0, -8.0, 72.0, 56.0]


def test_apply_deltas_linear_known_answer():
    anchors = torch.tensor([[0.0, 0.0, 32.0, 32.0], [-16.0, 8.0, 48.0, 72.0]])
    deltas = torch.tensor([[0.5, 0.25, 1.0, -3.0], [0.0, 1.0, 0.125, 2.0]])
    out = orpn.apply_deltas_linear(deltas, anchors)
    # ctr (16,16), size 32: l=16,t=8,r=32,b=relu(-3)=0
    assert out[0].tolist() == [0.0, 8.0, 48.0, 16.0]
    # ctr (16,40), size 64: l=0,t=64,r=8,b=128
    assert out[1].tolist() == [16.0, -24.0, 24.0, 168.0]


def test_stable_topk_equals_torch_topk_when_tie_free():
    ho = synth.make_head_outputs(2, (320, 480), seed=1)
    for c in ho.centerness:
        k = min(300, c.shape[1])
        v1, i1 = c.topk(k, dim=1)
        v2, i2 = orpn.topk_stable(c, k)
        assert torch.equal(v1, v2) and torch.equal(i1, i2)


def test_as_shipped_counts_at_800x1333():
    """7 323 pre-filter proposals/img (train), 4 273 (test)."""
    ho = synth.make_head_outputs(1, (800, 1333), seed=2, neg_frac=0.0)
    anchors = [Boxes(a) for a in ho.anchors]
    r = orpn.predict_proposals(anchors, ho.deltas, ho.centerness, ho.image_sizes, pre_nms_topk=2000, training=True)
    assert len(r[0]) <= 7323
    ho2 = synth.make_head_outputs(1, (800, 1333), seed=2, neg_frac=0.0)
    # with no negated deltas every box has positive extent before clipping; clipping can still empty a few
    r2 = orpn.predict_proposals(anchors, ho2.deltas, ho2.centerness, ho2.image_sizes, pre_nms_topk=1000, training=False)
    assert 4000 < len(r2[0]) <= 4273
    lv = r2[0].level_ids
    assert (lv[1:] >= lv[:-1]).all()          # level-major order
    for l in range(5):
        s = r2[0].objectness_logits[lv == l]
        assert (s[1:] <= s[:-1]).all()        # score-descending inside a level


def test_nonfinite_behaviour():
    ho = synth.make_head_outputs(2, (320, 480), seed=9, nonfinite=3)
    anchors = [Boxes(a) for a in ho.anchors]
    import pytest
    with pytest.raises(FloatingPointError):
        orpn.predict_proposals(anchors, ho.deltas, ho.centerness, ho.image_sizes, pre_nms_topk=300, training=True)
    r = orpn.predict_proposals(anchors, ho.deltas, ho.centerness, ho.image_sizes, pre_nms_topk=300, training=False)
    for inst in r:
        assert torch.isfinite(inst.proposal_boxes.tensor).all()


def test_nominal_mode_runs_and_is_sorted():
    ho = synth.make_head_outputs(1, (320, 480), seed=4)
    anchors = [Boxes(a) for a in ho.anchors]
    r = orpn.predict_proposals(anchors, ho.deltas, ho.centerness, ho.image_sizes, pre_nms_topk=300,
                               post_nms_topk=200, nms_thresh=0.7, training=True, mode="nominal")
    s = r[0].objectness_logits
    assert len(s) <= 200 and (s[1:] <= s[:-1]).all()
