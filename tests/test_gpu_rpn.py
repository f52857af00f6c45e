"""GPU parity: osr_rpn_select_decode (through the C ABI / python drop-ins) vs the oracle.

Bit-exact classes: selected anchor indices, decoded+clipped boxes, scores, per-level counts.
"""
import pytest
import torch

from oracle import rpn as orpn
from oracle.structures import Boxes as OBoxes

pytestmark = pytest.mark.gpu


def _oracle(ho, pre_k, training, device="cpu", min_box_size=0.0):
    anchors = [OBoxes(a.to(device)) for a in ho.anchors]
    return orpn.predict_proposals(
        anchors, [d.to(device) for d in ho.deltas], [c.to(device) for c in ho.centerness], ho.image_sizes,
        pre_nms_topk=pre_k, post_nms_topk=pre_k, min_box_size=min_box_size, training=training, mode="as_shipped")


def _ours(ho, pre_k, training, min_box_size=0.0, **kw):
    from osr_b200 import proposals as P
    dev = "cuda:0"
    return P.predict_proposals(
        [a.to(dev) for a in ho.anchors], [d.to(dev) for d in ho.deltas], [c.to(dev) for c in ho.centerness],
        ho.image_sizes, pre_nms_topk=pre_k, post_nms_topk=pre_k, min_box_size=min_box_size, training=training, **kw)


def _assert_same(ours, ref):
    assert len(ours) == len(ref)
    for o, r in zip(ours, ref):
        assert tuple(o.image_size) == tuple(r.image_size)
        assert len(o) == len(r), (len(o), len(r))
        assert torch.equal(o.proposal_boxes.tensor.cpu(), r.proposal_boxes.tensor.cpu())
        assert torch.equal(o.objectness_logits.cpu(), r.objectness_logits.cpu())


@pytest.mark.parametrize("hw,pre_k,n", [((800, 1333), 2000, 3), ((800, 1333), 1000, 2), ((320, 480), 300, 4),
                                        ((64, 96), 50, 2)])
def test_tie_free_bit_exact(hw, pre_k, n):
    from osr_b200 import synth
    ho = synth.make_head_outputs(n, hw, seed=11, mixed_sizes=True)
    ref = _oracle(ho, pre_k, training=True)
    ours = _ours(ho, pre_k, training=True)
    _assert_same(ours, ref)


def test_matches_torch_topk_on_gpu_tie_free():
    """Same inputs through the oracle *on the GPU* with torch.topk (the reference's own op)."""
    from osr_b200 import synth
    ho = synth.make_head_outputs(2, (800, 1333), seed=5)
    dev = "cuda:0"
    anchors = [OBoxes(a.to(dev)) for a in ho.anchors]
    ref = orpn.predict_proposals(anchors, [d.to(dev) for d in ho.deltas], [c.to(dev) for c in ho.centerness],
                                 ho.image_sizes, pre_nms_topk=2000, training=True, topk_impl="torch")
    ours = _ours(ho, 2000, training=True)
    _assert_same(ours, ref)


def test_tie_heavy_matches_stable_contract():
    """Ties: contract = score desc, lower anchor index first (oracle topk_stable)."""
    from osr_b200 import synth
    ho = synth.make_head_outputs(2, (480, 640), seed=3, ties="heavy")
    ref = _oracle(ho, 500, training=False)
    ours = _ours(ho, 500, training=False)
    _assert_same(ours, ref)


def test_selection_indices_and_counts():
    from osr_b200 import proposals as P, synth
    ho = synth.make_head_outputs(2, (800, 1333), seed=21)
    dev = "cuda:0"
    sel = P.rpn_select_decode([a.to(dev) for a in ho.anchors], [d.to(dev) for d in ho.deltas],
                              [c.to(dev) for c in ho.centerness], ho.image_sizes, 1000)
    counts = sel.counts.cpu()
    L = sel.num_levels
    assert (counts[:, :L].sum(1) == counts[:, L]).all()
    assert (counts[:, L + 1] == 0).all()
    ref = _oracle(ho, 1000, training=False)
    for n in range(2):
        c = int(counts[n, L])
        assert c == len(ref[n])
        lv = sel.level[n, :c].cpu().long()
        assert torch.equal(lv, ref[n].level_ids)
        # index check: score of the flat anchor index equals the output score
        for l in range(L):
            m = lv == l
            idx = sel.index[n, :c].cpu().long()[m]
            assert torch.equal(ho.centerness[l][n][idx], sel.scores[n, :c].cpu()[m])


def test_nonfinite_training_raises_eval_drops():
    from osr_b200 import synth
    ho = synth.make_head_outputs(2, (320, 480), seed=9, nonfinite=3)
    with pytest.raises(FloatingPointError):
        _ours(ho, 300, training=True)
    ref = _oracle(ho, 300, training=False)
    ours = _ours(ho, 300, training=False)
    _assert_same(ours, ref)


def test_min_box_size_filter():
    from osr_b200 import synth
    ho = synth.make_head_outputs(2, (320, 480), seed=10)
    ref = _oracle(ho, 300, training=False, min_box_size=16.0)
    ours = _ours(ho, 300, training=False, min_box_size=16.0)
    _assert_same(ours, ref)


def test_find_top_rpn_proposals_signature_predecoded():
    from osr_b200 import proposals as P, synth
    ho = synth.make_head_outputs(2, (320, 480), seed=12)
    anchors = [OBoxes(a) for a in ho.anchors]
    dec = orpn.decode_proposals(anchors, ho.deltas)
    ref = orpn.find_top_rpn_proposals(dec, ho.centerness, ho.image_sizes, 1.0, 300, 300, 0.0, True)
    ours = P.find_top_rpn_proposals([d.cuda() for d in dec], [c.cuda() for c in ho.centerness], ho.image_sizes,
                                    1.0, 300, 300, 0.0, True)
    _assert_same(ours, ref)


def test_raw_conv_layout_strided_view():
    """Head outputs in the raw (N, A*4, H, W) / (N, A, H, W) conv layout, A=1, consumed as permuted views
    (no copy) - classification_free_rpn.py:518-529 makes contiguous copies instead."""
    from osr_b200 import proposals as P, synth
    ho = synth.make_head_outputs(2, (320, 480), seed=13)
    dev = "cuda:0"
    deltas_v, ctr_v = [], []
    for (h, w), d, c in zip(ho.grid_sizes, ho.deltas, ho.centerness):
        raw = d.view(2, h, w, 4).permute(0, 3, 1, 2).contiguous().to(dev)          # (N, 4, H, W)
        deltas_v.append(raw.permute(0, 2, 3, 1).flatten(1, 2))                         # strided view (N, HW, 4)
        ctr_v.append(c.to(dev))
        assert not deltas_v[-1].is_contiguous()
    ours = P.predict_proposals([a.to(dev) for a in ho.anchors], deltas_v, ctr_v, ho.image_sizes,
                               pre_nms_topk=300, training=True)
    ref = _oracle(ho, 300, training=True)
    _assert_same(ours, ref)


def test_deterministic_run_to_run():
    from osr_b200 import synth
    ho = synth.make_head_outputs(2, (800, 1333), seed=31, ties="heavy")
    a = _ours(ho, 2000, training=False)
    b = _ours(ho, 2000, training=False)
    _assert_same(a, b)
