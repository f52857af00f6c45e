"""CPU pins of the ROIAlign oracle: loop restatement and separable regrouping vs the real torchvision
binary, plus the known-answer vectors of SURVEY.md Appendix C."""
import numpy as np
import torch
import torchvision

from oracle import roi_align as ora
from oracle.structures import Boxes


def test_appendix_c1_known_answer():
    f = torch.arange(25, dtype=torch.float32).view(1, 1, 5, 5)
    r = torch.tensor([[0.0, 1.0, 1.0, 3.0, 3.0]])
    exp = torch.tensor([[4.5, 5, 5.5, 6], [7, 7.5, 8, 8.5], [9.5, 10, 10.5, 11], [12, 12.5, 13, 13.5]])
    got = torchvision.ops.roi_align(f, r, (4, 4), 1.0, 0, True)[0, 0]
    assert torch.allclose(got, exp)


def test_loops_and_separable_match_torchvision():
    g = torch.Generator().manual_seed(0)
    feat = torch.randn(2, 3, 20, 30, generator=g)
    rois = torch.tensor([
        [0, 2.0, 3.0, 50.0, 40.0], [1, 0.0, 0.0, 120.0, 80.0], [0, 10.0, 10.0, 10.4, 10.4], [1, 100.0, 60.0, 119.0, 79.0],
        [0, -20.0, -20.0, 30.0, 30.0], [1, 5.0, 5.0, 5.0, 5.0], [0, 110.0, 70.0, 200.0, 140.0], [1, 7.0, 1.0, 9.0, 78.0],
    ])
    tv = torchvision.ops.roi_align(feat, rois, (7, 7), 0.25, 0, True)
    loops = ora.roi_align_loops(feat, rois, 7, 0.25, 0)
    sep = ora.roi_align_separable(feat, rois, 7, 0.25, 0)
    torch.testing.assert_close(loops, tv, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(sep, tv, rtol=1e-5, atol=1e-5)


def test_level_assignment_table():
    """Appendix C.3."""
    sizes = [0, 1, 111.99, 112, 223.99, 224, 447.99, 448, 1333]
    boxes = Boxes(torch.tensor([[0.0, 0.0, s, s] for s in sizes]))
    lv = ora.assign_boxes_to_levels([boxes], 2, 5, 224, 4)
    assert lv.tolist() == [0, 0, 0, 1, 1, 2, 2, 3, 3]


def test_pooler_matches_per_level_calls():
    from osr_b200 import synth
    feats = synth.make_features(2, (224, 320), 8, seed=1)
    rois = synth.make_rois(2, 20, (224, 320), seed=2)
    p = ora.ROIPooler(7, synth.POOL_SCALES, 0)
    out = p.forward(feats, [Boxes(r) for r in rois])
    lv = p.level_assignments([Boxes(r) for r in rois])
    fmt = ora.convert_boxes_to_pooler_format([Boxes(r) for r in rois])
    for m in range(fmt.shape[0]):
        l = int(lv[m])
        one = torchvision.ops.roi_align(feats[l], fmt[m:m + 1], (7, 7), synth.POOL_SCALES[l], 0, True)
        assert torch.equal(one[0], out[m])


def test_touched_pixels_bounds():
    from osr_b200 import synth
    rois = synth.make_rois(1, 30, (224, 320), seed=3)
    p = ora.ROIPooler(7, synth.POOL_SCALES, 0)
    lv = p.level_assignments([Boxes(r) for r in rois])
    fmt = ora.convert_boxes_to_pooler_format([Boxes(r) for r in rois])
    shapes = synth.fpn_grid_sizes(224, 320)[:4]
    u = ora.touched_pixels(shapes, synth.POOL_SCALES, fmt, lv, 1)
    assert 0 < u <= sum(h * w for h, w in shapes)
