"""The product's roofline byte model (osr_b200/roofline.py) equals the oracle's (oracle/bytes_model.py), reproduces the
SURVEY.md section 8(d) / BASELINE.md figures, and the vectorised touched-pixel count agrees with the oracle's."""
import torch

from oracle import bytes_model as ob
from oracle import roi_align as ora
from oracle.structures import Boxes
from osr_b200 import roofline as rf, synth


def test_models_equal_and_match_survey_numbers():
    grids = synth.fpn_grid_sizes(800, 1333)
    assert rf.s1_bytes_per_image(grids, 2000) == ob.s1_bytes_per_image(grids, 2000) == 621720
    assert rf.s1_bytes_per_image(grids, 1000) == ob.s1_bytes_per_image(grids, 1000) == 511920
    pooled = grids[:4]
    assert rf.s3_bwd_bytes(8192, 256, 7, 16, pooled) == ob.s3_bwd_bytes(8192, 256, 7, 16, pooled)
    assert abs(rf.s3_bwd_bytes(512, 256, 7, 1, pooled) / 1e6 - 117.1) < 0.1
    assert rf.s3_fwd_bytes(512, 256, 7, 89250) == ob.s3_fwd_bytes(512, 256, 7, 89250)
    assert abs(rf.s3_fwd_bytes(512, 256, 7, 89250) / 1e6 - 117.1) < 0.1
    assert rf.s5_fwd_bytes(8192, 1024, 256, 20) == ob.s5_fwd_bytes(8192, 1024, 256, 20)
    assert abs(rf.s5_fwd_bytes(8192, 1024, 256, 20) / 1e6 - 43.1) < 0.1
    assert rf.s5_bwd_bytes(8192, 256, 20) == ob.s5_bwd_bytes(8192, 256, 20)


def test_touched_pixels_close_to_oracle():
    rois = synth.make_rois(2, 60, (320, 480), seed=5)
    boxes = [Boxes(r) for r in rois]
    p = ora.ROIPooler(7, synth.POOL_SCALES, 0)
    lv = p.level_assignments(boxes)
    fmt = ora.convert_boxes_to_pooler_format(boxes)
    shapes = synth.fpn_grid_sizes(320, 480)[:4]
    exact = ora.touched_pixels(shapes, synth.POOL_SCALES, fmt, lv, 2)
    fast = rf.touched_pixels(shapes, synth.POOL_SCALES, fmt, lv, 2)
    # the vectorised version takes the rectangle spanned by the first..last sample (+1), the oracle the rectangle of
    # non-zero weights: they differ only where a sample sits exactly on a pixel centre
    assert abs(fast - exact) <= 0.03 * exact, (fast, exact)
