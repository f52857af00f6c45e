"""Oracle pins for the ROI-head inference post-processing: closed-form vectors, CPU only."""
import math

import torch

from oracle import rcnn_inference as oinf
from oracle.structures import Boxes, Instances


def test_apply_deltas_known_answers():
    boxes = torch.tensor([[10., 20., 30., 60.]])              # w = 20, h = 40, centre (20, 40)
    zero = oinf.apply_deltas(torch.zeros(1, 4), boxes)
    assert torch.equal(zero, boxes)
    d = torch.tensor([[10., -5., 5. * math.log(2.0), 0.]])    # dx = 1 -> +w, dy = -0.5 -> -h/2, dw = log 2 -> 2w
    out = oinf.apply_deltas(d, boxes)
    torch.testing.assert_close(out, torch.tensor([[20., 0., 60., 40.]]), rtol=0, atol=1e-4)
    big = oinf.apply_deltas(torch.tensor([[0., 0., 1000., 1000.]]), boxes)   # clamped at log(1000/16)
    torch.testing.assert_close(big[0, 2] - big[0, 0], torch.tensor(20. * 1000. / 16.), rtol=1e-5, atol=0)


def test_single_image_filter_nms_topk():
    boxes = torch.tensor([[0., 0., 10., 10.], [1., 1., 11., 11.], [50., 50., 60., 60.], [float("nan"), 0., 1., 1.]])
    scores = torch.tensor([[0.9], [0.8], [0.7], [0.99]])
    feats = torch.arange(4.)[:, None]
    res, idx = oinf.fast_rcnn_inference_single_image(boxes, scores, (100, 100), feats, 0.75, 0.5, 10)
    # NaN row removed first; 0.7 is below the threshold; box 1 overlaps box 0 with IoU 0.68 > 0.5 -> suppressed
    assert idx.tolist() == [0]
    assert res.get("scores").tolist() == [0.9] or abs(float(res.get("scores")[0]) - 0.9) < 1e-6
    assert res.get("features")[:, 0].tolist() == [0.0]
