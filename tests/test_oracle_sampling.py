"""Oracle pins for the sampling glue: hand-checked vectors, no GPU."""
import torch

from oracle import sampling as osamp
from oracle.structures import Boxes, Instances, pairwise_iou


def _inst(boxes, logits=None, hw=(100, 100)):
    i = Instances(hw)
    i.set("proposal_boxes", Boxes(torch.tensor(boxes, dtype=torch.float32)))
    i.set("objectness_logits", torch.zeros(len(boxes)) if logits is None else torch.tensor(logits))
    return i


def _tgt(boxes, classes, hw=(100, 100)):
    t = Instances(hw)
    t.set("gt_boxes", Boxes(torch.tensor(boxes, dtype=torch.float32).reshape(-1, 4)))
    t.set("gt_classes", torch.tensor(classes, dtype=torch.int64))
    return t


def test_matcher_threshold_and_first_argmax():
    gt = Boxes(torch.tensor([[0., 0., 10., 10.], [0., 0., 10., 10.], [20., 20., 40., 40.]]))
    pr = Boxes(torch.tensor([[0., 0., 10., 10.], [0., 0., 10., 20.], [50., 50., 60., 60.], [20., 20., 40., 30.]]))
    m = pairwise_iou(gt, pr)
    idx, lab = osamp.matcher(m, 0.5)
    assert idx.tolist() == [0, 0, 0, 2]          # duplicate GT: first maximal index; no overlap: index 0
    assert lab.tolist() == [1, 1, 0, 1]          # IoU 1.0, exactly 0.5 (>= thr), 0, 0.5
    assert m[idx, torch.arange(4)].tolist() == [1.0, 0.5, 0.0, 0.5]


def test_label_and_sample_counts_and_fields():
    g = torch.Generator().manual_seed(0)
    props = torch.rand(300, 2, generator=g) * 60
    wh = torch.rand(300, 2, generator=g) * 30 + 5
    p = _inst(torch.cat((props, props + wh), 1).tolist())
    t = _tgt([[10., 10., 40., 40.], [50., 50., 80., 90.]], [3, 7])
    out = osamp.label_and_sample_proposals([p], [t], num_classes=80, batch_size_per_image=64,
                                           positive_fraction=0.25, randperm=lambda n: torch.arange(n))[0]
    gc = out.get("gt_classes")
    n_fg = int((gc != 80).sum())
    assert len(out) == 64 and 2 <= n_fg <= 16              # the two appended GT boxes are positives
    assert bool(((out.get("ious") >= 0.5) == (gc != 80)).all())
    assert out.get("gt_boxes").tensor.shape == (64, 4)
    assert abs(float(out.get("objectness_logits").max()) - osamp.GT_LOGIT) < 1e-5
    # positives come first, then negatives (cat([fg, bg]))
    first_bg = int((gc == 80).nonzero()[0])
    assert bool((gc[first_bg:] == 80).all())
