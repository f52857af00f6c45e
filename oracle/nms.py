"""Oracle for NMS.

Live call sites in the reference: ``osrcnn_fast_rcnn.py:135`` (thr 1.0),
``softmax_classifier.py:93`` and ``:154`` (thr 0.5); nominal-mode site
``find_top_proposals.py:112`` (commented out in the shipped code).

``detectron2.layers.batched_nms`` (v0.6) is
``torchvision.ops.boxes.batched_nms(boxes.float(), scores, idxs, thr)``; the real
torchvision 0.26 binary is called here.  ``nms_loops`` is the loop-level
restatement of ``torchvision.ops.nms`` used to pin the algorithm (and the two
IoU arithmetics - CPU kernel vs the FMA-contracted CUDA kernel).
"""
from __future__ import annotations

import numpy as np
import torch
import torchvision
from torchvision.ops import boxes as tv_boxes


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    return torchvision.ops.nms(boxes, scores, iou_threshold)


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """detectron2.layers.batched_nms (v0.6)."""
    assert boxes.shape[-1] == 4
    return tv_boxes.batched_nms(boxes.float(), scores, idxs, iou_threshold)


def coordinate_trick_boxes(boxes: torch.Tensor, idxs: torch.Tensor) -> torch.Tensor:
    """The offset boxes torchvision's ``_batched_nms_coordinate_trick`` feeds to ``nms``."""
    max_coordinate = boxes.max()
    offsets = idxs.to(boxes) * (max_coordinate + torch.tensor(1).to(boxes))
    return boxes + offsets[:, None]


def _f32(x):
    return np.float32(x)


def iou_gpu_arith(a: np.ndarray, b: np.ndarray) -> np.float32:
    """IoU exactly as the sm_100 SASS of torchvision's ``nms_kernel_impl<float>`` computes it.

    ``a`` = row (earlier / higher-score) box, ``b`` = column box:
    ``inter / (fma(bw, bh, Sa) - inter)`` with ``Sa = rn(aw*ah)``.
    """
    f = np.float32
    left = max(a[0], b[0]); right = min(a[2], b[2])
    top = max(a[1], b[1]); bottom = min(a[3], b[3])
    w = max(f(right - left), f(0)); h = max(f(bottom - top), f(0))
    inter = f(w * h)
    sa = f(f(a[2] - a[0]) * f(a[3] - a[1]))
    bw = f(b[2] - b[0]); bh = f(b[3] - b[1])
    # fma: exact product in float64 (24+24 bits), one rounding to fp32
    union = f(f(np.float64(bw) * np.float64(bh) + np.float64(sa)) - inter)
    with np.errstate(divide="ignore", invalid="ignore"):
        return f(inter / union)


def iou_cpu_arith(a: np.ndarray, b: np.ndarray) -> np.float32:
    """IoU as torchvision's CPU ``nms_kernel_impl`` computes it: pre-rounded areas."""
    f = np.float32
    left = max(a[0], b[0]); right = min(a[2], b[2])
    top = max(a[1], b[1]); bottom = min(a[3], b[3])
    w = max(f(right - left), f(0)); h = max(f(bottom - top), f(0))
    inter = f(w * h)
    sa = f(f(a[2] - a[0]) * f(a[3] - a[1]))
    sb = f(f(b[2] - b[0]) * f(b[3] - b[1]))
    with np.errstate(divide="ignore", invalid="ignore"):
        return f(inter / f(f(sa + sb) - inter))


def nms_loops(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float, arithmetic: str = "cpu") -> torch.Tensor:
    """Greedy NMS restated with explicit loops (small inputs only).

    order = stable sort by score descending; suppress j iff iou(i, j) > (float)thr
    (strict); NaN IoU (zero-area pairs) never suppresses; result = original
    indices in score order.
    """
    b = boxes.detach().cpu().numpy().astype(np.float32)
    s = scores.detach().cpu()
    order = torch.sort(s, descending=True, stable=True)[1].numpy()
    iou = iou_gpu_arith if arithmetic == "gpu" else iou_cpu_arith
    thr = np.float32(iou_threshold)
    n = len(order)
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    for _i in range(n):
        if suppressed[_i]:
            continue
        i = order[_i]
        keep.append(int(i))
        for _j in range(_i + 1, n):
            if suppressed[_j]:
                continue
            j = order[_j]
            if iou(b[i], b[j]) > thr:
                suppressed[_j] = True
    return torch.tensor(keep, dtype=torch.int64)
