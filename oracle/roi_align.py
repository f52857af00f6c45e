"""Oracle for FPN level assignment + multi-level ROIAlignV2.

Call site in the reference: ``osrcnn_roi_heads.py:306`` (``self.box_pooler(features,
[x.proposal_boxes for x in proposals])``), pooler built at ``:108-113`` with
``ROIAlignV2``, 7x7, scales 1/4..1/32, ``sampling_ratio=0``.

Restates detectron2 v0.6 ``ROIPooler.forward`` / ``assign_boxes_to_levels`` /
``convert_boxes_to_pooler_format`` and calls the real
``torchvision.ops.roi_align(aligned=True)`` binary for the kernel.
``roi_align_loops`` is the loop-level restatement of that kernel
(torchvision ``roi_align_kernel.cpp``) used to pin it; ``separable_weights`` is the
exact algebraic regrouping the CUDA kernel uses (tested equal to the loops).
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np
import torch
import torchvision

from .structures import Boxes


def convert_boxes_to_pooler_format(box_lists: List[Boxes]) -> torch.Tensor:
    """detectron2 ``convert_boxes_to_pooler_format``: (M,5) = [batch_idx, x1, y1, x2, y2]."""
    boxes = torch.cat([x.tensor for x in box_lists], dim=0)
    sizes = torch.tensor([len(x) for x in box_lists], device=boxes.device)
    indices = torch.repeat_interleave(torch.arange(len(box_lists), dtype=boxes.dtype, device=boxes.device), sizes)
    return torch.cat([indices[:, None], boxes], dim=1)


def assign_boxes_to_levels(
    box_lists: List[Boxes],
    min_level: int,
    max_level: int,
    canonical_box_size: int,
    canonical_level: int,
) -> torch.Tensor:
    """detectron2 ``assign_boxes_to_levels``: floor(canonical_level + log2(sqrt(area)/size + 1e-8)),
    clamped to [min_level, max_level], minus min_level.  All ops fp32; eps inside the log."""
    box_sizes = torch.sqrt(torch.cat([boxes.area() for boxes in box_lists]))
    level_assignments = torch.floor(canonical_level + torch.log2(box_sizes / canonical_box_size + 1e-8))
    level_assignments = torch.clamp(level_assignments, min=min_level, max=max_level)
    return level_assignments.to(torch.int64) - min_level


class ROIPooler:
    """detectron2 ``ROIPooler`` restricted to ``pooler_type='ROIAlignV2'``."""

    def __init__(self, output_size, scales: Sequence[float], sampling_ratio: int = 0,
                 pooler_type: str = "ROIAlignV2", canonical_box_size: int = 224, canonical_level: int = 4):
        if isinstance(output_size, int):
            output_size = (output_size, output_size)
        assert pooler_type == "ROIAlignV2"
        self.output_size = output_size
        self.scales = list(scales)
        self.sampling_ratio = sampling_ratio
        min_level = -(math.log2(scales[0]))
        max_level = -(math.log2(scales[-1]))
        assert math.isclose(min_level, int(min_level)) and math.isclose(max_level, int(max_level))
        self.min_level = int(min_level)
        self.max_level = int(max_level)
        assert len(scales) == self.max_level - self.min_level + 1
        self.canonical_level = canonical_level
        self.canonical_box_size = canonical_box_size

    def level_assignments(self, box_lists: List[Boxes]) -> torch.Tensor:
        return assign_boxes_to_levels(box_lists, self.min_level, self.max_level,
                                      self.canonical_box_size, self.canonical_level)

    def forward(self, x: List[torch.Tensor], box_lists: List[Boxes]) -> torch.Tensor:
        num_level_assignments = len(self.scales)
        assert len(x) == num_level_assignments
        assert len(box_lists) == x[0].size(0)
        if len(box_lists) == 0:
            return torch.zeros((0, x[0].shape[1]) + tuple(self.output_size), device=x[0].device, dtype=x[0].dtype)
        pooler_fmt_boxes = convert_boxes_to_pooler_format(box_lists)
        if num_level_assignments == 1:
            return self._level(x[0], pooler_fmt_boxes, self.scales[0])
        level_assignments = self.level_assignments(box_lists)
        num_boxes = pooler_fmt_boxes.size(0)
        num_channels = x[0].shape[1]
        output = torch.zeros((num_boxes, num_channels, self.output_size[0], self.output_size[1]),
                             dtype=x[0].dtype, device=x[0].device)
        for level, scale in enumerate(self.scales):
            inds = torch.nonzero(level_assignments == level, as_tuple=True)[0]
            pooler_fmt_boxes_level = pooler_fmt_boxes[inds]
            output.index_put_((inds,), self._level(x[level], pooler_fmt_boxes_level, scale))
        return output

    __call__ = forward

    def _level(self, feat, rois, scale):
        # detectron2 ROIAlign(aligned=True).forward -> torchvision.ops.roi_align
        return torchvision.ops.roi_align(feat, rois.to(dtype=feat.dtype), self.output_size, scale,
                                         self.sampling_ratio, True)


# ---------------------------------------------------------------------------
# loop-level restatement of torchvision's roi_align CPU/CUDA kernel (aligned=True)
# ---------------------------------------------------------------------------

def _bilinear_terms(H: int, W: int, y: float, x: float):
    """torchvision ``bilinear_interpolate``: returns [(yy, xx, w)] * 4 or [] if out of range."""
    f = np.float32
    if y < -1.0 or y > H or x < -1.0 or x > W:
        return []
    y = f(max(y, f(0))); x = f(max(x, f(0)))
    y_low = int(y); x_low = int(x)
    if y_low >= H - 1:
        y_high = y_low = H - 1; y = f(y_low)
    else:
        y_high = y_low + 1
    if x_low >= W - 1:
        x_high = x_low = W - 1; x = f(x_low)
    else:
        x_high = x_low + 1
    ly = f(y - f(y_low)); lx = f(x - f(x_low))
    hy = f(f(1) - ly); hx = f(f(1) - lx)
    return [(y_low, x_low, f(hy * hx)), (y_low, x_high, f(hy * lx)),
            (y_high, x_low, f(ly * hx)), (y_high, x_high, f(ly * lx))]


def _roi_geometry(roi: np.ndarray, scale: float, P: int, sampling_ratio: int):
    f = np.float32
    off = f(0.5)
    sw = f(f(roi[0]) * f(scale)) - off
    sh = f(f(roi[1]) * f(scale)) - off
    ew = f(f(roi[2]) * f(scale)) - off
    eh = f(f(roi[3]) * f(scale)) - off
    roi_w = f(ew - sw); roi_h = f(eh - sh)
    bin_h = f(roi_h / f(P)); bin_w = f(roi_w / f(P))
    gh = sampling_ratio if sampling_ratio > 0 else int(math.ceil(roi_h / f(P)))
    gw = sampling_ratio if sampling_ratio > 0 else int(math.ceil(roi_w / f(P)))
    count = f(max(gh * gw, 1))
    return sw, sh, bin_w, bin_h, gw, gh, count


def roi_align_loops(feat: torch.Tensor, rois: torch.Tensor, P: int, scale: float, sampling_ratio: int = 0) -> torch.Tensor:
    """feat (N,C,H,W) fp32, rois (M,5).  Small inputs only (pure-Python loops)."""
    f = np.float32
    x = feat.detach().cpu().numpy().astype(np.float32)
    r = rois.detach().cpu().numpy().astype(np.float32)
    N, C, H, W = x.shape
    out = np.zeros((r.shape[0], C, P, P), dtype=np.float32)
    for m in range(r.shape[0]):
        n = int(r[m, 0])
        sw, sh, bin_w, bin_h, gw, gh, count = _roi_geometry(r[m, 1:], scale, P, sampling_ratio)
        for ph in range(P):
            for pw in range(P):
                acc = np.zeros(C, dtype=np.float32)
                for iy in range(gh):
                    yy = f(sh + f(f(ph) * bin_h) + f(f(f(iy) + f(0.5)) * bin_h / f(gh)))
                    for ix in range(gw):
                        xx = f(sw + f(f(pw) * bin_w) + f(f(f(ix) + f(0.5)) * bin_w / f(gw)))
                        for (py, px, w) in _bilinear_terms(H, W, yy, xx):
                            acc += w * x[n, :, py, px]
                out[m, :, ph, pw] = acc / count
    return torch.from_numpy(out)


def separable_weights(H: int, W: int, roi: np.ndarray, scale: float, P: int, sampling_ratio: int = 0):
    """The regrouping used by the CUDA kernels:

    out[ph,pw] = (1/count) * sum_y sum_x Wy[ph,y] * Wx[pw,x] * F[y,x]

    where Wy[ph,y] = sum over valid samples iy of (hy at y_low, ly at y_high).  Exact in
    real arithmetic because bilinear weights are products and the sampling grid is a
    Cartesian product (the out-of-range rule is also a product of a y- and an x-test).
    Returns dense (P,H), (P,W) fp32 matrices and count.
    """
    f = np.float32
    sw, sh, bin_w, bin_h, gw, gh, count = _roi_geometry(roi, scale, P, sampling_ratio)
    Wy = np.zeros((P, H), dtype=np.float32)
    Wx = np.zeros((P, W), dtype=np.float32)
    for (Wm, start, binsz, g, L) in ((Wy, sh, bin_h, gh, H), (Wx, sw, bin_w, gw, W)):
        for p in range(P):
            for i in range(g):
                c = f(start + f(f(p) * binsz) + f(f(f(i) + f(0.5)) * binsz / f(g)))
                if c < -1.0 or c > L:
                    continue
                c = f(max(c, f(0)))
                lo = int(c)
                if lo >= L - 1:
                    hi = lo = L - 1; c = f(lo)
                else:
                    hi = lo + 1
                l = f(c - f(lo)); h = f(f(1) - l)
                Wm[p, lo] += h
                Wm[p, hi] += l
    return Wy, Wx, count


def roi_align_separable(feat: torch.Tensor, rois: torch.Tensor, P: int, scale: float, sampling_ratio: int = 0) -> torch.Tensor:
    x = feat.detach().cpu().numpy().astype(np.float32)
    r = rois.detach().cpu().numpy().astype(np.float32)
    N, C, H, W = x.shape
    out = np.zeros((r.shape[0], C, P, P), dtype=np.float32)
    for m in range(r.shape[0]):
        Wy, Wx, count = separable_weights(H, W, r[m, 1:], scale, P, sampling_ratio)
        out[m] = np.einsum("py,cyx,qx->cpq", Wy, x[int(r[m, 0])], Wx) / count
    return torch.from_numpy(out)


def touched_pixels(level_shapes: Sequence[Tuple[int, int]], scales: Sequence[float], rois: torch.Tensor,
                   levels: torch.Tensor, num_images: int, P: int = 7, sampling_ratio: int = 0) -> int:
    """U of SURVEY.md section 8(d): number of distinct (image, level, y, x) feature pixels touched by any
    bilinear sample of any RoI.  Uses the separable footprint (rows x cols with non-zero weight)."""
    r = rois.detach().cpu().numpy().astype(np.float32)
    lv = levels.detach().cpu().numpy()
    masks = [np.zeros((num_images, h, w), dtype=bool) for (h, w) in level_shapes]
    for m in range(r.shape[0]):
        l = int(lv[m]); H, W = level_shapes[l]
        Wy, Wx, _ = separable_weights(H, W, r[m, 1:], scales[l], P, sampling_ratio)
        ys = np.nonzero((Wy != 0).any(axis=0))[0]
        xs = np.nonzero((Wx != 0).any(axis=0))[0]
        if len(ys) and len(xs):
            masks[l][int(r[m, 0]), ys.min():ys.max() + 1, xs.min():xs.max() + 1] = True
    return int(sum(mk.sum() for mk in masks))
