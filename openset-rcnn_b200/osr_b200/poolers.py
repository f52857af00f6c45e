"""Drop-in for detectron2's ``ROIPooler`` as the reference builds and calls it:

* constructed at ``openset_rcnn/modeling/roi_heads/osrcnn_roi_heads.py:108-113``
  (``output_size=7, scales=(1/4..1/32), sampling_ratio=0, pooler_type="ROIAlignV2"``)
* called at ``osrcnn_roi_heads.py:306``: ``box_pooler(features, [x.proposal_boxes for x in proposals])``
* differentiated w.r.t. the feature maps by ``losses.backward()`` (``train.py:145``)

Level assignment + all levels' ROIAlign run in ONE kernel launch (``osr_roi_align_fwd``), the backward in
one deterministic, atomic-free gather launch (``osr_roi_align_bwd``).  No host synchronisation.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import FeatLevel
from .structures import Boxes, cat_rows


def _feat_levels(x: Sequence[torch.Tensor], scales: Sequence[float]):
    L = len(x)
    arr = (FeatLevel * L)()
    N, C = x[0].shape[0], x[0].shape[1]
    for l, (f, s) in enumerate(zip(x, scales)):
        if f.dtype != torch.float32:
            raise _lib.OsrError("ROIPooler: fp32 feature maps required (the reference has no AMP)")
        assert f.dim() == 4 and f.shape[0] == N and f.shape[1] == C
        a = arr[l]
        a.data = f.data_ptr()
        a.sN, a.sC, a.sH, a.sW = f.stride()
        a.H, a.W = f.shape[2], f.shape[3]
        a.scale = float(s)
    return arr, N, C


def convert_boxes_to_pooler_format(box_lists: List[Boxes]) -> Tuple[torch.Tensor, torch.Tensor]:
    """(M,5) [image index, x1, y1, x2, y2] + (N+1) int32 image offsets (host-known lengths, no sync)."""
    tensors = [b.tensor if hasattr(b, "tensor") else b for b in box_lists]
    sizes = [int(t.shape[0]) for t in tensors]
    dev = tensors[0].device
    boxes = cat_rows(tensors)
    offs = [0]
    for s in sizes:
        offs.append(offs[-1] + s)
    offsets = torch.tensor(offs, dtype=torch.int32).to(dev, non_blocking=True)
    idx = torch.repeat_interleave(
        torch.arange(len(sizes), dtype=boxes.dtype, device=dev),
        torch.tensor(sizes, device=dev), output_size=offs[-1])
    return torch.cat([idx[:, None], boxes], dim=1), offsets


# NCHW callers (the reference's layout): stage the maps through a tiled transpose and run the channels_last kernels
# (forward: NCHW -> NHWC copy of each level; backward: the NHWC gradient is transposed back).  2.3 vs 3.0 ms per cfg-2
# step.  Set to False to run the NCHW-native kernels instead (tests cover both).
NCHW_STAGING = True


def _stage_ok(f: torch.Tensor) -> bool:
    return f.is_contiguous() and f.shape[1] % 32 == 0 and f.shape[1] <= 256 and not (f.shape[2] == 1 and f.shape[3] == 1)


def nchw_to_channels_last(f: torch.Tensor) -> torch.Tensor:
    """Same logical (N, C, H, W) tensor in channels_last memory, copied by ``osr_nchw_to_nhwc``."""
    N, C, H, W = f.shape
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=f.device, memory_format=torch.channels_last)
    _lib.check(_lib.lib().osr_nchw_to_nhwc(f.data_ptr(), out.data_ptr(), N, C, H * W, _lib.stream_ptr(f.device)), "osr_nchw_to_nhwc")
    return out


def channels_last_to_nchw(g: torch.Tensor) -> torch.Tensor:
    N, C, H, W = g.shape
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=g.device)
    _lib.check(_lib.lib().osr_nhwc_to_nchw(g.data_ptr(), out.data_ptr(), N, C, H * W, _lib.stream_ptr(g.device)), "osr_nhwc_to_nchw")
    return out


class _ROIAlignFPN(torch.autograd.Function):
    """(x_0..x_{L-1}) -> (M, C, P, P); backward writes one dense gradient per level."""

    @staticmethod
    def forward(ctx, rois, offsets, cfg, *feats):
        lib = _lib.lib()
        scales, P, sampling_ratio, canon_size, canon_level, min_level = cfg
        ctx.staged = bool(NCHW_STAGING and all(f.dtype == torch.float32 and _stage_ok(f) for f in feats))
        src = [nchw_to_channels_last(f) for f in feats] if ctx.staged else feats
        arr, N, C = _feat_levels(src, scales)
        M = rois.shape[0]
        dev = feats[0].device
        out = torch.empty((M, C, P, P), dtype=torch.float32, device=dev)
        lvl = torch.empty((M,), dtype=torch.int32, device=dev)
        ws = torch.empty(max(int(lib.osr_roi_align_fwd_workspace(M)), 256), dtype=torch.uint8, device=dev)
        rc = lib.osr_roi_align_fwd(arr, len(feats), N, C, rois.data_ptr(), M, P, sampling_ratio, 1,
                                   canon_size, canon_level, min_level, out.data_ptr(), lvl.data_ptr(),
                                   ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "osr_roi_align_fwd")
        ctx.cfg = cfg
        ctx.shapes = [tuple(f.shape) for f in feats]
        ctx.channels_last = [f.is_contiguous(memory_format=torch.channels_last) and not f.is_contiguous() for f in feats]
        ctx.save_for_backward(rois, offsets)
        ctx.mark_non_differentiable(lvl)
        return out, lvl

    @staticmethod
    def backward(ctx, grad_out, _grad_lvl):
        rois, offsets = ctx.saved_tensors
        if ctx.staged:   # gradients through the channels_last kernel, handed back in the caller's NCHW layout
            grads = roi_align_backward(grad_out, rois, offsets, ctx.shapes, [True] * len(ctx.shapes), ctx.cfg)
            grads = [channels_last_to_nchw(g) for g in grads]
        else:
            grads = roi_align_backward(grad_out, rois, offsets, ctx.shapes, ctx.channels_last, ctx.cfg)
        return (None, None, None) + tuple(grads)


def roi_align_backward_workspace(feats, M: int, cfg) -> torch.Tensor:
    lib = _lib.lib()
    arr, N, C = _feat_levels(list(feats), cfg[0])
    return torch.empty((max(int(lib.osr_roi_align_bwd_workspace(arr, len(feats), N, C, M)), 256),), dtype=torch.uint8,
                       device=feats[0].device)


def roi_align_backward_prepare(feats, rois: torch.Tensor, offsets: torch.Tensor, cfg, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``osr_roi_align_bwd_prepare``: the backward's per-RoI tables for maps shaped / laid out like ``feats`` (their values are
    not read).  Returns the workspace to hand to ``roi_align_backward(..., prepared=ws)``.  Depends on the RoIs only, so a
    training step can issue it on a side stream right after sampling."""
    lib = _lib.lib()
    scales, P, sampling_ratio, canon_size, canon_level, min_level = cfg
    arr, N, C = _feat_levels(list(feats), scales)
    M = rois.shape[0]
    dev = rois.device
    ws = out if out is not None else roi_align_backward_workspace(feats, M, cfg)
    rc = lib.osr_roi_align_bwd_prepare(arr, len(feats), N, C, rois.data_ptr(), offsets.data_ptr(), M, P, sampling_ratio, 1,
                                       canon_size, canon_level, min_level, ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
    _lib.check(rc, "osr_roi_align_bwd_prepare")
    return ws


def roi_align_backward(grad_out: torch.Tensor, rois: torch.Tensor, offsets: torch.Tensor, shapes, channels_last, cfg,
                       prepared: Optional[torch.Tensor] = None):
    """``osr_roi_align_bwd`` without autograd: dense gradient maps (one per level, fully written - no memset) for pooled
    gradients ``grad_out`` (M, C, P, P).  ``shapes`` / ``channels_last``: shape and memory format of each level's map.
    ``prepared``: a workspace from ``roi_align_backward_prepare`` for the same RoIs and map geometry."""
    lib = _lib.lib()
    scales, P, sampling_ratio, canon_size, canon_level, min_level = cfg
    dev = grad_out.device
    grad_out = grad_out.contiguous()
    grads = []
    for shp, cl in zip(shapes, channels_last):
        grads.append(torch.empty(tuple(shp), dtype=torch.float32, device=dev,
                                 memory_format=torch.channels_last if cl else torch.contiguous_format))
    arr, N, C = _feat_levels(grads, scales)
    M = rois.shape[0]
    if prepared is not None:
        rc = lib.osr_roi_align_bwd_prepared(arr, len(grads), N, C, grad_out.data_ptr(), rois.data_ptr(), offsets.data_ptr(),
                                            M, P, sampling_ratio, 1, canon_size, canon_level, min_level,
                                            prepared.data_ptr(), prepared.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "osr_roi_align_bwd_prepared")
        return grads
    ws_bytes = int(lib.osr_roi_align_bwd_workspace(arr, len(grads), N, C, M))
    ws = torch.empty((max(ws_bytes, 256),), dtype=torch.uint8, device=dev)
    rc = lib.osr_roi_align_bwd(arr, len(grads), N, C, grad_out.data_ptr(), rois.data_ptr(), offsets.data_ptr(),
                               M, P, sampling_ratio, 1, canon_size, canon_level, min_level,
                               ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
    _lib.check(rc, "osr_roi_align_bwd")
    return grads


class ROIPooler(torch.nn.Module):
    """Same constructor / forward contract as detectron2.modeling.poolers.ROIPooler (ROIAlignV2 only)."""

    def __init__(self, output_size, scales, sampling_ratio, pooler_type="ROIAlignV2",
                 canonical_box_size: int = 224, canonical_level: int = 4):
        super().__init__()
        if isinstance(output_size, int):
            output_size = (output_size, output_size)
        assert len(output_size) == 2 and output_size[0] == output_size[1]
        if pooler_type != "ROIAlignV2":
            raise ValueError("osr_b200.ROIPooler implements pooler_type='ROIAlignV2' (what the reference configures)")
        self.output_size = tuple(output_size)
        self.scales = tuple(float(s) for s in scales)
        self.sampling_ratio = int(sampling_ratio)
        min_level = -(math.log2(scales[0]))
        max_level = -(math.log2(scales[-1]))
        assert math.isclose(min_level, int(min_level)) and math.isclose(max_level, int(max_level)), \
            "Featuremap stride is not power of 2!"
        self.min_level = int(min_level)
        self.max_level = int(max_level)
        assert len(scales) == self.max_level - self.min_level + 1, \
            "[ROIPooler] Sizes of input featuremaps do not form a pyramid!"
        assert 0 <= self.min_level <= self.max_level
        self.canonical_level = canonical_level
        assert canonical_box_size > 0
        self.canonical_box_size = canonical_box_size

    def _cfg(self):
        return (self.scales, self.output_size[0], self.sampling_ratio, self.canonical_box_size,
                self.canonical_level, self.min_level)

    def forward_with_levels(self, x: List[torch.Tensor], box_lists: List[Boxes]):
        """Returns (pooled (M,C,P,P), level (M,) int32)."""
        num_level_assignments = len(self.scales)
        assert isinstance(x, list) and isinstance(box_lists, list), "Arguments to pooler must be lists"
        assert len(x) == num_level_assignments, \
            "unequal value, num_level_assignments={}, but x is list of {} Tensors".format(num_level_assignments, len(x))
        assert len(box_lists) == x[0].size(0), \
            "unequal value, x[0] batch dim 0 is {}, but box_list has length {}".format(x[0].size(0), len(box_lists))
        _lib.require_cuda(*x)
        if len(box_lists) == 0:
            z = torch.zeros((0, x[0].shape[1]) + self.output_size, device=x[0].device, dtype=x[0].dtype)
            return z, torch.zeros((0,), dtype=torch.int32, device=x[0].device)
        rois, offsets = convert_boxes_to_pooler_format(box_lists)
        return self.pool_rois(x, rois, offsets)

    def pool_rois_bf16(self, x: List[torch.Tensor], rois: torch.Tensor, offsets: torch.Tensor):
        """Forward only, pooled tile written as bf16 (``osr_roi_align_fwd_bf16``): the A operand of the tensor-core box
        head (``box_head.FastRCNNConvFCHead``).  Dense channels_last maps, C % 8 == 0.  Returns (pooled bf16, level).
        The gradient w.r.t. the maps is ``backward_rois`` (ROIAlign's backward never reads the pooled values)."""
        lib = _lib.lib()
        scales, P, sampling_ratio, canon_size, canon_level, min_level = self._cfg()
        feats = [f.detach() for f in x]
        arr, N, C = _feat_levels(feats, scales)
        rois = rois.contiguous().float()
        M = rois.shape[0]
        dev = feats[0].device
        out = torch.empty((M, C, P, P), dtype=torch.bfloat16, device=dev)
        lvl = torch.empty((M,), dtype=torch.int32, device=dev)
        if M == 0:
            return out, lvl
        ws = torch.empty(max(int(lib.osr_roi_align_fwd_workspace(M)), 256), dtype=torch.uint8, device=dev)
        rc = lib.osr_roi_align_fwd_bf16(arr, len(feats), N, C, rois.data_ptr(), M, P, sampling_ratio, 1, canon_size, canon_level,
                                        min_level, out.data_ptr(), lvl.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
        _lib.check(rc, "osr_roi_align_fwd_bf16")
        return out, lvl

    def pool_rois(self, x: List[torch.Tensor], rois: torch.Tensor, offsets: torch.Tensor):
        """Already-packed entry: rois (M,5) image-major, offsets (N+1) int32 (no python list handling)."""
        if rois.shape[0] == 0:
            z = torch.zeros((0, x[0].shape[1]) + self.output_size, device=x[0].device, dtype=x[0].dtype)
            return z, torch.zeros((0,), dtype=torch.int32, device=x[0].device)
        rois = rois.contiguous().float()
        return _ROIAlignFPN.apply(rois, offsets, self._cfg(), *x)

    def forward(self, x: List[torch.Tensor], box_lists: List[Boxes]) -> torch.Tensor:
        return self.forward_with_levels(x, box_lists)[0]

    def prepare_backward(self, x: List[torch.Tensor], rois: torch.Tensor, offsets: torch.Tensor,
                         out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The RoI-only part of ``backward_rois`` (``osr_roi_align_bwd_prepare``), to be issued early / on a side stream; pass
        the result as ``backward_rois(..., prepared=...)`` with the same ``x`` layouts, ``rois`` and ``offsets``."""
        # (only the maps' geometry - N, C, H, W, scale - enters the tables: any memory format of x will do)
        return roi_align_backward_prepare([f.detach() for f in x], rois.contiguous().float(), offsets, self._cfg(), out)

    def alloc_backward_workspace(self, x: List[torch.Tensor], rois: torch.Tensor) -> torch.Tensor:
        """Workspace for ``prepare_backward(..., out=)`` (allocate it on the stream that will run ``backward_rois``)."""
        return roi_align_backward_workspace([f.detach() for f in x], rois.shape[0], self._cfg())

    def backward_rois(self, grad_pooled: torch.Tensor, x: List[torch.Tensor], rois: torch.Tensor, offsets: torch.Tensor,
                      prepared: Optional[torch.Tensor] = None):
        """Gradient of ``pool_rois`` w.r.t. the maps ``x`` (only their shapes / layouts are read), without autograd."""
        cl = [f.is_contiguous(memory_format=torch.channels_last) and not f.is_contiguous() for f in x]
        if NCHW_STAGING and all(_stage_ok(f) for f in x):
            grads = roi_align_backward(grad_pooled, rois.contiguous().float(), offsets, [tuple(f.shape) for f in x],
                                       [True] * len(x), self._cfg(), prepared)
            return [channels_last_to_nchw(g) for g in grads]
        return roi_align_backward(grad_pooled, rois.contiguous().float(), offsets, [tuple(f.shape) for f in x], cl, self._cfg(),
                                  prepared)
