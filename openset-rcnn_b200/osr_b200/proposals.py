"""CF-RPN proposal stage: drop-ins for

* ``ClsFreeRPN.predict_proposals`` / ``_decode_proposals``
  (``openset_rcnn/modeling/proposal_generator/classification_free_rpn.py:558-610``)
* ``find_top_rpn_proposals`` (``openset_rcnn/modeling/find_top_proposals.py:22-128``)

Same names, argument meaning and error behaviour; the work is one call into
``osr_rpn_select_decode`` (+ ``osr_nms_segmented`` in ``nominal`` mode).  Exactly one D2H copy per batch
(per-image counts + non-finite flags) replaces the reference's >= 2 host syncs per image.

Modes: ``as_shipped`` = the reference as it is (NMS and post_nms_topk commented out,
``find_top_proposals.py:112-120``); ``nominal`` = stock detectron2 (per-level NMS, then the best
``post_nms_topk`` per image).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple, Union

import torch

from . import _lib
from ._lib import RpnLevel
from .structures import Boxes, Instances, boxes_view, make_instances


@dataclass
class RpnSelection:
    """Padded device-side result of ``osr_rpn_select_decode`` (no host sync has happened yet)."""
    boxes: torch.Tensor    # (N, Kmax, 4) fp32
    scores: torch.Tensor   # (N, Kmax)    fp32
    level: torch.Tensor    # (N, Kmax)    int32
    index: torch.Tensor    # (N, Kmax)    int32 flat anchor index inside the level
    counts: torch.Tensor   # (N, L+2)     int32 [per level..., total, flags]
    num_levels: int
    kmax: int
    max_level_k: int = 0  # largest per-level k (host-side bound for the NMS workspace)


def _tensor_of(x) -> torch.Tensor:
    return x.tensor if hasattr(x, "tensor") else x


def _image_hw_tensor(image_sizes: Sequence[Tuple[int, int]], device) -> torch.Tensor:
    return torch.tensor([[int(h), int(w)] for (h, w) in image_sizes], dtype=torch.int32).to(device, non_blocking=True)


def rpn_select_decode(
    anchors: Optional[Sequence[Union[torch.Tensor, Boxes]]],
    deltas: Sequence[torch.Tensor],
    scores: Sequence[torch.Tensor],
    image_sizes: Union[Sequence[Tuple[int, int]], torch.Tensor],
    pre_nms_topk: int,
    min_box_size: float = 0.0,
) -> RpnSelection:
    """Device-side proposal stage.  ``deltas[l]``: (N, HWA, 4) in any strides (a permuted view of the raw
    (N, A*4, H, W) conv output works when A == 1); ``scores[l]``: (N, HWA); ``anchors[l]``: (HWA, 4) or
    ``None`` for all levels when ``deltas`` already holds decoded boxes."""
    lib = _lib.lib()
    L = len(deltas)
    assert L == len(scores) and 1 <= L <= _lib.OSR_MAX_LEVELS
    dev = deltas[0].device
    _lib.require_cuda(*deltas, *scores)
    N = deltas[0].shape[0]
    levels = (RpnLevel * L)()
    keep_alive = []
    for l in range(L):
        d, s = deltas[l], scores[l]
        if d.dtype != torch.float32 or s.dtype != torch.float32:
            raise _lib.OsrError("rpn_select_decode: fp32 inputs required (the reference has no AMP)")
        assert d.dim() == 3 and d.shape[2] == 4 and s.dim() == 2 and s.shape[1] == d.shape[1]
        lv = levels[l]
        lv.deltas = d.data_ptr()
        lv.scores = s.data_ptr()
        if anchors is not None:
            a = _tensor_of(anchors[l])
            if not (a.is_contiguous() and a.dtype == torch.float32):
                a = a.contiguous().float()
            assert a.shape == (d.shape[1], 4)
            keep_alive.append(a)
            lv.anchors = a.data_ptr()
        else:
            lv.anchors = None
        lv.num_anchors = d.shape[1]
        lv.delta_stride_n, lv.delta_stride_a, lv.delta_stride_c = d.stride()
        lv.score_stride_n, lv.score_stride_a = s.stride()
    kmax = int(lib.osr_rpn_kmax(levels, L, int(pre_nms_topk)))
    if kmax < 0:
        _lib.check(-1, "osr_rpn_kmax")
    if isinstance(image_sizes, torch.Tensor):
        hw = image_sizes.to(device=dev, dtype=torch.int32)
    else:
        hw = _image_hw_tensor(image_sizes, dev)
    assert hw.shape == (N, 2)
    boxes = torch.empty((N, kmax, 4), dtype=torch.float32, device=dev)
    sc = torch.empty((N, kmax), dtype=torch.float32, device=dev)
    lvl = torch.empty((N, kmax), dtype=torch.int32, device=dev)
    idx = torch.empty((N, kmax), dtype=torch.int32, device=dev)
    counts = torch.empty((N, L + 2), dtype=torch.int32, device=dev)
    ws_bytes = int(lib.osr_rpn_select_decode_workspace(levels, L, N, int(pre_nms_topk)))
    ws = torch.empty((max(ws_bytes, 256),), dtype=torch.uint8, device=dev)
    rc = lib.osr_rpn_select_decode(levels, L, N, int(pre_nms_topk), float(min_box_size), hw.data_ptr(),
                                   boxes.data_ptr(), sc.data_ptr(), lvl.data_ptr(), idx.data_ptr(),
                                   counts.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
    _lib.check(rc, "osr_rpn_select_decode")
    max_level_k = max(min(int(d.shape[1]), int(pre_nms_topk)) for d in deltas)
    return RpnSelection(boxes, sc, lvl, idx, counts, L, kmax, max_level_k)


def _to_instances(sel: RpnSelection, image_sizes, training: bool, keep=None, keep_counts=None) -> List[Instances]:
    """The single host sync of the stage: counts + flags -> python lists -> ``Instances`` views."""
    L = sel.num_levels
    counts = sel.counts.cpu()  # one D2H copy
    if training and bool((counts[:, L + 1] != 0).any()):
        # find_top_proposals.py:96-101
        raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")
    totals = counts[:, L].tolist()
    results = []
    for n, image_size in enumerate(image_sizes):
        c = totals[n]
        results.append(make_instances(tuple(image_size), proposal_boxes=boxes_view(sel.boxes[n, :c]),
                                      objectness_logits=sel.scores[n, :c]))
    return results


def find_top_rpn_proposals(
    proposals: List[torch.Tensor],
    pred_objectness_logits: List[torch.Tensor],
    image_sizes: List[Tuple[int, int]],
    nms_thresh: float,
    pre_nms_topk: int,
    post_nms_topk: int,
    min_box_size: float,
    training: bool,
    mode: str = "as_shipped",
) -> List[Instances]:
    """Same signature as the reference (``find_top_proposals.py:22-31``): ``proposals`` are already decoded
    (N, Hi*Wi*A, 4) boxes.  Only the selected boxes are gathered."""
    sel = rpn_select_decode(None, proposals, pred_objectness_logits, image_sizes, pre_nms_topk, min_box_size)
    return _finish(sel, image_sizes, nms_thresh, post_nms_topk, training, mode)


def predict_proposals(
    anchors: Sequence[Union[torch.Tensor, Boxes]],
    pred_anchor_deltas: List[torch.Tensor],
    pred_centerness: List[torch.Tensor],
    image_sizes: List[Tuple[int, int]],
    *,
    nms_thresh: float = 1.0,
    pre_nms_topk: int = 2000,
    post_nms_topk: int = 2000,
    min_box_size: float = 0.0,
    training: bool = True,
    mode: str = "as_shipped",
) -> List[Instances]:
    """``ClsFreeRPN.predict_proposals`` (``classification_free_rpn.py:558-589``) with decode fused in."""
    with torch.no_grad():
        sel = rpn_select_decode(anchors, pred_anchor_deltas, pred_centerness, image_sizes, pre_nms_topk, min_box_size)
        return _finish(sel, image_sizes, nms_thresh, post_nms_topk, training, mode)


def _finish(sel: RpnSelection, image_sizes, nms_thresh, post_nms_topk, training, mode) -> List[Instances]:
    if mode == "as_shipped":
        return _to_instances(sel, image_sizes, training)
    if mode != "nominal":
        raise ValueError(mode)
    from .nms import rpn_nominal_nms  # local import: nms.py imports this module's RpnSelection

    return rpn_nominal_nms(sel, image_sizes, float(nms_thresh), int(post_nms_topk), training)


class ClsFreeRPNProposals:
    """The proposal half of ``ClsFreeRPN`` (``classification_free_rpn.py:165-316`` config, ``:558-589`` call):
    holds the (train, test) tuples the reference reads from cfg.MODEL.RPN.* and exposes ``predict_proposals``
    with the reference's positional signature."""

    def __init__(self, pre_nms_topk=(2000, 1000), post_nms_topk=(2000, 1000), nms_thresh=(1.0, 1.0),
                 min_box_size: float = 0.0, mode: str = "as_shipped"):
        self.pre_nms_topk = {True: pre_nms_topk[0], False: pre_nms_topk[1]}
        self.post_nms_topk = {True: post_nms_topk[0], False: post_nms_topk[1]}
        self.nms_thresh = {True: nms_thresh[0], False: nms_thresh[1]}
        self.min_box_size = float(min_box_size)
        self.mode = mode
        self.training = True

    def train(self, mode: bool = True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def predict_proposals(self, anchors, pred_anchor_deltas, pred_centerness, image_sizes):
        return predict_proposals(
            anchors, pred_anchor_deltas, pred_centerness, image_sizes,
            nms_thresh=self.nms_thresh[self.training], pre_nms_topk=self.pre_nms_topk[self.training],
            post_nms_topk=self.post_nms_topk[self.training], min_box_size=self.min_box_size,
            training=self.training, mode=self.mode)
