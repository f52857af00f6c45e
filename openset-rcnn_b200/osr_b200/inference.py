"""ROI-head inference post-processing.

Drop-in for ``OpensetFastRCNNOutputLayers.inference(predictions, proposals, box_features)``
(``osrcnn_fast_rcnn.py:380-404``) and the ``fast_rcnn_inference`` it calls (``:45-145``): box decode
(detectron2 ``Box2BoxTransform.apply_deltas``), objectness ``sqrt(iou * centerness)``, finite filter, clip, score
threshold, class-agnostic ``batched_nms`` and top-k - for ALL images with four kernel launches
(``osr_rcnn_decode_score`` + the three NMS kernels of ``osr_nms_segmented``) and one host sync, instead of a Python
loop over images with ~15 launches and 2 syncs each.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch

from . import _lib
from .nms import _MAX_SORT, _segmented
from .structures import Boxes, Instances

SCALE_CLAMP = math.log(1000.0 / 16)


def _decode(proposal_deltas, ious, proposals, weights, mean_type, score_thresh):
    lib = _lib.lib()
    pb = torch.cat([p.get("proposal_boxes").tensor for p in proposals], dim=0).contiguous().float()
    ctr = torch.cat([p.get("objectness_logits") for p in proposals]).contiguous().float()
    _lib.require_cuda(pb, proposal_deltas, ious)
    dev = pb.device
    deltas = proposal_deltas.contiguous().float()
    if deltas.shape[1] != 4:
        raise NotImplementedError("class-specific box regression (deltas of shape (R, K*4)) is not used by Openset R-CNN")
    iou = ious.reshape(-1).contiguous().float()
    lens = [len(p) for p in proposals]
    R = pb.shape[0]
    assert deltas.shape[0] == R and iou.shape[0] == R
    if mean_type not in ("geometric", "arithmetic"):
        raise ValueError(f"mean_type {mean_type!r}")   # the reference leaves `scores` undefined here (UnboundLocalError)
    off = torch.tensor([0] + torch.tensor(lens).cumsum(0).tolist(), dtype=torch.int32).to(dev, non_blocking=True)
    hw = torch.tensor([[int(p.image_size[0]), int(p.image_size[1])] for p in proposals], dtype=torch.int32).to(dev, non_blocking=True)
    out_boxes = torch.empty((max(R, 1), 4), dtype=torch.float32, device=dev)
    out_scores = torch.empty(max(R, 1), dtype=torch.float32, device=dev)
    out_eff = torch.empty(max(R, 1), dtype=torch.float32, device=dev)
    if R > 0:
        wx, wy, ww, wh = (float(w) for w in weights)
        rc = lib.osr_rcnn_decode_score(pb.data_ptr(), deltas.data_ptr(), iou.data_ptr(), ctr.data_ptr(), off.data_ptr(),
                                       hw.data_ptr(), len(lens), max(lens), wx, wy, ww, wh, float(SCALE_CLAMP),
                                       1 if mean_type == "geometric" else 0, float(score_thresh),
                                       out_boxes.data_ptr(), out_scores.data_ptr(), out_eff.data_ptr(),
                                       _lib.stream_ptr(dev))
        _lib.check(rc, "osr_rcnn_decode_score")
    return out_boxes[:R], out_scores[:R], out_eff[:R], lens, off


def inference(predictions: Tuple[torch.Tensor, torch.Tensor], proposals: List[Instances], box_features: torch.Tensor, *,
              weights: Sequence[float] = (10.0, 10.0, 5.0, 5.0), mean_type: str = "geometric",
              score_thresh: float = 0.0, nms_thresh: float = 0.5, topk_per_image: int = 100):
    """``OpensetFastRCNNOutputLayers.inference`` -> ``(List[Instances], List[kept indices])``; the keyword arguments are
    the module attributes the reference reads (``box2box_transform.weights``, ``mean_type``,
    ``test_objectness_score_thresh``, ``test_nms_thresh``, ``test_topk_per_image``).  Each ``Instances`` has
    ``pred_boxes``, ``scores``, ``pred_classes`` (all 0: class-agnostic) and ``features``."""
    if not len(proposals):
        return [], []
    proposal_deltas, ious = predictions
    boxes, scores, eff, lens, off = _decode(proposal_deltas, ious, proposals, weights, mean_type, score_thresh)
    dev = boxes.device
    N = len(lens)
    feats = box_features
    if max(lens) == 0:
        empty = torch.empty(0, dtype=torch.int64, device=dev)
        out = []
        for p in proposals:
            r = Instances(p.image_size)
            r.set("pred_boxes", Boxes(boxes[:0])); r.set("scores", scores[:0]); r.set("pred_classes", empty)
            r.set("features", feats[:0])
            out.append(r)
        return out, [empty for _ in proposals]
    if max(lens) > _MAX_SORT:
        raise NotImplementedError(f"more than {_MAX_SORT} proposals per image")
    seg_begin = off[:-1].contiguous()
    seg_len = (off[1:] - off[:-1]).contiguous()
    keep_idx, keep_cnt, _ = _segmented(boxes, eff, seg_begin, seg_len, max(lens), nms_thresh, False)
    # survivors are a prefix of each image's kept list (dropped rows carry score -inf and sort last)
    T = boxes.shape[0]
    pos = torch.arange(T, device=dev)
    seg = torch.searchsorted(off[1:].long(), pos, right=True).clamp_(max=N - 1)
    begin = off[:-1].long()[seg]
    within = (pos - begin) < keep_cnt.long()[seg]
    gidx = (keep_idx[:T] + begin).clamp_(0, T - 1)
    ok = within & (eff[gidx] > float("-inf"))
    n_ok = torch.zeros(N, dtype=torch.int64, device=dev).index_add_(0, seg, ok.long())
    # kept indices are positions in the finite-filtered list, as in the reference (:106-110)
    finite = torch.isfinite(scores)
    fin_pos = torch.cumsum(finite.long(), 0) - 1
    fin_base = torch.cat((torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(finite.long(), 0)))[off[:-1].long()]
    counts = n_ok.cpu().tolist()   # the one host sync
    results, kept = [], []
    for n, p in enumerate(proposals):
        b0 = int(sum(lens[:n]))
        k = counts[n] if topk_per_image < 0 else min(counts[n], topk_per_image)
        g = keep_idx[b0:b0 + k] + b0
        r = Instances(p.image_size)
        r.set("pred_boxes", Boxes(boxes[g]))
        r.set("scores", scores[g])
        r.set("pred_classes", torch.zeros(k, dtype=torch.int64, device=dev))
        r.set("features", feats[g])
        results.append(r)
        kept.append(fin_pos[g] - fin_base[n])
    return results, kept
