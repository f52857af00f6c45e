"""NMS drop-ins.

* ``batched_nms(boxes, scores, idxs, iou_threshold)`` - ``detectron2.layers.batched_nms`` as imported at
  ``find_top_proposals.py:7``, ``osrcnn_fast_rcnn.py:10``, ``softmax_classifier.py:11`` and called at
  ``osrcnn_fast_rcnn.py:135``, ``softmax_classifier.py:93/:154``.  The *Python dispatch* of torchvision is mirrored
  (coordinate trick below 100 000 box coordinates on CUDA, per-class otherwise - torchvision 0.26
  ``ops/boxes.py:batched_nms``) so the IoU rounding is identical to the reference's; the kernel itself is plain
  ``nms`` over segments (``osr_nms_segmented``).
* ``nms(boxes, scores, iou_threshold)`` - ``torchvision.ops.nms``.
* ``batched_nms_images`` - every image of a batch in ONE call (segments = images), replacing the per-image Python
  loops of ``fast_rcnn_inference`` / ``fast_rcnn_inference_single_image_known/unknown``.
* ``rpn_nominal_nms`` - the block commented out at ``find_top_proposals.py:112-120`` (stock detectron2
  behaviour), applied to the padded output of ``osr_rpn_select_decode`` without a host sync before the end.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from . import _lib
from .structures import Instances, boxes_view, flat_prefixes, make_instances

_MAX_SORT = 16384


def _segmented(boxes, scores, seg_begin, seg_len, max_len, thr, presorted, want_mask=False):
    lib = _lib.lib()
    dev = boxes.device
    T = boxes.shape[0]
    S = seg_begin.numel()
    keep_idx = torch.empty(max(T, 1), dtype=torch.int64, device=dev)
    keep_cnt = torch.empty(max(S, 1), dtype=torch.int32, device=dev)
    keep_mask = torch.empty(max(T, 1), dtype=torch.uint8, device=dev) if want_mask else None
    if want_mask:
        keep_mask.zero_()  # boxes outside every segment stay 0
    ws = torch.empty(max(int(lib.osr_nms_workspace(T, S, max_len)), 256), dtype=torch.uint8, device=dev)
    rc = lib.osr_nms_segmented(boxes.data_ptr(), scores.data_ptr(), T, seg_begin.data_ptr(), seg_len.data_ptr(), S,
                               int(max_len), float(thr), int(presorted), keep_idx.data_ptr(), keep_cnt.data_ptr(),
                               _lib.ptr(keep_mask), ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
    _lib.check(rc, "osr_nms_segmented")
    return keep_idx, keep_cnt, keep_mask


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """``torchvision.ops.nms``: int64 indices of kept boxes, sorted by decreasing score (stable)."""
    _lib.require_cuda(boxes, scores)
    K = boxes.shape[0]
    if K == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    boxes = boxes.contiguous().float()
    scores = scores.contiguous().float()
    dev = boxes.device
    seg_begin = torch.zeros(1, dtype=torch.int32, device=dev)
    seg_len = torch.full((1,), K, dtype=torch.int32, device=dev)
    if K <= _MAX_SORT:
        keep_idx, keep_cnt, _ = _segmented(boxes, scores, seg_begin, seg_len, K, iou_threshold, False)
        n = int(keep_cnt[0])  # the caller needs a dense tensor: same single sync torchvision's nms has
        return keep_idx[:n]
    # very long inputs: order with torch.sort (library), suppress with the kernel on the presorted boxes
    order = torch.sort(scores, descending=True, stable=True)[1]
    keep_idx, keep_cnt, _ = _segmented(boxes[order].contiguous(), scores[order].contiguous(), seg_begin, seg_len, K,
                                       iou_threshold, True)
    n = int(keep_cnt[0])
    return order[keep_idx[:n]]


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """``detectron2.layers.batched_nms`` (= torchvision ``batched_nms(boxes.float(), ...)``)."""
    assert boxes.shape[-1] == 4
    boxes = boxes.float()
    if boxes.numel() == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    if boxes.numel() > 100_000:
        # torchvision _batched_nms_vanilla: nms per class, then order the survivors by score
        keep_mask = torch.zeros_like(scores, dtype=torch.bool)
        for class_id in torch.unique(idxs):
            curr = torch.where(idxs == class_id)[0]
            keep_mask[curr[nms(boxes[curr], scores[curr], iou_threshold)]] = True
        keep = torch.where(keep_mask)[0]
        return keep[scores[keep].sort(descending=True, stable=True)[1]]
    # torchvision _batched_nms_coordinate_trick (same torch ops => same fp32 offsets and IoU rounding)
    max_coordinate = boxes.max()
    offsets = idxs.to(boxes) * (max_coordinate + torch.tensor(1).to(boxes))
    boxes_for_nms = boxes + offsets[:, None]
    return nms(boxes_for_nms, scores, iou_threshold)


def batched_nms_images(boxes_list: Sequence[torch.Tensor], scores_list: Sequence[torch.Tensor],
                       idxs_list: Sequence[torch.Tensor], iou_threshold: float,
                       topk_per_image: int = -1) -> List[torch.Tensor]:
    """``[batched_nms(b, s, i, thr)[:topk] for b, s, i in zip(...)]`` with one sort + mask + sweep launch for the
    whole batch (segments = images, coordinate trick applied per image exactly as torchvision does)."""
    n_img = len(boxes_list)
    if n_img == 0:
        return []
    dev = boxes_list[0].device
    lens = [int(b.shape[0]) for b in boxes_list]
    if max(lens) == 0:
        return [torch.empty((0,), dtype=torch.int64, device=dev) for _ in lens]
    if max(lens) > _MAX_SORT or max(lens) * 4 > 100_000:
        return [batched_nms(b, s, i, iou_threshold)[:topk_per_image if topk_per_image >= 0 else None]
                for b, s, i in zip(boxes_list, scores_list, idxs_list)]
    tricked = []
    for b, i in zip(boxes_list, idxs_list):
        b = b.float()
        if b.numel() == 0:
            tricked.append(b.reshape(0, 4))
            continue
        off = i.to(b) * (b.max() + torch.tensor(1).to(b))
        tricked.append(b + off[:, None])
    allb = torch.cat(tricked).contiguous()
    alls = torch.cat([s.float() for s in scores_list]).contiguous()
    begins, o = [], 0
    for l in lens:
        begins.append(o)
        o += l
    seg_begin = torch.tensor(begins, dtype=torch.int32).to(dev, non_blocking=True)
    seg_len = torch.tensor(lens, dtype=torch.int32).to(dev, non_blocking=True)
    keep_idx, keep_cnt, _ = _segmented(allb, alls, seg_begin, seg_len, max(lens), iou_threshold, False)
    cnt = keep_cnt.cpu().tolist()  # one sync for the whole batch
    out = []
    for n in range(n_img):
        k = cnt[n] if topk_per_image < 0 else min(cnt[n], topk_per_image)
        out.append(keep_idx[begins[n]:begins[n] + k])
    return out


def batched_nms_flat(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, img: torch.Tensor,
                     lens: Sequence[int], iou_threshold: float, topk_per_image: int = -1) -> Tuple[torch.Tensor, List[int]]:
    """``batched_nms_images`` on CONCATENATED inputs: ``boxes`` (T, 4) / ``scores`` / ``idxs`` hold the images one after
    the other (``lens[n]`` rows each, host-known), ``img`` (T,) int64 is the image of every row.  Returns the GLOBAL row
    indices of the kept boxes - image-major, score-descending inside an image, first ``topk_per_image`` of each - and the
    per-image counts.  No per-image launches: torchvision's coordinate trick needs every image's own ``boxes.max()``,
    which is one ``scatter_reduce(amax)`` here (max is exact, so the fp32 offsets - and with them the IoU rounding - are
    the ones ``[batched_nms(b, s, i, thr) for b, s, i in ...]`` computes); one host read (the keep counts)."""
    n_img = len(lens)
    dev = boxes.device
    lens = [int(l) for l in lens]
    begins, o = [], 0
    for l in lens:
        begins.append(o)
        o += l
    if n_img == 0 or o == 0:
        return torch.empty((0,), dtype=torch.int64, device=dev), [0] * n_img
    boxes = boxes.float()
    if max(lens) > _MAX_SORT or max(lens) * 4 > 100_000:   # torchvision switches to its per-class loop: per-image calls
        keeps = [batched_nms(boxes[b:b + l], scores[b:b + l], idxs[b:b + l], iou_threshold)[:topk_per_image if topk_per_image >= 0 else None] + b
                 for b, l in zip(begins, lens)]
        return torch.cat(keeps), [int(k.numel()) for k in keeps]
    mx = torch.full((n_img,), float("-inf"), dtype=torch.float32, device=dev)
    mx.scatter_reduce_(0, img, boxes.amax(dim=1), "amax", include_self=True)
    off = idxs.to(boxes) * (mx.index_select(0, img) + torch.tensor(1).to(boxes))
    allb = (boxes + off[:, None]).contiguous()
    seg = torch.tensor([begins, lens], dtype=torch.int32).to(dev, non_blocking=True)
    keep_idx, keep_cnt, _ = _segmented(allb, scores.float().contiguous(), seg[0], seg[1], max(lens), iou_threshold, False)
    cnt = keep_cnt.cpu().tolist()  # one sync for the whole batch
    ks = [c if topk_per_image < 0 else min(c, topk_per_image) for c in cnt[:n_img]]
    P, B = flat_prefixes(begins, ks, dev)
    return keep_idx.index_select(0, P) + B, ks


def rpn_nominal_nms(sel, image_sizes, nms_thresh: float, post_nms_topk: int, training: bool) -> List[Instances]:
    """Stock detectron2 tail of ``find_top_rpn_proposals`` (``find_top_proposals.py:112-120``, commented out in
    the reference): ``keep = batched_nms(boxes, scores, lvl, thr)[:post_nms_topk]`` per image.

    ``sel`` is the padded output of ``osr_rpn_select_decode``: per image the levels are contiguous runs that are
    already score-descending, so the kernel runs ``presorted`` on (image, level) segments - identical to
    torchvision's whole-image call because the coordinate trick makes different levels disjoint.  The per-image
    merge (score order across levels, first ``post_nms_topk``) is a masked stable sort.  One host sync at the end."""
    L, kmax = sel.num_levels, sel.kmax
    N = sel.boxes.shape[0]
    dev = sel.boxes.device
    counts = sel.counts[:, :L]                                            # (N, L) device
    total = sel.counts[:, L]
    pos = torch.arange(kmax, device=dev)[None, :]
    valid = pos < total[:, None]
    # torchvision coordinate trick, per image: boxes + level * (max_coordinate + 1), all fp32
    neg = torch.full((), float("-inf"), device=dev)
    max_coord = torch.where(valid[:, :, None], sel.boxes, neg).amax(dim=(1, 2))   # (N,)
    offsets = sel.level.to(torch.float32) * (max_coord + torch.tensor(1.0, device=dev))[:, None]
    boxes_for_nms = (sel.boxes + offsets[:, :, None]).view(-1, 4).contiguous()
    lvl_start = torch.cumsum(counts, dim=1) - counts                      # (N, L)
    seg_begin = (lvl_start + (torch.arange(N, device=dev) * kmax)[:, None]).to(torch.int32).reshape(-1).contiguous()
    seg_len = counts.to(torch.int32).reshape(-1).contiguous()
    # per-level upper bound known on the host: the largest per-level k
    _, _, keep_mask = _segmented(boxes_for_nms, sel.scores.reshape(-1).contiguous(), seg_begin, seg_len,
                                 _max_level_k(sel), nms_thresh, True, want_mask=True)
    keep_mask = keep_mask.view(N, kmax).bool() & valid
    # merge levels: survivors by score descending, ties by concatenated index (stable)
    masked = torch.where(keep_mask, sel.scores, neg)
    order = torch.sort(masked, dim=1, descending=True, stable=True)[1]
    n_keep = keep_mask.sum(dim=1).clamp(max=post_nms_topk)
    host = torch.stack((n_keep.to(torch.int32), sel.counts[:, L + 1])).cpu()   # the single host sync
    if training and bool((host[1] != 0).any()):
        raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")
    # all images' survivors in one gather per field; the per-image Instances are split views
    ks = host[0].tolist()
    P, B = flat_prefixes([n * kmax for n in range(N)], ks, dev)
    flat = order.reshape(-1).index_select(0, P) + B
    out_boxes = sel.boxes.reshape(-1, 4).index_select(0, flat).split(ks)
    out_scores = sel.scores.reshape(-1).index_select(0, flat).split(ks)
    results = []
    for n, image_size in enumerate(image_sizes):
        results.append(make_instances(tuple(image_size), proposal_boxes=boxes_view(out_boxes[n]), objectness_logits=out_scores[n]))
    return results


def _max_level_k(sel) -> int:
    # the widest level run can not exceed kmax; a tighter host-side bound keeps the mask workspace small
    return int(getattr(sel, "max_level_k", 0) or sel.kmax)
