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
from .structures import Boxes, Instances, boxes_view, cat_rows, flat_prefixes, make_instances

SCALE_CLAMP = math.log(1000.0 / 16)


def _decode(proposal_deltas, ious, proposals, weights, mean_type, score_thresh):
    lib = _lib.lib()
    pb = cat_rows([p.get("proposal_boxes").tensor for p in proposals]).contiguous().float()
    ctr = cat_rows([p.get("objectness_logits") for p in proposals]).contiguous().float()
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
    # every image's output rows in ONE gather per field (global indices from the host-known counts); the per-image
    # Instances are split views of those tensors, which the next stages recognise and do not copy again (cat_rows)
    ks = [c if topk_per_image < 0 else min(c, topk_per_image) for c in counts]
    begins, o = [], 0
    for l in lens:
        begins.append(o)
        o += l
    P, B = flat_prefixes(begins, ks, dev)
    g = keep_idx.index_select(0, P) + B
    out_boxes = boxes.index_select(0, g).split(ks)
    out_scores = scores.index_select(0, g).split(ks)
    out_feats = feats.index_select(0, g).split(ks)
    out_cls = torch.zeros(len(g), dtype=torch.int64, device=dev).split(ks)
    kept_all = fin_pos.index_select(0, g) - fin_base.index_select(0, seg.index_select(0, g))
    out_kept = kept_all.split(ks)
    results, kept = [], []
    for n, p in enumerate(proposals):
        results.append(make_instances(p.image_size, pred_boxes=boxes_view(out_boxes[n]), scores=out_scores[n],
                                      pred_classes=out_cls[n], features=out_feats[n]))
        kept.append(out_kept[n])
    return results, kept


def _clip(boxes: torch.Tensor, image_size) -> torch.Tensor:
    h, w = image_size
    return torch.stack((boxes[:, 0].clamp(min=0, max=w), boxes[:, 1].clamp(min=0, max=h),
                        boxes[:, 2].clamp(min=0, max=w), boxes[:, 3].clamp(min=0, max=h)), dim=-1)


def _nonzero_known_size(mask: torch.Tensor, size: int) -> torch.Tensor:
    """``mask.nonzero()`` when the number of hits is already known on the host: no device->host sync."""
    try:
        return torch.nonzero_static(mask, size=int(size))
    except (AttributeError, RuntimeError, NotImplementedError):
        return mask.nonzero()


def softmax_classifier_inference(fg_instances: List[Instances], cls_score, *, unknown_id: int = 80,
                                 known_score_thresh: float = 0.05, known_nms_thresh: float = 0.5, known_topk: int = 100,
                                 unknown_score_thresh: float = 0.05, unknown_nms_thresh: float = 0.5,
                                 unknown_topk: int = 50, class_id=None) -> List[Instances]:
    """``SoftMaxClassifier.inference(fg_instances)`` (``softmax_classifier.py:287-346``): known detections through the
    linear classifier + softmax + per-class NMS, unknown detections through class-agnostic NMS, unknown first.

    The whole batch is processed at once: the per-image Python loop of the reference (masks, finite filters, clip,
    threshold, ``nonzero``: ~25 launches and 3 host syncs per image) becomes one pass over the concatenated detections
    with per-row image ids, its two per-image NMS loops (``fast_rcnn_inference_single_image_known/unknown``,
    ``:47-168``) one segmented NMS call each (``batched_nms_flat``: segments = images, torchvision's per-image
    coordinate trick preserved), and the per-image result assembly one gather per field in a host-built output order
    (the per-image ``Instances`` are split views).  Four host reads per batch (known counts, candidate counts, two NMS
    keep counts).  Only the classifier GEMM stays per image, on exactly the rows the reference feeds it (a batched GEMM
    may pick another algorithm and round differently, and the score threshold / NMS order would see it); the softmax is
    row-wise and runs once on the concatenated logits."""
    from .nms import batched_nms_flat
    if not fg_instances:
        return []
    import numpy as np
    N = len(fg_instances)
    dev = fg_instances[0].get("scores").device
    sizes = [len(x) for x in fg_instances]
    T = int(sum(sizes))
    boxes = cat_rows([x.get("pred_boxes").tensor for x in fg_instances])
    scores = cat_rows([x.get("scores") for x in fg_instances])
    pcls = cat_rows([x.get("pred_classes") for x in fg_instances])
    feats = cat_rows([x.get("features") for x in fg_instances])
    img = torch.repeat_interleave(torch.arange(N, device=dev), torch.tensor(sizes, device=dev), output_size=T)
    hw = torch.tensor([[float(x.image_size[0]), float(x.image_size[1])] for x in fg_instances], device=dev)
    known = pcls != unknown_id
    n_known = torch.zeros(N, dtype=torch.int64, device=dev).index_add_(0, img, known.long()).cpu().tolist()   # host read 1
    n_unknown = [s - k for s, k in zip(sizes, n_known)]
    kidx = _nonzero_known_size(known, sum(n_known))[:, 0]          # image-major, ascending
    uidx = _nonzero_known_size(~known, sum(n_unknown))[:, 0]

    def clip(b, rows):   # Boxes.clip(image_size) with per-row image sizes
        h, w = hw[rows, 0], hw[rows, 1]
        zero = torch.zeros((), device=dev)
        return torch.stack((torch.minimum(torch.maximum(b[:, 0], zero), w), torch.minimum(torch.maximum(b[:, 1], zero), h),
                            torch.minimum(torch.maximum(b[:, 2], zero), w), torch.minimum(torch.maximum(b[:, 3], zero), h)), dim=-1)

    # ---- known detections: classifier (one gather, then per image), finite filter, clip, score threshold -> (row, class) candidates
    fk = feats.index_select(0, kidx)
    logits_l, o = [], 0
    for n in range(N):
        logits_l.append(cls_score(fk[o:o + n_known[n]]))
        o += n_known[n]
    probs = torch.softmax(torch.cat(logits_l, dim=0), dim=-1)
    kb = boxes.index_select(0, kidx)
    kimg = img.index_select(0, kidx)
    valid_k = torch.isfinite(kb).all(dim=1) & torch.isfinite(probs).all(dim=1)
    cand = valid_k[:, None] & (probs[:, :-1] > known_score_thresh)
    # ---- unknown detections: finite filter, clip, score threshold
    ub = boxes.index_select(0, uidx)
    us = scores.index_select(0, uidx)
    uimg = img.index_select(0, uidx)
    keep_u_mask = torch.isfinite(ub).all(dim=1) & torch.isfinite(us) & (us > unknown_score_thresh)
    cnt = torch.stack((torch.zeros(N, dtype=torch.int64, device=dev).index_add_(0, kimg, cand.sum(dim=1)),
                       torch.zeros(N, dtype=torch.int64, device=dev).index_add_(0, uimg, keep_u_mask.long()))).cpu().tolist()   # host read 2
    cnt_k, cnt_u = cnt
    inds = _nonzero_known_size(cand, sum(cnt_k))                   # (row, class), row-major = the reference's per-image order
    kb_c = clip(kb, kimg)
    c_boxes = kb_c.index_select(0, inds[:, 0])
    c_scores = probs[inds[:, 0], inds[:, 1]]
    c_cls = inds[:, 1]
    c_img = kimg.index_select(0, inds[:, 0])
    usel = _nonzero_known_size(keep_u_mask, sum(cnt_u))[:, 0]
    u_boxes = clip(ub, uimg).index_select(0, usel)
    u_scores = us.index_select(0, usel)
    u_img = uimg.index_select(0, usel)
    u_cls = torch.zeros(u_img.shape[0], dtype=torch.int64, device=dev)
    gk, kk = batched_nms_flat(c_boxes, c_scores, c_cls, c_img, cnt_k, known_nms_thresh, topk_per_image=known_topk)        # host read 3
    gu, ku = batched_nms_flat(u_boxes, u_scores, u_cls, u_img, cnt_u, unknown_nms_thresh, topk_per_image=unknown_topk)   # host read 4
    # ---- result: per image the unknown detections first, then the known ones (:330-342).  Images without unknown
    # FOREGROUND rows take the reference's known-only branch - the same tensors, since their unknown block is empty.
    kcls = c_cls.index_select(0, gk)
    if class_id is not None:
        kcls = class_id[kcls]
    all_boxes = torch.cat((u_boxes.index_select(0, gu), c_boxes.index_select(0, gk)))
    all_scores = torch.cat((u_scores.index_select(0, gu), c_scores.index_select(0, gk)))
    all_cls = torch.cat((torch.full((gu.shape[0],), int(unknown_id), dtype=torch.int64, device=dev), kcls))
    ku_a, kk_a = np.asarray(ku, dtype=np.int64), np.asarray(kk, dtype=np.int64)
    tot = ku_a + kk_a
    u_first, k_first = np.cumsum(ku_a) - ku_a, int(ku_a.sum()) + np.cumsum(kk_a) - kk_a
    out_first = np.cumsum(tot) - tot
    within = np.arange(int(tot.sum()), dtype=np.int64) - np.repeat(out_first, tot)
    ku_r = np.repeat(ku_a, tot)
    perm = np.where(within < ku_r, np.repeat(u_first, tot) + within, np.repeat(k_first, tot) + within - ku_r)
    perm = torch.from_numpy(perm).to(dev)
    tot_l = tot.tolist()
    ob = all_boxes.index_select(0, perm).split(tot_l)
    osc = all_scores.index_select(0, perm).split(tot_l)
    oc = all_cls.index_select(0, perm).split(tot_l)
    out = []
    for n, inst in enumerate(fg_instances):
        out.append(make_instances(inst.image_size, pred_boxes=boxes_view(ob[n]), scores=osc[n], pred_classes=oc[n]))
    return out
