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


def _clip(boxes: torch.Tensor, image_size) -> torch.Tensor:
    h, w = image_size
    return torch.stack((boxes[:, 0].clamp(min=0, max=w), boxes[:, 1].clamp(min=0, max=h),
                        boxes[:, 2].clamp(min=0, max=w), boxes[:, 3].clamp(min=0, max=h)), dim=-1)


def softmax_classifier_inference(fg_instances: List[Instances], cls_score, *, unknown_id: int = 80,
                                 known_score_thresh: float = 0.05, known_nms_thresh: float = 0.5, known_topk: int = 100,
                                 unknown_score_thresh: float = 0.05, unknown_nms_thresh: float = 0.5,
                                 unknown_topk: int = 50, class_id=None) -> List[Instances]:
    """``SoftMaxClassifier.inference(fg_instances)`` (``softmax_classifier.py:287-346``): known detections through the
    linear classifier + softmax + per-class NMS, unknown detections through class-agnostic NMS, unknown first.
    The classifier GEMM and the softmax stay torch (not on the RoI path); the two per-image NMS loops
    (``fast_rcnn_inference_single_image_known/unknown``, ``:47-168``) run as ONE segmented NMS call each for the whole
    batch (``batched_nms_images``: segments = images, torchvision's per-image coordinate trick preserved)."""
    from .nms import batched_nms_images
    if not fg_instances:
        return []
    dev = fg_instances[0].get("scores").device
    kb, ks, kc, ub, us, uc, has_unknown = [], [], [], [], [], [], []
    known_masks = [inst.get("pred_classes") != unknown_id for inst in fg_instances]
    n_known = torch.stack([k.sum() for k in known_masks]).cpu().tolist()   # host sync 1
    for inst, k, nk in zip(fg_instances, known_masks, n_known):
        boxes = inst.get("pred_boxes").tensor
        # the classifier runs per image like the reference (a batched GEMM could round differently)
        probs = torch.softmax(cls_score(inst.get("features")[k]), dim=-1)
        b = boxes[k]
        valid = torch.isfinite(b).all(dim=1) & torch.isfinite(probs).all(dim=1)
        b, probs = b[valid], probs[valid][:, :-1]
        b = _clip(b, inst.image_size)
        inds = (probs > known_score_thresh).nonzero()
        kb.append(b[inds[:, 0]]); ks.append(probs[inds[:, 0], inds[:, 1]]); kc.append(inds[:, 1])
        has_unknown.append(nk < len(inst))
        bu, su = boxes[~k], inst.get("scores")[~k]
        vu = torch.isfinite(bu).all(dim=1) & torch.isfinite(su)
        bu, su = _clip(bu[vu], inst.image_size), su[vu]
        fu = su > unknown_score_thresh
        ub.append(bu[fu]); us.append(su[fu]); uc.append(torch.zeros(bu[fu].shape[0], dtype=torch.int64, device=dev))
    keep_k = batched_nms_images(kb, ks, kc, known_nms_thresh, topk_per_image=known_topk)
    keep_u = batched_nms_images(ub, us, uc, unknown_nms_thresh, topk_per_image=unknown_topk)
    out = []
    for n, inst in enumerate(fg_instances):
        res = Instances(inst.image_size)
        kcls = kc[n][keep_k[n]]
        if class_id is not None:
            kcls = class_id[kcls]
        if has_unknown[n]:
            ucls = (torch.zeros(len(keep_u[n]), device=dev) + unknown_id).long()
            res.set("pred_boxes", Boxes(torch.cat([ub[n][keep_u[n]], kb[n][keep_k[n]]])))
            res.set("scores", torch.cat([us[n][keep_u[n]], ks[n][keep_k[n]]]))
            res.set("pred_classes", torch.cat([ucls, kcls]))
        else:
            res.set("pred_boxes", Boxes(kb[n][keep_k[n]]))
            res.set("scores", ks[n][keep_k[n]])
            res.set("pred_classes", kcls)
        out.append(res)
    return out
