"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): restatement of the ROI-head inference post-processing.

Follows ``openset_rcnn/modeling/roi_heads/osrcnn_fast_rcnn.py``: ``OpensetFastRCNNOutputLayers.inference`` (``:380-404``),
``predict_boxes`` (``:406-430``) -> detectron2 v0.6 ``Box2BoxTransform.apply_deltas``,
``predict_ious`` (``:432-452``), ``fast_rcnn_inference`` / ``fast_rcnn_inference_single_image`` (``:45-145``).
Device-agnostic torch; ``batched_nms`` is the real torchvision binary (oracle/nms.py).
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import torch

from .nms import batched_nms
from .structures import Boxes, Instances

SCALE_CLAMP = math.log(1000.0 / 16)


def apply_deltas(deltas: torch.Tensor, boxes: torch.Tensor, weights: Sequence[float] = (10.0, 10.0, 5.0, 5.0),
                 scale_clamp: float = SCALE_CLAMP) -> torch.Tensor:
    """detectron2 ``Box2BoxTransform.apply_deltas``."""
    deltas = deltas.float()
    boxes = boxes.to(deltas.dtype)
    widths = boxes[:, 2] - boxes[:, 0]
    heights = boxes[:, 3] - boxes[:, 1]
    ctr_x = boxes[:, 0] + 0.5 * widths
    ctr_y = boxes[:, 1] + 0.5 * heights
    wx, wy, ww, wh = weights
    dx = deltas[:, 0::4] / wx
    dy = deltas[:, 1::4] / wy
    dw = deltas[:, 2::4] / ww
    dh = deltas[:, 3::4] / wh
    dw = torch.clamp(dw, max=scale_clamp)
    dh = torch.clamp(dh, max=scale_clamp)
    pred_ctr_x = dx * widths[:, None] + ctr_x[:, None]
    pred_ctr_y = dy * heights[:, None] + ctr_y[:, None]
    pred_w = torch.exp(dw) * widths[:, None]
    pred_h = torch.exp(dh) * heights[:, None]
    x1 = pred_ctr_x - 0.5 * pred_w
    y1 = pred_ctr_y - 0.5 * pred_h
    x2 = pred_ctr_x + 0.5 * pred_w
    y2 = pred_ctr_y + 0.5 * pred_h
    return torch.stack((x1, y1, x2, y2), dim=-1).reshape(deltas.shape)


def predict_boxes(proposal_deltas, proposals: List[Instances], weights=(10.0, 10.0, 5.0, 5.0)):
    """``osrcnn_fast_rcnn.py:406-430``."""
    if not len(proposals):
        return []
    n = [len(p) for p in proposals]
    pb = torch.cat([p.get("proposal_boxes").tensor for p in proposals], dim=0)
    return apply_deltas(proposal_deltas, pb, weights).split(n)


def predict_ious(ious, proposals: List[Instances], mean_type: str = "geometric"):
    """``osrcnn_fast_rcnn.py:432-452``."""
    centerness = torch.cat([p.get("objectness_logits") for p in proposals]).unsqueeze(1)
    if mean_type == "geometric":
        scores = torch.sqrt(ious * centerness)
    if mean_type == "arithmetic":
        scores = (ious + centerness) / 2.0
    return scores.split([len(p) for p in proposals])


def fast_rcnn_inference_single_image(boxes, scores, image_shape, feats, score_thresh, nms_thresh, topk_per_image):
    """``osrcnn_fast_rcnn.py:89-145``."""
    valid_mask = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores).all(dim=1)
    if not valid_mask.all():
        boxes = boxes[valid_mask]
        scores = scores[valid_mask]
        feats = feats[valid_mask]
    num_bbox_reg_classes = boxes.shape[1] // 4
    b = Boxes(boxes.reshape(-1, 4))
    b.clip(image_shape)
    boxes = b.tensor.view(-1, num_bbox_reg_classes, 4)
    filter_mask = scores > score_thresh
    filter_inds = filter_mask.nonzero()
    if num_bbox_reg_classes == 1:
        boxes = boxes[filter_inds[:, 0], 0]
    else:
        boxes = boxes[filter_mask]
    scores = scores[filter_mask]
    feats = feats[filter_inds[:, 0]]
    keep = batched_nms(boxes, scores, filter_inds[:, 1], nms_thresh)
    if topk_per_image >= 0:
        keep = keep[:topk_per_image]
    boxes, scores, feats, filter_inds = boxes[keep], scores[keep], feats[keep], filter_inds[keep]
    result = Instances(image_shape)
    result.set("pred_boxes", Boxes(boxes))
    result.set("scores", scores)
    result.set("pred_classes", filter_inds[:, 1])
    result.set("features", feats)
    return result, filter_inds[:, 0]


def inference(predictions: Tuple[torch.Tensor, torch.Tensor], proposals: List[Instances], box_features: torch.Tensor, *,
              weights=(10.0, 10.0, 5.0, 5.0), mean_type="geometric", score_thresh=0.0, nms_thresh=0.5,
              topk_per_image=100):
    """``OpensetFastRCNNOutputLayers.inference`` (``:380-404``) -> ``fast_rcnn_inference`` (``:45-87``).
    NOTE the reference's finite filter compacts ``boxes`` before indices are taken, so the returned kept indices are
    positions in the finite-filtered list (as in detectron2)."""
    boxes = predict_boxes(predictions[0], proposals, weights)
    ious = predict_ious(predictions[1], proposals, mean_type)
    shapes = [p.image_size for p in proposals]
    feats = box_features.split([len(p) for p in proposals])
    res = [fast_rcnn_inference_single_image(b, s, sh, f, score_thresh, nms_thresh, topk_per_image)
           for s, b, sh, f in zip(ious, boxes, shapes, feats)]
    return [r[0] for r in res], [r[1] for r in res]


# ---------------------------------------------------------------------------------------------------------
# SoftMaxClassifier.inference (openset_rcnn/modeling/roi_heads/softmax_classifier.py:47-168, 287-346)
def _single_image_known(boxes, scores, image_shape, score_thresh, nms_thresh, topk_per_image):
    """``fast_rcnn_inference_single_image_known`` (``softmax_classifier.py:47-104``)."""
    valid_mask = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores).all(dim=1)
    if not valid_mask.all():
        boxes = boxes[valid_mask]
        scores = scores[valid_mask]
    scores = scores[:, :-1]
    num_bbox_reg_classes = boxes.shape[1] // 4
    b = Boxes(boxes.reshape(-1, 4))
    b.clip(image_shape)
    boxes = b.tensor.view(-1, num_bbox_reg_classes, 4)
    filter_mask = scores > score_thresh
    filter_inds = filter_mask.nonzero()
    if num_bbox_reg_classes == 1:
        boxes = boxes[filter_inds[:, 0], 0]
    else:
        boxes = boxes[filter_mask]
    scores = scores[filter_mask]
    keep = batched_nms(boxes, scores, filter_inds[:, 1], nms_thresh)
    if topk_per_image >= 0:
        keep = keep[:topk_per_image]
    boxes, scores, filter_inds = boxes[keep], scores[keep], filter_inds[keep]
    result = Instances(image_shape)
    result.set("pred_boxes", Boxes(boxes))
    result.set("scores", scores)
    result.set("pred_classes", filter_inds[:, 1])
    return result


def _single_image_unknown(boxes, scores, image_shape, score_thresh, nms_thresh, topk_per_image, unknown_id):
    """``fast_rcnn_inference_single_image_unknown`` (``softmax_classifier.py:106-168``)."""
    scores = scores.unsqueeze(dim=1)
    valid_mask = torch.isfinite(boxes).all(dim=1) & torch.isfinite(scores).all(dim=1)
    if not valid_mask.all():
        boxes = boxes[valid_mask]
        scores = scores[valid_mask]
    num_bbox_reg_classes = boxes.shape[1] // 4
    b = Boxes(boxes.reshape(-1, 4))
    b.clip(image_shape)
    boxes = b.tensor.view(-1, num_bbox_reg_classes, 4)
    filter_mask = scores > score_thresh
    filter_inds = filter_mask.nonzero()
    boxes = boxes[filter_inds[:, 0], 0] if num_bbox_reg_classes == 1 else boxes[filter_mask]
    scores = scores[filter_mask]
    keep = batched_nms(boxes, scores, filter_inds[:, 1], nms_thresh)
    if topk_per_image >= 0:
        keep = keep[:topk_per_image]
    boxes, scores = boxes[keep], scores[keep]
    result = Instances(image_shape)
    result.set("pred_boxes", Boxes(boxes))
    result.set("scores", scores)
    result.set("pred_classes", (torch.zeros(len(scores), device=scores.device) + unknown_id).long())
    return result


def softmax_classifier_inference(fg_instances: List[Instances], cls_score, *, unknown_id: int = 80,
                                 known_score_thresh: float = 0.05, known_nms_thresh: float = 0.5, known_topk: int = 100,
                                 unknown_score_thresh: float = 0.05, unknown_nms_thresh: float = 0.5,
                                 unknown_topk: int = 50, class_id=None) -> List[Instances]:
    """``SoftMaxClassifier.inference`` (``softmax_classifier.py:287-346``); ``cls_score`` is the module's linear layer,
    ``unknown_id`` = 80 (OpenDet benchmark) or 1000, ``class_id`` the GraspNet id table (None = benchmark)."""
    results = []
    for inst in fg_instances:
        known = inst.get("pred_classes") != unknown_id
        known_features = inst.get("features")[known]
        known_probs = torch.softmax(cls_score(known_features), dim=-1)
        result_k = _single_image_known(inst.get("pred_boxes")[known].tensor, known_probs, inst.image_size,
                                       known_score_thresh, known_nms_thresh, known_topk)
        res = Instances(inst.image_size)
        kc = result_k.get("pred_classes") if class_id is None else class_id[result_k.get("pred_classes")]
        if not known.all():
            result_unk = _single_image_unknown(inst.get("pred_boxes")[~known].tensor, inst.get("scores")[~known],
                                               inst.image_size, unknown_score_thresh, unknown_nms_thresh, unknown_topk,
                                               unknown_id)
            res.set("pred_boxes", Boxes(torch.cat([result_unk.get("pred_boxes").tensor, result_k.get("pred_boxes").tensor])))
            res.set("scores", torch.cat([result_unk.get("scores"), result_k.get("scores")]))
            res.set("pred_classes", torch.cat([result_unk.get("pred_classes"), kc]))
        else:
            res.set("pred_boxes", result_k.get("pred_boxes"))
            res.set("scores", result_k.get("scores"))
            res.set("pred_classes", kc)
        results.append(res)
    return results
