"""Oracle for the CF-RPN proposal stage.

Follows, line by line:

* ``openset_rcnn/modeling/proposal_generator/classification_free_rpn.py:558-610``
  (``predict_proposals`` / ``_decode_proposals``)
* ``openset_rcnn/modeling/find_top_proposals.py:22-128`` (``find_top_rpn_proposals``;
  the block commented out at ``:112-120`` is the ``nominal`` mode here)
* detectron2 v0.6 ``Box2BoxTransformLinear(normalize_by_size=True).apply_deltas``
  (used at ``classification_free_rpn.py:607``) and ``DefaultAnchorGenerator``
  (``classification_free_rpn.py:514``).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.nn.functional as F

from .structures import Boxes, Instances
from . import nms as _nms


def generate_anchors(
    grid_sizes: Sequence[Tuple[int, int]],
    strides: Sequence[int],
    sizes: Sequence[float],
    offset: float = 0.0,
    device="cpu",
) -> List[Boxes]:
    """detectron2 DefaultAnchorGenerator, 1 square anchor per cell (ASPECT_RATIOS [[1.0]]).

    Anchor at cell (y, x) of a level = [x*s - a/2, y*s - a/2, x*s + a/2, y*s + a/2];
    flat order (y*W + x)*A + a_idx (matches the permute at
    ``classification_free_rpn.py:518-529``).
    """
    out = []
    for (h, w), s, a in zip(grid_sizes, strides, sizes):
        # generate_cell_anchors: area = size**2; w = sqrt(area/ar); h = ar*w
        area = float(a) ** 2.0
        cw = (area / 1.0) ** 0.5
        ch = 1.0 * cw
        cell = torch.tensor([[-cw / 2.0, -ch / 2.0, cw / 2.0, ch / 2.0]], dtype=torch.float32, device=device)
        shifts_x = torch.arange(offset * s, w * s, step=s, dtype=torch.float32, device=device)
        shifts_y = torch.arange(offset * s, h * s, step=s, dtype=torch.float32, device=device)
        shift_y, shift_x = torch.meshgrid(shifts_y, shifts_x, indexing="ij")
        shift_x = shift_x.reshape(-1)
        shift_y = shift_y.reshape(-1)
        shifts = torch.stack((shift_x, shift_y, shift_x, shift_y), dim=1)
        out.append(Boxes((shifts.view(-1, 1, 4) + cell.view(1, -1, 4)).reshape(-1, 4)))
    return out


def apply_deltas_linear(deltas: torch.Tensor, boxes: torch.Tensor) -> torch.Tensor:
    """Box2BoxTransformLinear(normalize_by_size=True).apply_deltas (detectron2 v0.6).

    Call site: ``classification_free_rpn.py:607``.  Every op is a separately
    rounded fp32 op (no FMA) - the CUDA kernel spells them with ``__f*_rn``.
    """
    deltas = F.relu(deltas)
    boxes = boxes.to(deltas.dtype)

    ctr_x = 0.5 * (boxes[:, 0] + boxes[:, 2])
    ctr_y = 0.5 * (boxes[:, 1] + boxes[:, 3])
    stride_w = boxes[:, 2] - boxes[:, 0]
    stride_h = boxes[:, 3] - boxes[:, 1]
    strides = torch.stack([stride_w, stride_h, stride_w, stride_h], dim=1)
    deltas = deltas * strides

    l = deltas[:, 0::4]
    t = deltas[:, 1::4]
    r = deltas[:, 2::4]
    b = deltas[:, 3::4]

    pred_boxes = torch.zeros_like(deltas)
    pred_boxes[:, 0::4] = ctr_x[:, None] - l
    pred_boxes[:, 1::4] = ctr_y[:, None] - t
    pred_boxes[:, 2::4] = ctr_x[:, None] + r
    pred_boxes[:, 3::4] = ctr_y[:, None] + b
    return pred_boxes


def decode_proposals(anchors: List[Boxes], pred_anchor_deltas: List[torch.Tensor]) -> List[torch.Tensor]:
    """``ClsFreeRPN._decode_proposals`` (``classification_free_rpn.py:591-610``)."""
    N = pred_anchor_deltas[0].shape[0]
    proposals = []
    for anchors_i, deltas_i in zip(anchors, pred_anchor_deltas):
        B = anchors_i.tensor.size(1)
        deltas_i = deltas_i.reshape(-1, B)
        anchors_e = anchors_i.tensor.unsqueeze(0).expand(N, -1, -1).reshape(-1, B)
        proposals_i = apply_deltas_linear(deltas_i, anchors_e)
        proposals.append(proposals_i.view(N, -1, B))
    return proposals


def topk_stable(logits: torch.Tensor, k: int):
    """Contract used by the CUDA kernel: score descending, ties -> lower flat index first.

    ``torch.topk`` (``find_top_proposals.py:75``) leaves the order of equal
    scores unspecified; on tie-free inputs both agree exactly (tested).
    """
    vals, idx = torch.sort(logits, dim=1, descending=True, stable=True)
    return vals[:, :k], idx[:, :k]


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
    topk_impl: str = "stable",
) -> List[Instances]:
    """``find_top_rpn_proposals`` (``find_top_proposals.py:22-128``).

    mode="as_shipped": NMS + post_nms_topk are skipped, exactly like the
    reference (lines 112-120 are commented out there).
    mode="nominal": stock detectron2 behaviour (the commented block executed).
    """
    assert mode in ("as_shipped", "nominal")
    num_images = len(image_sizes)
    device = proposals[0].device

    topk_scores, topk_proposals, level_ids = [], [], []
    batch_idx = torch.arange(num_images, device=device)
    for level_id, (proposals_i, logits_i) in enumerate(zip(proposals, pred_objectness_logits)):
        Hi_Wi_A = logits_i.shape[1]
        num_proposals_i = min(Hi_Wi_A, pre_nms_topk)
        if topk_impl == "torch":
            topk_scores_i, topk_idx = logits_i.topk(num_proposals_i, dim=1)
        else:
            topk_scores_i, topk_idx = topk_stable(logits_i, num_proposals_i)
        topk_proposals_i = proposals_i[batch_idx[:, None], topk_idx]
        topk_proposals.append(topk_proposals_i)
        topk_scores.append(topk_scores_i)
        level_ids.append(torch.full((num_proposals_i,), level_id, dtype=torch.int64, device=device))

    topk_scores = torch.cat(topk_scores, dim=1)
    topk_proposals = torch.cat(topk_proposals, dim=1)
    level_ids = torch.cat(level_ids, dim=0)

    results: List[Instances] = []
    for n, image_size in enumerate(image_sizes):
        boxes = Boxes(topk_proposals[n])
        scores_per_img = topk_scores[n]
        lvl = level_ids

        valid_mask = torch.isfinite(boxes.tensor).all(dim=1) & torch.isfinite(scores_per_img)
        if not valid_mask.all():
            if training:
                raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")
            boxes = boxes[valid_mask]
            scores_per_img = scores_per_img[valid_mask]
            lvl = lvl[valid_mask]
        boxes.clip(image_size)

        keep = boxes.nonempty(threshold=min_box_size)
        if keep.sum().item() != len(boxes):
            boxes, scores_per_img, lvl = boxes[keep], scores_per_img[keep], lvl[keep]

        res = Instances(image_size)
        if mode == "nominal":
            keep = _nms.batched_nms(boxes.tensor, scores_per_img, lvl, nms_thresh)
            keep = keep[:post_nms_topk]
            res.proposal_boxes = boxes[keep]
            res.objectness_logits = scores_per_img[keep]
            res.level_ids = lvl[keep]
        else:
            res.proposal_boxes = boxes
            res.objectness_logits = scores_per_img
            res.level_ids = lvl  # extra (not in the reference): lets tests check the level split
        results.append(res)
    return results


def predict_proposals(
    anchors: List[Boxes],
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
    topk_impl: str = "stable",
) -> List[Instances]:
    """``ClsFreeRPN.predict_proposals`` (``classification_free_rpn.py:558-589``)."""
    with torch.no_grad():
        pred_proposals = decode_proposals(anchors, pred_anchor_deltas)
        return find_top_rpn_proposals(
            pred_proposals,
            pred_centerness,
            image_sizes,
            nms_thresh,
            pre_nms_topk,
            post_nms_topk,
            min_box_size,
            training,
            mode=mode,
            topk_impl=topk_impl,
        )
