"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): CPU/any-device restatement of the ROI-head sampling glue.

Follows ``OpensetROIHeads.label_and_sample_proposals`` (``openset_rcnn/modeling/roi_heads/osrcnn_roi_heads.py:136-230``)
and the detectron2 v0.6 helpers it calls (detectron2 is not vendored in the reference):
``add_ground_truth_to_proposals`` (``:177-178``), ``pairwise_iou`` (``:187-189``), ``Matcher([0.5],[0,1],
allow_low_quality_matches=False)`` (``:190``), the matched-IoU gather the reference added (``:193``),
``ROIHeads._sample_proposals`` + ``subsample_labels`` (``:195-197``) and the field copies (``:203-218``).
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Tuple

import torch

from .structures import Boxes, Instances, pairwise_iou

GT_LOGIT = math.log((1.0 - 1e-10) / (1e-10))   # add_ground_truth_to_proposals_single_image: objectness of appended GT


def add_ground_truth_to_proposals(targets: List[Instances], proposals: List[Instances]) -> List[Instances]:
    """detectron2 ``proposal_utils.add_ground_truth_to_proposals``: GT boxes appended AFTER the proposals with
    ``objectness_logits = log((1-1e-10)/1e-10)``."""
    out = []
    for t, p in zip(targets, proposals):
        gt_boxes = t.get("gt_boxes")
        dev = p.get("objectness_logits").device
        logits = GT_LOGIT * torch.ones(len(gt_boxes.tensor), device=dev)
        q = Instances(p.image_size)
        q.set("proposal_boxes", Boxes(torch.cat((p.get("proposal_boxes").tensor, gt_boxes.tensor.to(dev)), dim=0)))
        q.set("objectness_logits", torch.cat((p.get("objectness_logits"), logits), dim=0))
        out.append(q)
    return out


def matcher(match_quality_matrix: torch.Tensor, threshold: float = 0.5) -> Tuple[torch.Tensor, torch.Tensor]:
    """detectron2 ``Matcher([thr], [0, 1], allow_low_quality_matches=False).__call__`` on a (G, P) IoU matrix."""
    P = match_quality_matrix.shape[1]
    if match_quality_matrix.numel() == 0:
        return (match_quality_matrix.new_full((P,), 0, dtype=torch.int64),
                match_quality_matrix.new_full((P,), 0, dtype=torch.int8))
    matched_vals, matches = match_quality_matrix.max(dim=0)
    labels = matches.new_full(matches.size(), 1, dtype=torch.int8)
    for lab, low, high in ((0, -float("inf"), threshold), (1, threshold, float("inf"))):
        labels[(matched_vals >= low) & (matched_vals < high)] = lab
    return matches, labels


def subsample_labels(labels: torch.Tensor, num_samples: int, positive_fraction: float, bg_label: int,
                     randperm: Callable[[int], torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """detectron2 ``sampling.subsample_labels``; ``randperm(n)`` stands for ``torch.randperm(n, device=...)``."""
    positive = torch.nonzero((labels != -1) & (labels != bg_label), as_tuple=True)[0]
    negative = torch.nonzero(labels == bg_label, as_tuple=True)[0]
    num_pos = int(num_samples * positive_fraction)
    num_pos = min(positive.numel(), num_pos)
    num_neg = num_samples - num_pos
    num_neg = min(negative.numel(), num_neg)
    perm1 = randperm(positive.numel())[:num_pos].to(positive.device)
    perm2 = randperm(negative.numel())[:num_neg].to(negative.device)
    return positive[perm1], negative[perm2]


def keyed_randperm(keys_per_call):
    """The permutation stream under which ``subsample_labels`` keeps, per kind, the rows with the SMALLEST random keys in
    ascending (key, row) order - the contract of the one-launch CUDA sampler (``osr_sample_rois``): call i returns the
    stable arg-sort of ``keys_per_call[i]`` (the keys of the image's positives, then of its negatives, image after image:
    the order ``subsample_labels`` draws its permutations in)."""
    it = iter(keys_per_call)

    def rp(n):
        k = next(it)
        assert k.numel() == n
        return torch.argsort(k, stable=True)
    return rp


def sample_proposals(matched_idxs, matched_labels, gt_classes, *, num_classes, batch_size_per_image,
                     positive_fraction, randperm):
    """detectron2 ``ROIHeads._sample_proposals``."""
    has_gt = gt_classes.numel() > 0
    if has_gt:
        gt_classes = gt_classes[matched_idxs]
        gt_classes[matched_labels == 0] = num_classes
        gt_classes[matched_labels == -1] = -1
    else:
        gt_classes = torch.zeros_like(matched_idxs) + num_classes
    fg, bg = subsample_labels(gt_classes, batch_size_per_image, positive_fraction, num_classes, randperm)
    sampled = torch.cat([fg, bg], dim=0)
    return sampled, gt_classes[sampled]


def label_and_sample_proposals(proposals: List[Instances], targets: List[Instances], *, num_classes: int,
                               batch_size_per_image: int = 512, positive_fraction: float = 0.25,
                               iou_threshold: float = 0.5, proposal_append_gt: bool = True,
                               randperm: Optional[Callable[[int], torch.Tensor]] = None) -> List[Instances]:
    """``osrcnn_roi_heads.py:136-230`` (event-storage logging at :225-228 omitted)."""
    randperm = randperm or (lambda n: torch.randperm(n))
    if proposal_append_gt:
        proposals = add_ground_truth_to_proposals(targets, proposals)
    out = []
    for p, t in zip(proposals, targets):
        has_gt = len(t) > 0
        m = pairwise_iou(t.get("gt_boxes"), p.get("proposal_boxes"))
        matched_idxs, matched_labels = matcher(m, iou_threshold)
        matched_iou = m[matched_idxs, torch.arange(m.shape[1], device=m.device)]   # :193 (IndexError when G == 0)
        sampled, gt_classes = sample_proposals(matched_idxs, matched_labels, t.get("gt_classes"),
                                               num_classes=num_classes, batch_size_per_image=batch_size_per_image,
                                               positive_fraction=positive_fraction, randperm=randperm)
        q = p[sampled]
        q.set("gt_classes", gt_classes)
        q.set("ious", matched_iou[sampled])
        if has_gt:
            st = matched_idxs[sampled]
            for name, value in t.get_fields().items():
                if name.startswith("gt_") and not q.has(name):
                    q.set(name, value[st])
        out.append(q)
    return out
