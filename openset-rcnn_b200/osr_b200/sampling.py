"""ROI-head sampling glue: proposal <-> ground-truth matching and labelled sampling.

Drop-in for ``OpensetROIHeads.label_and_sample_proposals`` (``osrcnn_roi_heads.py:136-230``), same arguments and
returned ``List[Instances]`` (fields ``proposal_boxes``, ``objectness_logits``, ``gt_classes``, ``ious`` and the
targets' ``gt_*`` fields).  The per-image ``pairwise_iou`` matrix + ``Matcher`` + matched-IoU gather + class assignment
run as ONE kernel launch for the whole batch (``osr_match_label``: one thread per proposal, GT boxes in shared memory,
the G x P matrix is never materialised).  The random subsampling of the whole batch + the gather of every sampled field
is a second launch (``osr_sample_rois``: one random key per row, the rows with the smallest keys of each kind are kept -
the subset detectron2's ``subsample_labels`` keeps when its permutation is the arg-sort of those keys).  With an injected
``randperm`` the reference's per-image ``subsample_labels`` is replayed draw for draw instead (parity tests).

``match_proposals`` is the batched tensor-level entry (no ``Instances``), used by the pipeline / bench.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import _lib
from .structures import Boxes, Instances, boxes_view, make_instances

GT_LOGIT = math.log((1.0 - 1e-10) / (1e-10))


def match_proposals(boxes: torch.Tensor, box_offsets: torch.Tensor, gt_boxes: torch.Tensor, gt_classes: torch.Tensor,
                    gt_offsets: torch.Tensor, max_boxes_per_image: int, *, iou_threshold: float = 0.5,
                    background_label: int = 80, box_counts: Optional[torch.Tensor] = None,
                    box_counts_stride: int = 1) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """All images at once.  ``boxes`` (P,4) fp32 = the images' proposal lists concatenated, ``box_offsets`` (N+1)
    int32 on the device, likewise ``gt_boxes`` (G,4) / ``gt_classes`` (G) int64 / ``gt_offsets``.
    Returns ``(matched_idx int32, matched_iou fp32, matched_label int32, matched_class int64)``, each (P):
    detectron2 ``Matcher([thr],[0,1])`` on ``pairwise_iou(gt, proposals)`` + ``gt_classes[matched]`` / background.
    With ``box_counts`` (int32, device) image n owns ``[off[n], off[n] + box_counts[n * stride])`` - the padded
    ``RpnSelection`` layout - and rows outside every image are left untouched."""
    _lib.require_cuda(boxes, box_offsets, gt_boxes, gt_classes, gt_offsets)
    lib = _lib.lib()
    dev = boxes.device
    boxes = boxes.contiguous().float()
    gt_boxes = gt_boxes.contiguous().float()
    gt_classes = gt_classes.contiguous().to(torch.int64)
    assert box_offsets.dtype == torch.int32 and gt_offsets.dtype == torch.int32
    N = box_offsets.numel() - 1
    P = boxes.shape[0]
    midx = torch.empty(max(P, 1), dtype=torch.int32, device=dev)
    miou = torch.empty(max(P, 1), dtype=torch.float32, device=dev)
    mlab = torch.empty(max(P, 1), dtype=torch.int32, device=dev)
    mcls = torch.empty(max(P, 1), dtype=torch.int64, device=dev)
    if P > 0 and N > 0:
        rc = lib.osr_match_label(boxes.data_ptr(), box_offsets.data_ptr(), _lib.ptr(box_counts), int(box_counts_stride),
                                 _lib.ptr(gt_boxes if gt_boxes.numel() else None),
                                 _lib.ptr(gt_classes if gt_classes.numel() else None), gt_offsets.data_ptr(), int(N),
                                 int(max_boxes_per_image), float(iou_threshold), int(background_label),
                                 midx.data_ptr(), miou.data_ptr(), mlab.data_ptr(), mcls.data_ptr(),
                                 _lib.stream_ptr(dev))
        _lib.check(rc, "osr_match_label")
    return midx[:P], miou[:P], mlab[:P], mcls[:P]


def subsample_labels(labels: torch.Tensor, num_samples: int, positive_fraction: float, bg_label: int,
                     randperm: Optional[Callable[[int], torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """detectron2 ``sampling.subsample_labels`` (two ``torch.randperm`` draws on the labels' device)."""
    rp = randperm or (lambda n: torch.randperm(n, device=labels.device))
    positive = torch.nonzero((labels != -1) & (labels != bg_label), as_tuple=True)[0]
    negative = torch.nonzero(labels == bg_label, as_tuple=True)[0]
    num_pos = min(positive.numel(), int(num_samples * positive_fraction))
    num_neg = min(negative.numel(), num_samples - num_pos)
    perm1 = rp(positive.numel())[:num_pos].to(labels.device)
    perm2 = rp(negative.numel())[:num_neg].to(labels.device)
    return positive[perm1], negative[perm2]


def label_and_sample_proposals(proposals: List[Instances], targets: List[Instances], *, num_classes: int,
                               batch_size_per_image: int = 512, positive_fraction: float = 0.25,
                               iou_threshold: float = 0.5, proposal_append_gt: bool = True,
                               randperm: Optional[Callable[[int], torch.Tensor]] = None,
                               generator: Optional[torch.Generator] = None,
                               keys: Optional[Sequence[torch.Tensor]] = None) -> List[Instances]:
    """``OpensetROIHeads.label_and_sample_proposals(proposals, targets)`` (``osrcnn_roi_heads.py:136-230``); the
    keyword arguments are the module attributes it reads (``num_classes``, ``batch_size_per_image``,
    ``positive_fraction``, ``proposal_matcher`` threshold, ``proposal_append_gt``).  Like the reference it raises
    ``IndexError`` for an image without ground truth (the matched-IoU gather at ``:193`` indexes an empty matrix).

    Sampling: with ``randperm`` given (a callable standing for ``torch.randperm``) the reference's per-image
    ``subsample_labels`` is replayed draw for draw (parity tests inject the reference's own permutations).  Without it the
    whole batch is sampled at once on the device (``sample_rois`` = ``osr_sample_rois``: one ``torch.rand`` key per row, or
    ``keys`` = one fp32 tensor per image over its proposals followed by its appended ground truth; same distribution and
    order as the reference's two permutations per image, no per-image host syncs - the stage then has ONE host read, the
    per-image sample counts)."""
    N = len(proposals)
    assert len(targets) == N
    if N == 0:
        return []
    dev = proposals[0].get("proposal_boxes").tensor.device
    # ONE concatenation for the whole batch, laid out [proposals_0, gt_0, proposals_1, gt_1, ...] (add_ground_truth_to_proposals:
    # GT boxes after the proposals of their image, logit log((1-1e-10)/1e-10)); the per-image lists are views of it
    gts = [t.get("gt_boxes").tensor.to(dev) for t in targets]
    for g in gts:
        if len(g) == 0:
            raise IndexError("index 0 is out of bounds for dimension 0 with size 0 "
                             "(image without ground truth: osrcnn_roi_heads.py:193)")
    gt_logits = GT_LOGIT * torch.ones(sum(len(g) for g in gts), device=dev) if proposal_append_gt else None
    parts_b, parts_l, counts, gcounts = [], [], [], []
    go = 0
    for p, g in zip(proposals, gts):
        b = p.get("proposal_boxes").tensor
        parts_b.append(b)
        parts_l.append(p.get("objectness_logits"))
        c = b.shape[0]
        if proposal_append_gt:
            parts_b.append(g)
            parts_l.append(gt_logits[go:go + len(g)])
            c += len(g)
            go += len(g)
        counts.append(c)
        gcounts.append(len(g))
    boxes = torch.cat(parts_b, dim=0)
    logits = torch.cat(parts_l, dim=0)
    gt_boxes = torch.cat(gts, dim=0)
    gt_classes = torch.cat([t.get("gt_classes").to(dev) for t in targets], dim=0)
    offs, goffs = [0], [0]
    for c, gc in zip(counts, gcounts):
        offs.append(offs[-1] + c)
        goffs.append(goffs[-1] + gc)
    both = torch.tensor([offs, goffs], dtype=torch.int32).to(dev, non_blocking=True)   # one small H2D copy
    off, goff = both[0], both[1]
    midx, miou, mlab, mcls = match_proposals(boxes, off, gt_boxes, gt_classes, goff, max(counts),
                                             iou_threshold=iou_threshold, background_label=num_classes)
    if randperm is None:
        return _finish_batched(proposals, targets, boxes, logits, counts, offs, off, goff, midx, miou, mcls,
                               num_classes, batch_size_per_image, positive_fraction, proposal_append_gt, generator, keys)
    box_list = [boxes[offs[n]:offs[n + 1]] for n in range(N)]
    logit_list = [logits[offs[n]:offs[n + 1]] for n in range(N)]
    out = []
    b0 = 0
    for n, (p, t) in enumerate(zip(proposals, targets)):
        b1 = b0 + counts[n]
        cls_n = mcls[b0:b1]
        fg, bg = subsample_labels(cls_n, batch_size_per_image, positive_fraction, num_classes, randperm)
        sampled = torch.cat([fg, bg], dim=0)
        q = Instances(p.image_size)
        q.set("proposal_boxes", Boxes(box_list[n][sampled]))
        q.set("objectness_logits", logit_list[n][sampled])
        for name, value in p.get_fields().items():
            if name not in ("proposal_boxes", "objectness_logits") and not proposal_append_gt:
                q.set(name, value[sampled])
        q.set("gt_classes", cls_n[sampled])
        q.set("ious", miou[b0:b1][sampled])
        st = midx[b0:b1][sampled].long()
        for name, value in t.get_fields().items():
            if name.startswith("gt_") and not q.has(name):
                q.set(name, value.to(dev)[st] if hasattr(value, "to") else value[st])
        out.append(q)
        b0 = b1
    return out


def sample_rois(labels: torch.Tensor, keys: torch.Tensor, box_offsets: torch.Tensor, num_samples: int, num_pos_max: int,
                bg_label: int, *, box_counts: Optional[torch.Tensor] = None, box_counts_stride: int = 1,
                boxes: Optional[torch.Tensor] = None, logits: Optional[torch.Tensor] = None,
                ious: Optional[torch.Tensor] = None, matched_idx: Optional[torch.Tensor] = None,
                gt_offsets: Optional[torch.Tensor] = None, want_rois: bool = False, max_boxes_per_image: int = 0):
    """``osr_sample_rois``: detectron2 ``subsample_labels`` for all images in ONE launch (no host sync) + the gather of the
    sampled fields.  ``labels`` (P) int64 (= ``matched_class``), ``keys`` (P) fp32 random keys, ``box_offsets`` (N+1) int32.
    Per image the ``min(#pos, num_pos_max)`` positives and ``min(#neg, num_samples - kept positives)`` negatives with the
    smallest keys are kept (ties: lower row), positives first, each kind in ascending key order.  Returns a dict:
    ``index`` (N, num_samples) int32 (row inside the image, -1 beyond the count), ``count`` (N, 2) int32 = (kept positives,
    kept rows) and - for every source given - ``boxes`` / ``logits`` / ``classes`` / ``ious`` / ``gt`` (row in the
    concatenated targets), each (N, num_samples[, 4]), undefined beyond the count; ``want_rois``: also ``rois``
    (N, num_samples, 5) = (image, x1, y1, x2, y2), the rows ROIAlign consumes.  ``max_boxes_per_image``: host upper bound
    of an image's row count (lets the kernel keep the rows in shared memory; 0 = unknown, rows are re-read)."""
    _lib.require_cuda(labels, keys, box_offsets)
    lib = _lib.lib()
    dev = labels.device
    assert labels.dtype == torch.int64 and keys.dtype == torch.float32 and box_offsets.dtype == torch.int32
    labels, keys = labels.contiguous(), keys.contiguous()
    N = box_offsets.numel() - 1
    S = int(num_samples)
    out = dict(index=torch.empty((N, S), dtype=torch.int32, device=dev), count=torch.empty((N, 2), dtype=torch.int32, device=dev))
    if boxes is not None:
        boxes = boxes.contiguous().float()
        if not want_rois:
            out["boxes"] = torch.empty((N, S, 4), dtype=torch.float32, device=dev)
    if logits is not None:
        logits = logits.contiguous().float()
        out["logits"] = torch.empty((N, S), dtype=torch.float32, device=dev)
    if ious is not None:
        ious = ious.contiguous().float()
        out["ious"] = torch.empty((N, S), dtype=torch.float32, device=dev)
    if matched_idx is not None:
        assert matched_idx.dtype == torch.int32
        matched_idx = matched_idx.contiguous()
        out["gt"] = torch.empty((N, S), dtype=torch.int64, device=dev)
    out["classes"] = torch.empty((N, S), dtype=torch.int64, device=dev)
    if want_rois:
        assert boxes is not None
        out["rois"] = torch.empty((N, S, 5), dtype=torch.float32, device=dev)
    if N > 0:
        rc = lib.osr_sample_rois(labels.data_ptr(), keys.data_ptr(), box_offsets.data_ptr(), _lib.ptr(box_counts),
                                 int(box_counts_stride), int(N), int(max_boxes_per_image), S, int(num_pos_max), int(bg_label),
                                 _lib.ptr(boxes),
                                 _lib.ptr(logits), _lib.ptr(ious), _lib.ptr(matched_idx), _lib.ptr(gt_offsets),
                                 out["index"].data_ptr(), out["count"].data_ptr(), _lib.ptr(out.get("boxes")),
                                 _lib.ptr(out.get("logits")), out["classes"].data_ptr(), _lib.ptr(out.get("ious")),
                                 _lib.ptr(out.get("gt")), _lib.ptr(out.get("rois")), _lib.stream_ptr(dev))
        _lib.check(rc, "osr_sample_rois")
    return out


def sample_labels_batched(labels: torch.Tensor, valid: torch.Tensor, num_samples: int, positive_fraction: float,
                          bg_label: int, generator: Optional[torch.Generator] = None):
    """detectron2 ``subsample_labels`` for ALL images at once, without a host sync.  ``labels`` (N, Pmax) int64 padded,
    ``valid`` (N, Pmax) bool.  Per image: a uniformly random subset of at most ``int(num_samples * positive_fraction)``
    foreground rows followed by a uniformly random subset of background rows filling up to ``num_samples`` - the same
    distribution and the same (positives first) order as two ``torch.randperm`` draws, but from one ``torch.rand`` key per
    row and two per-image top-k selections.  Returns ``(index (N, num_samples) int64, count (N,) int64)``; entries at
    positions >= count[n] are undefined."""
    dev = labels.device
    N, Pmax = labels.shape
    pos = valid & (labels != -1) & (labels != bg_label)
    neg = valid & (labels == bg_label)
    keys = torch.rand((N, Pmax), device=dev, generator=generator)
    k = min(num_samples, Pmax)
    two = torch.full((), 2.0, device=dev)
    ip = torch.where(pos, keys, two).topk(k, dim=1, largest=False).indices
    ineg = torch.where(neg, keys, two).topk(k, dim=1, largest=False).indices
    npos = pos.sum(dim=1).clamp(max=int(num_samples * positive_fraction))
    nneg = torch.minimum(neg.sum(dim=1), num_samples - npos)
    slot = torch.arange(num_samples, device=dev)[None, :]
    from_pos = slot < npos[:, None]
    gp = ip.gather(1, slot.clamp(max=k - 1).expand(N, -1))
    gn = ineg.gather(1, (slot - npos[:, None]).clamp(min=0, max=k - 1))
    return torch.where(from_pos, gp, gn), npos + nneg


def _finish_batched(proposals, targets, boxes, logits, counts, offs, off, goff, midx, miou, mcls, num_classes,
                    batch_size_per_image, positive_fraction, proposal_append_gt, generator, keys=None):
    """Sampling + field gathering of ``label_and_sample_proposals`` for the whole batch: ONE launch (``osr_sample_rois``)
    draws the samples of every image and gathers boxes / logits / classes / IoUs / matched ground-truth rows, the targets'
    ``gt_*`` fields are concatenated once and gathered once, ONE host read of the per-image sample counts, then views."""
    N = len(proposals)
    dev = mcls.device
    total = offs[-1]
    if keys is None:
        keys_t = torch.rand(total, device=dev, generator=generator)
    else:
        keys_t = torch.cat([k.to(dev).float().reshape(-1) for k in keys]) if not torch.is_tensor(keys) else keys.to(dev).float()
        assert keys_t.numel() == total, "keys: one value per proposal (+ appended ground truth) of every image"
    S = int(batch_size_per_image)
    smp = sample_rois(mcls, keys_t, off, S, int(S * positive_fraction), num_classes, boxes=boxes, logits=logits, ious=miou,
                      matched_idx=midx, gt_offsets=goff, max_boxes_per_image=max(counts))
    cnt_dev = smp["count"][:, 1]
    idx, sb, sl, sc, si, sm_g = smp["index"], smp["boxes"], smp["logits"], smp["classes"], smp["ious"], smp["gt"]
    # the targets' gt_* fields, concatenated once and gathered once (tensor / Boxes fields; anything else per image below)
    gt_names = [name for name in targets[0].get_fields() if name.startswith("gt_")]
    batched, per_image = {}, []
    safe_g = None
    for name in gt_names:
        if name == "gt_classes":      # already there: the sampled class column (label_and_sample_proposals sets it first)
            continue
        vals = [t.get(name) for t in targets]
        if all(isinstance(v, Boxes) for v in vals) or all(torch.is_tensor(v) for v in vals):
            if safe_g is None:        # slots beyond an image's count hold garbage rows: clamp before gathering
                n_gt = sum(len(v) for v in vals)
                safe_g = sm_g.clamp(min=0, max=max(n_gt - 1, 0))
            if isinstance(vals[0], Boxes):
                batched[name] = ("boxes", torch.cat([v.tensor.to(dev) for v in vals], dim=0)[safe_g])
            else:
                batched[name] = ("tensor", torch.cat([v.to(dev) for v in vals], dim=0)[safe_g])
        else:
            per_image.append(name)
    n_s = cnt_dev.tolist()                                    # the one host sync of the stage
    out = []
    for n, (p, t) in enumerate(zip(proposals, targets)):
        k = n_s[n]
        f = {"proposal_boxes": boxes_view(sb[n, :k]), "objectness_logits": sl[n, :k]}   # k rows each, by construction
        if not proposal_append_gt:
            for name, value in p.get_fields().items():
                if name not in ("proposal_boxes", "objectness_logits"):
                    f[name] = value[idx[n, :k].long()]
        f["gt_classes"] = sc[n, :k]
        f["ious"] = si[n, :k]
        for name in gt_names:
            if name in f:
                continue
            if name in batched:
                kind, v = batched[name]
                f[name] = boxes_view(v[n, :k]) if kind == "boxes" else v[n, :k]
            else:
                value = t.get(name)
                st = sm_g[n, :k] - goff[n].long()
                f[name] = value.to(dev)[st] if hasattr(value, "to") else value[st]
        out.append(make_instances(p.image_size, **f))
    return out
