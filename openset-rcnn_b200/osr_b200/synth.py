"""Synthetic inputs at the RoI-path boundary.

The reference's random-init CF-RPN head emits sub-pixel boxes, so
benchmarks and parity tests inject synthetic *head outputs* (deltas + centerness) and FPN
feature maps of the real R50-FPN shapes instead of running a random-init head.
Everything is seeded; shapes follow ``classification_free_rpn.py:518-529`` (permuted
(N, HWA, 4) deltas and (N, HWA) centerness) and ``osrcnn_roi_heads.py:306`` (P2..P5 NCHW).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch

RPN_STRIDES = (4, 8, 16, 32, 64)
RPN_SIZES = (32, 64, 128, 256, 512)
POOL_SCALES = (1.0 / 4, 1.0 / 8, 1.0 / 16, 1.0 / 32)


def padded_hw(h: int, w: int, divisibility: int = 32) -> Tuple[int, int]:
    return (int(math.ceil(h / divisibility) * divisibility), int(math.ceil(w / divisibility) * divisibility))


def fpn_grid_sizes(h: int, w: int) -> List[Tuple[int, int]]:
    """(H_l, W_l) for p2..p6 of an image padded to a multiple of 32 (p6 = maxpool(k=1,s=2) of p5)."""
    ph, pw = padded_hw(h, w)
    sizes = [(ph // s, pw // s) for s in RPN_STRIDES[:4]]
    h5, w5 = sizes[-1]
    sizes.append(((h5 - 1) // 2 + 1, (w5 - 1) // 2 + 1))
    return sizes


def make_anchors(grid_sizes: Sequence[Tuple[int, int]], device="cpu") -> List[torch.Tensor]:
    """One square anchor per cell, sizes 32..512, strides 4..64, offset 0 (detectron2
    DefaultAnchorGenerator with ANCHOR_GENERATOR.SIZES [[32],[64],[128],[256],[512]],
    ASPECT_RATIOS [[1.0]] - ``configs/Base-RCNN-FPN.yaml:10`` + ``configs/VOC-COCO/*.yaml:7-8``).
    Returns L tensors (H_l*W_l, 4) fp32; all values are exact in fp32."""
    out = []
    for (h, w), s, a in zip(grid_sizes, RPN_STRIDES, RPN_SIZES):
        ys = torch.arange(h, dtype=torch.float32, device=device) * s
        xs = torch.arange(w, dtype=torch.float32, device=device) * s
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        half = a / 2.0
        out.append(torch.stack((xx - half, yy - half, xx + half, yy + half), dim=-1).reshape(-1, 4).contiguous())
    return out


@dataclass
class HeadOutputs:
    anchors: List[torch.Tensor]          # L x (HWA, 4)
    deltas: List[torch.Tensor]           # L x (N, HWA, 4)
    centerness: List[torch.Tensor]       # L x (N, HWA)
    image_sizes: List[Tuple[int, int]]   # N x (h, w), un-padded
    grid_sizes: List[Tuple[int, int]] = field(default_factory=list)


def make_head_outputs(
    num_images: int,
    image_hw: Tuple[int, int] = (800, 1333),
    *,
    seed: int = 1234,
    device="cpu",
    ties: str = "free",
    mixed_sizes: bool = False,
    neg_frac: float = 0.10,
    nonfinite: int = 0,
) -> HeadOutputs:
    """ties='free': centerness = (randperm(HWA)+0.5)/HWA per image (spacing >> ulp, no duplicates);
    ties='heavy': rounded to 1/256 (many duplicates).  deltas: each side U(0.05,1.2), ``neg_frac`` of the
    entries negated (ReLU -> 0 -> exercises empty boxes).  ``nonfinite`` > 0 plants that many inf/NaN."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    grids = fpn_grid_sizes(*image_hw)
    anchors = make_anchors(grids, device=device)
    deltas, ctr = [], []
    for (h, w) in grids:
        n = h * w
        d = torch.rand(num_images, n, 4, generator=g) * 1.15 + 0.05
        neg = torch.rand(num_images, n, 4, generator=g) < neg_frac
        d = torch.where(neg, -d, d)
        c = torch.stack([(torch.randperm(n, generator=g).to(torch.float32) + 0.5) / n for _ in range(num_images)])
        if ties == "heavy":
            c = torch.round(c * 256.0) / 256.0
        deltas.append(d)
        ctr.append(c)
    if nonfinite:
        for j in range(nonfinite):
            lvl = j % len(grids)
            n = grids[lvl][0] * grids[lvl][1]
            pos = int(torch.randint(0, n, (1,), generator=g))
            img = j % num_images
            ctr[lvl][img, pos] = 0.999999  # make sure it is selected
            deltas[lvl][img, pos, j % 4] = float("inf") if j % 2 == 0 else float("nan")
    if mixed_sizes:
        choices = [image_hw, (image_hw[0], image_hw[1] - image_hw[1] // 10), (image_hw[0] - image_hw[0] // 25, image_hw[1])]
        sizes = [choices[i % 3] for i in range(num_images)]
    else:
        sizes = [image_hw] * num_images
    return HeadOutputs(
        anchors=anchors,
        deltas=[d.to(device) for d in deltas],
        centerness=[c.to(device) for c in ctr],
        image_sizes=sizes,
        grid_sizes=grids,
    )


def make_features(num_images: int, image_hw=(800, 1333), channels: int = 256, *, seed: int = 4321,
                  device="cpu", channels_last: bool = False, dtype=torch.float32) -> List[torch.Tensor]:
    """p2..p5 N(0,1) feature maps (N, C, H_l, W_l)."""
    grids = fpn_grid_sizes(*image_hw)[:4]
    g = torch.Generator(device=device).manual_seed(seed)
    feats = []
    for (h, w) in grids:
        f = torch.randn(num_images, channels, h, w, generator=g, device=device, dtype=dtype)
        if channels_last:
            f = f.contiguous(memory_format=torch.channels_last)
        feats.append(f)
    return feats


def make_rois(num_images: int, rois_per_image: int, image_hw=(800, 1333), *, seed: int = 99,
              device="cpu") -> List[torch.Tensor]:
    """Random xyxy boxes spread over all pooler levels (log-uniform sqrt-area 8..700 px, aspect 1/3..3)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    h, w = image_hw
    out = []
    for _ in range(num_images):
        s = torch.exp(torch.rand(rois_per_image, generator=g) * (math.log(700.0) - math.log(8.0)) + math.log(8.0))
        ar = torch.exp((torch.rand(rois_per_image, generator=g) * 2 - 1) * math.log(3.0))
        bw = (s * torch.sqrt(ar)).clamp(max=w)
        bh = (s / torch.sqrt(ar)).clamp(max=h)
        cx = torch.rand(rois_per_image, generator=g) * w
        cy = torch.rand(rois_per_image, generator=g) * h
        x1 = (cx - bw / 2).clamp(0, w); x2 = (cx + bw / 2).clamp(0, w)
        y1 = (cy - bh / 2).clamp(0, h); y2 = (cy + bh / 2).clamp(0, h)
        out.append(torch.stack((x1, y1, x2, y2), dim=1).to(device))
    return out


def make_gt(num_images: int, gt_per_image: int = 8, image_hw=(800, 1333), *, num_known: int = 20, seed: int = 5,
            device="cpu") -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Synthetic ground truth: ``gt_per_image`` boxes per image, side U(32,512), class
    U{0..K-1}.  Returns concatenated ``(boxes (G,4), classes (G) int64, offsets (N+1) int32)``."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    h, w = image_hw
    G = num_images * gt_per_image
    side = torch.rand(G, 2, generator=g) * (512.0 - 32.0) + 32.0
    ctr = torch.rand(G, 2, generator=g) * torch.tensor([float(w), float(h)])
    x1 = (ctr[:, 0] - side[:, 0] / 2).clamp(0, w); x2 = (ctr[:, 0] + side[:, 0] / 2).clamp(0, w)
    y1 = (ctr[:, 1] - side[:, 1] / 2).clamp(0, h); y2 = (ctr[:, 1] + side[:, 1] / 2).clamp(0, h)
    boxes = torch.stack((x1, y1, x2, y2), dim=1)
    classes = torch.randint(0, num_known, (G,), generator=g)
    off = torch.arange(0, G + 1, gt_per_image, dtype=torch.int32)
    return boxes.to(device), classes.to(device), off.to(device)


def make_matched_gt(kept_boxes, gt_per_image: int = 8, *, num_known: int = 20, seed: int = 5, candidates: int = 512):
    """Ground truth that the proposals actually match: per image, ``gt_per_image`` of its kept proposals - those (among
    ``candidates`` random ones) that the most other kept proposals overlap at IoU >= 0.5 - with classes U{0..K-1}.
    Random ground truth (``make_gt``) leaves the labelled sampler almost no positives; with this one the positive quota of
    the reference (25 % of 512) fills, so the sampled labels / IoUs give the prototype loss its nominal foreground share.
    ``kept_boxes``: list of (P_n, 4) tensors.  Returns ``(boxes (G,4), classes (G) int64, offsets (N+1) int32)``."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    out_b, out_c = [], []
    for b in kept_boxes:
        P = b.shape[0]
        cand = b[torch.randperm(P, generator=g)[:min(candidates, P)].to(b.device)]
        area_c = (cand[:, 2] - cand[:, 0]) * (cand[:, 3] - cand[:, 1])
        area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
        lt = torch.maximum(cand[:, None, :2], b[None, :, :2])
        rb = torch.minimum(cand[:, None, 2:], b[None, :, 2:])
        wh = (rb - lt).clamp(min=0)
        inter = wh[..., 0] * wh[..., 1]
        iou = inter / (area_c[:, None] + area_b[None, :] - inter).clamp(min=1e-9)
        hits = (iou >= 0.5).sum(dim=1)
        out_b.append(cand[hits.topk(min(gt_per_image, cand.shape[0])).indices])
        out_c.append(torch.randint(0, num_known, (out_b[-1].shape[0],), generator=g).to(b.device))
    off = torch.tensor([0] + [int(x.shape[0]) for x in out_b]).cumsum(0).to(torch.int32).to(kept_boxes[0].device)
    return torch.cat(out_b).contiguous(), torch.cat(out_c), off


@dataclass
class PLNInputs:
    roi_features: torch.Tensor   # (R, feat_dim)
    gt_classes: torch.Tensor     # (R,) int64 in [0, num_classes]  (num_classes = background)
    ious: torch.Tensor           # (R,) fp32
    enc_w: torch.Tensor
    enc_b: torch.Tensor
    dec_w: torch.Tensor
    dec_b: torch.Tensor
    reps: torch.Tensor


def make_pln_inputs(R: int, *, feat_dim: int = 1024, emb_dim: int = 256, num_known: int = 20,
                    num_classes: int = 81, seed: int = 7, device="cpu", fg_frac: float = 0.25) -> PLNInputs:
    """``roi_features = relu(N(0,1))``; encoder/decoder N(0, 0.01^2), bias 0; representatives N(0,1)
    (``prototype_learning_network.py:67-78``); ``fg_frac`` of the rows are known-class foreground with
    iou U(0.5,1), the rest background (label = num_classes) or unknown-class with iou U(0,0.5)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.relu(torch.randn(R, feat_dim, generator=g))
    fg = torch.rand(R, generator=g) < fg_frac
    cls_fg = torch.randint(0, num_known, (R,), generator=g)
    cls_other = torch.where(torch.rand(R, generator=g) < 0.5,
                            torch.full((R,), num_classes, dtype=torch.int64),
                            torch.randint(num_known, num_classes, (R,), generator=g))
    gt = torch.where(fg, cls_fg, cls_other)
    iou = torch.where(fg, torch.rand(R, generator=g) * 0.5 + 0.5, torch.rand(R, generator=g) * 0.5)
    # a few fg-class rows below the iou threshold (must be ignored) and exact-threshold rows (strict >)
    if R >= 8:
        gt[:4] = torch.arange(4) % num_known
        iou[:2] = 0.5
        iou[2:4] = 0.49
    return PLNInputs(
        roi_features=x.to(device), gt_classes=gt.to(device), ious=iou.to(device),
        enc_w=(torch.randn(emb_dim, feat_dim, generator=g) * 0.01).to(device),
        enc_b=torch.zeros(emb_dim, device=device),
        dec_w=(torch.randn(feat_dim, emb_dim, generator=g) * 0.01).to(device),
        dec_b=torch.zeros(feat_dim, device=device),
        reps=torch.randn(num_known, emb_dim, generator=g).to(device),
    )
