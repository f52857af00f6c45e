"""Algorithmic (compulsory) byte model of the RoI path - SURVEY.md section 8(d).  ``bench.py`` divides these by the
CUDA-event duration of the matching kernels to get ``roofline.achieved``.  ``oracle/bytes_model.py`` holds the
same formulas for the tests (``tests/test_bytes_model.py`` keeps the two equal); this copy exists because the
product path must not import ``oracle/``.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch


def s1_bytes_per_image(grid_sizes: Sequence[Tuple[int, int]], pre_nms_topk: int, A: int = 1, nominal_post_k: int = 0) -> int:
    """4*sum(HWA) scores read + sum(k)*16 selected deltas read + sum(k)*(16+4) boxes+scores written."""
    hwa = [h * w * A for (h, w) in grid_sizes]
    ks = [min(n, pre_nms_topk) for n in hwa]
    b = 4 * sum(hwa) + sum(ks) * 16 + sum(ks) * 20
    if nominal_post_k:
        b += 8 * nominal_post_k
    return b


def s3_fwd_bytes(M: int, C: int, P: int, touched_px: int) -> int:
    """M*(20 + C*P*P*4) output + rois, plus C*4*U distinct feature pixels read."""
    return M * (20 + C * P * P * 4) + C * 4 * touched_px


def s3_bwd_bytes(M: int, C: int, P: int, num_images: int, pooled_level_shapes: Sequence[Tuple[int, int]]) -> int:
    """grad_out read + dense gradient written (zero fill included)."""
    px = sum(h * w for (h, w) in pooled_level_shapes)
    return M * C * P * P * 4 + 4 * C * px * num_images


def s5_fwd_bytes(R: int, feat_dim: int, emb_dim: int, K: int, encoder_fused: bool = True) -> int:
    b = R * 12 + K * emb_dim * 4 + R * emb_dim * 4
    if encoder_fused:
        b += R * feat_dim * 4 + emb_dim * feat_dim * 4
    else:
        b += R * emb_dim * 4
    return b


def s5_bwd_bytes(R: int, emb_dim: int, K: int) -> int:
    return R * emb_dim * 4 * 2 + K * emb_dim * 4 * 2 + R * 12


def touched_pixels(level_shapes: Sequence[Tuple[int, int]], scales: Sequence[float], rois: torch.Tensor,
                   levels: torch.Tensor, num_images: int, P: int = 7) -> int:
    """U: number of distinct (image, level, y, x) feature pixels inside the footprint rectangle of any RoI
    (rows/cols reached by its first..last bilinear sample; adaptive grid = ceil(roi/P)).  Vectorised torch
    restatement of the kernel's geometry, evaluated wherever ``rois`` lives (2-D difference array + cumsum)."""
    dev = rois.device
    total = 0
    img = rois[:, 0].long()
    for l, ((H, W), s) in enumerate(zip(level_shapes, scales)):
        m = levels.long() == l
        if not bool(m.any()):
            continue
        r = rois[m]
        n = img[m]
        lohi = []
        for (a, b, L) in ((r[:, 1], r[:, 3], W), (r[:, 2], r[:, 4], H)):
            start = a * s - 0.5
            size = (b * s - 0.5) - start
            binsz = size / P
            grid = torch.ceil(size / P).clamp(min=0)
            first = start + 0.5 * binsz / grid.clamp(min=1)
            last = start + (P - 1) * binsz + (grid - 0.5) * binsz / grid.clamp(min=1)
            empty = (grid <= 0) | (last < -1.0) | (first > L)
            lo = first.clamp(min=0).floor().clamp(max=L - 1).long()
            hi = (last.clamp(min=0, max=float(L)).floor().long() + 1).clamp(max=L - 1)
            lohi.append((lo, hi, empty))
        (x0, x1, ex), (y0, y1, ey) = lohi
        ok = ~(ex | ey)
        x0, x1, y0, y1, n = x0[ok], x1[ok], y0[ok], y1[ok], n[ok]
        diff = torch.zeros((num_images, H + 1, W + 1), dtype=torch.int32, device=dev)
        one = torch.ones_like(n, dtype=torch.int32)
        diff.index_put_((n, y0, x0), one, accumulate=True)
        diff.index_put_((n, y0, x1 + 1), -one, accumulate=True)
        diff.index_put_((n, y1 + 1, x0), -one, accumulate=True)
        diff.index_put_((n, y1 + 1, x1 + 1), one, accumulate=True)
        cover = diff.cumsum(1).cumsum(2)[:, :H, :W]
        total += int((cover > 0).sum())
    return total
