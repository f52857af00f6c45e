"""Algorithmic (compulsory) byte model of SURVEY.md section 8(d) - one definition shared by
``bench.py``'s roofline report and the tests.  The same formulas are duplicated (without the
oracle import) in ``osr_b200/roofline.py``; ``tests/test_bytes_model.py`` keeps them equal.
"""
from __future__ import annotations

from typing import Sequence, Tuple


def level_sizes(grid_sizes: Sequence[Tuple[int, int]], A: int = 1):
    return [h * w * A for (h, w) in grid_sizes]


def s1_bytes_per_image(grid_sizes, pre_nms_topk: int, A: int = 1, nominal_post_k: int = 0) -> int:
    """4*sum(HWA) scores read + sum(k)*16 selected deltas read + sum(k)*(16+4) boxes+scores written."""
    hwa = level_sizes(grid_sizes, A)
    ks = [min(n, pre_nms_topk) for n in hwa]
    b = 4 * sum(hwa) + sum(ks) * 16 + sum(ks) * 20
    if nominal_post_k:
        b += 8 * nominal_post_k
    return b


def s3_fwd_bytes(M: int, C: int, P: int, touched_px: int) -> int:
    """M*(20 + C*P*P*4) + C*4*U."""
    return M * (20 + C * P * P * 4) + C * 4 * touched_px


def s3_bwd_bytes(M: int, C: int, P: int, num_images: int, pooled_level_shapes) -> int:
    """grad_out read + dense grad written (zero-fill included)."""
    px = sum(h * w for (h, w) in pooled_level_shapes)
    return M * C * P * P * 4 + 4 * C * px * num_images


def s5_fwd_bytes(R: int, feat_dim: int, emb_dim: int, K: int, encoder_fused: bool = True) -> int:
    b = R * 12 + K * emb_dim * 4 + R * emb_dim * 4
    if encoder_fused:
        b += R * feat_dim * 4 + emb_dim * feat_dim * 4
    else:
        b += R * emb_dim * 4  # emb read instead of produced
    return b


def s5_bwd_bytes(R: int, emb_dim: int, K: int) -> int:
    return R * emb_dim * 4 * 2 + K * emb_dim * 4 * 2 + R * 12
