"""Oracle for the Prototype Learning Network loss / inference.

Follows ``openset_rcnn/modeling/roi_heads/prototype_learning_network.py``:
``PLN.loss`` ``:117-187``, ``PLN.inference`` ``:189-230``, ``PLN.encode`` ``:232-234``;
parameters ``:67-78``.  The reference hard-codes ``device='cuda'`` (``:67,71``); this
restatement is device-agnostic and functional (weights are arguments).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as F


def pln_distance(new_features: torch.Tensor, representatives: torch.Tensor, distance_type: str) -> torch.Tensor:
    # prototype_learning_network.py:156-161
    if distance_type == "L1":
        return torch.cdist(new_features, representatives, p=1.0)
    if distance_type == "L2":
        return torch.cdist(new_features, representatives)
    if distance_type == "COS":
        return 1.0 - torch.mm(new_features, representatives.transpose(0, 1))
    raise ValueError(distance_type)


def pln_loss_from_emb(
    emb_features: torch.Tensor,
    representatives_param: torch.Tensor,
    gt_classes: torch.Tensor,
    ious: torch.Tensor,
    *,
    num_known_classes: int,
    reps_per_class: int = 1,
    alpha: float = 0.1,
    beta: float = 0.9,
    loss_weight: float = 0.5,
    iou_threshold: float = 0.5,
    distance_type: str = "COS",
    id_map: Optional[torch.Tensor] = None,
    r_norm: Optional[float] = None,
    center_weight: float = 1.0,
) -> torch.Tensor:
    """Lines 134,137-187 of prototype_learning_network.py (everything after the encoder).

    ``r_norm`` / ``center_weight`` are the two knobs of the gathered multi-GPU variant
    (SURVEY.md 5.8): defaults reproduce the reference exactly.
    """
    new_features = F.normalize(emb_features)
    representatives = F.normalize(representatives_param)
    if id_map is not None:
        gt_classes = id_map[gt_classes]
    fg_inds = torch.nonzero(
        (gt_classes >= 0) & (gt_classes < num_known_classes) & (ious > iou_threshold), as_tuple=True
    )[0]
    new_features = new_features[fg_inds]

    dist = pln_distance(new_features, representatives, distance_type)
    min_dist, _ = torch.min(dist.reshape(-1, num_known_classes, reps_per_class), dim=2)
    ar = torch.arange(min_dist.shape[0], device=min_dist.device)
    intra_dist = min_dist[ar, gt_classes[fg_inds]]
    min_dist = min_dist.clone()  # the reference mutates a non-leaf in place; clone keeps autograd identical
    min_dist[ar, gt_classes[fg_inds]] = 1000
    inter_dist, _ = torch.min(min_dist, dim=1)

    center_dist = pln_distance(representatives, representatives, distance_type)
    center_dist_clone = center_dist.clone()
    for i in range(num_known_classes):
        center_dist_clone[i * reps_per_class:(i + 1) * reps_per_class,
                          i * reps_per_class:(i + 1) * reps_per_class] = 1000
    c_dist, _ = torch.min(center_dist_clone, dim=1)

    dml_loss = (
        torch.sum(torch.max(intra_dist - alpha, torch.zeros_like(intra_dist)))
        + torch.sum(torch.max(beta - inter_dist, torch.zeros_like(inter_dist)))
        + center_weight * torch.sum(torch.max(beta + alpha - c_dist, torch.zeros_like(c_dist)))
    )
    denom = max(gt_classes.numel(), 1.0) if r_norm is None else max(float(r_norm), 1.0)
    return dml_loss * loss_weight / denom


def pln_loss(
    roi_features: torch.Tensor,
    enc_w: torch.Tensor, enc_b: torch.Tensor,
    dec_w: torch.Tensor, dec_b: torch.Tensor,
    representatives_param: torch.Tensor,
    gt_classes: torch.Tensor,
    ious: torch.Tensor,
    **kw,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """``PLN.loss``: returns (emb_features, rec_features, loss)."""
    emb_features = F.linear(roi_features, enc_w, enc_b)   # :133
    rec_features = F.linear(emb_features, dec_w, dec_b)   # :135
    loss = pln_loss_from_emb(emb_features, representatives_param, gt_classes, ious, **kw)
    return emb_features, rec_features, loss


def pln_inference(
    features: torch.Tensor,
    enc_w, enc_b, dec_w, dec_b, representatives_param,
    *, num_known_classes: int, reps_per_class: int = 1, unk_thr: float = 0.23,
    distance_type: str = "COS", unknown_id: int = 80, class_id: Optional[torch.Tensor] = None,
):
    """``PLN.inference`` for one image (``:203-226``): returns (rec_features, pred_classes)."""
    representatives = F.normalize(representatives_param)
    emb = F.linear(features, enc_w, enc_b)
    rec = F.linear(emb, dec_w, dec_b)
    new = F.normalize(emb)
    dist = pln_distance(new, representatives, distance_type)
    min_dist, _ = torch.min(dist.reshape(-1, num_known_classes, reps_per_class), dim=2)
    min_dist, min_index = torch.min(min_dist, dim=1)
    unknown = min_dist > unk_thr
    if class_id is not None:
        min_index = class_id[min_index]
    min_index = min_index.clone()
    min_index[unknown] = unknown_id
    return rec, min_index


def pln_loss_grad_closed_form(
    emb: torch.Tensor, reps: torch.Tensor, gt_classes: torch.Tensor, ious: torch.Tensor,
    *, num_known_classes: int, alpha: float, beta: float, loss_weight: float, iou_threshold: float,
    r_norm: Optional[float] = None, center_weight: float = 1.0,
):
    """Closed-form d loss / d emb and d loss / d reps for COS distance, reps_per_class=1
    (SURVEY.md A.9).  Mirrors what ``osr_pln_loss_bwd`` computes; pinned against autograd in tests."""
    K = num_known_classes
    R = emb.shape[0]
    s = loss_weight / (max(R, 1.0) if r_norm is None else max(float(r_norm), 1.0))
    en = emb.norm(dim=1).clamp_min(1e-12)
    eh = emb / en[:, None]
    rn = reps.norm(dim=1).clamp_min(1e-12)
    rh = reps / rn[:, None]
    fg = (gt_classes >= 0) & (gt_classes < K) & (ious > iou_threshold)
    S = eh @ rh.t()
    d = 1.0 - S
    G = torch.zeros_like(S)
    idx = torch.nonzero(fg, as_tuple=True)[0]
    y = gt_classes[idx]
    d_fg = d[idx]
    intra = d_fg[torch.arange(len(idx)), y]
    d_m = d_fg.clone(); d_m[torch.arange(len(idx)), y] = 1000
    inter, cstar = d_m.min(dim=1)
    G[idx, y] -= (intra > alpha).to(S.dtype)
    G[idx, cstar] += (inter < beta).to(S.dtype)
    g_eh = G @ rh
    C = 1.0 - rh @ rh.t()
    Cm = C.clone(); Cm[torch.arange(K), torch.arange(K)] = 1000
    cd, jstar = Cm.min(dim=1)
    Gam = torch.zeros_like(C)
    Gam[torch.arange(K), jstar] = center_weight * (cd < alpha + beta).to(C.dtype)
    g_rh = G.t() @ eh + (Gam + Gam.t()) @ rh
    g_e = s * (g_eh - eh * (eh * g_eh).sum(1, keepdim=True)) / en[:, None]
    g_r = s * (g_rh - rh * (rh * g_rh).sum(1, keepdim=True)) / rn[:, None]
    return g_e, g_r
