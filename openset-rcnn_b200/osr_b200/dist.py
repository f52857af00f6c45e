"""Multi-GPU plumbing: one process per GPU (``torchrun``), images sharded across ranks (data parallel, like
``train.py:287-294`` + DDP at ``train.py:201-205``).  CF-RPN and ROIAlign need no exchange.  The only
collective on the path is the all-gather of PLN embeddings (+ labels, ious) that the north star adds
(the reference's PLN loss is per-rank).

Parity rule.  With W ranks and R_loc RoIs each, the reference optimises the DDP mean
``(1/W) sum_r L_r`` with ``L_r = w/R_loc * (A_r + B_r + C)``.  The gathered loss
``w/(W R_loc) * (sum A_r + sum B_r + W*C)`` is the same number; its gradient w.r.t. the local embeddings is
``1/W`` of the reference's per-rank gradient, so the local rows' gradient is scaled by W before it enters the
(DDP-averaged) encoder; ``representatives.grad`` is already identical on all ranks.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> tuple:
    """torchrun env -> (rank, local_rank, world_size); initialises the default group if WORLD_SIZE > 1."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


def shard_range(total: int, rank: int, world: int) -> range:
    """Contiguous, near-even split of ``total`` units (images) over ranks."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def all_gather_rows(x: torch.Tensor, group=None) -> torch.Tensor:
    """(R, ...) -> (W*R, ...) in rank order; equal R on every rank (the sampled-RoI count is fixed)."""
    world = dist.get_world_size(group)
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def gathered_pln_loss(emb: torch.Tensor, reps: torch.Tensor, labels: torch.Tensor, ious: torch.Tensor, *,
                      group=None, loss_fn: Optional[Callable] = None, **kw) -> torch.Tensor:
    """Global-batch PLN loss.  ``loss_fn(emb_all, reps, labels_all, ious_all, r_norm=, center_weight=,
    emb_grad_scale=, **kw)`` defaults to the CUDA op; tests pass a CPU implementation to exercise this
    host logic under gloo."""
    if loss_fn is None:
        from .pln import pln_loss_from_emb as loss_fn  # CUDA kernels
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    R = emb.shape[0]
    if world == 1:
        return loss_fn(emb, reps, labels, ious, r_norm=float(max(R, 1)), center_weight=1.0, emb_grad_scale=1.0, **kw)
    rank = dist.get_rank(group)
    with torch.no_grad():
        emb_all = all_gather_rows(emb.detach(), group)
        meta = torch.stack((labels.to(torch.float32), ious.to(torch.float32)), dim=1)  # labels < 2^24: exact in fp32
        meta_all = all_gather_rows(meta, group)
    labels_all = meta_all[:, 0].to(torch.int64)
    ious_all = meta_all[:, 1].contiguous()
    # local rows keep their autograd edge; remote rows are constants (their gradient lives on their own rank)
    emb_all = emb_all.clone()
    emb_all[rank * R:(rank + 1) * R] = emb
    return loss_fn(emb_all, reps, labels_all, ious_all, r_norm=float(max(world * R, 1)),
                   center_weight=float(world), emb_grad_scale=float(world), **kw)
