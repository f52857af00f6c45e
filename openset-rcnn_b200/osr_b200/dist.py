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
    """(R, ...) -> (W*R, ...) in rank order; R MUST be equal on every rank (see ``gather_row_counts`` /
    ``all_gather_rows_padded`` for the ragged case)."""
    world = dist.get_world_size(group)
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


def gather_row_counts(n: int, device, group=None) -> list:
    """Row count of every rank (one tiny all-gather + host read).  detectron2's ``subsample_labels`` returns fewer than
    ``batch_size_per_image`` rows when an image has too few negatives, so R can differ between ranks."""
    world = dist.get_world_size(group)
    mine = torch.tensor([int(n)], dtype=torch.int64, device=device)
    out = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, mine, group=group)
    return [int(v) for v in out.tolist()]


def all_gather_rows_padded(x: torch.Tensor, rmax: int, group=None, fill=0) -> torch.Tensor:
    """(R_r, ...) -> (W*rmax, ...): every rank's rows padded with ``fill`` to ``rmax`` rows, rank r at [r*rmax, ...)."""
    if x.shape[0] < rmax:
        pad = torch.full((rmax - x.shape[0],) + tuple(x.shape[1:]), fill, dtype=x.dtype, device=x.device)
        x = torch.cat((x, pad), dim=0)
    return all_gather_rows(x, group)


def _loss_on_global_rows(emb, emb_all_const, reps, labels_all, ious_all, rank, world, loss_fn, *, slot_rows=None,
                         global_rows=None, **kw):
    """Global-batch loss given the gathered rows: local rows keep their autograd edge, remote rows are constants
    (their gradient lives on their own rank).  ``slot_rows`` = rows per rank slot in the gathered layout (> R when the
    ranks' counts differ: the tail of a slot is padding with label -1 / iou 0, which the loss skips), ``global_rows`` =
    number of REAL rows over all ranks (the normaliser)."""
    R = emb.shape[0]
    slot = R if slot_rows is None else int(slot_rows)
    emb_all = emb_all_const.clone()
    emb_all[rank * slot:rank * slot + R] = emb
    n = world * R if global_rows is None else int(global_rows)
    return loss_fn(emb_all, reps, labels_all, ious_all, r_norm=float(max(n, 1)),
                   center_weight=float(world), emb_grad_scale=float(world), **kw)


def _gather_meta(labels, ious, group, rmax=None):
    meta = torch.stack((labels.to(torch.float32), ious.to(torch.float32)), dim=1)  # labels < 2^24: exact in fp32
    if rmax is not None and meta.shape[0] < rmax:   # padding rows: label -1 (never foreground), iou 0
        pad = torch.tensor([[-1.0, 0.0]], device=meta.device).expand(rmax - meta.shape[0], 2)
        meta = torch.cat((meta, pad), dim=0)
    meta_all = all_gather_rows(meta, group)
    return meta_all[:, 0].to(torch.int64), meta_all[:, 1].contiguous()


class FusedEncoderGather:
    """Encoder GEMM fused with the all-gather of its output (B200: tcgen05 GEMM whose epilogue stores every tile into
    all ranks' ``(W*R, E)`` buffers over NVLink - ``osr_pln_encode_gather_fwd``).  The buffers are torch symmetric
    memory (``torch.distributed._symmetric_memory``): allocated once, rendezvoused once, peer addresses handed to the
    kernel; with NVLS multicast support one ``multimem.st`` per 16 bytes is replicated by the NVSwitch.
    ``__call__(x, weight, bias)`` returns ``(emb_local, emb_all)``, views of this rank's buffer, valid until the
    call after next."""

    def __init__(self, rows_per_rank: int, emb_dim: int, device, group=None, multicast: Optional[bool] = None):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.R, self.E = int(rows_per_rank), int(emb_dim)
        # two buffer sets used alternately: a set is rewritten two calls later, and every rank has passed the barrier
        # of the call in between by then, so ONE barrier per call (after the stores) is enough
        self.bufs, self.hdls, self._ptrs, self._mc = [], [], [], []
        for _ in range(2):
            buf = symm_mem.empty(self.world * self.R, self.E, dtype=torch.float32, device=device)
            hdl = symm_mem.rendezvous(buf, self.group)
            self.bufs.append(buf)
            self.hdls.append(hdl)
            self._ptrs.append((ctypes.c_uint64 * self.world)(*[int(p) for p in hdl.buffer_ptrs]))
            use_mc = bool(hdl.has_multicast_support) if multicast is None else bool(multicast)
            self._mc.append(int(hdl.multicast_ptr) if use_mc and hdl.multicast_ptr else 0)
        self.multicast = self._mc[0] != 0
        self._ws = None
        self._flip = 0

    def __call__(self, x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None):
        import ctypes
        from . import _lib
        lib = _lib.lib()
        _lib.require_cuda(x, weight)
        xc = x.detach().contiguous().float()
        wc = weight.detach().contiguous().float()
        bc = None if bias is None else bias.detach().contiguous().float()
        R, Fd = xc.shape
        assert R == self.R and wc.shape[0] == self.E, (R, self.R, wc.shape, self.E)
        need = max(int(lib.osr_pln_encode_workspace(R, Fd, self.E)), 256)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=xc.device)
        k = self._flip
        self._flip ^= 1
        rc = lib.osr_pln_encode_gather_fwd(xc.data_ptr(), wc.data_ptr(), _lib.ptr(bc), R, Fd, self.E,
                                           ctypes.cast(self._ptrs[k], ctypes.c_void_p), self.world, self.rank,
                                           self._mc[k], self._ws.data_ptr(), self._ws.numel(),
                                           _lib.stream_ptr(xc.device))
        _lib.check(rc, "osr_pln_encode_gather_fwd")
        self.hdls[k].barrier(channel=0)   # all ranks' tiles have landed in this rank's buffer
        buf = self.bufs[k]
        return buf[self.rank * R:(self.rank + 1) * R], buf


def fused_gathered_pln_loss(enc: "FusedEncoderGather", x, weight, bias, reps, labels, ious, *, loss_fn=None, **kw):
    """Gathered PLN loss with the fused encoder -> all-gather kernel.  Returns ``(loss, emb_local)`` where
    ``emb_local`` is a leaf that receives d loss / d emb (the encoder's own backward GEMMs stay with the caller)."""
    if loss_fn is None:
        from .pln import pln_loss_from_emb as loss_fn
    with torch.no_grad():
        emb_loc, emb_all = enc(x, weight, bias)
        labels_all, ious_all = _gather_meta(labels, ious, enc.group)
    emb = emb_loc.clone().requires_grad_(True)
    loss = _loss_on_global_rows(emb, emb_all, reps, labels_all, ious_all, enc.rank, enc.world, loss_fn, **kw)
    return loss, emb


def gathered_pln_loss(emb: torch.Tensor, reps: torch.Tensor, labels: torch.Tensor, ious: torch.Tensor, *,
                      group=None, loss_fn: Optional[Callable] = None, rows_per_rank=None, **kw) -> torch.Tensor:
    """Global-batch PLN loss.  ``loss_fn(emb_all, reps, labels_all, ious_all, r_norm=, center_weight=,
    emb_grad_scale=, **kw)`` defaults to the CUDA op; tests pass a CPU implementation to exercise this
    host logic under gloo.

    ``rows_per_rank``: the row count of every rank if the caller knows it (a fixed-size sampler: pass
    ``[R] * world`` and no extra collective or host sync happens); ``None`` = exchange the counts first
    (``gather_row_counts``).  When the counts differ (``subsample_labels`` ran out of negatives on some rank) every rank
    pads to the largest count with label -1 / iou 0 rows, which the loss skips, and the normaliser is the number of
    REAL rows: the result is the loss of the global batch (every RoI weighs the same).  It equals the DDP mean of the
    reference's per-rank losses exactly when all counts are equal - with unequal counts that mean weighs the RoIs of a
    short rank more, which no single ``r_norm`` can express."""
    if loss_fn is None:
        from .pln import pln_loss_from_emb as loss_fn  # CUDA kernels
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    R = emb.shape[0]
    if world == 1:
        return loss_fn(emb, reps, labels, ious, r_norm=float(max(R, 1)), center_weight=1.0, emb_grad_scale=1.0, **kw)
    rank = dist.get_rank(group)
    counts = list(rows_per_rank) if rows_per_rank is not None else gather_row_counts(R, emb.device, group)
    if len(counts) != world or counts[rank] != R:
        raise ValueError(f"gathered_pln_loss: rows_per_rank {counts} does not describe this rank ({rank}: {R} rows, world {world})")
    rmax = max(counts)
    with torch.no_grad():
        if min(counts) == rmax:
            emb_all = all_gather_rows(emb.detach(), group)
            labels_all, ious_all = _gather_meta(labels, ious, group)
        else:
            emb_all = all_gather_rows_padded(emb.detach(), rmax, group)
            labels_all, ious_all = _gather_meta(labels, ious, group, rmax)
    return _loss_on_global_rows(emb, emb_all, reps, labels_all, ious_all, rank, world, loss_fn, slot_rows=rmax,
                                global_rows=sum(counts), **kw)


def check_uniform_requires_grad(t: torch.Tensor, group=None) -> None:
    """Raise on every rank if ``t.requires_grad`` differs between ranks (precondition of ``reduced_pln_loss``)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flag = torch.tensor([1.0 if t.requires_grad else 0.0], device=t.device)
    lo, hi = flag.clone(), flag.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    if float(lo) != float(hi):
        raise RuntimeError("reduced_pln_loss: representatives.requires_grad differs between ranks; the gradient "
                           "all-reduce in its hook would deadlock")


class _MeanOverRanks(torch.autograd.Function):
    """value: mean over ranks of the per-rank loss; backward: identity - the local loss keeps its full gradient, which is
    exactly the gathered rule for the local rows (global 1/W times the W scale that feeds a DDP-averaged encoder)."""

    @staticmethod
    def forward(ctx, x, group):
        y = x.detach().clone()
        dist.all_reduce(y, group=group)
        return y / dist.get_world_size(group)

    @staticmethod
    def backward(ctx, g):
        return g, None


def reduced_pln_loss(emb: torch.Tensor, reps: torch.Tensor, labels: torch.Tensor, ious: torch.Tensor, *,
                     group=None, loss_fn: Optional[Callable] = None, **kw) -> torch.Tensor:
    """The gathered PLN loss WITHOUT gathering: same value and the same gradients as ``gathered_pln_loss``, from the
    per-rank loss plus two tiny all-reduces.  Every row term of the loss depends only on (its embedding, the prototypes),
    so  (1/W) sum_r L_r  ==  w/(W R_loc) (sum A_r + sum B_r + W C)  is the global-batch loss; the local embedding
    gradient is the per-rank one, and ``representatives.grad`` is the rank mean of the per-rank gradients (all-reduced
    in a hook, so it is identical on all ranks as with the gathered formulation).  Cost per rank: the loss kernels on
    R_loc rows (not W R_loc) + all-reduce of 1 + K*D floats, instead of an all-gather of R_loc*D floats per peer.

    Requirements (collectives inside autograd): EVERY rank must call this with ``reps.requires_grad`` set the same way
    and must backpropagate through ``reps`` in the same step - the all-reduce of the prototype gradient runs in a tensor
    hook, so a rank that skips the loss, freezes ``reps`` or differentiates w.r.t. ``emb`` only would leave the others
    waiting (``check_uniform_requires_grad`` below is the cheap guard; call it once, not per step).  The row counts may
    differ between ranks: each rank normalises by its own count, i.e. this is the DDP mean of the per-rank losses.  If
    the module is ALSO wrapped in DDP the prototype gradient is averaged a second time, which leaves it unchanged."""
    if loss_fn is None:
        from .pln import pln_loss_from_emb as loss_fn  # CUDA kernels
    R = emb.shape[0]
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return loss_fn(emb, reps, labels, ious, r_norm=float(max(R, 1)), center_weight=1.0, emb_grad_scale=1.0, **kw)
    world = dist.get_world_size(group)
    reps_local = reps.view_as(reps)   # a non-leaf alias: its gradient hook averages over ranks before it reaches `reps`

    def _mean(g):
        g = g.contiguous().clone()
        dist.all_reduce(g, group=group)
        return g / world

    if reps_local.requires_grad:
        reps_local.register_hook(_mean)
    local = loss_fn(emb, reps_local, labels, ious, r_norm=float(max(R, 1)), center_weight=1.0, emb_grad_scale=1.0, **kw)
    return _MeanOverRanks.apply(local, group)
