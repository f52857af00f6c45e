"""Drop-in for the reference's ``PLN`` module
(``openset_rcnn/modeling/roi_heads/prototype_learning_network.py:17-234``).

Same constructor arguments, parameter names and shapes (``encoder`` 1024->256, ``decoder`` 256->1024,
``representatives`` (K*reps, 256) - checkpoints stay loadable), same ``loss`` /
``inference`` / ``encode`` methods.  Everything after the encoder in ``loss`` (lines 134,137-187) is two
kernel launches (``osr_pln_loss_fwd``) with a closed-form backward (``osr_pln_loss_bwd``); no
``nonzero`` host sync, no K-iteration Python loop.  ``inference`` classifies all images in one launch.

Multi-GPU: ``gather=True`` all-gathers (embedding, label, iou) over the process group and
evaluates the loss on the global batch with ``r_norm = global R`` and the prototype-separation term weighted
by ``world_size`` - which equals the DDP mean of the reference's per-rank losses (tested with gloo).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
from torch import nn
from torch.nn import functional as F

from . import _lib
from .structures import Instances, cat_rows


DISTANCE_TYPES = {"COS": 0, "L1": 1, "L2": 2}   # include/osr.h OSR_PLN_DIST_*  (MODEL.PLN.DISTANCE_TYPE)


def _dist_code(distance_type) -> int:
    if isinstance(distance_type, int):
        return distance_type
    try:
        return DISTANCE_TYPES[distance_type]
    except KeyError:
        raise ValueError(f"MODEL.PLN.DISTANCE_TYPE must be one of {sorted(DISTANCE_TYPES)}, got {distance_type!r}") from None


def _pln_fwd(emb, reps, labels, ious, cfg):
    """osr_pln_loss_fwd: returns (loss scalar tensor, saved tuple for ``_pln_bwd``)."""
    (K, rpc, alpha, beta, loss_weight, iou_thr, r_norm, center_weight, emb_grad_scale, dist) = cfg
    lib = _lib.lib()
    _lib.require_cuda(emb, reps, labels, ious)
    emb_c = emb.contiguous().float()
    reps_c = reps.contiguous().float()
    labels_c = labels.contiguous().to(torch.int64)
    ious_c = ious.contiguous().float()
    R, D = emb_c.shape
    dev = emb_c.device
    Kr = K * rpc
    assert reps_c.shape == (Kr, D)
    terms = torch.empty(4, dtype=torch.float32, device=dev)
    emb_inv = torch.empty(R, dtype=torch.float32, device=dev)
    rep_inv = torch.empty(Kr, dtype=torch.float32, device=dev)
    intra = torch.empty(R, dtype=torch.int32, device=dev)
    inter = torch.empty(R, dtype=torch.int32, device=dev)
    center = torch.empty(Kr, dtype=torch.int32, device=dev)
    sdist = torch.empty(2 * R + Kr, dtype=torch.float32, device=dev)   # distances saved for the (L2) backward
    ws = torch.empty(max(int(lib.osr_pln_workspace(R, D, K, rpc)), 256), dtype=torch.uint8, device=dev)
    rn = float(R) if r_norm is None else float(r_norm)
    rc = lib.osr_pln_loss_fwd(emb_c.data_ptr(), reps_c.data_ptr(), labels_c.data_ptr(), ious_c.data_ptr(),
                              R, D, K, rpc, dist, alpha, beta, loss_weight, iou_thr, rn, center_weight,
                              terms.data_ptr(), emb_inv.data_ptr(), rep_inv.data_ptr(), intra.data_ptr(),
                              inter.data_ptr(), center.data_ptr(), sdist.data_ptr() if dist == 2 else None, ws.data_ptr(),
                              ws.numel(), _lib.stream_ptr(dev))   # (only the L2 backward reads the saved distances)
    _lib.check(rc, "osr_pln_loss_fwd")
    saved = (emb_c, reps_c, labels_c, emb_inv, rep_inv, intra, inter, center, sdist)
    return terms[0], saved, (K, rpc, loss_weight, rn, center_weight, emb_grad_scale, dist)


def _pln_bwd(saved, bcfg, grad_loss):
    """osr_pln_loss_bwd (closed form): returns (grad_emb, grad_reps)."""
    lib = _lib.lib()
    emb, reps, labels, emb_inv, rep_inv, intra, inter, center, sdist = saved
    K, rpc, loss_weight, rn, center_weight, emb_grad_scale, dist = bcfg
    R, D = emb.shape
    dev = emb.device
    gl = grad_loss.reshape(1).contiguous().float()
    grad_emb = torch.empty_like(emb)
    grad_reps = torch.empty_like(reps)
    ws = torch.empty(max(int(lib.osr_pln_workspace(R, D, K, rpc)), 256), dtype=torch.uint8, device=dev)
    rc = lib.osr_pln_loss_bwd(emb.data_ptr(), reps.data_ptr(), labels.data_ptr(), emb_inv.data_ptr(),
                              rep_inv.data_ptr(), intra.data_ptr(), inter.data_ptr(), center.data_ptr(),
                              sdist.data_ptr() if dist == 2 else None, gl.data_ptr(), R, D, K, rpc, dist, loss_weight, rn, center_weight,
                              grad_emb.data_ptr(), grad_reps.data_ptr(), ws.data_ptr(), ws.numel(),
                              _lib.stream_ptr(dev))
    _lib.check(rc, "osr_pln_loss_bwd")
    if emb_grad_scale != 1.0:
        grad_emb = grad_emb * emb_grad_scale
    return grad_emb, grad_reps


class _PlnLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, reps, labels, ious, cfg):
        loss, saved, bcfg = _pln_fwd(emb, reps, labels, ious, cfg)
        ctx.save_for_backward(*saved)
        ctx.cfg = bcfg
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        grad_emb, grad_reps = _pln_bwd(ctx.saved_tensors, ctx.cfg, grad_loss)
        return grad_emb, grad_reps, None, None, None


def pln_loss_fwd_bwd(emb, reps, labels, ious, *, num_known_classes: int, reps_per_class: int = 1, alpha: float = 0.1,
                     beta: float = 0.9, loss_weight: float = 0.5, iou_threshold: float = 0.5, r_norm: Optional[float] = None,
                     center_weight: float = 1.0, emb_grad_scale: float = 1.0, grad_loss: Optional[torch.Tensor] = None,
                     distance_type="COS", split: bool = False):
    """Loss and its closed-form gradients in one call, WITHOUT autograd: ``(loss, d loss / d emb, d loss / d reps)``.
    Same kernels as ``pln_loss_from_emb`` + ``backward()``; used by the device-resident training step (no autograd engine
    hop, capturable in a CUDA graph).

    ``split=True`` runs only the row launch now (``d loss / d emb`` is complete when it returns) and returns
    ``(finish, d loss / d emb)``: calling ``finish()`` - typically on a side stream, next to the rest of the backward pass -
    runs the prototype-gradient launches and the loss reduction and returns ``(loss, d loss / d reps)``
    (``osr_pln_loss_fwd_bwd_phase``; same buffers, bit-identical to the one-call form)."""
    lib = _lib.lib()
    _lib.require_cuda(emb, reps, labels, ious)
    with torch.no_grad():
        emb_c = emb.detach().contiguous().float()
        reps_c = reps.detach().contiguous().float()
        labels_c = labels.contiguous().to(torch.int64)
        ious_c = ious.contiguous().float()
        R, D = emb_c.shape
        K, rpc = int(num_known_classes), int(reps_per_class)
        Kr = K * rpc
        dev = emb_c.device
        assert reps_c.shape == (Kr, D)
        gl = _ones_scalar(dev) if grad_loss is None else grad_loss.reshape(1).contiguous().float()
        # one allocation for the small per-row / per-prototype outputs
        small = torch.empty(4 + R + Kr + 2 * R + Kr, dtype=torch.float32, device=dev)
        terms, emb_inv, rep_inv, sdist = small[:4], small[4:4 + R], small[4 + R:4 + R + Kr], small[4 + R + Kr:]
        idx = torch.empty(2 * R + Kr, dtype=torch.int32, device=dev)
        intra, inter, center = idx[:R], idx[R:2 * R], idx[2 * R:]
        grad_emb = torch.empty_like(emb_c)
        grad_reps = torch.empty_like(reps_c)
        ws = torch.empty(max(int(lib.osr_pln_workspace(R, D, K, rpc)), 256), dtype=torch.uint8, device=dev)
        rn = float(R) if r_norm is None else float(r_norm)
        args = (emb_c.data_ptr(), reps_c.data_ptr(), labels_c.data_ptr(), ious_c.data_ptr(), gl.data_ptr(),
                R, D, K, rpc, _dist_code(distance_type), float(alpha), float(beta), float(loss_weight),
                float(iou_threshold), rn, float(center_weight), terms.data_ptr(), emb_inv.data_ptr(),
                rep_inv.data_ptr(), intra.data_ptr(), inter.data_ptr(), center.data_ptr(),
                sdist.data_ptr() if _dist_code(distance_type) == 2 else None, grad_emb.data_ptr(),
                grad_reps.data_ptr(), ws.data_ptr(), ws.numel())
        if split:
            _lib.check(lib.osr_pln_loss_fwd_bwd_phase(1, *args, _lib.stream_ptr(dev)), "osr_pln_loss_fwd_bwd_phase(1)")
            keep = (emb_c, reps_c, labels_c, ious_c, gl, small, idx, ws)   # the buffers phase 2 reads

            def finish():
                _lib.check(lib.osr_pln_loss_fwd_bwd_phase(2, *args, _lib.stream_ptr(dev)), "osr_pln_loss_fwd_bwd_phase(2)")
                return terms[0], grad_reps
            finish.keep = keep
            if emb_grad_scale != 1.0:
                grad_emb = grad_emb * float(emb_grad_scale)
            return finish, grad_emb
        rc = lib.osr_pln_loss_fwd_bwd(*args, _lib.stream_ptr(dev))
        _lib.check(rc, "osr_pln_loss_fwd_bwd")
        if emb_grad_scale != 1.0:
            grad_emb = grad_emb * float(emb_grad_scale)
    return terms[0], grad_emb, grad_reps


_ONES = {}


def _ones_scalar(dev) -> torch.Tensor:
    """A cached device scalar 1.0 (the upstream gradient of a loss that is backpropagated directly)."""
    key = (dev.type, dev.index)
    if key not in _ONES:
        _ONES[key] = torch.ones(1, dtype=torch.float32, device=dev)
    return _ONES[key]


class _EncodeTcFn(torch.autograd.Function):
    """emb = x @ W^T + b on tcgen05 tensor cores (bf16 operands, fp32 accumulate).  Backward: the two ordinary GEMMs
    (grad_x = g @ W, grad_W = g^T @ x) are library calls, like the reference's nn.Linear backward."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        lib = _lib.lib()
        _lib.require_cuda(x, weight)
        xc = x.contiguous().float()
        wc = weight.contiguous().float()
        bc = None if bias is None else bias.contiguous().float()
        R, Fd = xc.shape
        E = wc.shape[0]
        emb = torch.empty((R, E), dtype=torch.float32, device=xc.device)
        ws = torch.empty(max(int(lib.osr_pln_encode_workspace(R, Fd, E)), 256), dtype=torch.uint8, device=xc.device)
        rc = lib.osr_pln_encode_fwd(xc.data_ptr(), wc.data_ptr(), _lib.ptr(bc), R, Fd, E, emb.data_ptr(), ws.data_ptr(),
                                    ws.numel(), _lib.stream_ptr(xc.device))
        _lib.check(rc, "osr_pln_encode_fwd")
        ctx.save_for_backward(xc, wc)
        ctx.has_bias = bias is not None
        return emb

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g = g.contiguous()
        gx = g @ w if ctx.needs_input_grad[0] else None
        gw = g.t() @ x if ctx.needs_input_grad[1] else None
        gb = g.sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return gx, gw, gb


def pln_encode_tc(x, weight, bias=None):
    """Tensor-core encoder (``PLN.encoder`` forward): bf16 x bf16 -> fp32."""
    return _EncodeTcFn.apply(x, weight, bias)


def pln_loss_from_emb(emb, reps, labels, ious, *, num_known_classes: int, reps_per_class: int = 1,
                      alpha: float = 0.1, beta: float = 0.9, loss_weight: float = 0.5, iou_threshold: float = 0.5,
                      r_norm: Optional[float] = None, center_weight: float = 1.0, emb_grad_scale: float = 1.0,
                      distance_type="COS"):
    """Functional form: labels already id-mapped ([0,K) known, anything else ignored)."""
    cfg = (int(num_known_classes), int(reps_per_class), float(alpha), float(beta), float(loss_weight),
           float(iou_threshold), r_norm, float(center_weight), float(emb_grad_scale), _dist_code(distance_type))
    return _PlnLossFn.apply(emb, reps, labels, ious, cfg)


def pln_nearest(emb, reps, *, num_known_classes: int, reps_per_class: int, unk_thr: float, unknown_id: int,
                class_id: Optional[torch.Tensor] = None, distance_type="COS"):
    lib = _lib.lib()
    _lib.require_cuda(emb, reps)
    emb_c = emb.contiguous().float()
    reps_c = reps.detach().contiguous().float()
    R, D = emb_c.shape
    pred = torch.empty(R, dtype=torch.int64, device=emb_c.device)
    md = torch.empty(R, dtype=torch.float32, device=emb_c.device)
    cid = None if class_id is None else class_id.contiguous().to(torch.int64)
    rc = lib.osr_pln_nearest(emb_c.data_ptr(), reps_c.data_ptr(), R, D, num_known_classes, reps_per_class,
                             _dist_code(distance_type), float(unk_thr), int(unknown_id), _lib.ptr(cid), pred.data_ptr(), md.data_ptr(),
                             _lib.stream_ptr(emb_c.device))
    _lib.check(rc, "osr_pln_nearest")
    return pred, md


class PLN(nn.Module):
    """Prototype Learning Network (same signature as the reference's ``PLN.__init__``, ``:22-37``)."""

    def __init__(self, num_classes: int, num_known_classes: int, feature_dim: int, embedding_dim: int,
                 distance_type: str, reps_per_class: int, alpha: float, beta: float, loss_weight: float,
                 dataset_name: str = "", iou_threshold: float = 0.5, unk_thr: float = 0.23,
                 opendet_benchmark: bool = True, known_class_ids: Optional[Sequence[int]] = None,
                 device="cuda", gather: bool = False, process_group=None, encoder_impl: str = "fp32"):
        super().__init__()
        _dist_code(distance_type)   # 'COS' | 'L1' | 'L2' (:156-161); anything else is a config error
        self.num_classes = num_classes
        self.num_known_classes = num_known_classes
        self.feature_dim = feature_dim
        self.embedding_dim = embedding_dim
        self.distance_type = distance_type
        self.reps_per_class = reps_per_class
        self.alpha = alpha
        self.beta = beta
        self.loss_weight = loss_weight
        self.unk_thr = unk_thr
        self.opendet_benchmark = opendet_benchmark
        self.iou_threshold = iou_threshold
        self.gather = gather
        self.process_group = process_group
        # "fp32": nn.Linear exactly as the reference; "tcgen05": bf16 tensor-core GEMM (fp32 accumulate), ~3e-3 relative
        assert encoder_impl in ("fp32", "tcgen05")
        self.encoder_impl = encoder_impl

        self.encoder = nn.Linear(feature_dim, embedding_dim, device=device)
        nn.init.normal_(self.encoder.weight, std=0.01)
        nn.init.constant_(self.encoder.bias, 0)
        self.decoder = nn.Linear(embedding_dim, feature_dim, device=device)
        nn.init.normal_(self.decoder.weight, std=0.01)
        nn.init.constant_(self.decoder.bias, 0)
        self.representatives = nn.parameter.Parameter(
            torch.zeros(num_known_classes * reps_per_class, embedding_dim, device=device))
        nn.init.normal_(self.representatives)

        if not opendet_benchmark:
            # reference :80-95 builds these from MetadataCatalog + GRASPNET_KNOWN_IDS; here the contiguous ids of the
            # known classes are passed in directly (detectron2's catalog is not a dependency of this package)
            assert known_class_ids is not None and len(known_class_ids) == num_known_classes
            class_id = torch.sort(torch.tensor(list(known_class_ids), device=device))[0]
            id_map = torch.zeros(num_classes + 1, device=device) - 1
            for i, v in enumerate(class_id.tolist()):
                id_map[v] = i
            id_map[num_classes] = num_known_classes
            self.register_buffer("class_id", class_id.long(), persistent=False)
            self.register_buffer("id_map", id_map.long(), persistent=False)

    @classmethod
    def from_config(cls, cfg):
        """Mirror of ``PLN.from_config`` (``:99-115``) for a yacs-like cfg object."""
        return {
            "num_classes": cfg.MODEL.ROI_HEADS.NUM_CLASSES,
            "num_known_classes": cfg.MODEL.ROI_HEADS.NUM_KNOWN_CLASSES,
            "feature_dim": cfg.MODEL.ROI_BOX_HEAD.FC_DIM,
            "embedding_dim": cfg.MODEL.PLN.EMD_DIM,
            "distance_type": cfg.MODEL.PLN.DISTANCE_TYPE,
            "reps_per_class": cfg.MODEL.PLN.REPS_PER_CLASS,
            "alpha": cfg.MODEL.PLN.ALPHA,
            "beta": cfg.MODEL.PLN.BETA,
            "loss_weight": cfg.MODEL.PLN.LOSS_WEIGHT,
            "dataset_name": cfg.DATASETS.TRAIN[0],
            "iou_threshold": cfg.MODEL.PLN.IOU_THRESHOLD,
            "unk_thr": cfg.MODEL.PLN.UNK_THR,
            "opendet_benchmark": cfg.OPENDET_BENCHMARK,
        }

    # -- training ------------------------------------------------------------------------------------------
    def loss_from_tensors(self, roi_features: torch.Tensor, gt_classes: torch.Tensor, ious: torch.Tensor):
        if self.encoder_impl == "tcgen05":
            emb_features = pln_encode_tc(roi_features, self.encoder.weight, self.encoder.bias)   # :133 on tensor cores
        else:
            emb_features = self.encoder(roi_features)        # :133
        rec_features = self.decoder(emb_features)            # :135
        if not self.opendet_benchmark:
            gt_classes = self.id_map[gt_classes]             # :146-147
        R_local = gt_classes.numel()
        kw = dict(num_known_classes=self.num_known_classes, reps_per_class=self.reps_per_class, alpha=self.alpha,
                  beta=self.beta, loss_weight=self.loss_weight, iou_threshold=self.iou_threshold,
                  distance_type=self.distance_type)
        if self.gather:
            from .dist import gathered_pln_loss
            loss = gathered_pln_loss(emb_features, self.representatives, gt_classes, ious,
                                     group=self.process_group, **kw)
        else:
            loss = pln_loss_from_emb(emb_features, self.representatives, gt_classes, ious,
                                     r_norm=float(max(R_local, 1)), **kw)
        return emb_features, rec_features, loss

    def loss(self, roi_features: torch.Tensor, proposals: List[Instances]):
        """``PLN.loss(roi_features, proposals) -> (emb_features, rec_features, loss)`` (``:117-187``)."""
        dev = roi_features.device
        ious = torch.cat([p.ious for p in proposals], dim=0) if len(proposals) else torch.empty(0, device=dev)
        gt_classes = (torch.cat([p.gt_classes for p in proposals], dim=0) if len(proposals)
                      else torch.empty(0, dtype=torch.int64, device=dev))
        return self.loss_from_tensors(roi_features, gt_classes, ious)

    # -- inference -----------------------------------------------------------------------------------------
    def inference(self, fg_instances: List[Instances]):
        """``PLN.inference`` (``:189-230``): adds ``pred_classes`` (unknown = 80 / 1000) and replaces ``features`` by
        the reconstruction.  All images are classified in one launch."""
        if len(fg_instances) == 0:
            return []
        sizes = [len(x) for x in fg_instances]
        feats = cat_rows([x.features for x in fg_instances])   # no copy when the fields are split views of one tensor
        emb = self.encoder(feats)
        rec = self.decoder(emb)
        unknown_id = 80 if self.opendet_benchmark else 1000
        pred, _ = pln_nearest(emb, self.representatives, num_known_classes=self.num_known_classes,
                              reps_per_class=self.reps_per_class, unk_thr=self.unk_thr, unknown_id=unknown_id,
                              class_id=None if self.opendet_benchmark else self.class_id,
                              distance_type=self.distance_type)
        results = []
        rec_l, pred_l = rec.split(sizes), pred.split(sizes)
        for n, inst in enumerate(fg_instances):
            inst._fields["features"] = rec_l[n]        # same lengths by construction: the per-field assert of set() skipped
            inst._fields["pred_classes"] = pred_l[n]
            results.append(inst)
        return results

    def encode(self, roi_features):
        return F.normalize(self.encoder(roi_features))       # :232-234
