"""One pass of the RoI hot path over one batch - the "step" that ``bench.py`` times and ``smoke()`` runs.

Training step (BASELINE.json configs[1]):
  S1  CF-RPN proposal stage        osr_rpn_select_decode        (classification_free_rpn.py:545 -> :558-610)
  S2  RoI sampling glue            fixed pre-drawn indices      (osrcnn_roi_heads.py:136-230 is a section 8(f) "next" row;
                                                                 the timed step gathers 512 proposals/img with torch)
  S3  ROIPooler forward            osr_roi_align_fwd            (osrcnn_roi_heads.py:306)
  S4  box head FC                  NOT on the path (library GEMM, osrcnn_roi_heads.py:308): a fixed (R,1024) tensor
                                   stands in for its output and a fixed (M,C,7,7) tensor for its input gradient
  S5  PLN loss forward + backward  encoder GEMM (osr_pln_encode_fwd, tcgen05 bf16) + osr_pln_loss_fwd / _bwd (osrcnn_roi_heads.py:315)
  S3' ROIPooler backward           osr_roi_align_bwd            (train.py:145)
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

from . import _lib, synth
from .poolers import ROIPooler
from .pln import pln_encode_tc, pln_loss_from_emb
from .proposals import rpn_select_decode
from .sampling import match_proposals
from .dist import FusedEncoderGather, fused_gathered_pln_loss, gathered_pln_loss


@dataclass
class PathConfig:
    num_images: int = 16
    image_hw: Tuple[int, int] = (800, 1333)
    pre_nms_topk: int = 2000
    rois_per_image: int = 512
    channels: int = 256
    feat_dim: int = 1024
    emb_dim: int = 256
    num_known: int = 20
    num_classes: int = 81
    alpha: float = 0.1
    beta: float = 0.9
    loss_weight: float = 0.5
    iou_threshold: float = 0.5
    channels_last: bool = True    # FPN maps in channels_last (NHWC) memory format; False = NCHW as the reference
    encoder_impl: str = "tcgen05"   # PLN encoder: bf16 tensor cores (fp32 accumulate) | "fp32" = nn.Linear as the reference
    seed: int = 1234


class RoiPathStep:
    """Device-resident synthetic inputs + the timed step.  ``stage_events=True`` records CUDA events between
    stages so bench.py can attribute time to S1 / S3 / S5 / S3' without extra synchronisation."""

    STAGES = ("s1_proposals", "s2_sample_glue", "s3_roialign_fwd", "s5_pln_fwd_bwd", "s3_roialign_bwd")

    def __init__(self, cfg: PathConfig, device="cuda:0", host_inputs: bool = False):
        self.cfg = cfg
        self.device = torch.device(device)
        dev = self.device
        N = cfg.num_images
        ho = synth.make_head_outputs(N, cfg.image_hw, seed=cfg.seed)
        self.grid_sizes = ho.grid_sizes
        self.image_sizes = ho.image_sizes
        self.anchors = [a.to(dev) for a in ho.anchors]
        self.image_hw_dev = torch.tensor([[h, w] for (h, w) in ho.image_sizes], dtype=torch.int32, device=dev)
        feats_seed = cfg.seed + 1
        self.host_inputs = host_inputs
        if host_inputs:
            # pinned host copies of the step's inputs; two device buffer sets for copy/compute overlap
            self.h_deltas = [d.pin_memory() for d in ho.deltas]
            self.h_ctr = [c.pin_memory() for c in ho.centerness]
            self.h_feats = [f.pin_memory() for f in synth.make_features(N, cfg.image_hw, cfg.channels, seed=feats_seed,
                                                                        channels_last=cfg.channels_last)]
            self.dev_sets = []
            for _ in range(2):
                self.dev_sets.append(dict(
                    deltas=[torch.empty_like(d, device=dev) for d in self.h_deltas],
                    ctr=[torch.empty_like(c, device=dev) for c in self.h_ctr],
                    feats=[torch.empty_like(f, device=dev) for f in self.h_feats]))
            self.copy_stream = torch.cuda.Stream(device=dev)
            self.copy_done = [torch.cuda.Event() for _ in range(2)]
            self.compute_done = [torch.cuda.Event() for _ in range(2)]
            self.h_result = torch.empty(1 + N, dtype=torch.float32).pin_memory()
            self.deltas = self.dev_sets[0]["deltas"]; self.ctr = self.dev_sets[0]["ctr"]; self.feats = self.dev_sets[0]["feats"]
            for a, b in zip(self.deltas + self.ctr + self.feats, self.h_deltas + self.h_ctr + self.h_feats):
                a.copy_(b)
        else:
            self.deltas = [d.to(dev) for d in ho.deltas]
            self.ctr = [c.to(dev) for c in ho.centerness]
            self.feats = synth.make_features(N, cfg.image_hw, cfg.channels, seed=feats_seed, device=dev,
                                             channels_last=cfg.channels_last)
        R = N * cfg.rois_per_image
        pi = synth.make_pln_inputs(R, feat_dim=cfg.feat_dim, emb_dim=cfg.emb_dim, num_known=cfg.num_known,
                                   num_classes=cfg.num_classes, seed=cfg.seed + 2, device=dev)
        self.pln = pi
        g = torch.Generator(device=dev).manual_seed(cfg.seed + 3)
        self.grad_pooled = torch.randn(R, cfg.channels, 7, 7, device=dev, generator=g)
        self.pooler = ROIPooler(7, synth.POOL_SCALES, 0, "ROIAlignV2")
        self.roi_offsets = torch.arange(0, R + 1, cfg.rois_per_image, dtype=torch.int32, device=dev)
        self.img_col = torch.arange(N, dtype=torch.float32, device=dev).repeat_interleave(cfg.rois_per_image)[:, None]
        # S2 stand-in: pre-drawn sample positions inside each image's kept proposals (dry run gives the counts)
        sel = rpn_select_decode(self.anchors, self.deltas, self.ctr, self.image_hw_dev, cfg.pre_nms_topk)
        counts = sel.counts.cpu()
        L = sel.num_levels
        gs = torch.Generator().manual_seed(cfg.seed + 4)
        idx = []
        for n in range(N):
            c = int(counts[n, L])
            assert c >= cfg.rois_per_image, f"image {n}: only {c} proposals"
            idx.append(torch.randperm(c, generator=gs)[:cfg.rois_per_image] + n * sel.kmax)
        self.sample_idx = torch.cat(idx).to(dev)
        self.kmax = sel.kmax
        # S2 matching inputs: synthetic ground truth + the static offsets of the padded proposal layout
        self.gt_boxes, self.gt_classes, self.gt_off = synth.make_gt(N, 8, cfg.image_hw, num_known=cfg.num_known,
                                                                   seed=cfg.seed + 5, device=dev)
        self.prop_off = torch.arange(0, (N + 1) * sel.kmax, sel.kmax, dtype=torch.int32, device=dev)
        self.count_col = sel.num_levels
        self.last: Dict[str, torch.Tensor] = {}
        self._fused_enc = None
        self.fused_gather_error = None
        self.events: Optional[List[torch.cuda.Event]] = None

    # ------------------------------------------------------------------------------------------------
    def _mark(self, i):
        if self.events is not None:
            self.events[i].record()

    def step(self, deltas=None, ctr=None, feats=None, stage_events: bool = False, gather_pln: bool = False):
        cfg = self.cfg
        deltas = self.deltas if deltas is None else deltas
        ctr = self.ctr if ctr is None else ctr
        feats = self.feats if feats is None else feats
        self.events = [torch.cuda.Event(enable_timing=True) for _ in range(6)] if stage_events else None
        self._mark(0)
        # S1
        sel = rpn_select_decode(self.anchors, deltas, ctr, self.image_hw_dev, cfg.pre_nms_topk)
        self._mark(1)
        # S2 (glue): proposal <-> GT matching of ALL kept proposals (osr_match_label, as label_and_sample_proposals
        # does), then the sampling stand-in: pre-drawn indices (the reference's randperm needs a host sync per image)
        L2 = sel.counts.shape[1]
        match = match_proposals(sel.boxes.view(-1, 4), self.prop_off, self.gt_boxes, self.gt_classes, self.gt_off,
                                self.kmax, iou_threshold=cfg.iou_threshold, background_label=cfg.num_classes,
                                box_counts=sel.counts[:, self.count_col], box_counts_stride=L2)
        boxes = sel.boxes.view(-1, 4).index_select(0, self.sample_idx)
        rois = torch.cat((self.img_col, boxes), dim=1)
        self._mark(2)
        # S3 forward
        feats_g = [f.requires_grad_(True) for f in feats]
        pooled, lvl = self.pooler.pool_rois(feats_g, rois, self.roi_offsets)
        self._mark(3)
        # S5: encoder (nn.Linear) + prototype loss forward + backward to (emb, representatives)
        pi = self.pln
        reps = pi.reps.requires_grad_(True)
        kw = dict(num_known_classes=cfg.num_known, alpha=cfg.alpha, beta=cfg.beta, loss_weight=cfg.loss_weight,
                  iou_threshold=cfg.iou_threshold)
        if gather_pln and cfg.encoder_impl == "tcgen05":
            # north-star variant, B200-native: the encoder GEMM's epilogue stores its tiles into every rank's buffer
            # over NVLink (fused all-gather), then the global-batch loss (dist.py)
            if self._fused_enc is None:
                try:
                    self._fused_enc = FusedEncoderGather(pi.roi_features.shape[0], cfg.emb_dim, self.device)
                except Exception as e:  # noqa: BLE001 - no symmetric memory / P2P on this box: NCCL all-gather instead
                    self._fused_enc = False
                    self.fused_gather_error = repr(e)
        if gather_pln and cfg.encoder_impl == "tcgen05" and self._fused_enc:
            loss, emb = fused_gathered_pln_loss(self._fused_enc, pi.roi_features, pi.enc_w, pi.enc_b, reps,
                                                pi.gt_classes, pi.ious, **kw)
        else:
            if cfg.encoder_impl == "tcgen05":
                emb = pln_encode_tc(pi.roi_features, pi.enc_w, pi.enc_b).requires_grad_(True)
            else:
                emb = F.linear(pi.roi_features, pi.enc_w, pi.enc_b).requires_grad_(True)
            if gather_pln:   # encoder, then NCCL all-gather of (emb, label, iou), global-batch loss
                loss = gathered_pln_loss(emb, reps, pi.gt_classes, pi.ious, **kw)
            else:            # the reference's semantics: per-rank loss
                loss = pln_loss_from_emb(emb, reps, pi.gt_classes, pi.ious, **kw)
        g_emb, g_reps = torch.autograd.grad(loss, [emb, reps])
        self._mark(4)
        # S3 backward
        g_feats = torch.autograd.grad(pooled, feats_g, self.grad_pooled)
        self._mark(5)
        self.last = dict(sel=sel, match=match, rois=rois, pooled=pooled, level=lvl, loss=loss, g_emb=g_emb, g_reps=g_reps,
                         g_feats=g_feats)
        return loss, sel.counts

    def stage_ms(self) -> Dict[str, float]:
        ev = self.events
        return {name: ev[i].elapsed_time(ev[i + 1]) for i, name in enumerate(self.STAGES)}

    # ------------------------------------------------------------------------------------------------
    def e2e_prefetch(self, slot: int):
        """Start the H2D copy of one step's inputs (pinned host -> device set ``slot``) on the copy stream."""
        ds = self.dev_sets[slot]
        with torch.no_grad(), torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.compute_done[slot])
            for a, b in zip(ds["deltas"] + ds["ctr"] + ds["feats"], self.h_deltas + self.h_ctr + self.h_feats):
                a.copy_(b, non_blocking=True)
            self.copy_done[slot].record(self.copy_stream)

    def e2e_step(self, slot: int):
        """Compute on device set ``slot`` once its copy has landed; read the result (loss + per-image proposal
        counts) back to pinned host memory."""
        ds = self.dev_sets[slot]
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.copy_done[slot])
        loss, counts = self.step(ds["deltas"], ds["ctr"], ds["feats"])
        L = self.last["sel"].num_levels
        res = torch.cat((loss.reshape(1), counts[:, L].to(torch.float32)))
        self.h_result.copy_(res, non_blocking=True)
        self.compute_done[slot].record(cur)

    def h2d_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.h_deltas + self.h_ctr + self.h_feats)

    def d2h_bytes(self) -> int:
        return self.h_result.numel() * 4
