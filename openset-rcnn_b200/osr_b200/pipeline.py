"""One pass of the RoI hot path over one batch - the "step" that ``bench.py`` times and ``smoke()`` runs.

Training step (BASELINE.json configs[1]):
  S1  CF-RPN proposal stage        osr_rpn_select_decode        (classification_free_rpn.py:545 -> :558-610)
  S2  RoI sampling glue            osr_match_label + osr_sample_rois   (osrcnn_roi_heads.py:136-230: matcher labels of ALL kept
                                                                 proposals, then the labelled 512-per-image sample of the whole
                                                                 batch in one launch, fresh random keys every step)
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
from .pln import pln_encode_tc, pln_loss_from_emb, pln_loss_fwd_bwd
from .proposals import rpn_select_decode
from .sampling import match_proposals, sample_rois
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
    # inference-only path (BASELINE.json configs[3]) / nominal proposal mode
    post_nms_topk: int = 1000
    rpn_nms_thresh: float = 0.7
    unk_thr: float = 0.23
    gt_per_image: int = 8
    name: str = "cfg2"
    overlap_bwd_prep: bool = True  # the backward's RoI-only table kernel runs on a side stream next to S3 forward / S5
    box_head: bool = False        # S4 on the path: ROIAlign writes bf16, fc1 / fc2 on tcgen05 feed the PLN


def make_config(name: str = "cfg2", world: int = 1, **over) -> PathConfig:
    """BASELINE.json ``configs`` 2-5 (SURVEY.md section 8(d) "Config -> concrete sizes"); per-GPU sizes."""
    if name == "cfg2":    # R50-FPN VOC-COCO training step, 16 images / GPU, k = 2000, 512 RoIs / image, K = 20
        cfg = PathConfig(name=name)
    elif name == "cfg3":  # GraspNet: 720x1280 -> 750x1333, 8 images / GPU, K = 28, cross-image PLN all-gather at N > 1
        cfg = PathConfig(name=name, num_images=8, image_hw=(750, 1333), num_known=28, num_classes=88, alpha=0.05, beta=0.95,
                         loss_weight=2.0, unk_thr=0.09)
    elif name == "cfg4":  # inference-only: 32 images, k = 1000 / level, NMS -> 1000 proposals / image
        cfg = PathConfig(name=name, num_images=32, pre_nms_topk=1000, post_nms_topk=1000)
    elif name == "cfg5":  # stress: 1333x1333, k = 4000 / level, 1024 RoIs / image, 128 images over the GPUs
        # (the configuration is defined for 2 / 4 / 8 GPUs: 64 / 32 / 16 images per GPU; one GPU runs the 8-GPU share)
        cfg = PathConfig(name=name, num_images=max(1, 128 // max(world, 1)) if world > 1 else 16, image_hw=(1333, 1333),
                         pre_nms_topk=4000, rois_per_image=1024)
    else:
        raise ValueError(f"unknown config {name!r} (cfg2 | cfg3 | cfg4 | cfg5)")
    for k, v in over.items():
        if v is not None:
            setattr(cfg, k, v)
    return cfg


class RoiPathStep:
    """Device-resident synthetic inputs + the timed step.  ``stage_events=True`` records CUDA events between
    stages so bench.py can attribute time to S1 / S3 / S5 / S3' without extra synchronisation."""

    STAGES = ("s1_proposals", "s2_sample_glue", "s3_roialign_fwd", "s5_pln_fwd_bwd", "s3_roialign_bwd")

    def __init__(self, cfg: PathConfig, device="cuda:0", host_inputs: bool = False):
        self.cfg = cfg
        if cfg.box_head:
            self.STAGES = ("s1_proposals", "s2_sample_glue", "s3_roialign_fwd", "s4_box_head_fc_fwd", "s5_pln_fwd_bwd",
                           "s3_roialign_bwd")
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
        # S2 inputs: ground truth that the kept proposals actually match (a dry run of S1 gives them), so the labelled sampler
        # fills the reference's positive quota and the prototype loss sees its nominal foreground share
        sel = rpn_select_decode(self.anchors, self.deltas, self.ctr, self.image_hw_dev, cfg.pre_nms_topk)
        counts = sel.counts.cpu()
        L = sel.num_levels
        kept = []
        for n in range(N):
            c = int(counts[n, L])
            assert c >= cfg.rois_per_image, f"image {n}: only {c} proposals"
            kept.append(sel.boxes[n, :c])
        self.kmax = sel.kmax
        self.gt_boxes, self.gt_classes, self.gt_off = synth.make_matched_gt(kept, cfg.gt_per_image, num_known=cfg.num_known,
                                                                           seed=cfg.seed + 5)
        self.prop_off = torch.arange(0, (N + 1) * sel.kmax, sel.kmax, dtype=torch.int32, device=dev)
        self.count_col = sel.num_levels
        if cfg.box_head:
            # detectron2 FastRCNNConvFCHead weights (c2_xavier_fill), kept in bf16 for the tensor-core GEMMs
            gw = torch.Generator(device=dev).manual_seed(cfg.seed + 6)
            kin = cfg.channels * 49
            b1 = (3.0 / kin) ** 0.5
            b2 = (3.0 / cfg.feat_dim) ** 0.5
            self.fc1_w = ((torch.rand(cfg.feat_dim, kin, device=dev, generator=gw) * 2 - 1) * b1).to(torch.bfloat16)
            self.fc2_w = ((torch.rand(cfg.feat_dim, cfg.feat_dim, device=dev, generator=gw) * 2 - 1) * b2).to(torch.bfloat16)
            self.fc1_b = torch.zeros(cfg.feat_dim, device=dev)
            self.fc2_b = torch.zeros(cfg.feat_dim, device=dev)
        self.last: Dict[str, torch.Tensor] = {}
        self._fused_enc = None
        self._side = None
        self.fused_gather_error = None
        self.events: Optional[List[torch.cuda.Event]] = None

    # ------------------------------------------------------------------------------------------------
    def _mark(self, i):
        if self.events is not None:
            self.events[i].record()

    def step(self, deltas=None, ctr=None, feats=None, stage_events: bool = False, gather_pln: bool = False, keys=None):
        """One pass of the path.  ``keys``: the sampler's random keys, one per padded proposal row (N * kmax) - drawn afresh
        with ``torch.rand`` when omitted (tests pass fixed keys to compare an eager step with a graph replay)."""
        cfg = self.cfg
        deltas = self.deltas if deltas is None else deltas
        ctr = self.ctr if ctr is None else ctr
        feats = self.feats if feats is None else feats
        self.events = [torch.cuda.Event(enable_timing=True) for _ in range(len(self.STAGES) + 1)] if stage_events else None
        self._mark(0)
        # S1
        sel = rpn_select_decode(self.anchors, deltas, ctr, self.image_hw_dev, cfg.pre_nms_topk)
        self._mark(1)
        # S2 (glue): proposal <-> GT matching of ALL kept proposals (osr_match_label), then the labelled sampling of the
        # whole batch in one launch (osr_sample_rois: fresh random keys, the matcher's classes; positives first, 25 % quota)
        # which also emits ROIAlign's (image, box) rows and the sampled classes / IoUs the prototype loss consumes -
        # label_and_sample_proposals (osrcnn_roi_heads.py:136-230) without Instances and without a host sync
        L2 = sel.counts.shape[1]
        cnt_col = sel.counts[:, self.count_col]
        match = match_proposals(sel.boxes.view(-1, 4), self.prop_off, self.gt_boxes, self.gt_classes, self.gt_off,
                                self.kmax, iou_threshold=cfg.iou_threshold, background_label=cfg.num_classes,
                                box_counts=cnt_col, box_counts_stride=L2)
        if keys is None:
            keys = torch.rand(match[3].shape[0], device=match[3].device)
        smp = sample_rois(match[3], keys, self.prop_off, cfg.rois_per_image, int(cfg.rois_per_image * 0.25), cfg.num_classes,
                          box_counts=cnt_col, box_counts_stride=L2, boxes=sel.boxes.view(-1, 4), ious=match[1], want_rois=True,
                          max_boxes_per_image=self.kmax)
        rois = smp["rois"].view(-1, 5)
        s_cls, s_iou = smp["classes"].view(-1), smp["ious"].view(-1)
        self._mark(2)
        # the backward's per-RoI tables depend on the RoIs only: issue them now on a side stream (a parallel branch of the
        # captured graph), joined right before the backward gather
        bwd_ws = None
        if cfg.overlap_bwd_prep:
            cur = torch.cuda.current_stream(rois.device)
            if self._side is None:
                self._side = torch.cuda.Stream(device=rois.device)
            bwd_ws = self.pooler.alloc_backward_workspace(feats, rois)      # allocated on the main stream
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                self.pooler.prepare_backward(feats, rois, self.roi_offsets, out=bwd_ws)
        # S3 forward (kernels are called directly, without autograd: no engine thread hop, capturable in a CUDA graph;
        # tests/test_gpu_pipeline.py checks this path against the autograd one)
        k = 3
        pi = self.pln
        roi_features = pi.roi_features
        with torch.no_grad():
            if cfg.box_head:
                # S3 writes the pooled tile in bf16; S4 = fc1 + ReLU + fc2 + ReLU on tcgen05 produces the PLN's input
                from .box_head import linear_bf16
                pooled, lvl = self.pooler.pool_rois_bf16(feats, rois, self.roi_offsets)
                self._mark(k); k += 1
                h = linear_bf16(pooled.view(pooled.shape[0], -1), self.fc1_w, self.fc1_b, True, torch.bfloat16)
                roi_features = linear_bf16(h, self.fc2_w, self.fc2_b, True, torch.float32)
            else:
                pooled, lvl = self.pooler.pool_rois([f.detach() for f in feats], rois, self.roi_offsets)
        self._mark(k); k += 1
        # S5: encoder + prototype loss forward + backward to (emb, representatives)
        finish_reps = None
        reps = pi.reps
        kw = dict(num_known_classes=cfg.num_known, alpha=cfg.alpha, beta=cfg.beta, loss_weight=cfg.loss_weight,
                  iou_threshold=cfg.iou_threshold)
        if gather_pln and cfg.encoder_impl == "tcgen05":
            # north-star variant, B200-native: the encoder GEMM's epilogue stores its tiles into every rank's buffer
            # over NVLink (fused all-gather), then the global-batch loss (dist.py)
            if self._fused_enc is None:
                try:
                    self._fused_enc = FusedEncoderGather(roi_features.shape[0], cfg.emb_dim, self.device)
                except Exception as e:  # noqa: BLE001 - no symmetric memory / P2P on this box: NCCL all-gather instead
                    self._fused_enc = False
                    self.fused_gather_error = repr(e)
        if gather_pln:
            # global-batch loss, autograd-free (capturable): gathered rows -> loss + closed-form gradients of ALL rows, the
            # local block of d loss / d emb is this rank's (scaled by W: dist.py parity rule)
            import torch.distributed as tdist
            from .dist import _gather_meta, all_gather_rows
            W = tdist.get_world_size() if tdist.is_initialized() else 1
            rank = tdist.get_rank() if tdist.is_initialized() else 0
            R_loc = roi_features.shape[0]
            with torch.no_grad():
                if cfg.encoder_impl == "tcgen05" and self._fused_enc:
                    emb, emb_all = self._fused_enc(roi_features, pi.enc_w, pi.enc_b)
                else:
                    emb = (pln_encode_tc(roi_features, pi.enc_w, pi.enc_b) if cfg.encoder_impl == "tcgen05"
                           else F.linear(roi_features, pi.enc_w, pi.enc_b))
                    emb_all = all_gather_rows(emb) if W > 1 else emb
                if W > 1 and gather_pln != "reduce":
                    labels_all, ious_all = _gather_meta(s_cls, s_iou, None)
                else:
                    labels_all, ious_all = s_cls, s_iou
            if gather_pln == "reduce":
                # embeddings are gathered (fused into the encoder epilogue above) but every row term depends only on (its
                # embedding, the prototypes): per-rank loss + ONE all-reduce of (loss, representatives.grad) gives the same
                # value and gradients as evaluating the loss kernels on all W*R rows (dist.reduced_pln_loss; tested)
                loss, g_emb, g_reps = pln_loss_fwd_bwd(emb, reps, s_cls, s_iou, **kw)
                if W > 1:
                    packed = torch.cat((loss.reshape(1), g_reps.reshape(-1)))
                    tdist.all_reduce(packed)
                    packed = packed / W
                    loss, g_reps = packed[0], packed[1:].view_as(g_reps)
            else:
                loss, g_all, g_reps = pln_loss_fwd_bwd(emb_all, reps, labels_all, ious_all, r_norm=float(W * R_loc),
                                                       center_weight=float(W), emb_grad_scale=float(W), **kw)
                g_emb = g_all[rank * R_loc:(rank + 1) * R_loc]
        else:            # the reference's semantics: per-rank loss
            with torch.no_grad():
                if cfg.encoder_impl == "tcgen05":
                    emb = pln_encode_tc(roi_features, pi.enc_w, pi.enc_b)
                else:
                    emb = F.linear(roi_features, pi.enc_w, pi.enc_b)
            if cfg.overlap_bwd_prep:
                # d loss / d emb (what the rest of the backward pass waits for) is complete after the row launch; the
                # prototype-gradient launches + loss reduction are an independent branch of the backward graph and run on the
                # side stream next to the ROIAlign backward
                finish_reps, g_emb = pln_loss_fwd_bwd(emb, reps, s_cls, s_iou, split=True, **kw)
            else:
                loss, g_emb, g_reps = pln_loss_fwd_bwd(emb, reps, s_cls, s_iou, **kw)
        self._mark(k); k += 1
        # S3 backward
        cur = torch.cuda.current_stream(rois.device)
        if bwd_ws is not None:
            cur.wait_stream(self._side)          # the backward's tables (issued before the forward)
        if finish_reps is not None:
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                loss, g_reps = finish_reps()
        g_feats = self.pooler.backward_rois(self.grad_pooled, feats, rois, self.roi_offsets, prepared=bwd_ws)
        if finish_reps is not None:
            cur.wait_stream(self._side)
        self._mark(k)
        self.last = dict(sel=sel, match=match, sample=smp, keys=keys, rois=rois, pooled=pooled, level=lvl, loss=loss, g_emb=g_emb, g_reps=g_reps,
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

    def alg_bytes(self) -> Dict[str, int]:
        """Algorithmic bytes per stage of the last step (needs one step to have run)."""
        from . import roofline
        cfg = self.cfg
        N = cfg.num_images
        level_shapes = self.grid_sizes[:4]
        M = N * cfg.rois_per_image
        self.touched = roofline.touched_pixels(level_shapes, synth.POOL_SCALES, self.last["rois"], self.last["level"], N)
        return {
            "s1_proposals": N * roofline.s1_bytes_per_image(self.grid_sizes, cfg.pre_nms_topk),
            "s3_roialign_fwd": roofline.s3_fwd_bytes(M, cfg.channels, 7, self.touched),
            "s3_roialign_bwd": roofline.s3_bwd_bytes(M, cfg.channels, 7, N, level_shapes),
            "s5_pln_fwd_bwd": roofline.s5_fwd_bytes(M, cfg.feat_dim, cfg.emb_dim, cfg.num_known) +
                              roofline.s5_bwd_bytes(M, cfg.emb_dim, cfg.num_known),
        }

    def box_head_flops(self) -> int:
        """fc1 + fc2 forward FLOPs of one step (2 R K N each)."""
        cfg = self.cfg
        R = cfg.num_images * cfg.rois_per_image
        return 2 * R * (cfg.channels * 49) * cfg.feat_dim + 2 * R * cfg.feat_dim * cfg.feat_dim


# =====================================================================================================================
class ApiTrainStep:
    """The same training step through the DROP-IN API, object for object what the reference's ``GeneralizedRCNN`` does
    between the RPN head and the box head (``classification_free_rpn.py:545``, ``osrcnn_roi_heads.py:268-316``):

        proposals = ClsFreeRPNProposals.predict_proposals(anchors, deltas, centerness, image_sizes)   -> List[Instances]
        sampled   = label_and_sample_proposals(proposals, targets)       (matcher + torch.randperm on the device)
        pooled    = ROIPooler.forward(features, [x.proposal_boxes for x in sampled])
        _, _, l   = PLN.loss(box_features, sampled)                       (labels / ious are the MATCHER's, not synthetic)
        backward  : PLN loss -> (embedding, prototypes);  pooled -> feature maps

    ``Instances`` construction, the proposal-count host sync and the sampler's ``nonzero`` syncs are inside the step -
    this is what ``e2e_api`` in the bench line times.  Box-head FC: a fixed (R, 1024) tensor stands in (as in
    ``RoiPathStep``)."""

    STAGES = ("s1_proposals", "s2_label_and_sample", "s3_roialign_fwd", "s5_pln_fwd_bwd", "s3_roialign_bwd")

    def __init__(self, cfg: PathConfig, device="cuda:0"):
        from .pln import PLN
        from .proposals import ClsFreeRPNProposals
        from .structures import Boxes, Instances
        self.cfg = cfg
        self.device = dev = torch.device(device)
        N = cfg.num_images
        ho = synth.make_head_outputs(N, cfg.image_hw, seed=cfg.seed)
        self.image_sizes = ho.image_sizes
        self.anchors = [Boxes(a.to(dev)) for a in ho.anchors]
        self.deltas = [d.to(dev) for d in ho.deltas]
        self.ctr = [c.to(dev) for c in ho.centerness]
        self.feats = synth.make_features(N, cfg.image_hw, cfg.channels, seed=cfg.seed + 1, device=dev,
                                         channels_last=cfg.channels_last)
        self.rpn = ClsFreeRPNProposals(pre_nms_topk=(cfg.pre_nms_topk, cfg.pre_nms_topk),
                                       post_nms_topk=(cfg.post_nms_topk, cfg.post_nms_topk), nms_thresh=(1.0, 1.0))
        self.pooler = ROIPooler(7, synth.POOL_SCALES, 0, "ROIAlignV2")
        self.pln = PLN(cfg.num_classes, cfg.num_known, cfg.feat_dim, cfg.emb_dim, "COS", 1, cfg.alpha, cfg.beta,
                       cfg.loss_weight, "synthetic", cfg.iou_threshold, cfg.unk_thr, True, device=dev,
                       encoder_impl=cfg.encoder_impl)
        gtb, gtc, goff = synth.make_gt(N, cfg.gt_per_image, cfg.image_hw, num_known=cfg.num_known, seed=cfg.seed + 5, device=dev)
        self.targets = []
        for n in range(N):
            t = Instances(tuple(self.image_sizes[n]))
            t.set("gt_boxes", Boxes(gtb[n * cfg.gt_per_image:(n + 1) * cfg.gt_per_image]))
            t.set("gt_classes", gtc[n * cfg.gt_per_image:(n + 1) * cfg.gt_per_image])
            self.targets.append(t)
        R = N * cfg.rois_per_image
        g = torch.Generator(device=dev).manual_seed(cfg.seed + 3)
        self.box_features = torch.relu(torch.randn(R, cfg.feat_dim, device=dev, generator=g))
        self.grad_pooled = torch.randn(R, cfg.channels, 7, 7, device=dev, generator=g)
        self.events = None
        self.last = {}

    def _mark(self, i):
        if self.events is not None:
            self.events[i].record()

    def step(self, stage_events: bool = False):
        from .sampling import label_and_sample_proposals
        cfg = self.cfg
        self.events = [torch.cuda.Event(enable_timing=True) for _ in range(6)] if stage_events else None
        self._mark(0)
        proposals = self.rpn.train().predict_proposals(self.anchors, self.deltas, self.ctr, self.image_sizes)
        self._mark(1)
        sampled = label_and_sample_proposals(proposals, self.targets, num_classes=cfg.num_classes,
                                             batch_size_per_image=cfg.rois_per_image, positive_fraction=0.25,
                                             iou_threshold=0.5, proposal_append_gt=True)
        self._mark(2)
        feats = [f.requires_grad_(True) for f in self.feats]
        pooled = self.pooler(feats, [x.proposal_boxes for x in sampled])
        self._mark(3)
        M = pooled.shape[0]
        emb, rec, loss = self.pln.loss(self.box_features[:M], sampled)
        grads = torch.autograd.grad(loss, [self.pln.representatives, self.pln.encoder.weight])
        self._mark(4)
        g_feats = torch.autograd.grad(pooled, feats, self.grad_pooled[:M])
        self._mark(5)
        self.last = dict(proposals=proposals, sampled=sampled, pooled=pooled, loss=loss, grads=grads, g_feats=g_feats)
        return loss


class InferencePathStep:
    """Inference-only RoI path (BASELINE.json configs[3]) through the drop-in API:

        S1  predict_proposals(mode="nominal", training=False): top-k / level + decode + per-level NMS + best post_nms_topk
            (``classification_free_rpn.py:558-589`` with the block at ``find_top_proposals.py:112-120`` switched on - the
            only configuration that yields "1000 post-NMS proposals / image")
        S3  ROIPooler.forward on all kept proposals (<= 32 000 RoIs, ``osrcnn_roi_heads.py:306``)
        S6  ROI-head post-processing: box decode + objectness + NMS(1.0) + top-1000 (``osrcnn_fast_rcnn.py:380-404``),
            ``PLN.inference`` (``prototype_learning_network.py:189-230``), ``SoftMaxClassifier.inference`` per-class NMS
            (``softmax_classifier.py:287-346``) - fixed tensors stand in for the box head / predictor outputs."""

    STAGES = ("s1_proposals_nms", "s3_roialign_fwd", "s6_roi_head_postprocess")

    def __init__(self, cfg: PathConfig, device="cuda:0"):
        from .pln import PLN
        from .structures import Boxes
        self.cfg = cfg
        self.device = dev = torch.device(device)
        N = cfg.num_images
        ho = synth.make_head_outputs(N, cfg.image_hw, seed=cfg.seed)
        self.grid_sizes = ho.grid_sizes
        self.image_sizes = ho.image_sizes
        self.anchors = [Boxes(a.to(dev)) for a in ho.anchors]
        self.deltas = [d.to(dev) for d in ho.deltas]
        self.ctr = [c.to(dev) for c in ho.centerness]
        self.feats = synth.make_features(N, cfg.image_hw, cfg.channels, seed=cfg.seed + 1, device=dev,
                                         channels_last=cfg.channels_last)
        self.pooler = ROIPooler(7, synth.POOL_SCALES, 0, "ROIAlignV2")
        self.pln = PLN(cfg.num_classes, cfg.num_known, cfg.feat_dim, cfg.emb_dim, "COS", 1, cfg.alpha, cfg.beta,
                       cfg.loss_weight, "synthetic", cfg.iou_threshold, cfg.unk_thr, True, device=dev).eval()
        Rmax = N * cfg.post_nms_topk
        g = torch.Generator(device=dev).manual_seed(cfg.seed + 3)
        self.box_features = torch.relu(torch.randn(Rmax, cfg.feat_dim, device=dev, generator=g))
        self.pred_deltas = torch.randn(Rmax, 4, device=dev, generator=g) * 0.5
        self.pred_iou = torch.rand(Rmax, 1, device=dev, generator=g)
        self.cls_score = torch.nn.Linear(cfg.feat_dim, cfg.num_known + 1, device=dev)
        with torch.no_grad():   # prototypes near some embeddings so that PLN.inference yields known AND unknown detections
            e = self.pln.encoder(self.box_features[:cfg.num_known * 37:37])
            self.pln.representatives.copy_(e + 0.1 * torch.randn(e.shape, device=dev, generator=g) * e.norm(dim=1, keepdim=True) / 16)
            self.cls_score.weight.mul_(20.0)
        self.events = None
        self.last = {}

    def _mark(self, i):
        if self.events is not None:
            self.events[i].record()

    @torch.no_grad()
    def step(self, stage_events: bool = False):
        from .inference import inference, softmax_classifier_inference
        from .proposals import predict_proposals
        cfg = self.cfg
        self.events = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if stage_events else None
        self._mark(0)
        proposals = predict_proposals(self.anchors, self.deltas, self.ctr, self.image_sizes, nms_thresh=cfg.rpn_nms_thresh,
                                      pre_nms_topk=cfg.pre_nms_topk, post_nms_topk=cfg.post_nms_topk, training=False,
                                      mode="nominal")
        self._mark(1)
        boxes = [x.proposal_boxes for x in proposals]
        pooled, lvl = self.pooler.forward_with_levels(self.feats, boxes)
        self._mark(2)
        M = pooled.shape[0]
        fg, _ = inference((self.pred_deltas[:M], self.pred_iou[:M]), proposals, self.box_features[:M], score_thresh=0.05,
                          nms_thresh=1.0, topk_per_image=1000)
        fg = self.pln.inference(fg)
        dets = softmax_classifier_inference(fg, self.cls_score, unknown_id=80, known_score_thresh=0.05, known_nms_thresh=0.5,
                                            known_topk=50, unknown_score_thresh=0.0, unknown_nms_thresh=0.5, unknown_topk=50)
        self._mark(3)
        self.last = dict(proposals=proposals, pooled=pooled, level=lvl, dets=dets, M=M)
        return pooled

    def alg_bytes(self) -> Dict[str, int]:
        from . import roofline
        cfg = self.cfg
        N = cfg.num_images
        M = int(self.last["M"])
        rois = torch.cat([torch.cat((torch.full((len(p), 1), float(n), device=self.device), p.proposal_boxes.tensor), dim=1)
                          for n, p in enumerate(self.last["proposals"])])
        self.touched = roofline.touched_pixels(self.grid_sizes[:4], synth.POOL_SCALES, rois, self.last["level"], N)
        return {
            "s1_proposals_nms": N * roofline.s1_bytes_per_image(self.grid_sizes, cfg.pre_nms_topk, nominal_post_k=cfg.post_nms_topk),
            "s3_roialign_fwd": roofline.s3_fwd_bytes(M, cfg.channels, 7, self.touched),
        }
