#!/usr/bin/env python
"""bench.py - RoI-path images/s (CF-RPN proposals + ROIAlignV2 fwd/bwd + PLN loss fwd/bwd) on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W

Workload = BASELINE.json configs[1]: R50-FPN VOC-COCO training step, 16 synthetic 800x1333 images per GPU,
2000 pre-NMS CF-RPN proposals per FPN level (7 323/img as the reference ships it), 512 RoIs/img, PLN K=20.
One "step" = one pass of the hot path over that batch (osr_b200/pipeline.py).  Images are sharded across GPUs
(weak scaling, no data-path collective in this configuration: the reference's PLN loss is per-rank).

Prints ONE JSON line (rank 0).  `value` = device-resident throughput; `e2e` = same step fed from pinned HOST
buffers (H2D of head outputs + FPN maps inside the timed region, loss + proposal counts read back);
`roofline` = dominant kernel vs the measured HBM peak; `cpu_baseline` = the oracle (the reference's
torch/torchvision CPU code path) on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "openset-rcnn_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "roi_path_images_per_sec"
UNIT = "images/s"
WORKLOADS = {
    "cfg2": "cfg2_R50FPN_train_16img_per_gpu_800x1333_k2000_512rois_K20",
    "cfg3": "cfg3_GraspNet_train_8img_per_gpu_750x1333_k2000_512rois_K28_gathered_PLN",
    "cfg4": "cfg4_inference_32img_per_gpu_800x1333_k1000_nms_to_1000_proposals",
    "cfg5": "cfg5_stress_128img_total_1333x1333_k4000_1024rois",
}
WORKLOAD = WORKLOADS["cfg2"]


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML from a background thread every ~2 ms (the
    timed region of the default run is only ~30 ms, too short for an `nvidia-smi -lms` child process to start)."""

    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40), ("sw_power_cap", 0x4))

    def __init__(self, index: int):
        self.index = index
        self.sm, self.reasons, self.smax = [], set(), None
        self.h = None
        self._stop = threading.Event()
        self.t = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:  # noqa: BLE001
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
                self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.h = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for name, bit in self.REASONS:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def start(self):
        if self.h is None:
            return
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml unavailable"]}
        self._stop.set()
        self.t.join(timeout=1.0)
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.smax,
                "samples": len(self.sm), "reasons": sorted(self.reasons)}


def barrier(world):
    if world > 1:
        dist.barrier()


def max_over_ranks(x: float, world: int, device) -> float:
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ----------------------------------------------------------------------------------------------- CPU arm
def build_cpu_path(num_images: int, cfg):
    """The oracle (reference torch/torchvision CPU code path) on `num_images` images of the same workload."""
    from oracle.pipeline import CpuRoiPath  # the one place bench.py executes oracle/: as the measured CPU baseline
    from osr_b200 import synth
    ho = synth.make_head_outputs(num_images, cfg.image_hw, seed=cfg.seed)
    feats = synth.make_features(num_images, cfg.image_hw, cfg.channels, seed=cfg.seed + 1)
    R = num_images * cfg.rois_per_image
    pi = synth.make_pln_inputs(R, feat_dim=cfg.feat_dim, emb_dim=cfg.emb_dim, num_known=cfg.num_known,
                               num_classes=cfg.num_classes, seed=cfg.seed + 2)
    g = torch.Generator().manual_seed(cfg.seed + 3)
    grad_pooled = torch.randn(R, cfg.channels, 7, 7, generator=g)
    # ground truth as in the CUDA arm: matched to the kept proposals of a dry run of S1, so the labelled sampler fills its quota
    from oracle import rpn as orpn
    from oracle.structures import Boxes as OBoxes
    props = orpn.predict_proposals([OBoxes(a) for a in ho.anchors], ho.deltas, ho.centerness, ho.image_sizes,
                                   pre_nms_topk=cfg.pre_nms_topk, post_nms_topk=cfg.pre_nms_topk, training=True)
    gtb, gtc, goff = synth.make_matched_gt([p.proposal_boxes.tensor for p in props], cfg.gt_per_image, num_known=cfg.num_known,
                                           seed=cfg.seed + 5)
    targets = [(gtb[int(goff[n]):int(goff[n + 1])], gtc[int(goff[n]):int(goff[n + 1])]) for n in range(num_images)]
    path = CpuRoiPath(ho, feats, pi, grad_pooled, None, pre_nms_topk=cfg.pre_nms_topk,
                      rois_per_image=cfg.rois_per_image, num_known=cfg.num_known, alpha=cfg.alpha, beta=cfg.beta,
                      loss_weight=cfg.loss_weight, iou_threshold=cfg.iou_threshold, targets=targets,
                      num_classes=cfg.num_classes)
    return path


def cpu_baseline(cfg, budget_s: float = 20.0):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_img = 1
    path = build_cpu_path(n_img, cfg)
    path.step()  # warm-up
    times = []
    t_start = time.perf_counter()
    while len(times) < 3 and (time.perf_counter() - t_start) < budget_s:
        times.append(path.step()["total"])
    best = min(times)
    return {"value": n_img / best, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_img} image of the same workload (k=2000/level, 512 RoIs, fwd+bwd), 1 warm-up + best of {len(times)}, "
                      f"torch {torch.get_num_threads()} threads; oracle = reference torch/torchvision CPU code path"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (oracle port: detectron2 is not
    installable here), all host threads, same metric/config; each step = a bounded sample of the workload."""
    if rank != 0:
        return
    from osr_b200.pipeline import PathConfig
    cfg = PathConfig()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_img = 1
    path = build_cpu_path(n_img, cfg)
    for _ in range(max(args.warmup, 1)):
        path.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        path.step()
    dt = time.perf_counter() - t0
    value = n_img * args.steps / dt
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": WORKLOAD, "sample_images_per_step": n_img, "device": "cpu"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_img} image/step of the same workload; oracle port of the reference's "
                                   f"torch/torchvision CPU path, {torch.get_num_threads()} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU library baseline
def gpu_baseline(path, cfg, iters: int = 5):
    """The kernels the reference actually runs on a GPU, timed on this B200 on the SAME
    inputs in the same process: torchvision-CUDA roi_align forward / backward (per level + index_put, as detectron2's
    ROIPooler does, NCHW maps), ATen topk + the elementwise decode of ALL anchors, torchvision-CUDA nms, and the
    torch op chain of PLN.loss.  Library calls only (nothing from oracle/); the reference's per-image Python loops and
    host syncs are NOT included, so this flatters the reference."""
    import torchvision
    dev = path.device
    N = cfg.num_images

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / iters

    out = {}
    # S1: decode every anchor (Box2BoxTransformLinear) + per-level topk + gather
    def s1():
        res = []
        for a, d, c in zip(path.anchors, path.deltas, path.ctr):
            d = torch.relu(d)
            cx, cy = 0.5 * (a[:, 0] + a[:, 2]), 0.5 * (a[:, 1] + a[:, 3])
            st = torch.stack([a[:, 2] - a[:, 0], a[:, 3] - a[:, 1]] * 2, dim=1)
            d = d * st
            boxes = torch.stack((cx - d[..., 0], cy - d[..., 1], cx + d[..., 2], cy + d[..., 3]), dim=-1)
            k = min(c.shape[1], cfg.pre_nms_topk)
            v, i = c.topk(k, dim=1)
            res.append(torch.gather(boxes, 1, i[..., None].expand(-1, -1, 4)))
        return torch.cat(res, dim=1)
    out["s1_decode_topk_ms"] = timed(s1)
    # S3: detectron2 ROIPooler = per-level torchvision roi_align + index_put (NCHW maps), backward by autograd
    from osr_b200 import synth
    rois = path.last["rois"].detach()
    lvl = path.last["level"].long()
    feats = [f.detach().contiguous().requires_grad_(True) for f in path.feats]
    idx = [torch.nonzero(lvl == l, as_tuple=True)[0] for l in range(4)]

    def s3_fwd():
        o = torch.zeros((rois.shape[0], cfg.channels, 7, 7), device=dev)
        for l in range(4):
            o.index_put_((idx[l],), torchvision.ops.roi_align(feats[l], rois[idx[l]], (7, 7), synth.POOL_SCALES[l], 0, True))
        return o
    out["s3_roialign_fwd_ms"] = timed(s3_fwd)
    pooled = s3_fwd()
    out["s3_roialign_bwd_ms"] = timed(lambda: torch.autograd.grad(pooled, feats, path.grad_pooled, retain_graph=True))
    del pooled, feats
    # NMS: one image's kept proposals (level-major, as the nominal mode / ROI-head sites see them), thr 0.7
    sel = path.last["sel"]
    c = int(sel.counts[0, sel.num_levels])
    b, sc = sel.boxes[0, :c].contiguous(), sel.scores[0, :c].contiguous()
    out["nms_boxes"] = c
    out["nms_torchvision_ms"] = timed(lambda: torchvision.ops.nms(b, sc, 0.7))
    from osr_b200 import nms as onms
    out["nms_ours_ms"] = timed(lambda: onms.nms(b, sc, 0.7))
    # S5: the torch op chain of PLN.loss (encoder Linear fp32 + loss) forward + backward
    import torch.nn.functional as F
    pi = path.pln

    def s5():
        x = pi.roi_features
        w = pi.enc_w.detach().requires_grad_(True)
        reps = pi.reps.detach().requires_grad_(True)
        emb = F.linear(x, w, pi.enc_b)
        nf, r = F.normalize(emb), F.normalize(reps)
        fg = torch.nonzero((pi.gt_classes >= 0) & (pi.gt_classes < cfg.num_known) & (pi.ious > cfg.iou_threshold), as_tuple=True)[0]
        dist = 1.0 - nf[fg] @ r.t()
        ar = torch.arange(fg.numel(), device=dev)
        intra = dist[ar, pi.gt_classes[fg]]
        dist = dist.clone()
        dist[ar, pi.gt_classes[fg]] = 1000
        inter = dist.min(dim=1)[0]
        cd = (1.0 - r @ r.t()).clone()
        cd.fill_diagonal_(1000)
        cmin = cd.min(dim=1)[0]
        loss = ((intra - cfg.alpha).clamp(min=0).sum() + (cfg.beta - inter).clamp(min=0).sum() +
                (cfg.beta + cfg.alpha - cmin).clamp(min=0).sum()) * cfg.loss_weight / x.shape[0]
        return torch.autograd.grad(loss, [w, reps])
    out["s5_pln_fwd_bwd_ms"] = timed(s5)
    out["path_ms"] = out["s1_decode_topk_ms"] + out["s3_roialign_fwd_ms"] + out["s3_roialign_bwd_ms"] + out["s5_pln_fwd_bwd_ms"]
    out["images_per_s"] = N * 1e3 / out["path_ms"]
    out["note"] = ("torchvision 0.26 CUDA roi_align fwd/bwd (NCHW, per level + index_put), ATen topk + full decode, "
                   "torchvision nms, torch PLN.loss op chain (fp32) on the same inputs; Python per-image loops / host syncs "
                   "of the reference excluded")
    return out


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, local_rank, world):
    from osr_b200 import _lib, roofline, synth
    from osr_b200.pipeline import PathConfig, RoiPathStep

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _lib.lib()  # fail loudly if the extension is missing
    from osr_b200.pipeline import ApiTrainStep, InferencePathStep, make_config
    cfg = make_config(args.config, world, channels_last=not args.nchw, seed=1234 + 1000 * int(args.config[3:]) + rank,
                      num_images=args.images)
    N = cfg.num_images
    infer = args.config == "cfg4"
    gather_headline = (args.config == "cfg3") and world > 1   # cfg 3 IS the gathered-PLN configuration
    path = InferencePathStep(cfg, dev) if infer else RoiPathStep(cfg, dev)
    step_kw = {"gather_pln": "reduce"} if gather_headline else {}

    for _ in range(max(args.warmup, 3)):
        path.step(**step_kw)
    torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    stage_events = []
    barrier(world)
    torch.cuda.synchronize(dev)
    if sampler:
        sampler.start()
    # ---- eager launches: K timed steps with CUDA events between the stages --------------------------------------------
    _lib.reset_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        path.step(stage_events=True, **step_kw)
        stage_events.append(path.events)
    ev1.record()
    torch.cuda.synchronize(dev)
    launches = _lib.launch_count()
    barrier(world)
    eager_ms = max_over_ranks(ev0.elapsed_time(ev1), world, dev) / args.steps
    eager = {"ms_per_step": eager_ms, "value": world * N * 1e3 / eager_ms}
    ms_step = eager_ms

    # ---- one CUDA graph replay per step: the fixed-shape device-resident step is captured once; the
    # headline is the graph-launched step when the capture succeeds, the eager one otherwise ------------------------------
    graph, graph_err = None, None
    if not args.no_graph and not infer:
        # (the fused encoder + all-gather alternates between two symmetric buffer sets: capture two steps per graph there)
        spg = 2 if gather_headline else 1
        reps_n = max(1, args.steps // spg)
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(spg):
                    path.step(**step_kw)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            barrier(world)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for _ in range(spg):
                    path.step(**step_kw)
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize(dev)
            barrier(world)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(reps_n):
                graph.replay()
            g1.record()
            torch.cuda.synchronize(dev)
            barrier(world)
            ms_step = max_over_ranks(g0.elapsed_time(g1), world, dev) / (reps_n * spg)
        except Exception as e:  # noqa: BLE001 - keep the eager number, say so in the line
            graph, graph_err = None, repr(e)[:300]
            try:
                torch.cuda.synchronize(dev)
            except Exception:  # noqa: BLE001
                pass
    clocks = sampler.stop() if sampler else None
    value = world * N * 1e3 / ms_step
    if graph is None:
        eager = None

    # per-stage device time (CUDA events on the launch stream, averaged over the timed steps)
    stages = {}
    for i, name in enumerate(path.STAGES):
        stages[name] = sum(ev[i].elapsed_time(ev[i + 1]) for ev in stage_events) / len(stage_events)

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    peak, peak_src = measured_peak()
    alg = path.alg_bytes()
    U = path.touched
    per_stage = {k: {"ms": stages[k], "alg_bytes": alg[k], "gbs": alg[k] / (stages[k] * 1e-3) / 1e9,
                     "frac": alg[k] / (stages[k] * 1e-3) / 1e9 / peak} for k in alg}
    dom = max(alg, key=lambda k: stages[k])
    path_bytes = sum(alg.values())
    path_ms = sum(stages[k] for k in alg)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r2_dram_traffic.json")
    if cfg.channels_last and args.config == "cfg2" and os.path.exists(tpath):   # ncu capture of the same workload (per launch), committed
        with open(tpath) as f:
            tj = json.load(f)
        if dom in tj:
            traffic, traffic_src = tj[dom]["dram_bytes_per_launch"], tj[dom]["source"]
    roof = {"bound": "hbm", "kernel": dom, "achieved": per_stage[dom]["gbs"], "peak": peak, "unit": "GB/s",
            "frac": per_stage[dom]["frac"], "traffic": traffic, "traffic_source": traffic_src,
            "alg_bytes": alg[dom], "peak_source": peak_src,
            "path_aggregate": {"alg_bytes": path_bytes, "ms": path_ms, "gbs": path_bytes / (path_ms * 1e-3) / 1e9,
                               "frac": path_bytes / (path_ms * 1e-3) / 1e9 / peak},
            "stages": per_stage, "touched_feature_pixels": U}

    if not infer and path.last.get("sample") is not None:   # what the labelled sampler drew in the last step
        _c = path.last["sample"]["count"].cpu()
        _cls, _iou = path.last["sample"]["classes"], path.last["sample"]["ious"]
        _EXTRA["sampled_rows"] = {"per_image_min": int(_c[:, 1].min()), "positives_per_image_mean": float(_c[:, 0].float().mean()),
                                  "pln_foreground_rows": int(((_cls < cfg.num_known) & (_iou > cfg.iou_threshold)).sum())}

    # ---- the library kernels the reference runs on a GPU, same inputs, same process ---------------------
    gbase = None
    if rank == 0 and not args.no_gpu_baseline and not infer:
        try:
            gbase = gpu_baseline(path, cfg)
        except Exception as e:  # noqa: BLE001 - a side measurement must never take the headline line down
            gbase = {"error": repr(e)}
        torch.cuda.empty_cache()

    # ---- the other feature layout, short run (same inputs, same code path selection rules) --------------
    alt = None
    if (rank == 0 or world > 1) and not infer and not args.quick:
        alt_feats = [f.contiguous() if cfg.channels_last else f.contiguous(memory_format=torch.channels_last) for f in path.feats]
        for _ in range(3):
            path.step(feats=alt_feats)
        torch.cuda.synchronize(dev)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(5):
            path.step(feats=alt_feats)
        a1.record()
        torch.cuda.synchronize(dev)
        alt_ms = a0.elapsed_time(a1) / 5
        alt = {"feature_layout": "NCHW" if cfg.channels_last else "channels_last", "ms_per_step": alt_ms,
               "value_per_gpu": N * 1e3 / alt_ms, "steps": 5}
        del alt_feats

    # ---- north-star variant at N > 1: PLN loss over the global batch (NCCL all-gather of the embeddings) ---------
    gathered = None
    if world > 1 and not infer and not gather_headline:
        try:
            for _ in range(3):
                path.step(gather_pln=True)
            torch.cuda.synchronize(dev)
            barrier(world)
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(10):
                path.step(gather_pln=True)
            g1.record()
            torch.cuda.synchronize(dev)
            g_ms = max_over_ranks(g0.elapsed_time(g1), world, dev) / 10
            R_loc = N * cfg.rois_per_image
            fused = bool(path._fused_enc)
            gathered = {"ms_per_step": g_ms, "value": world * N * 1e3 / g_ms, "steps": 10,
                        "all_gather_bytes_received_per_rank": (world - 1) * R_loc * (cfg.emb_dim * 4 + 8),
                        "fused_encoder_gather": fused,
                        "note": ("same step with the PLN loss evaluated over the global batch: the tcgen05 encoder GEMM's "
                                 "epilogue stores its tiles into every rank's symmetric-memory buffer over NVLink (fused "
                                 "all-gather, osr_pln_encode_gather_fwd), " if fused else
                                 "same step with the PLN loss evaluated over the global batch: NCCL all_gather_into_tensor "
                                 "of the embeddings (symmetric memory unavailable: %s), " % path.fused_gather_error) +
                                "labels/ious by NCCL all_gather, every rank then runs the loss kernels on W*R rows "
                                "(osr_b200/dist.py); the headline value keeps the reference's per-rank loss"}
            # fused gather + per-rank loss terms + one all-reduce of (loss, representatives.grad): same numbers, W-independent cost
            for _ in range(3):
                path.step(gather_pln="reduce")
            torch.cuda.synchronize(dev)
            barrier(world)
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            for _ in range(10):
                path.step(gather_pln="reduce")
            r1.record()
            torch.cuda.synchronize(dev)
            r_ms = max_over_ranks(r0.elapsed_time(r1), world, dev) / 10
            gathered["reduced"] = {"ms_per_step": r_ms, "value": world * N * 1e3 / r_ms,
                                   "note": "fused encoder + all-gather as above; the loss kernels run on the LOCAL rows and (loss, "
                                           "representatives.grad) are all-reduced (1 + K*256 floats): identical value / gradients "
                                           "(tests/test_dist_gloo.py), cost independent of the world size"}
            # the same gathered step as ONE CUDA graph per rank (peer stores, the symmetric-memory barrier and the NCCL
            # all-gather of the labels are all stream-ordered and capturable)
            try:
                if args.no_gather_graph:
                    raise RuntimeError("skipped (--no-gather-graph)")
                gg = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    path.step(gather_pln="reduce")
                    path.step(gather_pln="reduce")
                torch.cuda.current_stream(dev).wait_stream(side)
                torch.cuda.synchronize(dev)
                barrier(world)
                with torch.cuda.graph(gg):
                    path.step(gather_pln="reduce")
                    path.step(gather_pln="reduce")   # two steps: the fused kernel alternates between two symmetric buffer sets
                for _ in range(2):
                    gg.replay()
                torch.cuda.synchronize(dev)
                barrier(world)
                h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                h0.record()
                for _ in range(5):
                    gg.replay()
                h1.record()
                torch.cuda.synchronize(dev)
                gg_ms = max_over_ranks(h0.elapsed_time(h1), world, dev) / 10
                gathered["reduced"]["graph_ms_per_step"] = gg_ms
                gathered["reduced"]["graph_value"] = world * N * 1e3 / gg_ms
                del gg
            except Exception as e:  # noqa: BLE001
                gathered["graph_error"] = repr(e)[:200]
                try:
                    torch.cuda.synchronize(dev)
                except Exception:  # noqa: BLE001
                    pass
        except Exception as e:  # noqa: BLE001 - the side measurement must never take the headline line down
            gathered = {"error": repr(e)}
        barrier(world)

    # ---- the same training step through the drop-in API objects (Instances, matcher labels feeding PLN) -------------
    api = None
    if not infer and not args.quick:
        try:
            del graph
            api_path = ApiTrainStep(cfg, dev)
            for _ in range(3):
                api_path.step()
            torch.cuda.synchronize(dev)
            barrier(world)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            api_events = []
            a0.record()
            for _ in range(10):
                api_path.step(stage_events=True)
                api_events.append(api_path.events)
            a1.record()
            torch.cuda.synchronize(dev)
            api_ms = max_over_ranks(a0.elapsed_time(a1), world, dev) / 10
            api = {"ms_per_step": api_ms, "value": world * N * 1e3 / api_ms, "steps": 10,
                   "stage_ms": {name: sum(ev[i].elapsed_time(ev[i + 1]) for ev in api_events) / len(api_events)
                                for i, name in enumerate(api_path.STAGES)},
                   "sampled_rois": int(api_path.last["pooled"].shape[0]),
                   "note": "predict_proposals -> label_and_sample_proposals (matcher + device randperm) -> ROIPooler.forward("
                           "list, list) -> PLN.loss(box_features, sampled) with the matcher's labels / IoUs -> backward; "
                           "Instances construction and the two host reads (proposal counts, sample counts) are inside"}
            del api_path
        except Exception as e:  # noqa: BLE001
            api = {"error": repr(e)}
        barrier(world)

    # ---- S4 on the path: ROIAlign writes bf16, fc1 / fc2 run on tcgen05 and feed the PLN ---------------
    bh = None
    if not infer and not args.quick and cfg.channels_last:
        try:
            import dataclasses
            bh_path = RoiPathStep(dataclasses.replace(cfg, box_head=True), dev)
            for _ in range(3):
                bh_path.step()
            torch.cuda.synchronize(dev)
            barrier(world)
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            bh_events = []
            b0.record()
            for _ in range(10):
                bh_path.step(stage_events=True)
                bh_events.append(bh_path.events)
            b1.record()
            torch.cuda.synchronize(dev)
            bh_ms = max_over_ranks(b0.elapsed_time(b1), world, dev) / 10
            st = {name: sum(ev[i].elapsed_time(ev[i + 1]) for ev in bh_events) / len(bh_events) for i, name in enumerate(bh_path.STAGES)}
            fl = bh_path.box_head_flops()
            # the library GEMMs on the same operands (what nn.Linear under bf16 autocast would run)
            x1 = bh_path.last["pooled"].view(bh_path.last["pooled"].shape[0], -1)
            for _ in range(2):
                torch.relu(torch.relu(x1 @ bh_path.fc1_w.t()) @ bh_path.fc2_w.t())
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(5):
                torch.relu(torch.relu(x1 @ bh_path.fc1_w.t()) @ bh_path.fc2_w.t())
            c1.record()
            torch.cuda.synchronize(dev)
            lib_ms = c0.elapsed_time(c1) / 5
            bh = {"ms_per_step": bh_ms, "value": world * N * 1e3 / bh_ms, "steps": 10, "stage_ms": st, "s4_flops": fl,
                  "s4_tflops": fl / (st["s4_box_head_fc_fwd"] * 1e-3) / 1e12, "s4_cublas_bf16_ms": lib_ms,
                  "pooled_bytes_bf16": int(x1.numel() * 2),
                  "note": "same step with S4 on the path: osr_roi_align_fwd_bf16 writes the pooled tile in bf16, fc1 (12544->1024) and "
                          "fc2 (1024->1024) + bias + ReLU run on tcgen05 (osr_linear_bf16_fwd) and their output feeds the PLN encoder; the "
                          "FC backward GEMMs are library calls outside the path (fixed pooled gradient as in the headline)"}
            del bh_path, x1
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            bh = {"error": repr(e)[:300]}
        barrier(world)

    # ---- the same step with the PLN encoder as the reference runs it: a plain fp32 nn.Linear (library GEMM, no TF32) ------------
    f32enc = None
    if not infer and not args.quick:
        try:
            import dataclasses
            f_path = RoiPathStep(dataclasses.replace(cfg, encoder_impl="fp32"), dev)
            for _ in range(3):
                f_path.step()
            torch.cuda.synchronize(dev)
            barrier(world)
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(10):
                f_path.step()
            f1.record()
            torch.cuda.synchronize(dev)
            f_ms = max_over_ranks(f0.elapsed_time(f1), world, dev) / 10
            f32enc = {"ms_per_step": f_ms, "value": world * N * 1e3 / f_ms, "steps": 10, "launch": "eager",
                      "note": "encoder = fp32 library GEMM (torch, allow_tf32 off) instead of the tcgen05 kind::tf32 kernel; everything "
                              "else identical; compare with `eager`"}
            del f_path
            torch.cuda.empty_cache()
        except Exception as e:  # noqa: BLE001
            f32enc = {"error": repr(e)[:300]}
        barrier(world)
    _EXTRA["fp32_encoder"] = f32enc

    if infer or args.quick:
        if rank == 0:
            line = _line(args, world, cfg, N, value, ms_step, clocks, None, launches, roof, stages, alt, gbase, gathered, None,
                         eager, graph_err, api, bh)
            print(json.dumps(line), flush=True)
        return

    # ---- end to end: inputs start in pinned host memory every step ----------------------------------
    del path
    torch.cuda.empty_cache()
    e2e_path = RoiPathStep(cfg, dev, host_inputs=True)
    for i in range(3):
        e2e_path.e2e_prefetch(i % 2)
        e2e_path.e2e_step(i % 2)
    torch.cuda.synchronize(dev)
    barrier(world)
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    e2e_path.e2e_prefetch(0)
    for i in range(args.steps):
        if i + 1 < args.steps:
            e2e_path.e2e_prefetch((i + 1) % 2)   # overlaps with step i's compute (second device buffer set)
        e2e_path.e2e_step(i % 2)
    t1.record()
    torch.cuda.synchronize(dev)
    e2e_ms = max_over_ranks(t0.elapsed_time(t1), world, dev) / args.steps
    e2e = {"value": world * N * 1e3 / e2e_ms, "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": e2e_path.h2d_bytes(), "d2h_bytes_per_step": e2e_path.d2h_bytes(),
           "note": "H2D of head outputs + FPN maps from pinned memory overlapped with the previous step's compute "
                   "(double-buffered); PCIe-bound"}
    barrier(world)

    if rank != 0:
        return
    line = _line(args, world, cfg, N, value, ms_step, clocks, e2e, launches, roof, stages, alt, gbase, gathered,
                 cpu_baseline(cfg) if (world == 1 and not args.no_cpu_baseline) else None, eager, graph_err, api, bh)
    print(json.dumps(line), flush=True)


_EXTRA = {}   # optional keys of the JSON line filled in by the measurement blocks


def _line(args, world, cfg, N, value, ms_step, clocks, e2e, launches, roof, stages, alt, gbase, gathered, cpu, eager,
          graph_err, api, bh=None):
    infer = args.config == "cfg4"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.config],
                   "pln_encoder": "tcgen05 kind::tf32 on the fp32 operands, fp32 accumulate (all other arithmetic fp32)",
                   "images_per_gpu": N, "global_batch": world * N, "image_hw": list(cfg.image_hw),
                   "pre_nms_topk_per_level": cfg.pre_nms_topk, "rois_per_image": cfg.rois_per_image,
                   "feature_layout": "channels_last" if cfg.channels_last else "NCHW",
                   "proposal_mode": ("nominal (per-level NMS %.2f + best %d per image: the block at find_top_proposals.py:112-120 "
                                     "switched on)" % (cfg.rpn_nms_thresh, cfg.post_nms_topk)) if infer else
                                    "as_shipped (find_top_proposals.py:112-120 commented out)",
                   "l2": "inputs larger than L2 (FPN maps %.2f GB per step vs 126 MB L2)" % (N * 0.0914),
                   "launch": ("one CUDA graph replay per step" if not (args.config == "cfg3" and world > 1) else "one CUDA graph replay per two steps") if (eager is not None) else "eager launches",
                   "timed_stages": ("S1 proposals + NMS, S3 ROIAlign fwd, S6 ROI-head post-processing (decode + NMS + PLN.inference + "
                                    "classifier NMS); box head / predictor outputs are fixed tensors") if infer else
                                   ("S1 proposals, S2 labelled sampling (osr_match_label on all kept proposals + torch.rand keys + osr_sample_rois: "
                                    "512 RoIs / image, 25 % positive quota, fresh draw every step, no host sync), S3 ROIAlign fwd, "
                                    "S5 encoder+PLN loss fwd/bwd" + (" over the GLOBAL batch (encoder fused with the all-gather of the embeddings over NVLink; loss terms on local rows + one all-reduce of loss and prototype gradient)" if (args.config == "cfg3" and world > 1) else "") +
                                    " on the SAMPLED rows' matcher classes / IoUs, S3 ROIAlign bwd; box-head FC excluded (SURVEY.md 8(d))"),
                   "sampled_rows": _EXTRA.get("sampled_rows"),
                   "side_stream": ("the ROIAlign backward's RoI-only table kernel (issued after S2) and the PLN prototype-gradient launches "
                                   "+ loss reduction (issued after the row launch of S5) run on a side stream = parallel branches of the "
                                   "graph; their time is inside value / ms_per_step / eager but not inside stage_ms") if not infer else None,
                   "parallelism": f"dp{world}"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof,
        "stage_ms": stages, "alt_layout": alt, "gpu_baseline": gbase,
    }
    if eager is not None:
        line["eager"] = eager
    if graph_err is not None:
        line["cuda_graph_error"] = graph_err
    if api is not None:
        line["api_step"] = api
    if bh is not None:
        line["with_box_head"] = bh
    if _EXTRA.get("fp32_encoder") is not None:
        line["fp32_encoder"] = _EXTRA["fp32_encoder"]
    if gathered is not None:
        line["gathered_pln"] = gathered
    if cpu is not None:
        line["cpu_baseline"] = cpu
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nchw", action="store_true",
                    help="feed NCHW-contiguous FPN maps (the reference's layout) instead of channels_last (the north star's "
                         "'coalesced NHWC reads'; what a channels_last backbone emits); the other layout is reported under 'alt_layout'")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--config", default="cfg2", choices=sorted(WORKLOADS),
                    help="BASELINE.json configs[1..4]; cfg2 is the headline (the driver's default)")
    ap.add_argument("--images", type=int, default=None, help="images per GPU (default: the config's)")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of one CUDA graph replay per step")
    ap.add_argument("--no-gather-graph", action="store_true", help="do not time the gathered-PLN step as a CUDA graph")
    ap.add_argument("--quick", action="store_true", help="device-resident measurement only (skip alt layout, API step, e2e, baselines)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, local_rank, world)
    finally:
        if world > 1 and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
