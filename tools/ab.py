"""A/B timing of kernel variants on the cfg-2 step (device-resident, CUDA events between stages).

usage: python tools/ab.py bwd=0,1,2 [fwd=0,2] [--steps 20] [--images 16] [--nchw]
Each `key=v1,v2,...` sweeps one OSR_TUNE_* switch (include/osr.h) with the others at their defaults; prints one JSON line
per setting with the per-stage milliseconds.  Also checks that every variant of a key produces the same loss / gradients
as the first one to rtol 1e-4 (cheap guard against timing a broken kernel).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "openset-rcnn_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from osr_b200 import _lib  # noqa: E402
from osr_b200.pipeline import PathConfig, RoiPathStep  # noqa: E402


def time_steps(path, steps):
    for _ in range(3):
        path.step()
    torch.cuda.synchronize()
    acc = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    evs = []
    for _ in range(steps):
        path.step(stage_events=True)
        evs.append(path.events)
    e1.record()
    torch.cuda.synchronize()
    acc = {name: sum(ev[i].elapsed_time(ev[i + 1]) for ev in evs) / steps for i, name in enumerate(path.STAGES)}
    acc["step"] = e0.elapsed_time(e1) / steps
    return acc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("sweeps", nargs="*")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--images", type=int, default=16)
    ap.add_argument("--nchw", action="store_true")
    ap.add_argument("--topk", type=int, default=2000)
    ap.add_argument("--rois", type=int, default=512)
    ap.add_argument("--seed", type=int, default=3234)
    a = ap.parse_args()
    path = RoiPathStep(PathConfig(num_images=a.images, channels_last=not a.nchw, seed=a.seed, pre_nms_topk=a.topk,
                                  rois_per_image=a.rois), "cuda:0")
    print(json.dumps({"setting": "defaults", **time_steps(path, a.steps)}), flush=True)
    for sw in a.sweeps:
        key, vals = sw.split("=")
        ref = None
        for v in vals.split(","):
            prev = _lib.set_tuning(key, int(v, 0))
            try:
                t = time_steps(path, a.steps)
                cur = dict(loss=path.last["loss"].detach().clone(), g=[g.clone() for g in path.last["g_feats"]],
                           pooled=path.last["pooled"].detach().clone())
                ok = True
                if ref is None:
                    ref = cur
                else:
                    ok = bool(torch.allclose(cur["loss"], ref["loss"], rtol=1e-4)) and \
                        bool(torch.allclose(cur["pooled"], ref["pooled"], rtol=1e-4, atol=1e-4)) and \
                        all(torch.allclose(x, y, rtol=1e-4, atol=1e-4 * float(y.abs().max())) for x, y in zip(cur["g"], ref["g"]))
                print(json.dumps({"setting": f"{key}={v}", "matches_first": ok, **t}), flush=True)
            finally:
                _lib.set_tuning(key, prev)


if __name__ == "__main__":
    main()
