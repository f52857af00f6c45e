"""Profiling driver: a few device-resident RoI-path steps, nothing else (for `ncu ... python tools/prof_step.py`)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "openset-rcnn_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from osr_b200.pipeline import PathConfig, RoiPathStep  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--images", type=int, default=None, help="images per step (default: the configuration's own, 16 without --config)")
ap.add_argument("--nchw", action="store_true")
ap.add_argument("--config", default=None, help="cfg2 | cfg3 | cfg5 (training path shapes of that configuration)")
ap.add_argument("--box-head", action="store_true", help="S4 on the path (bf16 pooled tile, fc1 / fc2 on tcgen05)")
ap.add_argument("--infer", action="store_true", help="inference path (cfg 4): nominal proposals + NMS, ROIAlign fwd, post-processing")
ap.add_argument("--tune", action="append", default=[], help="key=value of an OSR_TUNE_* switch (bwd, fwd, pln, rpn, nms)")
a = ap.parse_args()
from osr_b200 import _lib  # noqa: E402
for kv in a.tune:
    k, v = kv.split("=")
    _lib.set_tuning(k, int(v))
if a.infer:
    from osr_b200.pipeline import InferencePathStep, make_config
    path = InferencePathStep(make_config("cfg4", num_images=a.images, seed=3234), "cuda:0")
elif a.config:
    from osr_b200.pipeline import make_config
    path = RoiPathStep(make_config(a.config, num_images=a.images, channels_last=not a.nchw, seed=3234), "cuda:0")
else:
    path = RoiPathStep(PathConfig(num_images=a.images or 16, channels_last=not a.nchw, seed=3234, box_head=a.box_head), "cuda:0")
for _ in range(a.steps):
    path.step()
torch.cuda.synchronize()
print("done")
