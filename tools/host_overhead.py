"""Host-side enqueue time of one RoI-path step vs its GPU time (is the step launch-bound?)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "openset-rcnn_b200"))
import torch
from osr_b200.pipeline import PathConfig, RoiPathStep

path = RoiPathStep(PathConfig(), device="cuda:0")
for _ in range(5):
    path.step()
torch.cuda.synchronize()
n = 50
t0 = time.perf_counter()
for _ in range(n):
    path.step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/n:.3f} ms/step, wall incl. drain {1e3*(t2-t0)/n:.3f} ms/step")
