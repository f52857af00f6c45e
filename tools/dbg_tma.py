import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "openset-rcnn_b200")); sys.path.insert(0, ROOT)
import torch
from osr_b200 import _lib, synth
from osr_b200.poolers import _feat_levels
lib = _lib.lib()
feats = synth.make_features(1, (320, 480), 16, seed=3, device="cuda:0")
arr, N, Cc = _feat_levels(feats, synth.POOL_SCALES)
out = torch.zeros(2048, device="cuda:0")
fn = lib.osr_debug_tma_selftest
fn.restype = C.c_int
fn.argtypes = [C.POINTER(_lib.FeatLevel), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
for level in (0, 1, 2):
    rc = fn(arr, 4, N, Cc, level, 5, 3, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    f = feats[level]
    exp = f[0, :8, 3:11, 5:37]
    got = out.view(8, 8, 32)[:, :, :exp.shape[2]]
    print("level", level, "rc", rc, "match", bool(torch.equal(got, exp)))
