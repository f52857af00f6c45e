import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "openset-rcnn_b200")); sys.path.insert(0, ROOT)
import torch
from osr_b200.poolers import ROIPooler
from osr_b200 import synth
from oracle.structures import Boxes
p = ROIPooler(7, synth.POOL_SCALES, 0, "ROIAlignV2")
feats = synth.make_features(1, (320, 480), 16, seed=3, device="cuda:0")
rois = synth.make_rois(1, 8, (320, 480), seed=17)
out = p.forward(feats, [Boxes(r.cuda()) for r in rois])
torch.cuda.synchronize()
print("ok", float(out.abs().sum()))
