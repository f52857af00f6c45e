"""Generates tests/golden/golden_v2.npz - fixtures for the section 8(f) rows (sampling glue matching, ROI-head inference
decode), produced in the build container (no GPU) by torch 2.11.0 CPU ops driven through the detectron2 v0.6 stand-in
(tests/golden/d2shim.py: pairwise_iou / Matcher / Box2BoxTransform.apply_deltas / Boxes.clip) - nothing comes from oracle/.
(The reference-executed versions of the same rows - label_and_sample_proposals, OpensetFastRCNNOutputLayers.inference -
are in golden_ref_v1.npz.)
Run:  python tests/golden/make_golden_v2.py     (deterministic; commit the .npz with this script)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "openset-rcnn_b200"))

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import d2shim  # noqa: E402
from d2shim import Boxes, pairwise_iou  # noqa: E402
from osr_b200 import synth  # noqa: E402

out = {}
g = torch.Generator().manual_seed(20261018)

# ---- matching: 2 images, proposals = jittered GT + random boxes ---------------------------------------------------
gt_all, cls_all, prop_all = [], [], []
for n, (P, G) in enumerate([(400, 6), (123, 3)]):
    c = torch.rand(G, 2, generator=g) * torch.tensor([1000.0, 600.0])
    wh = torch.rand(G, 2, generator=g) * 300 + 30
    gt = torch.cat((c, c + wh), 1)
    src = gt[torch.randint(0, G, (P // 2,), generator=g)]
    jit = src + (torch.rand(P // 2, 4, generator=g) - 0.5) * 0.5 * (src[:, 2:] - src[:, :2]).repeat(1, 2)
    rnd = synth.make_rois(1, P - P // 2, (800, 1333), seed=70 + n)[0]
    props = torch.cat((jit, rnd, gt[:1]), 0)   # last row: an exact GT duplicate (IoU 1)
    m = pairwise_iou(Boxes(gt), Boxes(props))
    idx, lab = d2shim.Matcher([0.5], [0, 1], allow_low_quality_matches=False)(m)
    out[f"match_gt{n}"], out[f"match_cls{n}"], out[f"match_props{n}"] = gt.numpy(), torch.randint(0, 20, (G,), generator=g).numpy(), props.numpy()
    out[f"match_idx{n}"], out[f"match_lab{n}"] = idx.numpy(), lab.numpy()
    out[f"match_iou{n}"] = m[idx, torch.arange(m.shape[1])].numpy()

# ---- ROI-head inference decode: apply_deltas (weights 10,10,5,5), sqrt(iou * centerness) ---------------------------
R = 300
pb = synth.make_rois(1, R, (800, 1333), seed=90)[0]
deltas = torch.randn(R, 4, generator=g) * torch.tensor([1.5, 1.5, 1.0, 1.0])
deltas[::13, 3] = 40.0
ious = torch.rand(R, 1, generator=g)
ctr = torch.rand(R, generator=g)
dec = d2shim.Box2BoxTransform(weights=(10.0, 10.0, 5.0, 5.0)).apply_deltas(deltas, pb)
b = Boxes(dec.clone()); b.clip((800, 1333))
out["dec_boxes_in"], out["dec_deltas"], out["dec_ious"], out["dec_ctr"] = pb.numpy(), deltas.numpy(), ious.numpy(), ctr.numpy()
out["dec_boxes_clipped"] = b.tensor.numpy()
out["dec_scores"] = torch.sqrt(ious[:, 0] * ctr).numpy()

np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v2.npz"), **out)
print("wrote golden_v2.npz:", {k: v.shape for k, v in out.items()})
