"""Generates tests/golden/golden_v1.npz - small input/output vectors for every row of the hot path, produced in the
build container (no GPU) by the REAL third-party binaries the reference bottoms out in:
  torchvision 0.26.0+cu128 (ops.roi_align aligned=True, ops.nms, ops.boxes.batched_nms) and torch 2.11.0 (topk, mm,
  autograd), driven through the oracle's restatement of the detectron2 glue.
The reference itself ships no tests or fixtures and cannot be imported here (detectron2 absent).
Run:  python tests/golden/make_golden.py     (deterministic; commit the .npz with this script)
"""
import os
import sys

import numpy as np
import torch
import torchvision

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "openset-rcnn_b200"))

from oracle import nms as onms, pln as opln, roi_align as ora, rpn as orpn  # noqa: E402
from oracle.structures import Boxes  # noqa: E402
from osr_b200 import synth  # noqa: E402

out = {}
g = torch.Generator().manual_seed(20261017)

# ---- ROIAlign / ROIPooler (torchvision CPU kernel + autograd) ------------------------------------------------
hw = (128, 160)
feats = [f.clone().requires_grad_(True) for f in synth.make_features(2, hw, 6, seed=11)]
rois = synth.make_rois(2, 12, hw, seed=12)
rois[0] = torch.cat([rois[0], torch.tensor([[0.0, 0.0, 160.0, 128.0], [10.0, 10.0, 10.5, 10.5], [5.0, 5.0, 5.0, 5.0],
                                            [0.0, 0.0, 112.0, 112.0], [3.0, 100.0, 150.0, 104.0]])])
pooler = ora.ROIPooler(7, synth.POOL_SCALES, 0)
boxes = [Boxes(r) for r in rois]
pooled = pooler.forward(feats, boxes)
gout = torch.randn(pooled.shape, generator=g)
grads = torch.autograd.grad(pooled, feats, gout)
for l in range(4):
    out[f"roi_feat{l}"] = feats[l].detach().numpy()
    out[f"roi_grad{l}"] = grads[l].numpy()
out["roi_rois0"], out["roi_rois1"] = rois[0].numpy(), rois[1].numpy()
out["roi_pooled"] = pooled.detach().numpy()
out["roi_gout"] = gout.numpy()
out["roi_levels"] = pooler.level_assignments(boxes).numpy()

# ---- NMS (torchvision CPU kernel) ---------------------------------------------------------------------------
c = torch.rand(400, 2, generator=g) * 200
wh = torch.rand(400, 2, generator=g) * 70 + 4
nb = torch.cat([c - wh / 2, c + wh / 2], 1)
ns = torch.rand(400, generator=g)
ni = torch.randint(0, 3, (400,), generator=g)
out["nms_boxes"], out["nms_scores"], out["nms_idxs"] = nb.numpy(), ns.numpy(), ni.numpy()
for thr in (0.5, 0.7, 1.0):
    out[f"nms_keep_{thr}"] = torchvision.ops.nms(nb, ns, thr).numpy()
out["nms_batched_keep_0.5"] = onms.batched_nms(nb, ns, ni, 0.5).numpy()

# ---- CF-RPN proposal stage (ATen topk + elementwise ops through the oracle) ------------------------------------
ho = synth.make_head_outputs(2, (96, 128), seed=21, mixed_sizes=True)
props = orpn.predict_proposals([Boxes(a) for a in ho.anchors], ho.deltas, ho.centerness, ho.image_sizes,
                               pre_nms_topk=60, post_nms_topk=60, training=False, topk_impl="torch")
for n, p in enumerate(props):
    out[f"rpn_boxes{n}"] = p.proposal_boxes.tensor.numpy()
    out[f"rpn_scores{n}"] = p.objectness_logits.numpy()
    out[f"rpn_levels{n}"] = p.level_ids.numpy()
nom = orpn.predict_proposals([Boxes(a) for a in ho.anchors], ho.deltas, ho.centerness, ho.image_sizes, nms_thresh=0.7,
                             pre_nms_topk=60, post_nms_topk=40, training=False, mode="nominal", topk_impl="torch")
for n, p in enumerate(nom):
    out[f"rpn_nominal_boxes{n}"] = p.proposal_boxes.tensor.numpy()
    out[f"rpn_nominal_scores{n}"] = p.objectness_logits.numpy()

# ---- PLN loss + gradients (ATen mm + autograd) ------------------------------------------------------------------
pi = synth.make_pln_inputs(96, feat_dim=64, emb_dim=256, num_known=20, seed=31)
emb = (pi.roi_features @ pi.enc_w.t()).detach().requires_grad_(True)
reps = pi.reps.clone().requires_grad_(True)
kw = dict(num_known_classes=20, alpha=0.1, beta=0.9, loss_weight=0.5, iou_threshold=0.5)
loss = opln.pln_loss_from_emb(emb, reps, pi.gt_classes, pi.ious, **kw)
ge, gr = torch.autograd.grad(loss, [emb, reps])
out["pln_emb"], out["pln_reps"] = emb.detach().numpy(), pi.reps.numpy()
out["pln_labels"], out["pln_ious"] = pi.gt_classes.numpy(), pi.ious.numpy()
out["pln_loss"] = loss.detach().numpy()
out["pln_grad_emb"], out["pln_grad_reps"] = ge.numpy(), gr.numpy()

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes;", "torch", torch.__version__, "torchvision", torchvision.__version__)
