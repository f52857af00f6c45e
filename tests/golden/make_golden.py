"""Generates tests/golden/golden_v1.npz - small input/output vectors for every row of the hot path, produced in the
build container (no GPU) by the REAL third-party binaries the reference bottoms out in:
  torchvision 0.26.0+cu128 (ops.roi_align aligned=True, ops.nms, ops.boxes.batched_nms) and torch 2.11.0 (topk, mm,
  autograd), driven through the UNMODIFIED reference code where the reference has code for the row
  (ClsFreeRPN.predict_proposals, PLN.loss - imported from /root/reference with the detectron2 stand-in d2shim.py) and
  through the stand-in's detectron2 v0.6 glue (ROIPooler, batched_nms) elsewhere.  Nothing comes from oracle/.
The `nominal` proposal mode is the block the reference ships commented out (find_top_proposals.py:112-120): the
generator re-enables exactly those lines in memory (the file on disk is untouched, nothing is copied into the repo).
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

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import d2shim  # noqa: E402
from make_golden_ref import CudaToCpu  # noqa: E402
from osr_b200 import synth  # noqa: E402  (seeded synthetic INPUTS only)

ref = d2shim.import_reference()
Boxes = d2shim.Boxes


def make_rpn(pre, post, thr, module=None):
    m = module or ref.classification_free_rpn
    return m.ClsFreeRPN(in_features=["p2", "p3", "p4", "p5", "p6"], head=torch.nn.Identity(), anchor_generator=None,
                        anchor_matcher=None, objectness_anchor_matcher=None,
                        box2box_transform=d2shim.Box2BoxTransformLinear(normalize_by_size=True), batch_size_per_image=256,
                        positive_fraction=0.5, objectness_positive_fraction=1.0, pre_nms_topk=(pre, pre),
                        post_nms_topk=(post, post), nms_thresh=(thr, thr), min_box_size=0.0).eval()


def nominal_find_top_rpn_proposals():
    """find_top_rpn_proposals with the reference's commented-out NMS block (:112-120, :123-124) switched back on."""
    path = os.path.join("/root/reference", "openset_rcnn", "modeling", "find_top_proposals.py")
    src = open(path).read()
    for a, b in (("        # keep = batched_nms(", "        keep = batched_nms("),
                 ("        # keep = keep[:post_nms_topk]", "        keep = keep[:post_nms_topk]"),
                 ("        # res.proposal_boxes = boxes[keep]", "        res.proposal_boxes = boxes[keep]"),
                 ("        # res.objectness_logits = scores_per_img[keep]", "        res.objectness_logits = scores_per_img[keep]"),
                 ("        res.proposal_boxes = boxes\n", "\n"), ("        res.objectness_logits = scores_per_img\n", "\n")):
        assert src.count(a) == 1, a
        src = src.replace(a, b)
    ns = {}
    exec(compile(src, path + " [nominal]", "exec"), ns)
    return ns["find_top_rpn_proposals"]

out = {}
g = torch.Generator().manual_seed(20261017)

# ---- ROIAlign / ROIPooler (torchvision CPU kernel + autograd) ------------------------------------------------
hw = (128, 160)
feats = [f.clone().requires_grad_(True) for f in synth.make_features(2, hw, 6, seed=11)]
rois = synth.make_rois(2, 12, hw, seed=12)
rois[0] = torch.cat([rois[0], torch.tensor([[0.0, 0.0, 160.0, 128.0], [10.0, 10.0, 10.5, 10.5], [5.0, 5.0, 5.0, 5.0],
                                            [0.0, 0.0, 112.0, 112.0], [3.0, 100.0, 150.0, 104.0]])])
pooler = d2shim.ROIPooler(7, synth.POOL_SCALES, 0, "ROIAlignV2")
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
out["roi_levels"] = d2shim.assign_boxes_to_levels(boxes, 2, 5, 224, 4).numpy()

# ---- NMS (torchvision CPU kernel) ---------------------------------------------------------------------------
c = torch.rand(400, 2, generator=g) * 200
wh = torch.rand(400, 2, generator=g) * 70 + 4
nb = torch.cat([c - wh / 2, c + wh / 2], 1)
ns = torch.rand(400, generator=g)
ni = torch.randint(0, 3, (400,), generator=g)
out["nms_boxes"], out["nms_scores"], out["nms_idxs"] = nb.numpy(), ns.numpy(), ni.numpy()
for thr in (0.5, 0.7, 1.0):
    out[f"nms_keep_{thr}"] = torchvision.ops.nms(nb, ns, thr).numpy()
out["nms_batched_keep_0.5"] = d2shim.batched_nms(nb, ns, ni, 0.5).numpy()

# ---- CF-RPN proposal stage (the reference's ClsFreeRPN.predict_proposals; ATen topk + elementwise ops) ----------------
ho = synth.make_head_outputs(2, (96, 128), seed=21, mixed_sizes=True)
anch = [Boxes(a) for a in ho.anchors]
props = make_rpn(60, 60, 1.0).predict_proposals(anch, ho.deltas, ho.centerness, ho.image_sizes)
# level ids are not returned by the reference: recover them by running each level alone (the path is per level)
per_level = [make_rpn(60, 60, 1.0).predict_proposals([anch[l]], [ho.deltas[l]], [ho.centerness[l]], ho.image_sizes)
             for l in range(len(anch))]
for n, p in enumerate(props):
    out[f"rpn_boxes{n}"] = p.proposal_boxes.tensor.numpy()
    out[f"rpn_scores{n}"] = p.objectness_logits.numpy()
    lv = torch.cat([torch.full((len(per_level[l][n]),), l, dtype=torch.int64) for l in range(len(anch))])
    assert torch.equal(torch.cat([per_level[l][n].proposal_boxes.tensor for l in range(len(anch))]), p.proposal_boxes.tensor)
    out[f"rpn_levels{n}"] = lv.numpy()
rpn_nom = make_rpn(60, 40, 0.7)
fn = nominal_find_top_rpn_proposals()
with torch.no_grad():
    dec = rpn_nom._decode_proposals(anch, ho.deltas)
    nom = fn(dec, ho.centerness, ho.image_sizes, 0.7, 60, 40, 0.0, False)
for n, p in enumerate(nom):
    out[f"rpn_nominal_boxes{n}"] = p.proposal_boxes.tensor.numpy()
    out[f"rpn_nominal_scores{n}"] = p.objectness_logits.numpy()

# ---- PLN loss + gradients (the reference's PLN.loss with an identity encoder so that emb == the stored input) --------
pi = synth.make_pln_inputs(96, feat_dim=64, emb_dim=256, num_known=20, seed=31)
emb = (pi.roi_features @ pi.enc_w.t()).detach().requires_grad_(True)
with CudaToCpu():
    pln = ref.prototype_learning_network.PLN(num_classes=81, num_known_classes=20, feature_dim=256, embedding_dim=256,
                                             distance_type="COS", reps_per_class=1, alpha=0.1, beta=0.9, loss_weight=0.5,
                                             dataset_name="voc_2007_train", iou_threshold=0.5, unk_thr=0.23,
                                             opendet_benchmark=True)
with torch.no_grad():
    pln.encoder.weight.copy_(torch.eye(256))
    pln.representatives.copy_(pi.reps)
q = d2shim.Instances((1, 1))
q.gt_classes, q.ious = pi.gt_classes, pi.ious
with CudaToCpu():
    e2, _, loss = pln.loss(emb, [q])
assert torch.equal(e2, emb)
ge, gr = torch.autograd.grad(loss, [emb, pln.representatives])
out["pln_emb"], out["pln_reps"] = emb.detach().numpy(), pi.reps.numpy()
out["pln_labels"], out["pln_ious"] = pi.gt_classes.numpy(), pi.ious.numpy()
out["pln_loss"] = loss.detach().numpy()
out["pln_grad_emb"], out["pln_grad_reps"] = ge.numpy(), gr.numpy()

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes;", "torch", torch.__version__, "torchvision", torchvision.__version__)
