"""Generates tests/golden/golden_ref_v1.npz by EXECUTING THE UNMODIFIED REFERENCE (/root/reference) in this container.

How: ``d2shim.import_reference()`` puts a minimal detectron2-v0.6 / fvcore stand-in into ``sys.modules`` (containers,
Matcher, Box2BoxTransform*, ROIPooler -> the real torchvision ``roi_align``, ``batched_nms`` -> the real torchvision
binary, registries, event storage) and imports the reference files as they lie on disk:

  openset_rcnn/modeling/find_top_proposals.py                      find_top_rpn_proposals            (:22-128)
  openset_rcnn/modeling/proposal_generator/classification_free_rpn.py  ClsFreeRPN.predict_proposals / _decode_proposals (:558-610)
  openset_rcnn/modeling/roi_heads/osrcnn_roi_heads.py              OpensetROIHeads.label_and_sample_proposals (:136-230),
                                                                   _forward_box (:282-329)
  openset_rcnn/modeling/roi_heads/prototype_learning_network.py    PLN.loss / inference / encode      (:117-234)
  openset_rcnn/modeling/roi_heads/osrcnn_fast_rcnn.py              OpensetFastRCNNOutputLayers.inference, fast_rcnn_inference (:45-145, :380-450)
  openset_rcnn/modeling/roi_heads/softmax_classifier.py            SoftMaxClassifier.inference        (:47-168, :287-346)

The reference hard-codes ``device='cuda'`` in PLN / SoftMaxClassifier (prototype_learning_network.py:67,71,85-92;
softmax_classifier.py:165-167,224-242); there is no GPU here and the reference cannot travel to the GPU box, so the
script runs it under a ``TorchFunctionMode`` that rewrites a requested 'cuda' device to 'cpu'.  No reference source is
modified or copied; nothing is imported from ``oracle/`` or from the product package.

Run:  python tests/golden/make_golden_ref.py     (deterministic; commit the .npz with this script)
"""
import math
import os
import sys

import numpy as np
import torch
import torchvision
from torch import nn
from torch.overrides import TorchFunctionMode

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import d2shim  # noqa: E402

REF_ROOT = os.environ.get("OSR_REFERENCE_ROOT", "/root/reference")


class CudaToCpu(TorchFunctionMode):
    """device='cuda' -> 'cpu' for every torch call made by the reference while the mode is active."""

    @staticmethod
    def _is_cuda(d):
        return (isinstance(d, str) and d.startswith("cuda")) or (isinstance(d, torch.device) and d.type == "cuda")

    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = dict(kwargs or {})
        if self._is_cuda(kwargs.get("device", None)):
            kwargs["device"] = "cpu"
        args = tuple("cpu" if self._is_cuda(a) else a for a in args)
        return func(*args, **kwargs)


def main():
    ref = d2shim.import_reference(REF_ROOT)
    from detectron2.structures import Boxes, Instances  # the shim's containers, as the reference sees them
    out = {}
    g = torch.Generator().manual_seed(20261101)

    # ==================================================================================== A. CF-RPN proposal stage
    IMG = (160, 224)
    image_sizes = [(160, 224), (150, 200)]
    N = len(image_sizes)
    strides = [4, 8, 16, 32, 64]
    grids = [(IMG[0] // s, IMG[1] // s) for s in strides[:4]]
    grids.append(((grids[-1][0] - 1) // 2 + 1, (grids[-1][1] - 1) // 2 + 1))
    ag = d2shim.DefaultAnchorGenerator(sizes=[[32], [64], [128], [256], [512]], aspect_ratios=[[1.0]], strides=strides,
                                       offset=0.0)
    anchors = ag([torch.zeros(N, 1, h, w) for (h, w) in grids])
    rpn = ref.classification_free_rpn.ClsFreeRPN(
        in_features=["p2", "p3", "p4", "p5", "p6"], head=nn.Identity(), anchor_generator=ag,
        anchor_matcher=d2shim.Matcher([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True),
        objectness_anchor_matcher=d2shim.Matcher([0.1, 0.3], [0, -1, 1], allow_low_quality_matches=True),
        box2box_transform=d2shim.Box2BoxTransformLinear(normalize_by_size=True),
        batch_size_per_image=256, positive_fraction=0.5, objectness_positive_fraction=1.0,
        pre_nms_topk=(300, 200), post_nms_topk=(300, 200), nms_thresh=(1.0, 1.0), min_box_size=0.0)
    out["rpn_pre_nms_topk"] = np.array([300, 200])
    out["rpn_image_sizes"] = np.array(image_sizes)
    deltas, ctr = [], []
    for l, (h, w) in enumerate(grids):
        n = h * w
        d = torch.rand(N, n, 4, generator=g) * 1.15 + 0.05
        d = torch.where(torch.rand(N, n, 4, generator=g) < 0.10, -d, d)   # ReLU -> 0 -> empty boxes
        c = torch.stack([(torch.randperm(n, generator=g).float() + 0.5) / n for _ in range(N)])  # tie-free
        deltas.append(d)
        ctr.append(c)
        out[f"rpn_anchors{l}"] = anchors[l].tensor.numpy()
        out[f"rpn_deltas{l}"] = d.numpy()
        out[f"rpn_ctr{l}"] = c.numpy()
    props = {}
    for mode in ("train", "eval"):
        rpn.train(mode == "train")
        res = rpn.predict_proposals(anchors, [d.clone() for d in deltas], [c.clone() for c in ctr], image_sizes)
        props[mode] = res
        for n, p in enumerate(res):
            out[f"rpn_{mode}_boxes{n}"] = p.proposal_boxes.tensor.numpy()
            out[f"rpn_{mode}_scores{n}"] = p.objectness_logits.numpy()
    # non-finite inputs: dropped in eval, FloatingPointError in training (find_top_proposals.py:94-104)
    d_bad = [d.clone() for d in deltas]
    c_bad = [c.clone() for c in ctr]
    c_bad[0][0, 17] = 0.99999
    d_bad[0][0, 17, 2] = float("inf")
    c_bad[2][1, 5] = 0.99999
    d_bad[2][1, 5, 0] = float("nan")
    for l in (0, 2):
        out[f"rpn_bad_deltas{l}"] = d_bad[l].numpy()
        out[f"rpn_bad_ctr{l}"] = c_bad[l].numpy()
    rpn.eval()
    res = rpn.predict_proposals(anchors, [d.clone() for d in d_bad], [c.clone() for c in c_bad], image_sizes)
    for n, p in enumerate(res):
        out[f"rpn_bad_eval_boxes{n}"] = p.proposal_boxes.tensor.numpy()
        out[f"rpn_bad_eval_scores{n}"] = p.objectness_logits.numpy()
    rpn.train()
    try:
        rpn.predict_proposals(anchors, d_bad, c_bad, image_sizes)
        raised = ""
    except FloatingPointError as e:
        raised = str(e)
    out["rpn_bad_train_error"] = np.array(raised)
    assert raised

    # ================================================================ B. label_and_sample_proposals (+ C/D heads)
    C, FC, EMB = 6, 64, 256
    NUM_CLASSES, NUM_KNOWN = 81, 20
    perms = []
    real_randperm = torch.randperm

    def recording_randperm(n, *a, **k):
        p = real_randperm(n, *a, **k)
        perms.append(p.clone())
        return p

    def build_heads(opendet_benchmark, num_classes, num_known, dataset_name, alpha, beta, unk_thr, loss_weight):
        torch.manual_seed(77)
        with CudaToCpu():
            heads = ref.osrcnn_roi_heads.OpensetROIHeads(
                box_in_features=["p2", "p3", "p4", "p5"],
                box_pooler=d2shim.ROIPooler(output_size=7, scales=(1 / 4, 1 / 8, 1 / 16, 1 / 32), sampling_ratio=0,
                                            pooler_type="ROIAlignV2"),
                box_head=d2shim.FastRCNNConvFCHead(d2shim.ShapeSpec(channels=C, height=7, width=7), conv_dims=[],
                                                   fc_dims=[FC, FC]),
                box_predictor=ref.osrcnn_fast_rcnn.OpensetFastRCNNOutputLayers(
                    d2shim.ShapeSpec(channels=FC), box2box_transform=d2shim.Box2BoxTransform(weights=(10.0, 10.0, 5.0, 5.0)),
                    num_classes=num_classes, test_objectness_score_thresh=0.05, test_nms_thresh=1.0,
                    test_topk_per_image=1000, mean_type="geometric", cls_agnostic_bbox_reg=True,
                    loss_weight={"loss_box_reg": 0.5, "loss_iou": 0.5}),
                dml=ref.prototype_learning_network.PLN(
                    num_classes=num_classes, num_known_classes=num_known, feature_dim=FC, embedding_dim=EMB,
                    distance_type="COS", reps_per_class=1, alpha=alpha, beta=beta, loss_weight=loss_weight,
                    dataset_name=dataset_name, iou_threshold=0.5, unk_thr=unk_thr, opendet_benchmark=opendet_benchmark),
                softmaxcls=ref.softmax_classifier.SoftMaxClassifier(
                    num_classes=num_classes, num_known_classes=num_known, dataset_name=dataset_name,
                    opendet_benchmark=opendet_benchmark, input_size=FC, known_score_thresh=0.05, known_nms_thresh=0.5,
                    known_topk=50, unknown_score_thresh=0.0, unknown_nms_thresh=0.5, unknown_topk=50,
                    cls_loss_weight=0.9),
                num_classes=num_classes, batch_size_per_image=64, positive_fraction=0.25,
                proposal_matcher=d2shim.Matcher([0.5], [0, 1], allow_low_quality_matches=False), proposal_append_gt=True)
        return heads

    heads = build_heads(True, NUM_CLASSES, NUM_KNOWN, "voc_2007_train", 0.1, 0.9, 0.23, 0.5)
    # ground truth: jittered copies of well-sized training proposals (so that IoU > 0.5 matches exist) + random classes
    targets = []
    for n, p in enumerate(props["train"]):
        b = p.proposal_boxes.tensor
        wh = b[:, 2:] - b[:, :2]
        cand = torch.nonzero((wh.min(dim=1).values > 24) & (wh.max(dim=1).values < 140)).flatten()
        pick = cand[torch.randperm(len(cand), generator=g)[: 4 - n]]
        gt = b[pick] + (torch.rand(len(pick), 4, generator=g) - 0.5) * 6.0
        # the training sets only carry known classes (ids < NUM_KNOWN): SoftMaxClassifier.loss maps anything else to -1
        cls = torch.tensor([3, 17, 11, 7][: len(pick)]) if n == 0 else torch.tensor([19, 0, 5][: len(pick)])
        t = Instances(image_sizes[n])
        t.gt_boxes = Boxes(gt)
        t.gt_classes = cls
        targets.append(t)
        out[f"gt_boxes{n}"], out[f"gt_classes{n}"] = gt.numpy(), cls.numpy()
    torch.manual_seed(1234)
    torch.randperm = recording_randperm
    try:
        sampled = heads.label_and_sample_proposals(props["train"], targets)
    finally:
        torch.randperm = real_randperm
    assert len(perms) == 2 * N
    for i, p in enumerate(perms):
        out[f"sample_perm{i}"] = p.numpy()
    for n, s in enumerate(sampled):
        out[f"sample_boxes{n}"] = s.proposal_boxes.tensor.numpy()
        out[f"sample_logits{n}"] = s.objectness_logits.numpy()
        out[f"sample_gt_classes{n}"] = s.gt_classes.numpy()
        out[f"sample_ious{n}"] = s.ious.numpy()
        out[f"sample_gt_boxes{n}"] = s.gt_boxes.tensor.numpy()
        assert int(((s.gt_classes < NUM_KNOWN) & (s.ious > 0.5)).sum()) > 0, "need known-class foreground rows"

    # ================================================== C. training _forward_box: ROIPooler -> box_head -> PLN.loss
    feats = {}
    for l, (h, w) in enumerate(grids[:4]):
        feats[f"p{l + 2}"] = torch.randn(N, C, h, w, generator=g).requires_grad_(True)
        out[f"feat{l}"] = feats[f"p{l + 2}"].detach().numpy()
    heads.train()
    grab = {}
    heads.box_pooler.register_forward_hook(lambda m, i, o: grab.__setitem__("pooled", o))
    heads.box_head.register_forward_hook(lambda m, i, o: grab.__setitem__("box_features", o))
    with CudaToCpu():
        losses = heads._forward_box(feats, sampled)
        emb, rec, loss_dml = heads.dml.loss(grab["box_features"], sampled)
    assert torch.equal(loss_dml, losses["loss_dml"])
    for k, v in heads.state_dict().items():
        out["w_" + k] = v.detach().clone().numpy()   # clone: the parameters are edited in place below
    out["train_pooled"] = grab["pooled"].detach().numpy()
    out["train_box_features"] = grab["box_features"].detach().numpy()
    out["train_emb"], out["train_rec"] = emb.detach().numpy(), rec.detach().numpy()
    out["train_loss_dml"] = loss_dml.detach().numpy()
    for k, v in losses.items():
        out["train_" + k] = v.detach().numpy()
    params = [heads.dml.representatives, heads.dml.encoder.weight, heads.dml.encoder.bias]
    grads = torch.autograd.grad(loss_dml, [grab["pooled"], grab["box_features"]] + params + [feats[f"p{l}"] for l in (2, 3, 4, 5)],
                                allow_unused=True)
    names = ["pooled", "box_features", "reps", "enc_w", "enc_b", "feat0", "feat1", "feat2", "feat3"]
    for nm, gr in zip(names, grads):
        out["train_grad_" + nm] = (torch.zeros(1) if gr is None else gr).numpy()
    out["train_levels"] = d2shim.assign_boxes_to_levels([s.proposal_boxes for s in sampled], 2, 5, 224, 4).numpy()
    with CudaToCpu():
        out["train_encode"] = heads.dml.encode(grab["box_features"].detach()).detach().numpy()

    # ============================== D. inference _forward_box: decode + objectness + NMS + PLN.inference + classifier
    def run_inference(h, tag, full=True):
        h.eval()
        grab.clear()
        stage = {}
        hooks = [h.box_predictor.register_forward_hook(lambda m, i, o: stage.__setitem__("pred", o)),
                 h.box_head.register_forward_hook(lambda m, i, o: stage.__setitem__("box_features", o))]
        orig_pln_inf = h.dml.inference

        def pln_inf(fg):
            stage["fg_boxes"] = [x.pred_boxes.tensor.clone() for x in fg]
            stage["fg_scores"] = [x.scores.clone() for x in fg]
            stage["fg_feats"] = [x.features.clone() for x in fg]
            r = orig_pln_inf(fg)
            stage["pln_classes"] = [x.pred_classes.clone() for x in r]
            stage["pln_rec"] = [x.features.clone() for x in r]
            return r
        h.dml.inference = pln_inf
        with CudaToCpu(), torch.no_grad():
            res = h._forward_box({k: v.detach() for k, v in feats.items()}, props["eval"])
        for hk in hooks:
            hk.remove()
        h.dml.inference = orig_pln_inf
        if full:   # (the GraspNet-style heads share the box head / predictor weights: only the PLN + classifier outputs differ)
            out[f"{tag}_pred_deltas"] = stage["pred"][0].numpy()
            out[f"{tag}_pred_iou"] = stage["pred"][1].numpy()
            out[f"{tag}_box_features"] = stage["box_features"].numpy()
        for n in range(N):
            if full:
                out[f"{tag}_fg_boxes{n}"] = stage["fg_boxes"][n].numpy()
                out[f"{tag}_fg_scores{n}"] = stage["fg_scores"][n].numpy()
                out[f"{tag}_fg_feats{n}"] = stage["fg_feats"][n].numpy()
                if n == 0:
                    out[f"{tag}_pln_rec{n}"] = stage["pln_rec"][n].numpy()
            else:
                assert np.array_equal(out[f"inf_fg_feats{n}"], stage["fg_feats"][n].numpy())
                assert np.array_equal(out[f"inf_fg_boxes{n}"], stage["fg_boxes"][n].numpy())
            out[f"{tag}_pln_classes{n}"] = stage["pln_classes"][n].numpy()
            out[f"{tag}_final_boxes{n}"] = res[n].pred_boxes.tensor.numpy()
            out[f"{tag}_final_scores{n}"] = res[n].scores.numpy()
            out[f"{tag}_final_classes{n}"] = res[n].pred_classes.numpy()
        return stage, res

    def plant_prototypes(h, num_known):
        """'Checkpoint' values that make the inference case non-trivial: prototypes near some embeddings (so that a part
        of the detections is closer than UNK_THR), a classifier with spread-out probabilities."""
        with torch.no_grad(), CudaToCpu():
            bf = h.box_head(h.box_pooler([feats[f"p{l}"].detach() for l in (2, 3, 4, 5)],
                                         [x.proposal_boxes for x in props["eval"]]))
            e = h.dml.encoder(bf)
            idx = torch.randperm(e.shape[0], generator=g)[:num_known]
            h.dml.representatives.copy_(e[idx] + 0.15 * e[idx].norm(dim=1, keepdim=True) * torch.randn(num_known, EMB, generator=g) / math.sqrt(EMB))
            h.softmaxcls.cls_score.weight.mul_(60.0)
            h.box_predictor.iou_pred.weight.mul_(8.0)
            h.box_predictor.bbox_pred.weight.mul_(150.0)

    plant_prototypes(heads, NUM_KNOWN)
    for k, v in heads.state_dict().items():
        out["winf_" + k] = v.detach().clone().numpy()
    stage, res = run_inference(heads, "inf")
    kn = sum(int((c != 80).sum()) for c in stage["pln_classes"])
    un = sum(int((c == 80).sum()) for c in stage["pln_classes"])
    assert kn > 20 and un > 20, (kn, un)

    # GraspNet-style variant (OPENDET_BENCHMARK False): id_map / class_id tables built from the MetadataCatalog
    meta = ref.graspnet_meta.get_graspnet_instances_meta()
    d2shim.MetadataCatalog.get("graspnet_train").thing_dataset_id_to_contiguous_id = meta["thing_dataset_id_to_contiguous_id"]
    gh = build_heads(False, 88, 28, "graspnet_train", 0.05, 0.95, 0.09, 2.0)
    out["gn_class_id"] = gh.dml.class_id.numpy()
    out["gn_id_map"] = gh.dml.id_map.numpy()
    known_ids = gh.dml.class_id.tolist()
    gn_sampled = []
    for n, s in enumerate(sampled):   # same boxes, GraspNet label space: known contiguous ids / background 88
        q = Instances(s.image_size)
        q.proposal_boxes = s.proposal_boxes
        q.objectness_logits = s.objectness_logits
        cls = s.gt_classes.clone()
        fg = cls != NUM_CLASSES
        new = torch.tensor(known_ids)[(cls.clamp(max=NUM_KNOWN - 1) * 3 + n) % 28]
        q.gt_classes = torch.where(fg, new, torch.full_like(cls, 88))
        q.ious = s.ious
        q.gt_boxes = s.gt_boxes
        gn_sampled.append(q)
        out[f"gn_gt_classes{n}"] = q.gt_classes.numpy()
    with CudaToCpu():
        bf = torch.from_numpy(out["train_box_features"]).requires_grad_(True)
        emb, rec, loss = gh.dml.loss(bf, gn_sampled)
    for k, v in gh.dml.state_dict().items():
        out["gn_w_" + k] = v.detach().clone().numpy()
    out["gn_loss_dml"] = loss.detach().numpy()
    ge, gr = torch.autograd.grad(loss, [bf, gh.dml.representatives])
    out["gn_grad_box_features"], out["gn_grad_reps"] = ge.numpy(), gr.numpy()
    plant_prototypes(gh, 28)
    for k, v in gh.state_dict().items():
        if k.startswith(("dml.", "softmaxcls.")):
            out["gninf_" + k] = v.detach().clone().numpy()
        else:
            assert np.array_equal(out["winf_" + k], v.detach().numpy()), k
    stage, res = run_inference(gh, "gninf", full=False)
    kn = sum(int((c != 1000).sum()) for c in stage["pln_classes"])
    assert kn > 10, kn

    out["versions"] = np.array(f"torch {torch.__version__} torchvision {torchvision.__version__}")
    path = os.path.join(HERE, "golden_ref_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
