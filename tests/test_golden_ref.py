"""Parity pinned to the reference's OWN code: ``tests/golden/golden_ref_v1.npz`` was produced by executing the unmodified
/root/reference Python (``tests/golden/make_golden_ref.py`` + the detectron2 stand-in ``tests/golden/d2shim.py``; no
fixture value comes from ``oracle/`` or from the product).  The oracle must reproduce it on the CPU; the CUDA path
(through the C ABI) must reproduce it on the GPU: bit-exact for proposals, sampling, labels, levels and kept sets,
stated fp32 tolerances for ROIAlign features / PLN loss / gradients / exp-decoded boxes."""
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import pln as opln, rcnn_inference as oinf, roi_align as ora, rpn as orpn, sampling as osamp
from oracle import structures as ost
from osr_b200 import synth

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_ref_v1.npz"))
N = 2
IMAGE_SIZES = [tuple(int(v) for v in r) for r in G["rpn_image_sizes"]]
K_TRAIN, K_EVAL = (int(v) for v in G["rpn_pre_nms_topk"])
SCALES = (1 / 4, 1 / 8, 1 / 16, 1 / 32)
VOC = dict(num_known_classes=20, alpha=0.1, beta=0.9, loss_weight=0.5, iou_threshold=0.5)
GN = dict(num_known_classes=28, alpha=0.05, beta=0.95, loss_weight=2.0, iou_threshold=0.5)


def t(name, dev="cpu"):
    return torch.from_numpy(G[name]).to(dev)


def rpn_inputs(dev="cpu", bad=False):
    anchors = [t(f"rpn_anchors{l}", dev) for l in range(5)]
    deltas = [t(f"rpn_bad_deltas{l}" if bad and l in (0, 2) else f"rpn_deltas{l}", dev) for l in range(5)]
    ctr = [t(f"rpn_bad_ctr{l}" if bad and l in (0, 2) else f"rpn_ctr{l}", dev) for l in range(5)]
    return anchors, deltas, ctr


def proposals_of(mod, tag, dev="cpu"):
    out = []
    for n in range(N):
        p = mod.Instances(IMAGE_SIZES[n])
        p.set("proposal_boxes", mod.Boxes(t(f"rpn_{tag}_boxes{n}", dev)))
        p.set("objectness_logits", t(f"rpn_{tag}_scores{n}", dev))
        out.append(p)
    return out


def targets_of(mod, dev="cpu"):
    out = []
    for n in range(N):
        q = mod.Instances(IMAGE_SIZES[n])
        q.set("gt_boxes", mod.Boxes(t(f"gt_boxes{n}", dev)))
        q.set("gt_classes", t(f"gt_classes{n}", dev))
        out.append(q)
    return out


def sampled_of(mod, dev="cpu", classes_key="sample_gt_classes"):
    out = []
    for n in range(N):
        q = mod.Instances(IMAGE_SIZES[n])
        q.set("proposal_boxes", mod.Boxes(t(f"sample_boxes{n}", dev)))
        q.set("objectness_logits", t(f"sample_logits{n}", dev))
        q.set("gt_classes", t(f"{classes_key}{n}", dev))
        q.set("ious", t(f"sample_ious{n}", dev))
        out.append(q)
    return out


def replay_randperm(dev="cpu"):
    it = iter(range(2 * N))

    def rp(n):
        p = t(f"sample_perm{next(it)}", dev)
        assert p.numel() == n
        return p
    return rp


def check_sampled(res):
    for n, s in enumerate(res):
        assert torch.equal(s.get("proposal_boxes").tensor.cpu(), t(f"sample_boxes{n}"))
        assert torch.equal(s.get("objectness_logits").cpu(), t(f"sample_logits{n}"))
        assert torch.equal(s.get("gt_classes").cpu(), t(f"sample_gt_classes{n}"))
        assert torch.equal(s.get("ious").cpu(), t(f"sample_ious{n}"))
        assert torch.equal(s.get("gt_boxes").tensor.cpu(), t(f"sample_gt_boxes{n}"))


def box_head(x, prefix, dev="cpu"):
    """detectron2 FastRCNNConvFCHead (flatten, fc1, relu, fc2, relu) with the fixture's weights."""
    x = torch.flatten(x, 1)
    for k in (1, 2):
        x = F.relu(F.linear(x, t(f"{prefix}box_head.fc{k}.weight", dev), t(f"{prefix}box_head.fc{k}.bias", dev)))
    return x


# ================================================================================================ CPU: the oracle
def test_fixture_was_not_generated_from_the_oracle():
    here = os.path.dirname(os.path.abspath(__file__))
    for f in ("make_golden_ref.py", "d2shim.py", "make_golden.py", "make_golden_v2.py"):
        src = open(os.path.join(here, "golden", f)).read()
        assert not re.search(r"^\s*(import oracle|from oracle[ .])", src, flags=re.M), f
        if f in ("make_golden_ref.py", "d2shim.py"):
            assert "osr_b200" not in src, f


def test_anchor_generators_match_reference_run():
    grids = [(40, 56), (20, 28), (10, 14), (5, 7), (3, 4)]
    for l, (a, b) in enumerate(zip(orpn.generate_anchors(grids, synth.RPN_STRIDES, synth.RPN_SIZES),
                                   synth.make_anchors(grids))):
        assert torch.equal(a.tensor, t(f"rpn_anchors{l}"))
        assert torch.equal(b, t(f"rpn_anchors{l}"))


@pytest.mark.parametrize("topk_impl", ["torch", "stable"])
def test_oracle_rpn_matches_reference(topk_impl):
    anchors, deltas, ctr = rpn_inputs()
    ab = [ost.Boxes(a) for a in anchors]
    for tag, k, training in (("train", K_TRAIN, True), ("eval", K_EVAL, False)):
        res = orpn.predict_proposals(ab, deltas, ctr, IMAGE_SIZES, pre_nms_topk=k, post_nms_topk=k, training=training,
                                     topk_impl=topk_impl)
        for n, p in enumerate(res):
            assert torch.equal(p.proposal_boxes.tensor, t(f"rpn_{tag}_boxes{n}"))
            assert torch.equal(p.objectness_logits, t(f"rpn_{tag}_scores{n}"))
    _, d_bad, c_bad = rpn_inputs(bad=True)
    res = orpn.predict_proposals(ab, d_bad, c_bad, IMAGE_SIZES, pre_nms_topk=K_EVAL, post_nms_topk=K_EVAL, training=False,
                                 topk_impl=topk_impl)
    for n, p in enumerate(res):
        assert torch.equal(p.proposal_boxes.tensor, t(f"rpn_bad_eval_boxes{n}"))
        assert torch.equal(p.objectness_logits, t(f"rpn_bad_eval_scores{n}"))
    with pytest.raises(FloatingPointError) as e:
        orpn.predict_proposals(ab, d_bad, c_bad, IMAGE_SIZES, pre_nms_topk=K_TRAIN, post_nms_topk=K_TRAIN, training=True)
    assert str(e.value) == str(G["rpn_bad_train_error"])


def test_oracle_sampling_matches_reference():
    res = osamp.label_and_sample_proposals(proposals_of(ost, "train"), targets_of(ost), num_classes=81,
                                           batch_size_per_image=64, positive_fraction=0.25, randperm=replay_randperm())
    check_sampled(res)


def test_oracle_training_forward_matches_reference():
    feats = [t(f"feat{l}").requires_grad_(True) for l in range(4)]
    boxes = [ost.Boxes(t(f"sample_boxes{n}")) for n in range(N)]
    pooler = ora.ROIPooler(7, SCALES, 0)
    pooled = pooler.forward(feats, boxes)
    assert torch.equal(pooler.level_assignments(boxes), t("train_levels"))
    assert torch.equal(pooled.detach(), t("train_pooled"))
    bf = box_head(pooled, "w_")
    torch.testing.assert_close(bf.detach(), t("train_box_features"), rtol=0, atol=0)
    reps = t("w_dml.representatives").requires_grad_(True)
    enc_w = t("w_dml.encoder.weight").requires_grad_(True)
    enc_b = t("w_dml.encoder.bias").requires_grad_(True)
    gt = torch.cat([t(f"sample_gt_classes{n}") for n in range(N)])
    iou = torch.cat([t(f"sample_ious{n}") for n in range(N)])
    emb, rec, loss = opln.pln_loss(bf, enc_w, enc_b, t("w_dml.decoder.weight"), t("w_dml.decoder.bias"), reps, gt, iou, **VOC)
    assert torch.equal(emb.detach(), t("train_emb")) and torch.equal(rec.detach(), t("train_rec"))
    torch.testing.assert_close(loss.detach(), t("train_loss_dml"), rtol=1e-6, atol=0)
    grads = torch.autograd.grad(loss, [pooled, bf, reps, enc_w, enc_b] + feats, allow_unused=True)
    for nm, g in zip(["pooled", "box_features", "reps", "enc_w", "enc_b", "feat0", "feat1", "feat2", "feat3"], grads):
        ref = t("train_grad_" + nm)
        if g is None:
            assert ref.numel() == 1 and float(ref) == 0.0
        else:
            torch.testing.assert_close(g, ref, rtol=1e-5, atol=1e-9)
    enc = F.normalize(F.linear(bf.detach(), enc_w.detach(), enc_b.detach()))
    assert torch.equal(enc, t("train_encode"))
    # closed-form gradient (the formula the CUDA backward implements) against the reference's autograd
    ge, gr = opln.pln_loss_grad_closed_form(t("train_emb").double(), t("w_dml.representatives").double(), gt, iou.double(), **VOC)
    ge_ref = torch.autograd.grad(opln.pln_loss_from_emb((e := t("train_emb").requires_grad_(True)), reps.detach(), gt, iou, **VOC), e)[0]
    torch.testing.assert_close(ge.float(), ge_ref, rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(gr.float(), t("train_grad_reps"), rtol=1e-4, atol=1e-7)


def test_oracle_graspnet_label_space_matches_reference():
    bf = t("train_box_features").requires_grad_(True)
    reps = t("gn_w_representatives").requires_grad_(True)
    gt = torch.cat([t(f"gn_gt_classes{n}") for n in range(N)])
    iou = torch.cat([t(f"sample_ious{n}") for n in range(N)])
    emb, rec, loss = opln.pln_loss(bf, t("gn_w_encoder.weight"), t("gn_w_encoder.bias"), t("gn_w_decoder.weight"),
                                   t("gn_w_decoder.bias"), reps, gt, iou, id_map=t("gn_id_map"), **GN)
    torch.testing.assert_close(loss.detach(), t("gn_loss_dml"), rtol=1e-6, atol=0)
    ge, gr = torch.autograd.grad(loss, [bf, reps])
    torch.testing.assert_close(ge, t("gn_grad_box_features"), rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(gr, t("gn_grad_reps"), rtol=1e-5, atol=1e-9)


def _oracle_inference(prefix, tag):
    props = proposals_of(ost, "eval")
    bf = t("inf_box_features")
    res, _ = oinf.inference((t("inf_pred_deltas"), t("inf_pred_iou")), props, bf, score_thresh=0.05, nms_thresh=1.0,
                            topk_per_image=1000)
    return res


def test_oracle_inference_matches_reference():
    feats = [t(f"feat{l}") for l in range(4)]
    props = proposals_of(ost, "eval")
    pooled = ora.ROIPooler(7, SCALES, 0).forward(feats, [p.get("proposal_boxes") for p in props])
    bf = box_head(pooled, "winf_")
    assert torch.equal(bf, t("inf_box_features"))
    deltas = F.linear(bf, t("winf_box_predictor.bbox_pred.weight"), t("winf_box_predictor.bbox_pred.bias"))
    iou = torch.sigmoid(F.linear(bf, t("winf_box_predictor.iou_pred.weight"), t("winf_box_predictor.iou_pred.bias")))
    assert torch.equal(deltas, t("inf_pred_deltas")) and torch.equal(iou, t("inf_pred_iou"))
    res = _oracle_inference("winf_", "inf")
    for n, r in enumerate(res):
        assert torch.equal(r.get("pred_boxes").tensor, t(f"inf_fg_boxes{n}"))
        assert torch.equal(r.get("scores"), t(f"inf_fg_scores{n}"))
        assert torch.equal(r.get("features"), t(f"inf_fg_feats{n}"))
    for tag, pre, kw, unknown_id, class_id in (("inf", "winf_dml.", dict(num_known_classes=20, unk_thr=0.23), 80, None),
                                               ("gninf", "gninf_dml.", dict(num_known_classes=28, unk_thr=0.09), 1000,
                                                t("gn_class_id"))):
        res = _oracle_inference("winf_", "inf")
        for n, r in enumerate(res):
            rec, cls = opln.pln_inference(r.get("features"), t(pre + "encoder.weight"), t(pre + "encoder.bias"),
                                          t(pre + "decoder.weight"), t(pre + "decoder.bias"), t(pre + "representatives"),
                                          unknown_id=unknown_id, class_id=class_id, **kw)
            assert torch.equal(cls, t(f"{tag}_pln_classes{n}"))
            if tag == "inf" and n == 0:
                assert torch.equal(rec, t("inf_pln_rec0"))
            r.set("features", rec)
            r.set("pred_classes", cls)
        spre = pre.replace("dml.", "softmaxcls.")
        w, b = t(spre + "cls_score.weight"), t(spre + "cls_score.bias")
        final = oinf.softmax_classifier_inference(res, lambda x: F.linear(x, w, b), unknown_id=unknown_id,
                                                  known_score_thresh=0.05, known_nms_thresh=0.5, known_topk=50,
                                                  unknown_score_thresh=0.0, unknown_nms_thresh=0.5, unknown_topk=50,
                                                  class_id=class_id)
        for n, r in enumerate(final):
            assert torch.equal(r.get("pred_boxes").tensor, t(f"{tag}_final_boxes{n}"))
            assert torch.equal(r.get("scores"), t(f"{tag}_final_scores{n}"))
            assert torch.equal(r.get("pred_classes"), t(f"{tag}_final_classes{n}"))


# ================================================================================================ GPU: the CUDA path
@pytest.mark.gpu
def test_gpu_rpn_matches_reference():
    from osr_b200 import proposals as P
    anchors, deltas, ctr = rpn_inputs("cuda")
    for tag, k, training in (("train", K_TRAIN, True), ("eval", K_EVAL, False)):
        res = P.predict_proposals(anchors, deltas, ctr, IMAGE_SIZES, pre_nms_topk=k, post_nms_topk=k, training=training)
        for n, p in enumerate(res):
            assert torch.equal(p.proposal_boxes.tensor.cpu(), t(f"rpn_{tag}_boxes{n}"))
            assert torch.equal(p.objectness_logits.cpu(), t(f"rpn_{tag}_scores{n}"))
    # the free function on pre-decoded boxes (find_top_proposals.py:22)
    dec = orpn.decode_proposals([ost.Boxes(a.cpu()) for a in anchors], [d.cpu() for d in deltas])
    res = P.find_top_rpn_proposals([d.cuda() for d in dec], ctr, IMAGE_SIZES, 1.0, K_TRAIN, K_TRAIN, 0.0, True)
    for n, p in enumerate(res):
        assert torch.equal(p.proposal_boxes.tensor.cpu(), t(f"rpn_train_boxes{n}"))
    _, d_bad, c_bad = rpn_inputs("cuda", bad=True)
    res = P.predict_proposals(anchors, d_bad, c_bad, IMAGE_SIZES, pre_nms_topk=K_EVAL, post_nms_topk=K_EVAL, training=False)
    for n, p in enumerate(res):
        assert torch.equal(p.proposal_boxes.tensor.cpu(), t(f"rpn_bad_eval_boxes{n}"))
        assert torch.equal(p.objectness_logits.cpu(), t(f"rpn_bad_eval_scores{n}"))
    with pytest.raises(FloatingPointError) as e:
        P.predict_proposals(anchors, d_bad, c_bad, IMAGE_SIZES, pre_nms_topk=K_TRAIN, post_nms_topk=K_TRAIN, training=True)
    assert str(e.value) == str(G["rpn_bad_train_error"])


@pytest.mark.gpu
def test_gpu_sampling_matches_reference():
    from osr_b200 import sampling as S, structures as st
    res = S.label_and_sample_proposals(proposals_of(st, "train", "cuda"), targets_of(st, "cuda"), num_classes=81,
                                       batch_size_per_image=64, positive_fraction=0.25, randperm=replay_randperm("cuda"))
    check_sampled(res)


@pytest.mark.gpu
def test_gpu_keyed_sampling_kernel_matches_reference():
    """The one-launch sampler (osr_sample_rois, the default path of label_and_sample_proposals) against the unmodified
    reference's run: the random keys are built so that their arg-sort over an image's positives / negatives is the
    permutation the reference drew (fixture ``sample_perm*``); the sampled Instances must then be the reference's, bit for bit."""
    from osr_b200 import sampling as S, structures as st
    keys = []
    for n in range(N):
        boxes = torch.cat((t(f"rpn_train_boxes{n}"), t(f"gt_boxes{n}")))          # proposals + appended ground truth
        m = ost.pairwise_iou(ost.Boxes(t(f"gt_boxes{n}")), ost.Boxes(boxes))
        idx, lab = osamp.matcher(m, 0.5)
        cls = t(f"gt_classes{n}")[idx].clone()
        cls[lab == 0] = 81
        k = torch.full((boxes.shape[0],), 2.0)
        for rows, perm in ((torch.nonzero(cls != 81).flatten(), t(f"sample_perm{2 * n}")),
                           (torch.nonzero(cls == 81).flatten(), t(f"sample_perm{2 * n + 1}"))):
            assert perm.numel() == rows.numel()
            k[rows[perm]] = (torch.arange(rows.numel(), dtype=torch.float32) + 0.5) / max(rows.numel(), 1)
        keys.append(k)
    res = S.label_and_sample_proposals(proposals_of(st, "train", "cuda"), targets_of(st, "cuda"), num_classes=81,
                                       batch_size_per_image=64, positive_fraction=0.25, keys=keys)
    check_sampled(res)


@pytest.mark.gpu
@pytest.mark.parametrize("channels_last", [False, True])
def test_gpu_training_forward_matches_reference(channels_last):
    from osr_b200 import structures as st
    from osr_b200.pln import PLN
    from osr_b200.poolers import ROIPooler
    fmt = torch.channels_last if channels_last else torch.contiguous_format
    feats = [t(f"feat{l}", "cuda").contiguous(memory_format=fmt).requires_grad_(True) for l in range(4)]
    sampled = sampled_of(st, "cuda")
    pooler = ROIPooler(7, SCALES, 0, "ROIAlignV2")
    pooled, lvl = pooler.forward_with_levels(feats, [s.get("proposal_boxes") for s in sampled])
    assert torch.equal(lvl.cpu().long(), t("train_levels"))
    torch.testing.assert_close(pooled.detach().cpu(), t("train_pooled"), rtol=1e-5, atol=1e-4)
    bf = box_head(pooled, "w_", "cuda")
    torch.testing.assert_close(bf.detach().cpu(), t("train_box_features"), rtol=1e-4, atol=1e-4)
    pln = PLN(81, 20, 64, 256, "COS", 1, 0.1, 0.9, 0.5, "voc_2007_train", 0.5, 0.23, True)
    pln.load_state_dict({k[len("w_dml."):]: t(k, "cuda") for k in G.files if k.startswith("w_dml.")})
    emb, rec, loss = pln.loss(bf, sampled)
    torch.testing.assert_close(emb.detach().cpu(), t("train_emb"), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rec.detach().cpu(), t("train_rec"), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(loss.detach().cpu(), t("train_loss_dml"), rtol=1e-5, atol=1e-7)
    grads = torch.autograd.grad(loss, [pooled, bf, pln.representatives, pln.encoder.weight] + feats[:2])
    for nm, g in zip(["pooled", "box_features", "reps", "enc_w", "feat0", "feat1"], grads):
        ref = t("train_grad_" + nm)
        torch.testing.assert_close(g.cpu(), ref, rtol=1e-4, atol=1e-4 * float(ref.abs().max()))
    torch.testing.assert_close(pln.encode(bf.detach()).cpu(), t("train_encode"), rtol=1e-4, atol=1e-6)


@pytest.mark.gpu
def test_gpu_graspnet_label_space_matches_reference():
    from osr_b200 import structures as st
    from osr_b200.pln import PLN
    pln = PLN(88, 28, 64, 256, "COS", 1, 0.05, 0.95, 2.0, "graspnet_train", 0.5, 0.09, False,
              known_class_ids=[int(v) for v in G["gn_class_id"]])
    assert torch.equal(pln.id_map.cpu(), t("gn_id_map")) and torch.equal(pln.class_id.cpu(), t("gn_class_id"))
    pln.load_state_dict({k[len("gn_w_"):]: t(k, "cuda") for k in G.files if k.startswith("gn_w_")})
    bf = t("train_box_features", "cuda").requires_grad_(True)
    _, _, loss = pln.loss(bf, sampled_of(st, "cuda", classes_key="gn_gt_classes"))
    torch.testing.assert_close(loss.detach().cpu(), t("gn_loss_dml"), rtol=1e-5, atol=1e-7)
    ge, gr = torch.autograd.grad(loss, [bf, pln.representatives])
    torch.testing.assert_close(ge.cpu(), t("gn_grad_box_features"), rtol=1e-4, atol=1e-4 * float(t("gn_grad_box_features").abs().max()))
    torch.testing.assert_close(gr.cpu(), t("gn_grad_reps"), rtol=1e-4, atol=1e-4 * float(t("gn_grad_reps").abs().max()))


@pytest.mark.gpu
def test_gpu_inference_matches_reference():
    from osr_b200 import inference as I, structures as st
    from osr_b200.pln import PLN
    props = proposals_of(st, "eval", "cuda")
    res, _ = I.inference((t("inf_pred_deltas", "cuda"), t("inf_pred_iou", "cuda")), props, t("inf_box_features", "cuda"),
                         score_thresh=0.05, nms_thresh=1.0, topk_per_image=1000)
    for n, r in enumerate(res):
        # kept set and order are exact; scores sqrt(iou * centerness) to an ulp or two (the CPU reference rounds the
        # product and the square root separately); the decoded boxes go through expf, whose CUDA and glibc versions
        # differ by an ulp: 1e-3 px
        assert r.get("scores").shape == t(f"inf_fg_scores{n}").shape
        torch.testing.assert_close(r.get("scores").cpu(), t(f"inf_fg_scores{n}"), rtol=1e-6, atol=0)
        assert torch.equal(r.get("features").cpu(), t(f"inf_fg_feats{n}"))
        torch.testing.assert_close(r.get("pred_boxes").tensor.cpu(), t(f"inf_fg_boxes{n}"), rtol=0, atol=1e-3)
    for tag, pre, ctor, unknown_id in (
            ("inf", "winf_dml.", dict(num_classes=81, num_known_classes=20, unk_thr=0.23, opendet_benchmark=True), 80),
            ("gninf", "gninf_dml.", dict(num_classes=88, num_known_classes=28, unk_thr=0.09, opendet_benchmark=False,
                                         known_class_ids=[int(v) for v in G["gn_class_id"]]), 1000)):
        pln = PLN(feature_dim=64, embedding_dim=256, distance_type="COS", reps_per_class=1, alpha=0.1, beta=0.9,
                  loss_weight=0.5, **ctor)
        pln.load_state_dict({k[len(pre):]: t(k, "cuda") for k in G.files if k.startswith(pre)})
        fg = []
        for n in range(N):   # the reference's own intermediate detections as input: isolates PLN.inference
            q = st.Instances(IMAGE_SIZES[n])
            q.set("pred_boxes", st.Boxes(t(f"inf_fg_boxes{n}", "cuda")))
            q.set("scores", t(f"inf_fg_scores{n}", "cuda"))
            q.set("features", t(f"inf_fg_feats{n}", "cuda"))
            fg.append(q)
        with torch.no_grad():
            out = pln.inference(fg)
        # nearest-prototype classes: exact except where the reference's own distance is within 1e-5 of UNK_THR or of the
        # runner-up prototype (fp32 GEMM rounding differs between cuBLAS and the CPU)
        for n, r in enumerate(out):
            ref = t(f"{tag}_pln_classes{n}")
            got = r.get("pred_classes").cpu()
            e = F.normalize(F.linear(t(f"inf_fg_feats{n}"), t(pre + "encoder.weight"), t(pre + "encoder.bias")))
            d = 1.0 - e @ F.normalize(t(pre + "representatives")).t()
            top2 = torch.topk(d, 2, dim=1, largest=False).values
            safe = ((top2[:, 0] - ctor["unk_thr"]).abs() > 1e-5) & ((top2[:, 1] - top2[:, 0]) > 1e-5)
            assert safe.float().mean() > 0.95
            assert torch.equal(got[safe], ref[safe])
            if tag == "inf" and n == 0:
                torch.testing.assert_close(r.get("features").cpu(), t("inf_pln_rec0"), rtol=1e-4, atol=1e-5)
        # SoftMaxClassifier.inference on the reference's own PLN output
        for n in range(N):
            out[n].set("pred_classes", t(f"{tag}_pln_classes{n}", "cuda"))
        spre = pre.replace("dml.", "softmaxcls.")
        w, b = t(spre + "cls_score.weight", "cuda"), t(spre + "cls_score.bias", "cuda")
        final = I.softmax_classifier_inference(out, lambda x: F.linear(x, w, b), unknown_id=unknown_id,
                                               known_score_thresh=0.05, known_nms_thresh=0.5, known_topk=50,
                                               unknown_score_thresh=0.0, unknown_nms_thresh=0.5, unknown_topk=50,
                                               class_id=None if unknown_id == 80 else t("gn_class_id", "cuda"))
        for n, r in enumerate(final):
            assert torch.equal(r.get("pred_classes").cpu(), t(f"{tag}_final_classes{n}"))
            torch.testing.assert_close(r.get("scores").cpu(), t(f"{tag}_final_scores{n}"), rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(r.get("pred_boxes").tensor.cpu(), t(f"{tag}_final_boxes{n}"), rtol=0, atol=1e-3)
