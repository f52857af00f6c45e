"""GPU parity of the ROI-head inference post-processing (osr_rcnn_decode_score + segmented NMS) against the oracle's
restatement of osrcnn_fast_rcnn.py run on the same GPU (torch's own ops + torchvision CUDA nms): bit-exact boxes,
scores, kept indices and gathered features."""
import pytest
import torch

from oracle import rcnn_inference as oinf
from oracle.structures import Boxes as OBoxes, Instances as OInstances

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _inputs(lens, seed, hw=(800, 1333), bad=True):
    from osr_b200 import synth
    from osr_b200.structures import Boxes, Instances
    g = torch.Generator().manual_seed(seed)
    ours, ref = [], []
    for n, k in enumerate(lens):
        b = synth.make_rois(1, k, hw, seed=seed + n)[0] if k else torch.empty(0, 4)
        ctr = torch.rand(k, generator=g)
        a = Instances(hw); a.set("proposal_boxes", Boxes(b.to(DEV))); a.set("objectness_logits", ctr.to(DEV))
        c = OInstances(hw); c.set("proposal_boxes", OBoxes(b.to(DEV))); c.set("objectness_logits", ctr.to(DEV))
        ours.append(a); ref.append(c)
    R = sum(lens)
    deltas = torch.randn(R, 4, generator=g) * torch.tensor([1.5, 1.5, 1.0, 1.0])
    deltas[::17, 2] = 30.0          # exercises the scale clamp
    ious = torch.rand(R, 1, generator=g)
    if bad and R > 50:
        deltas[5, 0] = float("nan"); deltas[11, 3] = float("inf"); ious[23, 0] = float("nan"); ious[31, 0] = -0.5
    feats = torch.randn(R, 32, generator=g)
    return ours, ref, (deltas.to(DEV), ious.to(DEV)), feats.to(DEV)


@pytest.mark.parametrize("lens,thr,nms,topk,mean", [
    ([1000, 1000, 873, 1000], 0.05, 0.5, 100, "geometric"),
    ([4273, 4273], 0.0, 1.0, 1000, "geometric"),       # the reference's shipped thresholds (nms 1.0: sort + top-1000)
    ([300, 0, 7], 0.3, 0.7, -1, "arithmetic"),
    ([64], 0.99, 0.5, 100, "geometric"),               # almost everything filtered
])
def test_inference_matches_oracle_on_gpu(lens, thr, nms, topk, mean):
    from osr_b200.inference import inference
    ours_p, ref_p, preds, feats = _inputs(lens, seed=21)
    kw = dict(mean_type=mean, score_thresh=thr, nms_thresh=nms, topk_per_image=topk)
    got, got_idx = inference(preds, ours_p, feats, **kw)
    exp, exp_idx = oinf.inference(preds, ref_p, feats, **kw)
    assert len(got) == len(exp) == len(lens)
    for a, b, ia, ib in zip(got, exp, got_idx, exp_idx):
        assert len(a) == len(b), (len(a), len(b))
        assert torch.equal(a.get("pred_boxes").tensor, b.get("pred_boxes").tensor)
        assert torch.equal(a.get("scores"), b.get("scores"))
        assert torch.equal(a.get("pred_classes"), b.get("pred_classes"))
        assert torch.equal(a.get("features"), b.get("features"))
        assert torch.equal(ia, ib)


def test_decode_close_to_cpu_reference():
    """Same path against the oracle on the CPU: boxes within 1e-4 px (expf / reciprocal-multiply differ by ulps
    between the CPU and CUDA math libraries), identical survivor sets for well-separated scores."""
    from osr_b200.inference import inference
    ours_p, ref_p, preds, feats = _inputs([500, 321], seed=4, bad=False)
    got, _ = inference(preds, ours_p, feats, score_thresh=0.0, nms_thresh=1.0, topk_per_image=-1)
    ref_cpu = []
    for p in ref_p:
        q = OInstances(p.image_size)
        q.set("proposal_boxes", OBoxes(p.get("proposal_boxes").tensor.cpu()))
        q.set("objectness_logits", p.get("objectness_logits").cpu())
        ref_cpu.append(q)
    exp, _ = oinf.inference((preds[0].cpu(), preds[1].cpu()), ref_cpu, feats.cpu(), score_thresh=0.0, nms_thresh=1.0,
                            topk_per_image=-1)
    for a, b in zip(got, exp):
        assert len(a) == len(b)
        torch.testing.assert_close(a.get("pred_boxes").tensor.cpu(), b.get("pred_boxes").tensor, rtol=1e-6, atol=1e-3)
        torch.testing.assert_close(a.get("scores").cpu(), b.get("scores"), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("class_table", [False, True])
def test_softmax_classifier_inference_matches_oracle(class_table):
    """SoftMaxClassifier.inference (known: per-class NMS over K classes; unknown: class-agnostic NMS; unknown first)."""
    from osr_b200.inference import softmax_classifier_inference
    from osr_b200 import synth
    from osr_b200.structures import Boxes, Instances
    torch.manual_seed(0)
    K = 20
    cls = torch.nn.Linear(32, K + 1).to(DEV)
    g = torch.Generator().manual_seed(8)
    ours, ref = [], []
    for n, (k, frac_unk) in enumerate([(300, 0.3), (100, 0.0), (57, 1.0), (0, 0.0)]):
        b = synth.make_rois(1, k, (800, 1333), seed=40 + n)[0] if k else torch.empty(0, 4)
        sc = torch.rand(k, generator=g)
        pc = torch.where(torch.rand(k, generator=g) < frac_unk, torch.full((k,), 80), torch.randint(0, K, (k,), generator=g))
        f = torch.randn(k, 32, generator=g) * 3
        for cont, Bx, Inst in ((ours, Boxes, Instances), (ref, OBoxes, OInstances)):
            i = Inst((800, 1333))
            i.set("pred_boxes", Bx(b.to(DEV))); i.set("scores", sc.to(DEV)); i.set("pred_classes", pc.to(DEV))
            i.set("features", f.to(DEV))
            cont.append(i)
    table = (torch.arange(K, device=DEV) * 3 + 1) if class_table else None
    kw = dict(unknown_id=80, known_score_thresh=0.05, known_nms_thresh=0.5, known_topk=100, unknown_score_thresh=0.2,
              unknown_nms_thresh=0.5, unknown_topk=50, class_id=table)
    with torch.no_grad():
        got = softmax_classifier_inference(ours, cls, **kw)
        exp = oinf.softmax_classifier_inference(ref, cls, **kw)
    for a, b in zip(got, exp):
        assert len(a) == len(b)
        assert torch.equal(a.get("pred_boxes").tensor, b.get("pred_boxes").tensor)
        assert torch.equal(a.get("scores"), b.get("scores"))
        assert torch.equal(a.get("pred_classes"), b.get("pred_classes"))


def test_chained_stages_on_split_views_equal_the_copying_path():
    """inference -> PLN.inference -> SoftMaxClassifier.inference: the first stage hands out split views of its batched
    gathers and the next stages recognise them (cat_rows, no copy).  The results must equal the run in which every
    field is cloned in between (so that the stages take the torch.cat path)."""
    from osr_b200.inference import inference, softmax_classifier_inference
    from osr_b200.pln import PLN
    from osr_b200.structures import Boxes, Instances, cat_rows
    torch.manual_seed(3)
    K = 20
    ours_p, _, preds, _ = _inputs([400, 0, 250, 31], seed=9, bad=True)
    R = 400 + 250 + 31
    feats = torch.relu(torch.randn(R, 64, device=DEV))
    pln = PLN(81, K, 64, 32, "COS", 1, 0.1, 0.9, 0.5, "synthetic", 0.5, 0.23, True, device=DEV).eval()
    with torch.no_grad():
        pln.representatives.copy_(pln.encoder(feats[:K * 7:7]))
    cls = torch.nn.Linear(64, K + 1).to(DEV)
    kw = dict(unknown_id=80, known_score_thresh=0.05, known_nms_thresh=0.5, known_topk=50, unknown_score_thresh=0.0,
              unknown_nms_thresh=0.5, unknown_topk=50)

    def cloned(insts, copy=True):   # PLN.inference writes its fields into the Instances it is given: hand it fresh ones
        out = []
        for x in insts:
            y = Instances(x.image_size)
            for k, v in x.get_fields().items():
                if copy:
                    v = Boxes(v.tensor.clone()) if hasattr(v, "tensor") else v.clone()
                y.set(k, v)
            out.append(y)
        return out

    with torch.no_grad():
        fg, kept = inference(preds, ours_p, feats, score_thresh=0.05, nms_thresh=0.9, topk_per_image=300)
        f_views = [x.get("features") for x in fg]
        assert cat_rows(f_views).data_ptr() == f_views[0].data_ptr()      # the views really are recognised
        a = softmax_classifier_inference(pln.inference(cloned(fg, copy=False)), cls, **kw)          # view path
        b = softmax_classifier_inference(cloned(pln.inference(cloned(fg))), cls, **kw)              # torch.cat path
    assert sum(len(x) for x in a) > 0
    for x, y in zip(a, b):
        assert len(x) == len(y)
        assert torch.equal(x.get("pred_boxes").tensor, y.get("pred_boxes").tensor)
        assert torch.equal(x.get("scores"), y.get("scores"))
        assert torch.equal(x.get("pred_classes"), y.get("pred_classes"))
