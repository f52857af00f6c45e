"""GPU parity: osr_nms_segmented (through nms / batched_nms / nominal RPN mode) vs torchvision's CUDA nms
(the kernel the reference runs).  Keep indices must be bit-exact."""
import pytest
import torch
import torchvision

from oracle import nms as onms
from oracle import rpn as orpn
from oracle.structures import Boxes as OBoxes

pytestmark = pytest.mark.gpu


def _rand_boxes(n, seed, spread=800.0, size=120.0, device="cuda:0"):
    g = torch.Generator().manual_seed(seed)
    c = torch.rand(n, 2, generator=g) * spread
    wh = torch.rand(n, 2, generator=g) * size + 2.0
    b = torch.cat([c - wh / 2, c + wh / 2], dim=1)
    s = torch.rand(n, generator=g)
    return b.to(device), s.to(device)


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 1000, 4273, 7323])
@pytest.mark.parametrize("thr", [0.5, 0.7, 1.0])
def test_nms_matches_torchvision_cuda(n, thr):
    from osr_b200.nms import nms
    b, s = _rand_boxes(n, seed=n)
    got = nms(b, s, thr)
    exp = torchvision.ops.nms(b, s, thr)
    assert torch.equal(got, exp)


def test_nms_dense_overlaps_and_long_input():
    from osr_b200.nms import nms
    b, s = _rand_boxes(20000, seed=7, spread=300.0, size=80.0)   # > 16384: torch.sort + presorted kernel path
    assert torch.equal(nms(b, s, 0.5), torchvision.ops.nms(b, s, 0.5))
    b, s = _rand_boxes(6000, seed=8, spread=100.0, size=60.0)     # heavy suppression
    got = nms(b, s, 0.3)
    assert torch.equal(got, torchvision.ops.nms(b, s, 0.3)) and len(got) < 1500


def test_appendix_c4_golden_vectors():
    from osr_b200.nms import nms
    dev = "cuda:0"
    disjoint = torch.tensor([[i * 10.0, 0.0, i * 10.0 + 5.0, 5.0] for i in range(20)], device=dev)
    assert nms(disjoint, torch.ones(20, device=dev), 0.5).tolist() == list(range(20))          # stable for ties
    sc = torch.tensor([.5, .9, .5, .9, .5, .9], device=dev)
    for thr in (0.5, 1.0):
        assert nms(disjoint[:6], sc, thr).tolist() == [1, 3, 5, 0, 2, 4]
    same = torch.tensor([[0.0, 0.0, 10.0, 10.0], [0.0, 0.0, 10.0, 10.0]], device=dev)
    assert nms(same, torch.tensor([0.1, 0.2], device=dev), 1.0).tolist() == [1, 0]              # thr 1.0 keeps identical
    assert nms(same, torch.tensor([0.1, 0.2], device=dev), 0.5).tolist() == [1]
    zero = torch.tensor([[5.0, 5.0, 5.0, 5.0], [5.0, 5.0, 5.0, 5.0]], device=dev)
    assert sorted(nms(zero, torch.tensor([0.3, 0.2], device=dev), 0.5).tolist()) == [0, 1]       # NaN IoU never suppresses
    assert nms(torch.empty(0, 4, device=dev), torch.empty(0, device=dev), 0.5).numel() == 0


def test_iou_within_one_ulp_of_threshold_follows_gpu_arithmetic():
    """Pairs whose IoU is within rounding of 0.5: the decision must follow torchvision's CUDA kernel
    (FMA-contracted union), which the oracle's iou_gpu_arith restates."""
    from osr_b200.nms import nms
    g = torch.Generator().manual_seed(3)
    n = 3000
    # box B = A shifted so that IoU ~ 0.5: intersection w*(h) / union ... use w=h=s, shift d along x: iou=(s-d)/(s+d)=.5 -> d=s/3
    s_ = torch.rand(n, generator=g) * 300 + 3
    x = torch.rand(n, generator=g) * 500 + (torch.arange(n) * 2000.0)   # pairs far apart from each other
    y = torch.rand(n, generator=g) * 500
    a = torch.stack([x, y, x + s_, y + s_], 1)
    b = a.clone(); b[:, 0] += s_ / 3; b[:, 2] += s_ / 3
    boxes = torch.cat([a, b]).cuda()
    scores = torch.cat([torch.full((n,), 0.9), torch.full((n,), 0.8)]).cuda()
    got = nms(boxes, scores, 0.5)
    exp = torchvision.ops.nms(boxes, scores, 0.5)
    assert torch.equal(got, exp)
    nk = len(got) - n
    assert 0 < nk < n  # some pairs above, some below the threshold: the test really sits on the boundary


@pytest.mark.parametrize("ncls", [1, 20])
def test_batched_nms_matches_detectron2_semantics(ncls):
    from osr_b200.nms import batched_nms
    b, s = _rand_boxes(3000, seed=11, spread=400.0)
    idxs = torch.randint(0, ncls, (3000,), generator=torch.Generator().manual_seed(1)).cuda()
    got = batched_nms(b, s, idxs, 0.5)
    exp = onms.batched_nms(b, s, idxs, 0.5)
    assert torch.equal(got, exp)


def test_batched_nms_images_one_call():
    from osr_b200.nms import batched_nms_images
    bl, sl, il = [], [], []
    for n, k in enumerate([500, 0, 1200, 37]):
        b, s = _rand_boxes(k, seed=20 + n, spread=300.0)
        bl.append(b); sl.append(s)
        il.append(torch.randint(0, 5, (k,), generator=torch.Generator().manual_seed(n)).cuda())
    got = batched_nms_images(bl, sl, il, 0.5, topk_per_image=100)
    for g_, b, s, i in zip(got, bl, sl, il):
        exp = onms.batched_nms(b, s, i, 0.5)[:100] if len(b) else torch.empty(0, dtype=torch.int64, device="cuda:0")
        assert torch.equal(g_, exp)


@pytest.mark.parametrize("topk", [-1, 100])
def test_batched_nms_flat_equals_the_per_image_calls(topk):
    """Concatenated inputs, per-image coordinate trick from one scatter-max: global keep indices equal
    [batched_nms(b, s, i, thr)[:topk] + first row of the image]."""
    from osr_b200.nms import batched_nms_flat
    bl, sl, il, lens = [], [], [], [500, 0, 1200, 37, 0]
    for n, k in enumerate(lens):
        b, s = _rand_boxes(k, seed=30 + n, spread=300.0 + 100.0 * n)
        bl.append(b); sl.append(s)
        il.append(torch.randint(0, 5, (k,), generator=torch.Generator().manual_seed(n)).cuda())
    img = torch.repeat_interleave(torch.arange(len(lens)), torch.tensor(lens)).cuda()
    keep, counts = batched_nms_flat(torch.cat(bl), torch.cat(sl), torch.cat(il), img, lens, 0.5, topk_per_image=topk)
    exp, o = [], 0
    for b, s, i in zip(bl, sl, il):
        e = onms.batched_nms(b, s, i, 0.5) if len(b) else torch.empty(0, dtype=torch.int64, device="cuda:0")
        exp.append((e[:topk] if topk >= 0 else e) + o)
        o += len(b)
    assert counts == [int(e.numel()) for e in exp]
    assert torch.equal(keep, torch.cat(exp))
    keep, counts = batched_nms_flat(torch.empty(0, 4).cuda(), torch.empty(0).cuda(), torch.empty(0, dtype=torch.int64).cuda(),
                                    torch.empty(0, dtype=torch.int64).cuda(), [0, 0], 0.5)
    assert keep.numel() == 0 and counts == [0, 0]


@pytest.mark.parametrize("thr", [0.7, 1.0])
def test_rpn_nominal_mode_matches_oracle_on_gpu(thr):
    """find_top_proposals.py:112-120 executed (stock detectron2): ours vs oracle run on the same GPU
    (torch.topk + torchvision CUDA nms).  Tie-free scores => bit-exact."""
    from osr_b200 import proposals as P, synth
    ho = synth.make_head_outputs(2, (800, 1333), seed=41, mixed_sizes=True)
    dev = "cuda:0"
    ours = P.predict_proposals([a.to(dev) for a in ho.anchors], [d.to(dev) for d in ho.deltas],
                               [c.to(dev) for c in ho.centerness], ho.image_sizes, nms_thresh=thr,
                               pre_nms_topk=2000, post_nms_topk=1000, training=True, mode="nominal")
    ref = orpn.predict_proposals([OBoxes(a.to(dev)) for a in ho.anchors], [d.to(dev) for d in ho.deltas],
                                 [c.to(dev) for c in ho.centerness], ho.image_sizes, nms_thresh=thr,
                                 pre_nms_topk=2000, post_nms_topk=1000, training=True, mode="nominal",
                                 topk_impl="torch")
    for o, r in zip(ours, ref):
        assert len(o) == len(r) and len(o) <= 1000
        assert torch.equal(o.proposal_boxes.tensor, r.proposal_boxes.tensor)
        assert torch.equal(o.objectness_logits, r.objectness_logits)
