"""GPU parity of the sampling glue (osr_match_label + label_and_sample_proposals) against the oracle."""
import pytest
import torch

from oracle import sampling as osamp
from oracle.structures import Boxes as OBoxes, Instances as OInstances, pairwise_iou

pytestmark = pytest.mark.gpu


def _make(n_img, n_prop, n_gt, seed, hw=(800, 1333)):
    g = torch.Generator().manual_seed(seed)
    props, tgts = [], []
    for i in range(n_img):
        h, w = hw
        G = n_gt if isinstance(n_gt, int) else n_gt[i]
        gc = torch.rand(G, 2, generator=g) * torch.tensor([w * 0.8, h * 0.8])
        gwh = torch.rand(G, 2, generator=g) * 480 + 32
        gtb = torch.cat((gc, (gc + gwh).minimum(torch.tensor([float(w), float(h)]))), 1)
        P = n_prop if isinstance(n_prop, int) else n_prop[i]
        # proposals: jittered copies of GT boxes (so many IoUs straddle 0.5) + random boxes + exact duplicates
        k = P // 2
        src = gtb[torch.randint(0, G, (k,), generator=g)]
        jit = src + (torch.rand(k, 4, generator=g) - 0.5) * 0.6 * (src[:, 2:] - src[:, :2]).repeat(1, 2)
        c = torch.rand(P - k, 2, generator=g) * torch.tensor([float(w), float(h)])
        rwh = torch.rand(P - k, 2, generator=g) * 300 + 4
        rnd = torch.cat((c, c + rwh), 1)
        pb = torch.cat((jit, rnd), 0)
        pb = torch.stack((pb[:, 0].clamp(0, w), pb[:, 1].clamp(0, h), pb[:, 2].clamp(0, w), pb[:, 3].clamp(0, h)), 1)
        pb[-1] = gtb[0]                       # exact GT duplicate: IoU == 1
        if P >= 2:
            pb[-2] = torch.tensor([5., 5., 5., 9.])  # zero-area box
        props.append(pb)
        tgts.append((gtb, torch.randint(0, 20, (G,), generator=g)))
    return props, tgts


@pytest.mark.parametrize("n_img,n_prop,n_gt", [(4, 7323, 8), (3, [100, 2000, 1], [1, 300, 5]), (1, 33, 2)])
def test_match_label_bit_exact(n_img, n_prop, n_gt):
    from osr_b200.sampling import match_proposals
    props, tgts = _make(n_img, n_prop, n_gt, seed=3)
    dev = "cuda:0"
    counts = [p.shape[0] for p in props]
    gcounts = [t[0].shape[0] for t in tgts]
    off = torch.tensor([0] + torch.tensor(counts).cumsum(0).tolist(), dtype=torch.int32, device=dev)
    goff = torch.tensor([0] + torch.tensor(gcounts).cumsum(0).tolist(), dtype=torch.int32, device=dev)
    midx, miou, mlab, mcls = match_proposals(torch.cat(props).to(dev), off, torch.cat([t[0] for t in tgts]).to(dev),
                                             torch.cat([t[1] for t in tgts]).to(dev), goff, max(counts),
                                             iou_threshold=0.5, background_label=80)
    b0 = 0
    for p, (gb, gc) in zip(props, tgts):
        for where in ("cpu", dev):   # torch's own op chain on the CPU and on the GPU
            m = pairwise_iou(OBoxes(gb.to(where)), OBoxes(p.to(where)))
            idx, lab = osamp.matcher(m, 0.5)
            iou = m[idx, torch.arange(m.shape[1], device=where)]
            sl = slice(b0, b0 + p.shape[0])
            assert torch.equal(midx[sl].cpu().long(), idx.cpu())
            assert torch.equal(miou[sl].cpu(), iou.cpu()), "matched IoU must be bit-exact"
            assert torch.equal(mlab[sl].cpu().to(torch.int8), lab.cpu())
            cls = gc.to(where)[idx]
            cls[lab == 0] = 80
            assert torch.equal(mcls[sl].cpu(), cls.cpu())
        b0 += p.shape[0]


def test_label_and_sample_proposals_matches_oracle():
    from osr_b200.sampling import label_and_sample_proposals
    from osr_b200.structures import Boxes, Instances
    props, tgts = _make(3, [7323, 500, 40], 8, seed=11)
    dev = "cuda:0"
    ours_p, ours_t, ref_p, ref_t = [], [], [], []
    for p, (gb, gc) in zip(props, tgts):
        lg = torch.linspace(3, -3, p.shape[0])
        a = Instances((800, 1333)); a.set("proposal_boxes", Boxes(p.to(dev))); a.set("objectness_logits", lg.to(dev))
        b = Instances((800, 1333)); b.set("gt_boxes", Boxes(gb.to(dev))); b.set("gt_classes", gc.to(dev))
        ours_p.append(a); ours_t.append(b)
        c = OInstances((800, 1333)); c.set("proposal_boxes", OBoxes(p)); c.set("objectness_logits", lg)
        d = OInstances((800, 1333)); d.set("gt_boxes", OBoxes(gb)); d.set("gt_classes", gc)
        ref_p.append(c); ref_t.append(d)
    # identical "random" draws on both sides: a seeded CPU permutation stream consumed in the same order
    def stream(seed):
        g = torch.Generator().manual_seed(seed)
        return lambda n: torch.randperm(n, generator=g)
    kw = dict(num_classes=80, batch_size_per_image=512, positive_fraction=0.25)
    ours = label_and_sample_proposals(ours_p, ours_t, randperm=stream(5), **kw)
    ref = osamp.label_and_sample_proposals(ref_p, ref_t, randperm=stream(5), **kw)
    for a, b in zip(ours, ref):
        assert len(a) == len(b)
        assert torch.equal(a.get("proposal_boxes").tensor.cpu(), b.get("proposal_boxes").tensor)
        assert torch.equal(a.get("gt_classes").cpu(), b.get("gt_classes"))
        assert torch.equal(a.get("ious").cpu(), b.get("ious"))
        assert torch.equal(a.get("gt_boxes").tensor.cpu(), b.get("gt_boxes").tensor)
        assert torch.equal(a.get("objectness_logits").cpu(), b.get("objectness_logits"))


def test_image_without_gt_raises_like_the_reference():
    from osr_b200.sampling import label_and_sample_proposals
    from osr_b200.structures import Boxes, Instances
    a = Instances((100, 100)); a.set("proposal_boxes", Boxes(torch.rand(5, 4).cuda())); a.set("objectness_logits", torch.zeros(5).cuda())
    b = Instances((100, 100)); b.set("gt_boxes", Boxes(torch.empty(0, 4).cuda())); b.set("gt_classes", torch.empty(0, dtype=torch.int64).cuda())
    with pytest.raises(IndexError):
        label_and_sample_proposals([a], [b], num_classes=80)


def test_padded_layout_with_counts_column():
    """The RpnSelection layout: (N, Kmax, 4) padded boxes + a strided counts column; rows past the count untouched."""
    from osr_b200.sampling import match_proposals
    props, tgts = _make(3, [50, 7, 31], 4, seed=9)
    dev = "cuda:0"
    kmax = 64
    padded = torch.full((3, kmax, 4), float("nan"))
    counts = torch.zeros(3, 7, dtype=torch.int32)
    for n, p in enumerate(props):
        padded[n, :p.shape[0]] = p
        counts[n, 5] = p.shape[0]
    counts = counts.to(dev)
    off = torch.arange(0, 4 * kmax, kmax, dtype=torch.int32, device=dev)
    goff = torch.arange(0, 13, 4, dtype=torch.int32, device=dev)
    midx, miou, mlab, mcls = match_proposals(padded.view(-1, 4).to(dev), off, torch.cat([t[0] for t in tgts]).to(dev),
                                             torch.cat([t[1] for t in tgts]).to(dev), goff, kmax,
                                             background_label=80, box_counts=counts[:, 5], box_counts_stride=7)
    for n, (p, (gb, gc)) in enumerate(zip(props, tgts)):
        m = pairwise_iou(OBoxes(gb), OBoxes(p))
        idx, lab = osamp.matcher(m, 0.5)
        sl = slice(n * kmax, n * kmax + p.shape[0])
        assert torch.equal(midx[sl].cpu().long(), idx)
        assert torch.equal(miou[sl].cpu(), m[idx, torch.arange(p.shape[0])])
        assert torch.equal(mlab[sl].cpu().to(torch.int8), lab)


def test_label_and_sample_batched_path_is_consistent_with_the_matcher():
    """Default (no injected randperm): the whole batch is sampled on the device with one host read.  Every sampled row
    must carry exactly the label / IoU / GT box the reference's matcher assigns to that box, foreground rows first."""
    from osr_b200 import sampling as S, structures as st
    props, tgts = _make(3, [900, 2500, 40], [4, 8, 2], seed=12)
    P, T = [], []
    for pb, (gb, gc) in zip(props, tgts):
        p = st.Instances((800, 1333)); p.set("proposal_boxes", st.Boxes(pb.cuda())); p.set("objectness_logits", torch.rand(len(pb)).cuda())
        t = st.Instances((800, 1333)); t.set("gt_boxes", st.Boxes(gb.cuda())); t.set("gt_classes", gc.cuda())
        P.append(p); T.append(t)
    out = S.label_and_sample_proposals(P, T, num_classes=80, batch_size_per_image=128, positive_fraction=0.25)
    for q, pb, (gb, gc) in zip(out, props, tgts):
        allb = torch.cat((pb, gb))                       # proposals + appended GT
        m = pairwise_iou(OBoxes(gb), OBoxes(allb))
        idx, lab = osamp.matcher(m, 0.5)
        iou = m[idx, torch.arange(m.shape[1])]
        cls = gc[idx].clone(); cls[lab == 0] = 80
        sb = q.get("proposal_boxes").tensor.cpu()
        # locate every sampled box in the candidate list (boxes are distinct except planted duplicates: compare values)
        pos = [(allb == b).all(dim=1).nonzero()[0, 0].item() for b in sb]
        assert torch.equal(q.get("gt_classes").cpu(), cls[pos]) and torch.equal(q.get("ious").cpu(), iou[pos])
        assert torch.equal(q.get("gt_boxes").tensor.cpu(), gb[idx[pos]])
        n_fg = int((q.get("gt_classes") != 80).sum())
        assert n_fg <= 32 and bool((q.get("gt_classes")[:n_fg] != 80).all()) and bool((q.get("gt_classes")[n_fg:] == 80).all())
        assert len(q) == min(128, n_fg + int((cls == 80).sum()))


def _keyed_randperm(keys_of_kind):
    """The permutation detectron2's subsample_labels would have to draw so that it keeps the rows with the smallest keys in
    ascending (key, row) order: the stable arg-sort of the kind's keys.  `keys_of_kind` is consumed in call order
    (positives, then negatives, image after image) - the order subsample_labels calls randperm in."""
    it = iter(keys_of_kind)

    def rp(n):
        k = next(it)
        assert k.numel() == n
        return torch.argsort(k, stable=True)
    return rp


@pytest.mark.parametrize("case", ["mixed", "ties", "few_rows", "all_positive", "no_positive_quota", "counts_column"])
def test_sample_rois_equals_subsample_labels_with_the_keyed_permutation(case):
    """osr_sample_rois against the restatement of detectron2's subsample_labels (oracle/sampling.py), fed with the
    permutation that the keys define; bit-exact indices, counts and gathered fields, no tolerance."""
    from osr_b200.sampling import sample_rois
    g = torch.Generator().manual_seed(hash(case) % 1000 + 3)
    S, frac, bg = 128, 0.25, 80
    sizes = {"mixed": [7323, 500, 40, 0, 2000], "ties": [3000, 700], "few_rows": [50, 5, 1],
             "all_positive": [600, 90], "no_positive_quota": [900], "counts_column": [300, 2000, 17]}[case]
    if case == "no_positive_quota":
        frac = 0.0
    labels, keys = [], []
    for P in sizes:
        r = torch.rand(P, generator=g)
        lab = torch.full((P,), bg, dtype=torch.int64)
        lab[r < 0.15] = torch.randint(0, 20, (int((r < 0.15).sum()),), generator=g)
        lab[(r >= 0.15) & (r < 0.2)] = -1
        if case == "all_positive":
            lab = torch.randint(0, 20, (P,), generator=g)
        k = torch.rand(P, generator=g)
        if case == "ties":
            k = torch.floor(k * 16) / 16          # 16 distinct keys: the threshold bucket always holds ties
            k[::7] = -k[::7]                      # negative keys order below positive ones
        labels.append(lab); keys.append(k)
    N = len(sizes)
    kmax = max(sizes) + 5
    padded = case == "counts_column"
    if padded:   # (N, kmax) padded rows + a strided counts column, as osr_rpn_select_decode leaves them
        lab_t = torch.full((N, kmax), 7, dtype=torch.int64); key_t = torch.zeros(N, kmax)
        for n, (l, k) in enumerate(zip(labels, keys)):
            lab_t[n, :len(l)] = l; key_t[n, :len(k)] = k
        off = torch.arange(0, (N + 1) * kmax, kmax, dtype=torch.int32)
        cnt = torch.zeros(N, 3, dtype=torch.int32); cnt[:, 1] = torch.tensor(sizes, dtype=torch.int32)
        lab_t, key_t = lab_t.view(-1), key_t.view(-1)
    else:
        lab_t, key_t = torch.cat(labels), torch.cat(keys)
        off = torch.tensor([0] + list(torch.tensor(sizes).cumsum(0)), dtype=torch.int32)
        cnt = None
    T = lab_t.numel()
    boxes = torch.rand(T, 4, generator=g) * 100
    logits, ious = torch.randn(T, generator=g), torch.rand(T, generator=g)
    midx = torch.randint(0, 4, (T,), generator=g, dtype=torch.int32)
    goff = torch.arange(0, 4 * (N + 1), 4, dtype=torch.int32)
    dev = "cuda:0"
    # both instantiations: rows cached in shared memory (host bound given) / re-read from global memory in every pass
    bound = max(sizes) if case in ("mixed", "ties", "all_positive", "counts_column") else 0
    out = sample_rois(lab_t.to(dev), key_t.to(dev), off.to(dev), S, int(S * frac), bg,
                      box_counts=None if cnt is None else cnt.to(dev)[:, 1], box_counts_stride=3,
                      boxes=boxes.to(dev), logits=logits.to(dev), ious=ious.to(dev), matched_idx=midx.to(dev),
                      gt_offsets=goff.to(dev), max_boxes_per_image=bound)
    other = sample_rois(lab_t.to(dev), key_t.to(dev), off.to(dev), S, int(S * frac), bg,
                        box_counts=None if cnt is None else cnt.to(dev)[:, 1], box_counts_stride=3,
                        max_boxes_per_image=0 if bound else max(sizes))
    assert torch.equal(other["index"], out["index"]) and torch.equal(other["count"], out["count"])
    rois = sample_rois(lab_t.to(dev), key_t.to(dev), off.to(dev), S, int(S * frac), bg,
                       box_counts=None if cnt is None else cnt.to(dev)[:, 1], box_counts_stride=3,
                       boxes=boxes.to(dev), want_rois=True)["rois"].cpu()
    index, count = out["index"].cpu(), out["count"].cpu()
    for n, (lab, k) in enumerate(zip(labels, keys)):
        pos = torch.nonzero((lab != -1) & (lab != bg)).flatten()
        neg = torch.nonzero(lab == bg).flatten()
        fg, bgi = osamp.subsample_labels(lab, S, frac, bg, _keyed_randperm([k[pos], k[neg]]))
        want = torch.cat([fg, bgi])
        c = len(want)
        assert count[n].tolist() == [len(fg), c]
        assert torch.equal(index[n, :c].long(), want) and bool((index[n, c:] == -1).all())
        b0 = int(off[n])
        assert torch.equal(out["boxes"][n, :c].cpu(), boxes[b0 + want])
        assert torch.equal(out["logits"][n, :c].cpu(), logits[b0 + want])
        assert torch.equal(out["ious"][n, :c].cpu(), ious[b0 + want])
        assert torch.equal(out["classes"][n, :c].cpu(), lab[want])
        assert torch.equal(out["gt"][n, :c].cpu(), midx[b0 + want].long() + 4 * n)
        assert torch.equal(rois[n, :c, 1:], boxes[b0 + want]) and bool((rois[n, :c, 0] == n).all())


def test_sample_rois_is_a_uniform_draw():
    """Distribution: over many key draws every negative row is kept with probability quota / population (5 sigma)."""
    from osr_b200.sampling import sample_rois
    dev = "cuda:0"
    P, S, trials = 400, 64, 600
    lab = torch.full((P,), 80, dtype=torch.int64, device=dev)
    lab[:40] = 3                                   # 40 positives, quota 16
    off = torch.tensor([0, P], dtype=torch.int32, device=dev)
    g = torch.Generator(device=dev).manual_seed(5)
    hits = torch.zeros(P, device=dev)
    for _ in range(trials):
        idx = sample_rois(lab, torch.rand(P, device=dev, generator=g), off, S, 16, 80)["index"][0].long()
        hits[idx] += 1
    p_pos, p_neg = 16 / 40, 48 / 360
    for sl, pr in ((slice(0, 40), p_pos), (slice(40, P), p_neg)):
        sd = (trials * pr * (1 - pr)) ** 0.5
        assert float((hits[sl] - trials * pr).abs().max()) < 5 * sd
