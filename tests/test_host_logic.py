"""Host-side logic that needs no GPU: sampling (subsample_labels is pure torch), sharding, the byte model's stage
accounting, and the product's refusal to run its matching / inference ops on CPU tensors."""
import pytest
import torch

from oracle import sampling as osamp


def test_subsample_labels_matches_oracle_with_the_same_permutations():
    from osr_b200.sampling import subsample_labels
    g = torch.Generator().manual_seed(3)
    labels = torch.where(torch.rand(700, generator=g) < 0.3, torch.randint(0, 20, (700,), generator=g), torch.full((700,), 80))
    labels[::50] = -1

    def stream(seed):
        gg = torch.Generator().manual_seed(seed)
        return lambda n: torch.randperm(n, generator=gg)

    a = subsample_labels(labels, 512, 0.25, 80, stream(9))
    b = osamp.subsample_labels(labels, 512, 0.25, 80, stream(9))
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert len(a[0]) == min(128, int(((labels != -1) & (labels != 80)).sum())) and len(a[0]) + len(a[1]) == 512
    assert bool((labels[a[0]] != 80).all()) and bool((labels[a[1]] == 80).all())


def test_subsample_labels_few_negatives():
    from osr_b200.sampling import subsample_labels
    labels = torch.tensor([1, 2, 80, 3, 80, -1, 4])
    fg, bg = subsample_labels(labels, 4, 0.5, 80, lambda n: torch.arange(n))
    assert fg.tolist() == [0, 1] and bg.tolist() == [2, 4]


def test_shard_range_covers_everything_once():
    from osr_b200.dist import shard_range
    for total, world in ((16, 8), (17, 4), (3, 8), (128, 3)):
        seen = []
        for r in range(world):
            seen += list(shard_range(total, r, world))
        assert seen == list(range(total))


def test_matching_and_inference_ops_refuse_cpu_tensors():
    from osr_b200 import _lib
    from osr_b200.inference import inference
    from osr_b200.sampling import match_proposals
    from osr_b200.structures import Boxes, Instances
    off = torch.tensor([0, 2], dtype=torch.int32)
    with pytest.raises(_lib.OsrError):
        match_proposals(torch.rand(2, 4), off, torch.rand(1, 4), torch.zeros(1, dtype=torch.int64),
                        torch.tensor([0, 1], dtype=torch.int32), 2)
    p = Instances((10, 10)); p.set("proposal_boxes", Boxes(torch.rand(2, 4))); p.set("objectness_logits", torch.rand(2))
    with pytest.raises(_lib.OsrError):
        inference((torch.rand(2, 4), torch.rand(2, 1)), [p], torch.rand(2, 8))


def test_sample_labels_batched_contract():
    """The sync-free batched sampler: per image at most int(B * frac) foreground rows FIRST, then background rows up to B,
    no padding / ignored rows, no duplicates, and the counts detectron2's subsample_labels would give."""
    from osr_b200.sampling import sample_labels_batched
    g = torch.Generator().manual_seed(11)
    N, P, B = 5, 400, 64
    labels = torch.where(torch.rand(N, P, generator=g) < 0.2, torch.randint(0, 20, (N, P), generator=g), torch.full((N, P), 81))
    labels[:, ::37] = -1
    valid = torch.ones(N, P, dtype=torch.bool)
    valid[1, 100:] = False
    labels[2] = 81            # no foreground at all
    labels[3, 40:] = -1       # few candidates: fewer than B samples
    labels[4, :] = 7          # only foreground: int(B * frac) samples
    idx, cnt = sample_labels_batched(labels, valid, B, 0.25, 81, generator=g)
    for n in range(N):
        lab = labels[n][valid[n]]
        n_pos = min(int(((lab != -1) & (lab != 81)).sum()), int(B * 0.25))
        n_neg = min(int((lab == 81).sum()), B - n_pos)
        k = int(cnt[n])
        assert k == n_pos + n_neg
        ii = idx[n, :k]
        assert bool(valid[n, ii].all()) and len(set(ii.tolist())) == k
        got = labels[n, ii]
        assert bool(((got[:n_pos] != 81) & (got[:n_pos] != -1)).all()) and bool((got[n_pos:] == 81).all())
    # a different generator state draws a different subset (it is a random sample, not a prefix)
    idx2, _ = sample_labels_batched(labels, valid, B, 0.25, 81, generator=g)
    assert not torch.equal(idx, idx2)


def test_cat_rows_recognises_consecutive_views_and_falls_back_otherwise():
    from osr_b200.structures import cat_rows
    x = torch.arange(40.0).reshape(10, 4)
    parts = list(x.split([3, 0, 5, 2]))
    y = cat_rows(parts)
    assert y.data_ptr() == x.data_ptr() and torch.equal(y, x)                     # all rows: the buffer itself
    y = cat_rows(parts[1:])
    assert y.data_ptr() == x[3:].data_ptr() and torch.equal(y, x[3:])             # a suffix, empty view in front
    y = cat_rows([parts[0], parts[3]])                                            # rows missing in between: a real cat
    assert y.data_ptr() != x.data_ptr() and torch.equal(y, torch.cat([parts[0], parts[3]]))
    y = cat_rows([parts[2], parts[0]])                                            # wrong order
    assert torch.equal(y, torch.cat([parts[2], parts[0]]))
    y = cat_rows([x[:3], torch.empty(0, 4), x[3:]])                               # a foreign (empty) tensor
    assert torch.equal(y, x)
    y = cat_rows([x[:3].clone(), x[3:]])
    assert torch.equal(y, x) and y.data_ptr() != x.data_ptr()
    xt = x.t()                                                                    # non-contiguous views
    assert torch.equal(cat_rows([xt[:2], xt[2:]]), xt)
    z = torch.arange(10)
    assert torch.equal(cat_rows(list(z.split([4, 6]))), z)                        # 1-D fields (scores, classes)
    w = x.clone().requires_grad_()
    y = cat_rows([w[:3], w[3:]])                                                  # autograd inputs are never aliased
    assert y.requires_grad and y.data_ptr() != w.data_ptr()
    assert cat_rows([x]) is x


def test_flat_prefixes_lists_the_first_rows_of_every_segment():
    from osr_b200.structures import flat_prefixes
    P, B = flat_prefixes([0, 10, 20, 30], [2, 0, 3, 1], "cpu")
    assert P.tolist() == [0, 1, 20, 21, 22, 30] and B.tolist() == [0, 0, 20, 20, 20, 30]
    P, B = flat_prefixes([0, 5], [0, 0], "cpu")
    assert P.numel() == 0 and B.numel() == 0 and P.dtype == torch.int64


def test_make_instances_and_boxes_view_build_the_same_objects_as_the_constructors():
    from osr_b200.structures import Boxes, Instances, boxes_view, make_instances
    t = torch.arange(12.0).reshape(3, 4)
    a = Instances((5, 7))
    a.set("proposal_boxes", Boxes(t))
    a.set("objectness_logits", torch.ones(3))
    b = make_instances((5, 7), proposal_boxes=boxes_view(t), objectness_logits=torch.ones(3))
    assert type(b) is type(a) and b.image_size == a.image_size and len(b) == len(a) == 3
    assert list(b.get_fields()) == list(a.get_fields())
    assert type(b.proposal_boxes) is type(a.proposal_boxes) and b.proposal_boxes.tensor.data_ptr() == t.data_ptr()
    assert torch.equal(b.proposal_boxes.tensor, a.proposal_boxes.tensor) and len(b.proposal_boxes) == 3
    assert torch.equal(b[1:].objectness_logits, a[1:].objectness_logits)      # the result behaves like any Instances
    assert len(make_instances((5, 7), x=torch.zeros(0))) == 0
