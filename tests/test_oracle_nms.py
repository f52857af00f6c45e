"""CPU pins of the NMS oracle: loop restatement vs the torchvision binary; Appendix C.4 vectors."""
import torch
import torchvision

from oracle import nms as onms


def _rand(n, seed):
    g = torch.Generator().manual_seed(seed)
    c = torch.rand(n, 2, generator=g) * 200
    wh = torch.rand(n, 2, generator=g) * 60 + 2
    return torch.cat([c - wh / 2, c + wh / 2], 1), torch.rand(n, generator=g)


def test_loops_match_torchvision_cpu():
    for seed, thr in ((1, 0.5), (2, 0.7), (3, 1.0), (4, 0.3)):
        b, s = _rand(300, seed)
        assert torch.equal(onms.nms_loops(b, s, thr, "cpu"), torchvision.ops.nms(b, s, thr))


def test_appendix_c4():
    disjoint = torch.tensor([[i * 10.0, 0.0, i * 10.0 + 5.0, 5.0] for i in range(20)])
    assert torchvision.ops.nms(disjoint, torch.ones(20), 0.5).tolist() == list(range(20))
    sc = torch.tensor([.5, .9, .5, .9, .5, .9])
    assert torchvision.ops.nms(disjoint[:6], sc, 0.5).tolist() == [1, 3, 5, 0, 2, 4]
    same = torch.tensor([[0.0, 0.0, 10.0, 10.0], [0.0, 0.0, 10.0, 10.0]])
    assert torchvision.ops.nms(same, torch.tensor([0.1, 0.2]), 1.0).tolist() == [1, 0]
    assert onms.nms_loops(same, torch.tensor([0.1, 0.2]), 1.0, "gpu").tolist() == [1, 0]


def test_batched_nms_dispatch_paths_agree_on_disjoint_classes():
    b, s = _rand(500, 9)
    idxs = torch.randint(0, 4, (500,), generator=torch.Generator().manual_seed(2))
    a = onms.batched_nms(b, s, idxs, 0.5)
    from torchvision.ops import boxes as tvb
    v = tvb._batched_nms_vanilla(b, s, idxs, 0.5)
    assert sorted(a.tolist()) == sorted(v.tolist())
