"""The device-resident training step (osr_b200/pipeline.py: kernels called directly, no autograd, CUDA-graph capturable)
must produce exactly what the autograd formulation of the same ops produces, and the captured graph must replay it."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _path():
    from osr_b200.pipeline import PathConfig, RoiPathStep
    cfg = PathConfig(num_images=2, image_hw=(320, 480), pre_nms_topk=300, rois_per_image=64, channels=64)
    return cfg, RoiPathStep(cfg, "cuda:0")


def test_direct_step_equals_autograd_formulation():
    from osr_b200.pln import pln_encode_tc, pln_loss_from_emb
    cfg, path = _path()
    loss, _ = path.step()
    last = path.last
    feats = [f.detach().clone().requires_grad_(True) for f in path.feats]
    pooled, _ = path.pooler.pool_rois(feats, last["rois"], path.roi_offsets)
    assert torch.equal(pooled.detach(), last["pooled"])
    g_feats = torch.autograd.grad(pooled, feats, path.grad_pooled)
    for a, b in zip(g_feats, last["g_feats"]):
        assert torch.equal(a, b)
    pi = path.pln
    emb = pln_encode_tc(pi.roi_features, pi.enc_w, pi.enc_b).requires_grad_(True)
    reps = pi.reps.detach().clone().requires_grad_(True)
    s_cls, s_iou = last["sample"]["classes"].view(-1), last["sample"]["ious"].view(-1)   # the matcher's, via the sampler
    assert int(((s_cls < cfg.num_known) & (s_iou > cfg.iou_threshold)).sum()) > 0
    l2 = pln_loss_from_emb(emb, reps, s_cls, s_iou, num_known_classes=cfg.num_known, alpha=cfg.alpha, beta=cfg.beta,
                           loss_weight=cfg.loss_weight, iou_threshold=cfg.iou_threshold)
    ge, gr = torch.autograd.grad(l2, [emb, reps])
    assert torch.equal(l2.detach(), loss.detach())
    assert torch.equal(ge, last["g_emb"]) and torch.equal(gr, last["g_reps"])


def test_step_replays_from_a_cuda_graph():
    cfg, path = _path()
    keys = torch.rand(cfg.num_images * path.kmax, device="cuda:0")   # fixed sampler keys: eager step and replay draw the same sample
    loss, _ = path.step(keys=keys)
    ref = dict(loss=loss.detach().clone(), g=[g.clone() for g in path.last["g_feats"]], pooled=path.last["pooled"].clone())
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        path.step(keys=keys)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loss_g, _ = path.step(keys=keys)
    out = path.last
    for t in out["g_feats"]:
        t.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(loss_g, ref["loss"]) and torch.equal(out["pooled"], ref["pooled"])
    for a, b in zip(out["g_feats"], ref["g"]):
        assert torch.equal(a, b)


def test_graph_replay_draws_a_fresh_sample_every_step():
    """Without fixed keys the captured step holds the torch.rand launch: every replay samples anew (a graph-safe generator
    offset), always 512-style full samples with the positives first."""
    cfg, path = _path()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        path.step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        path.step()
    out = path.last
    seen = []
    for _ in range(3):
        g.replay()
        torch.cuda.synchronize()
        seen.append(out["sample"]["index"].clone())
        cnt = out["sample"]["count"]
        assert int(cnt[:, 1].min()) == cfg.rois_per_image
        cls = out["sample"]["classes"]
        for n in range(cfg.num_images):
            k = int(cnt[n, 0])
            assert bool((cls[n, :k] != cfg.num_classes).all()) and bool((cls[n, k:] == cfg.num_classes).all())
    assert not torch.equal(seen[0], seen[1]) and not torch.equal(seen[1], seen[2])
