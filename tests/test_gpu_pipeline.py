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
    l2 = pln_loss_from_emb(emb, reps, pi.gt_classes, pi.ious, num_known_classes=cfg.num_known, alpha=cfg.alpha, beta=cfg.beta,
                           loss_weight=cfg.loss_weight, iou_threshold=cfg.iou_threshold)
    ge, gr = torch.autograd.grad(l2, [emb, reps])
    assert torch.equal(l2.detach(), loss.detach())
    assert torch.equal(ge, last["g_emb"]) and torch.equal(gr, last["g_reps"])


def test_step_replays_from_a_cuda_graph():
    cfg, path = _path()
    loss, _ = path.step()
    ref = dict(loss=loss.detach().clone(), g=[g.clone() for g in path.last["g_feats"]], pooled=path.last["pooled"].clone())
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        path.step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        loss_g, _ = path.step()
    out = path.last
    for t in out["g_feats"]:
        t.zero_()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(loss_g, ref["loss"]) and torch.equal(out["pooled"], ref["pooled"])
    for a, b in zip(out["g_feats"], ref["g"]):
        assert torch.equal(a, b)
