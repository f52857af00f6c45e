"""world_size=2 gloo test (CPU) of the multi-GPU host logic: sharding + the gathered PLN loss rule.
The CUDA loss op is replaced by the oracle's CPU implementation through ``loss_fn`` - what is tested here is
dist.py (gather order, local-row autograd edge, r_norm / center_weight / gradient scaling)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KW = dict(num_known_classes=20, alpha=0.1, beta=0.9, loss_weight=0.5, iou_threshold=0.5)


def _oracle_loss_fn(emb, reps, labels, ious, *, r_norm, center_weight, emb_grad_scale, **kw):
    from oracle import pln as opln

    class Scale(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            return x.view_as(x)

        @staticmethod
        def backward(ctx, g):
            return g * emb_grad_scale

    return opln.pln_loss_from_emb(Scale.apply(emb), reps, labels, ious, r_norm=r_norm, center_weight=center_weight, **kw)


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "openset-rcnn_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from osr_b200 import dist as odist, synth
    from oracle import pln as opln
    r, lr, w = odist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    Rl = 64
    pi = synth.make_pln_inputs(Rl, seed=100 + rank)
    enc_w = synth.make_pln_inputs(4, seed=1).enc_w
    reps0 = synth.make_pln_inputs(4, seed=1).reps
    emb = (pi.roi_features @ enc_w.t()).requires_grad_(True)
    reps = reps0.clone().requires_grad_(True)
    loss = odist.gathered_pln_loss(emb, reps, pi.gt_classes, pi.ious, loss_fn=_oracle_loss_fn, **KW)
    loss.backward()
    # reference behaviour: per-rank local loss; DDP averages parameter grads
    emb_r = emb.detach().clone().requires_grad_(True); reps_r = reps0.clone().requires_grad_(True)
    local = opln.pln_loss_from_emb(emb_r, reps_r, pi.gt_classes, pi.ious, **KW)
    local.backward()
    mean_loss = local.detach().clone(); dist.all_reduce(mean_loss); mean_loss /= world
    mean_reps_grad = reps_r.grad.clone(); dist.all_reduce(mean_reps_grad); mean_reps_grad /= world
    ok = (torch.allclose(loss.detach(), mean_loss, rtol=1e-5, atol=1e-7)
          and torch.allclose(emb.grad, emb_r.grad, rtol=1e-4, atol=1e-8)
          and torch.allclose(reps.grad, mean_reps_grad, rtol=1e-4, atol=1e-8))
    # the same loss without gathering: per-rank loss + all-reduce of the value and of the prototype gradient
    emb2 = emb.detach().clone().requires_grad_(True); reps2 = reps0.clone().requires_grad_(True)
    loss2 = odist.reduced_pln_loss(emb2, reps2, pi.gt_classes, pi.ious, loss_fn=_oracle_loss_fn, **KW)
    loss2.backward()
    ok = ok and (torch.allclose(loss2.detach(), loss.detach(), rtol=1e-5, atol=1e-7)
                 and torch.allclose(emb2.grad, emb.grad, rtol=1e-4, atol=1e-8)
                 and torch.allclose(reps2.grad, reps.grad, rtol=1e-4, atol=1e-8))
    shards = [list(odist.shard_range(19, k, world)) for k in range(world)]
    ok = ok and sorted(sum(shards, [])) == list(range(19))
    q.put((rank, bool(ok), float(loss), float(mean_loss)))
    dist.barrier()
    dist.destroy_process_group()


def test_gathered_pln_loss_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert abs(res[0][2] - res[1][2]) < 1e-9  # identical global loss on both ranks


def _worker_ragged(rank, world, port, q):
    """Unequal row counts per rank (detectron2's subsample_labels can return fewer than batch_size_per_image rows):
    the gathered loss must equal the single-process loss over the concatenated batch, value and gradients."""
    for p in (ROOT, os.path.join(ROOT, "openset-rcnn_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from osr_b200 import dist as odist, synth
    from oracle import pln as opln
    odist.init_from_env("gloo")
    counts = [64, 41]
    pis = [synth.make_pln_inputs(counts[k], seed=300 + k) for k in range(world)]
    enc_w = synth.make_pln_inputs(4, seed=1).enc_w
    reps0 = synth.make_pln_inputs(4, seed=1).reps
    embs = [(p.roi_features @ enc_w.t()) for p in pis]
    emb = embs[rank].clone().requires_grad_(True)
    reps = reps0.clone().requires_grad_(True)
    pi = pis[rank]
    loss = odist.gathered_pln_loss(emb, reps, pi.gt_classes, pi.ious, loss_fn=_oracle_loss_fn, **KW)   # counts exchanged inside
    loss.backward()
    # single-process reference: the global batch in one call; local-row gradients scaled by W (DDP averages them)
    e_all = torch.cat(embs).requires_grad_(True)
    r_all = reps0.clone().requires_grad_(True)
    ref = opln.pln_loss_from_emb(e_all, r_all, torch.cat([p.gt_classes for p in pis]), torch.cat([p.ious for p in pis]),
                                 r_norm=float(sum(counts)), center_weight=float(world), **KW)
    ref.backward()
    o = sum(counts[:rank])
    ok = (torch.allclose(loss.detach(), ref.detach(), rtol=1e-5, atol=1e-7)
          and torch.allclose(emb.grad, world * e_all.grad[o:o + counts[rank]], rtol=1e-4, atol=1e-8)
          and torch.allclose(reps.grad, r_all.grad, rtol=1e-4, atol=1e-8))
    # a wrong explicit count list is rejected before any collective on the rows
    try:
        odist.gathered_pln_loss(emb.detach(), reps0, pi.gt_classes, pi.ious, loss_fn=_oracle_loss_fn,
                                rows_per_rank=[counts[rank] + 1] * world, **KW)
        ok = False
    except ValueError:
        pass
    odist.check_uniform_requires_grad(reps)
    q.put((rank, bool(ok), float(loss), float(ref)))
    dist.barrier()
    dist.destroy_process_group()


def test_gathered_pln_loss_unequal_rows_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29950 + (os.getpid() % 40)
    procs = [ctx.Process(target=_worker_ragged, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    assert abs(res[0][2] - res[1][2]) < 1e-9
