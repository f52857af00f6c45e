"""Multi-GPU GPU tests (need >= 2 visible GPUs; skipped on a single-GPU box): the fused encoder -> all-gather kernel
(tcgen05 GEMM epilogue storing into every rank's symmetric-memory buffer over NVLink) against encoder + NCCL
all_gather_into_tensor, including the gathered PLN loss and its embedding gradient - all bit-identical."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("multicast", ["1", "0"])
def test_fused_encoder_gather_two_gpus(multicast):
    env = dict(os.environ, OSR_MC=multicast)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29541" if multicast == "1" else "29542",
           os.path.join(ROOT, "tools", "symm_probe", "fused_gather_check.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "bit-identical to encoder + NCCL all-gather (incl. loss, grads): True" in out.stdout, out.stdout[-2000:]


@pytest.mark.parametrize("world,rank", [(2, 0), (2, 1), (4, 2)])
def test_encode_gather_epilogue_on_one_gpu(world, rank):
    """The fused encoder -> all-gather kernel with every "peer" buffer on THIS GPU (plain device allocations instead of
    NVLink-mapped ones): verifies on a 1-GPU box that the multi-destination epilogue writes this rank's (R, E) block,
    bit-identical to the single-destination encoder, at rows [rank*R, (rank+1)*R) of every destination and nothing else."""
    import ctypes
    from osr_b200 import _lib, synth
    from osr_b200.pln import pln_encode_tc
    lib = _lib.lib()
    dev = torch.device("cuda:0")
    R, Fd, E = 300, 1024, 256          # R not a multiple of the 128-row tile: the last tile is partial
    pi = synth.make_pln_inputs(R, seed=5, device=dev)
    ref = pln_encode_tc(pi.roi_features, pi.enc_w, pi.enc_b + 0.25)
    bufs = [torch.full((world * R, E), -7.0, device=dev) for _ in range(world)]
    ptrs = (ctypes.c_uint64 * world)(*[b.data_ptr() for b in bufs])
    ws = torch.empty(max(int(lib.osr_pln_encode_workspace(R, Fd, E)), 256), dtype=torch.uint8, device=dev)
    bias = (pi.enc_b + 0.25).contiguous()
    rc = lib.osr_pln_encode_gather_fwd(pi.roi_features.data_ptr(), pi.enc_w.data_ptr(), bias.data_ptr(), R, Fd, E,
                                       ctypes.cast(ptrs, ctypes.c_void_p), world, rank, 0, ws.data_ptr(), ws.numel(),
                                       _lib.stream_ptr(dev))
    _lib.check(rc, "osr_pln_encode_gather_fwd")
    torch.cuda.synchronize()
    for b in bufs:
        assert torch.equal(b[rank * R:(rank + 1) * R], ref)
        rest = torch.cat((b[:rank * R], b[(rank + 1) * R:]))
        assert bool((rest == -7.0).all()), "rows of other ranks must not be touched"
    # argument checks of the entry point
    assert lib.osr_pln_encode_gather_fwd(pi.roi_features.data_ptr(), pi.enc_w.data_ptr(), None, R, Fd, E,
                                         ctypes.cast(ptrs, ctypes.c_void_p), world, world, 0, ws.data_ptr(), ws.numel(),
                                         _lib.stream_ptr(dev)) < 0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_ops_on_a_non_current_device():
    """Tensors on cuda:1 while the current device is cuda:0: every entry point launches on the device that owns its
    output pointer (osr::DeviceGuard) and the caller's current device is restored."""
    from osr_b200 import synth
    from osr_b200.pln import pln_loss_from_emb
    from osr_b200.poolers import ROIPooler
    from osr_b200.proposals import rpn_select_decode
    torch.cuda.set_device(0)
    dev = torch.device("cuda:1")
    ho = synth.make_head_outputs(2, (320, 480), seed=3, device=dev)
    sel = rpn_select_decode(ho.anchors, ho.deltas, ho.centerness, ho.image_sizes, 300)
    ref = rpn_select_decode([a.to("cuda:0") for a in ho.anchors], [d.to("cuda:0") for d in ho.deltas],
                            [c.to("cuda:0") for c in ho.centerness], ho.image_sizes, 300)
    assert sel.boxes.device == dev and torch.equal(sel.counts.cpu(), ref.counts.cpu())
    L = sel.num_levels
    for n in range(2):
        k = int(sel.counts[n, L])
        assert torch.equal(sel.boxes[n, :k].cpu(), ref.boxes[n, :k].cpu())
    feats = synth.make_features(2, (320, 480), 64, seed=4, device=dev, channels_last=True)
    rois = [r.to(dev) for r in synth.make_rois(2, 50, (320, 480), seed=5)]
    pooler = ROIPooler(7, synth.POOL_SCALES, 0, "ROIAlignV2")
    from oracle.structures import Boxes as OBoxes
    fg = [f.requires_grad_(True) for f in feats]
    out = pooler(fg, [OBoxes(r) for r in rois])
    out0 = pooler([f.detach().to("cuda:0") for f in feats], [OBoxes(r.to("cuda:0")) for r in rois])
    assert torch.equal(out.detach().cpu(), out0.cpu())
    g = torch.autograd.grad(out, fg, torch.ones_like(out))
    assert all(x.device == dev and torch.isfinite(x).all() for x in g)
    pi = synth.make_pln_inputs(256, seed=6, device=dev)
    emb = (pi.roi_features @ pi.enc_w.t()).requires_grad_(True)
    loss = pln_loss_from_emb(emb, pi.reps.clone().requires_grad_(True), pi.gt_classes, pi.ious, num_known_classes=20)
    loss.backward()
    assert loss.device == dev and torch.cuda.current_device() == 0
