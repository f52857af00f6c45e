"""N-GPU check + timing of the fused encoder -> all-gather kernel against encoder + NCCL all_gather_into_tensor.
torchrun --nproc-per-node N tools/symm_probe/fused_gather_check.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "openset-rcnn_b200"))
import torch, torch.distributed as dist
from osr_b200 import synth
from osr_b200.dist import FusedEncoderGather, fused_gathered_pln_loss, gathered_pln_loss, reduced_pln_loss, all_gather_rows
from osr_b200.pln import pln_encode_tc

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
R = 8192
pi = synth.make_pln_inputs(R, seed=100 + rank, device=dev)
w = synth.make_pln_inputs(8, seed=7, device=dev)      # same parameters on every rank
enc_w, enc_b, reps = w.enc_w, w.enc_b + 0.01, w.reps
kw = dict(num_known_classes=20, alpha=0.1, beta=0.9, loss_weight=0.5, iou_threshold=0.5)

enc = FusedEncoderGather(R, enc_w.shape[0], dev, multicast=(os.environ.get("OSR_MC", "1") == "1") and None)
emb_loc, emb_all = enc(pi.roi_features, enc_w, enc_b)
ref_loc = pln_encode_tc(pi.roi_features, enc_w, enc_b)
ref_all = all_gather_rows(ref_loc)
torch.cuda.synchronize()
ok = torch.equal(emb_all, ref_all) and torch.equal(emb_loc, ref_loc)

# loss + gradient parity of the two gathered formulations
loss_f, emb_f = fused_gathered_pln_loss(enc, pi.roi_features, enc_w, enc_b, reps.clone().requires_grad_(True),
                                        pi.gt_classes, pi.ious, **kw)
g_f = torch.autograd.grad(loss_f, emb_f)[0]
emb_n = ref_loc.clone().requires_grad_(True)
loss_n = gathered_pln_loss(emb_n, reps.clone().requires_grad_(True), pi.gt_classes, pi.ious, **kw)
g_n = torch.autograd.grad(loss_n, emb_n)[0]
ok = ok and torch.equal(loss_f, loss_n) and torch.equal(g_f, g_n)

# the same loss from per-rank losses + two small all-reduces (no gather): value / gradients within fp32 summation order
emb_r = ref_loc.clone().requires_grad_(True)
reps_r = reps.clone().requires_grad_(True)
loss_r = reduced_pln_loss(emb_r, reps_r, pi.gt_classes, pi.ious, **kw)
g_r, gr_r = torch.autograd.grad(loss_r, [emb_r, reps_r])
reps_n = reps.clone().requires_grad_(True)
emb_n2 = ref_loc.clone().requires_grad_(True)
loss_n2 = gathered_pln_loss(emb_n2, reps_n, pi.gt_classes, pi.ious, **kw)
g_n2, gr_n2 = torch.autograd.grad(loss_n2, [emb_n2, reps_n])
ok_reduced = (torch.allclose(loss_r, loss_n2, rtol=1e-5, atol=1e-7) and torch.allclose(g_r, g_n2, rtol=1e-4, atol=1e-9)
              and torch.allclose(gr_r, gr_n2, rtol=1e-4, atol=1e-8))


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

t_fused = timeit(lambda: enc(pi.roi_features, enc_w, enc_b))
t_nccl = timeit(lambda: all_gather_rows(pln_encode_tc(pi.roi_features, enc_w, enc_b)))
t_enc = timeit(lambda: pln_encode_tc(pi.roi_features, enc_w, enc_b))
def _gathered_step():
    e = ref_loc.clone().requires_grad_(True); r = reps.clone().requires_grad_(True)
    torch.autograd.grad(gathered_pln_loss(e, r, pi.gt_classes, pi.ious, **kw), [e, r])
def _reduced_step():
    e = ref_loc.clone().requires_grad_(True); r = reps.clone().requires_grad_(True)
    torch.autograd.grad(reduced_pln_loss(e, r, pi.gt_classes, pi.ious, **kw), [e, r])
t_gl = timeit(_gathered_step, 20)
t_rl = timeit(_reduced_step, 20)
flag2 = torch.tensor([1 if ok_reduced else 0], device=dev)
dist.all_reduce(flag2, op=dist.ReduceOp.MIN)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world {world} (NVLS multicast stores: {enc.multicast}): fused encoder+gather bit-identical to encoder + NCCL all-gather (incl. loss, grads): {bool(flag.item())}")
    print(f"reduced_pln_loss (per-rank loss + 2 all-reduces) equals the gathered loss (value, grads): {bool(flag2.item())}; loss fwd+bwd ms: gathered (NCCL) {t_gl:.4f} | reduced {t_rl:.4f}")
    print(f"ms (max over ranks): fused {t_fused:.4f}  | encoder + NCCL all_gather {t_nccl:.4f}  | encoder alone {t_enc:.4f}")
dist.destroy_process_group()
