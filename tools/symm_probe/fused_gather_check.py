"""N-GPU check + timing of the fused encoder -> all-gather kernel against encoder + NCCL all_gather_into_tensor.
torchrun --nproc-per-node N tools/symm_probe/fused_gather_check.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "openset-rcnn_b200"))
import torch, torch.distributed as dist
from osr_b200 import synth
from osr_b200.dist import FusedEncoderGather, fused_gathered_pln_loss, gathered_pln_loss, all_gather_rows
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
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world {world} (NVLS multicast stores: {enc.multicast}): fused encoder+gather bit-identical to encoder + NCCL all-gather (incl. loss, grads): {bool(flag.item())}")
    print(f"ms (max over ranks): fused {t_fused:.4f}  | encoder + NCCL all_gather {t_nccl:.4f}  | encoder alone {t_enc:.4f}")
dist.destroy_process_group()
