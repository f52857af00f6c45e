import sys, torch
sys.path.insert(0, "openset-rcnn_b200")
from osr_b200 import synth
from osr_b200.pln import pln_loss_fwd_bwd, pln_encode_tc, _pln_fwd, _dist_code
pi = synth.make_pln_inputs(8192, num_known=20, num_classes=81, seed=1, device="cuda:0")
emb = pln_encode_tc(pi.roi_features, pi.enc_w, pi.enc_b)
kw = dict(num_known_classes=20, alpha=0.1, beta=0.9, loss_weight=0.5, iou_threshold=0.5)
def timeit(f, n=200):
    for _ in range(10): f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
print("encode us", timeit(lambda: pln_encode_tc(pi.roi_features, pi.enc_w, pi.enc_b)))
print("fwd_bwd us", timeit(lambda: pln_loss_fwd_bwd(emb, pi.reps, pi.gt_classes, pi.ious, **kw)))
cfg = (20, 1, 0.1, 0.9, 0.5, 0.5, None, 1.0, 1.0, 0)
print("fwd only us", timeit(lambda: _pln_fwd(emb, pi.reps, pi.gt_classes, pi.ious, cfg)))
print("empty-ish op us", timeit(lambda: torch.empty(16, device="cuda:0")))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = pln_loss_fwd_bwd(emb, pi.reps, pi.gt_classes, pi.ious, **kw)
print("fwd_bwd graph us", timeit(lambda: g.replay()))
g2 = torch.cuda.CUDAGraph()
with torch.cuda.graph(g2):
    e2 = pln_encode_tc(pi.roi_features, pi.enc_w, pi.enc_b)
print("encode graph us", timeit(lambda: g2.replay()))
