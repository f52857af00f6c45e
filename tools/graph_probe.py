"""Which stage of the step breaks CUDA-graph capture?  (debug helper)"""
import os, sys, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "openset-rcnn_b200")); sys.path.insert(0, ROOT)
import torch
from osr_b200 import _lib
from osr_b200.pipeline import PathConfig, RoiPathStep
from osr_b200.proposals import rpn_select_decode
from osr_b200.sampling import match_proposals
from osr_b200.pln import pln_encode_tc, pln_loss_from_emb

dev = torch.device("cuda:0")
cfg = PathConfig(num_images=4)
path = RoiPathStep(cfg, dev)
for _ in range(2):
    path.step()
torch.cuda.synchronize()

def try_capture(name, fn, mode="thread_local"):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, capture_error_mode=mode):
            fn()
        g.replay(); torch.cuda.synchronize()
        print(name, mode, "OK", flush=True)
    except Exception as e:
        print(name, mode, "FAILED:", repr(e)[:160], "| osr:", _lib.lib().osr_last_error(), flush=True)
        try:
            torch.cuda.synchronize()
        except Exception as e2:
            print("  sync after failure:", repr(e2)[:200])

sel = [None]
def s1():
    sel[0] = rpn_select_decode(path.anchors, path.deltas, path.ctr, path.image_hw_dev, cfg.pre_nms_topk)
def s2():
    s = sel[0]
    match_proposals(s.boxes.view(-1, 4), path.prop_off, path.gt_boxes, path.gt_classes, path.gt_off, path.kmax,
                    iou_threshold=0.5, background_label=81, box_counts=s.counts[:, path.count_col], box_counts_stride=s.counts.shape[1])
rois = path.last["rois"].detach()
def s3():
    feats = [f.requires_grad_(True) for f in path.feats]
    pooled, lvl = path.pooler.pool_rois(feats, rois, path.roi_offsets)
    torch.autograd.grad(pooled, feats, path.grad_pooled)
def s5():
    pi = path.pln
    reps = pi.reps.requires_grad_(True)
    emb = pln_encode_tc(pi.roi_features, pi.enc_w, pi.enc_b).requires_grad_(True)
    loss = pln_loss_from_emb(emb, reps, pi.gt_classes, pi.ious, num_known_classes=20)
    torch.autograd.grad(loss, [emb, reps])

def s3f():
    path.pooler.pool_rois([f.detach() for f in path.feats], rois, path.roi_offsets)
def s3_nograd_bwd():
    from osr_b200.poolers import _feat_levels
    import ctypes
    lib = _lib.lib()
    grads = [torch.empty_like(f) for f in path.feats]
    arr, N, C = _feat_levels(grads, path.pooler.scales)
    M = rois.shape[0]
    ws = torch.empty(int(lib.osr_roi_align_bwd_workspace(arr, 4, N, C, M)), dtype=torch.uint8, device=dev)
    rc = lib.osr_roi_align_bwd(arr, 4, N, C, path.grad_pooled.data_ptr(), rois.data_ptr(), path.roi_offsets.data_ptr(), M, 7, 0, 1,
                               224, 4, 2, ws.data_ptr(), ws.numel(), _lib.stream_ptr(dev))
    _lib.check(rc, "bwd")
def s5enc():
    pi = path.pln
    pln_encode_tc(pi.roi_features, pi.enc_w, pi.enc_b)
emb0 = pln_encode_tc(path.pln.roi_features, path.pln.enc_w, path.pln.enc_b).detach()
def s5loss():
    pi = path.pln
    pln_loss_from_emb(emb0, pi.reps.detach(), pi.gt_classes, pi.ious, num_known_classes=20)
def s5lossbwd():
    pi = path.pln
    e = emb0.clone().requires_grad_(True)
    loss = pln_loss_from_emb(e, pi.reps.detach(), pi.gt_classes, pi.ious, num_known_classes=20)
    torch.autograd.grad(loss, [e])
def plain_autograd():
    x = torch.randn(64, 64, device=dev, requires_grad=True)
    torch.autograd.grad((x @ x).sum(), [x])
table = dict(s1=s1, s2=s2, s3=s3, s3f=s3f, s3b=s3_nograd_bwd, s5=s5, s5enc=s5enc, s5loss=s5loss, s5lossbwd=s5lossbwd,
             plain=plain_autograd, step=path.step)
s1()
for name in sys.argv[1:]:
    try_capture(name, table[name], "thread_local")
