"""Where the cfg 4 (inference-only) step spends its time: wall clock per section with a synchronize on both sides, and the
torch profiler's kernel / CPU-op totals for one step.  usage: python tools/infer_time.py [--images 32]"""
import argparse
import sys
import time

import torch

sys.path.insert(0, "openset-rcnn_b200")
from osr_b200 import pipeline  # noqa: E402
from osr_b200.inference import inference, softmax_classifier_inference  # noqa: E402
from osr_b200.proposals import predict_proposals  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--images", type=int, default=32)
ap.add_argument("--profile", type=int, default=1)
args = ap.parse_args()
cfg = pipeline.make_config("cfg4", num_images=args.images)
st = pipeline.InferencePathStep(cfg)
for _ in range(3):
    st.step()
torch.cuda.synchronize()


def section(f, n=10):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = f()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    return out, ts[len(ts) // 2]


with torch.no_grad():
    proposals, t1 = section(lambda: predict_proposals(st.anchors, st.deltas, st.ctr, st.image_sizes, nms_thresh=cfg.rpn_nms_thresh,
                                                      pre_nms_topk=cfg.pre_nms_topk, post_nms_topk=cfg.post_nms_topk,
                                                      training=False, mode="nominal"))
    boxes = [x.proposal_boxes for x in proposals]
    (pooled, lvl), t3 = section(lambda: st.pooler.forward_with_levels(st.feats, boxes))
    M = pooled.shape[0]
    (fg, _), t6a = section(lambda: inference((st.pred_deltas[:M], st.pred_iou[:M]), proposals, st.box_features[:M], score_thresh=0.05,
                                             nms_thresh=1.0, topk_per_image=1000))
    fg2, t6b = section(lambda: st.pln.inference(fg))
    dets, t6c = section(lambda: softmax_classifier_inference(fg2, st.cls_score, unknown_id=80, known_score_thresh=0.05,
                                                             known_nms_thresh=0.5, known_topk=50, unknown_score_thresh=0.0,
                                                             unknown_nms_thresh=0.5, unknown_topk=50))
    _, tall = section(lambda: st.step())
print(f"M={M} fg={sum(len(x) for x in fg)} dets={sum(len(x) for x in dets)}")
print(f"wall ms (median, sync both sides): predict_proposals {t1:.3f} | pooler {t3:.3f} | inference {t6a:.3f} | "
      f"pln.inference {t6b:.3f} | softmax_classifier_inference {t6c:.3f} | whole step {tall:.3f}")
if args.profile:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        st.step()
        torch.cuda.synchronize()
    ka = prof.key_averages()
    print(ka.table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
    print(ka.table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=60))
