// ROI-head inference post-processing, stage 1: box decode + objectness score + validity,
// one thread per proposal, all images in one launch.  Replaces
//   OpensetFastRCNNOutputLayers.predict_boxes  -> detectron2 Box2BoxTransform(weights).apply_deltas  (osrcnn_fast_rcnn.py:423)
//   OpensetFastRCNNOutputLayers.predict_ious   -> sqrt(iou * centerness) | (iou + centerness) / 2    (:443-450)
//   fast_rcnn_inference_single_image           -> isfinite filter, Boxes.clip, score > thresh         (:106-126)
// Arithmetic follows torch's CUDA elementwise chain op by op (no FMA contraction):  d / w is a multiply by the fp32
// reciprocal (ATen divides by a scalar that way), dw clamped to log(1000/16), exp = expf, centre = x1 + 0.5 * w,
// pred = d * size + centre as two rounded ops.  Rows that the reference would drop (non-finite box or score, or
// score <= thresh) get score -inf and a zero box: they sort behind every survivor, so the class-agnostic NMS that
// follows (osr_nms_segmented over whole images) returns the survivors first, in the reference's order.
#include <math_constants.h>

#include "osr_common.cuh"

namespace {

constexpr int kThreads = 256;

struct DecodeParams {
  const float* boxes;
  const float* deltas;
  const float* ious;
  const float* ctr;
  const int32_t* off;
  const int32_t* image_hw;
  float rwx, rwy, rww, rwh, clampv, thr;
  int geometric;
  float* out_boxes;
  float* out_scores;
  float* out_eff;
};

__global__ void __launch_bounds__(kThreads) rcnn_decode_kernel(const __grid_constant__ DecodeParams p) {
  const int n = blockIdx.y;
  const int b0 = p.off[n], b1 = p.off[n + 1];
  const int i = b0 + blockIdx.x * kThreads + threadIdx.x;
  if (i >= b1) return;
  const float4 b = __ldg(reinterpret_cast<const float4*>(p.boxes) + i);
  const float4 d = __ldg(reinterpret_cast<const float4*>(p.deltas) + i);
  const float w = __fsub_rn(b.z, b.x), h = __fsub_rn(b.w, b.y);
  const float cx = __fadd_rn(b.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(b.y, __fmul_rn(0.5f, h));
  const float dx = __fmul_rn(d.x, p.rwx), dy = __fmul_rn(d.y, p.rwy);
  float dw = __fmul_rn(d.z, p.rww), dh = __fmul_rn(d.w, p.rwh);
  dw = dw > p.clampv ? p.clampv : dw;   // torch.clamp(max=): NaN stays NaN (the row is then dropped as non-finite)
  dh = dh > p.clampv ? p.clampv : dh;
  const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
  const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
  float x1 = __fsub_rn(pcx, __fmul_rn(0.5f, pw)), y1 = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
  float x2 = __fadd_rn(pcx, __fmul_rn(0.5f, pw)), y2 = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
  const float iou = __ldg(p.ious + i), c = __ldg(p.ctr + i);
  const float score = p.geometric ? __fsqrt_rn(__fmul_rn(iou, c)) : __fmul_rn(__fadd_rn(iou, c), 0.5f);
  const bool finite = isfinite(x1) && isfinite(y1) && isfinite(x2) && isfinite(y2) && isfinite(score);
  const float ih = (float)p.image_hw[2 * n], iw = (float)p.image_hw[2 * n + 1];
  x1 = fminf(fmaxf(x1, 0.f), iw); y1 = fminf(fmaxf(y1, 0.f), ih);
  x2 = fminf(fmaxf(x2, 0.f), iw); y2 = fminf(fmaxf(y2, 0.f), ih);
  const bool keep = finite && (score > p.thr);
  reinterpret_cast<float4*>(p.out_boxes)[i] = keep ? make_float4(x1, y1, x2, y2) : make_float4(0.f, 0.f, 0.f, 0.f);
  p.out_scores[i] = finite ? score : CUDART_NAN_F;   // NaN marks the rows the reference's isfinite filter removes
  p.out_eff[i] = keep ? score : -CUDART_INF_F;
}

}  // namespace

extern "C" int osr_rcnn_decode_score(const float* proposal_boxes, const float* deltas, const float* ious,
                                     const float* centerness, const int32_t* box_offsets, const int32_t* image_hw,
                                     int num_images, int max_boxes_per_image, float wx, float wy, float ww, float wh,
                                     float scale_clamp, int geometric_mean, float score_thresh, float* out_boxes,
                                     float* out_scores, float* out_effective_scores, void* stream) {
  osr::DeviceGuard device_guard(out_boxes);
  if (num_images < 0 || max_boxes_per_image < 0) return osr::fail_arg(OSR_E_ARG, "rcnn_decode: negative size");
  if (num_images == 0 || max_boxes_per_image == 0) return 0;
  if (!proposal_boxes || !deltas || !ious || !centerness || !box_offsets || !image_hw || !out_boxes || !out_scores ||
      !out_effective_scores)
    return osr::fail_arg(OSR_E_ARG, "rcnn_decode: null pointer argument");
  if ((reinterpret_cast<uintptr_t>(proposal_boxes) & 15) || (reinterpret_cast<uintptr_t>(deltas) & 15) ||
      (reinterpret_cast<uintptr_t>(out_boxes) & 15))
    return osr::fail_arg(OSR_E_ARG, "rcnn_decode: box / delta arrays must be 16-byte aligned");
  DecodeParams p;
  p.boxes = proposal_boxes; p.deltas = deltas; p.ious = ious; p.ctr = centerness; p.off = box_offsets;
  p.image_hw = image_hw;
  p.rwx = 1.0f / wx; p.rwy = 1.0f / wy; p.rww = 1.0f / ww; p.rwh = 1.0f / wh;   // ATen: tensor / scalar = tensor * (1 / scalar)
  p.clampv = scale_clamp; p.thr = score_thresh; p.geometric = geometric_mean;
  p.out_boxes = out_boxes; p.out_scores = out_scores; p.out_eff = out_effective_scores;
  dim3 grid(osr::ceil_div(max_boxes_per_image, kThreads), num_images);
  rcnn_decode_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  OSR_LAUNCH_CHECK();
  return 0;
}
