// Proposal <-> ground-truth matching for the ROI-head sampling glue.
// Replaces, per image, detectron2's  pairwise_iou(gt, proposals)  (a G x P matrix),  Matcher([0.5], [0, 1])
// (max / argmax over G, threshold), the gather  M[matched_idxs, arange(P)]  added by the reference
// (osrcnn_roi_heads.py:187-193) and the class assignment of ROIHeads._sample_proposals (gt_classes[matched_idxs],
// background where unmatched) - called from label_and_sample_proposals (osrcnn_roi_heads.py:136-230).
// One thread per proposal, the image's GT boxes staged in shared memory; the G x P matrix is never materialised.
// Arithmetic is torch's elementwise chain with every op rounded separately (no FMA contraction):
//   area = (x2-x1)*(y2-y1);  w = max(min(x2a,x2b) - max(x1a,x1b), 0);  inter = w*h;
//   iou = inter > 0 ? inter / ((area_gt + area_p) - inter) : 0;   arg-max = FIRST maximal GT (torch.max semantics).
// => matched index, IoU and label are bit-exact against the reference ops on CPU and GPU.
#include "osr_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kGtChunk = 256;

struct MatchParams {
  const float* boxes;
  const int32_t* box_off;
  const int32_t* box_cnt;   // optional: image n owns [off[n], off[n] + cnt[n * cnt_stride]) (padded layouts)
  int cnt_stride;
  const float* gt_boxes;
  const int64_t* gt_classes;
  const int32_t* gt_off;
  float thr;
  int64_t bg;
  int32_t* midx;
  float* miou;
  int32_t* mlabel;
  int64_t* mclass;
};

__global__ void __launch_bounds__(kThreads) match_label_kernel(const __grid_constant__ MatchParams p) {
  __shared__ float4 s_gt[kGtChunk];
  __shared__ float s_area[kGtChunk];
  const int n = blockIdx.y;
  const int b0 = p.box_off[n], b1 = p.box_cnt ? b0 + p.box_cnt[n * p.cnt_stride] : p.box_off[n + 1];
  const int g0 = p.gt_off[n], g1 = p.gt_off[n + 1];
  const int c = b0 + blockIdx.x * kThreads + threadIdx.x;
  if (blockIdx.x * kThreads >= b1 - b0) return;   // whole block past this image's proposals
  const bool live = c < b1;
  float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
  if (live) box = __ldg(reinterpret_cast<const float4*>(p.boxes) + c);
  const float area_p = __fmul_rn(__fsub_rn(box.z, box.x), __fsub_rn(box.w, box.y));
  float best = 0.f;
  int best_g = 0;
  bool first = true;
  for (int base = g0; base < g1; base += kGtChunk) {
    const int ng = min(kGtChunk, g1 - base);
    __syncthreads();
    if (threadIdx.x < ng) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(p.gt_boxes) + base + threadIdx.x);
      s_gt[threadIdx.x] = g;
      s_area[threadIdx.x] = __fmul_rn(__fsub_rn(g.z, g.x), __fsub_rn(g.w, g.y));
    }
    __syncthreads();
    for (int k = 0; k < ng; ++k) {
      const float4 g = s_gt[k];
      const float w = fmaxf(__fsub_rn(fminf(g.z, box.z), fmaxf(g.x, box.x)), 0.f);
      const float h = fmaxf(__fsub_rn(fminf(g.w, box.w), fmaxf(g.y, box.y)), 0.f);
      const float inter = __fmul_rn(w, h);
      const float iou = inter > 0.f ? __fdiv_rn(inter, __fsub_rn(__fadd_rn(s_area[k], area_p), inter)) : 0.f;
      if (first || iou > best) {   // strict >: the first maximal GT wins, as torch.max(dim=0)
        best = iou;
        best_g = base - g0 + k;
        first = false;
      }
    }
  }
  if (!live) return;
  const bool has_gt = g1 > g0;
  const int label = (has_gt && best >= p.thr) ? 1 : 0;   // Matcher: thresholds [-inf, thr, inf] -> labels [0, 1]
  p.midx[c] = best_g;
  p.miou[c] = best;
  p.mlabel[c] = label;
  p.mclass[c] = label ? __ldg(p.gt_classes + g0 + best_g) : p.bg;
}

}  // namespace

extern "C" int osr_match_label(const float* boxes, const int32_t* box_offsets, const int32_t* box_counts,
                               int box_counts_stride, const float* gt_boxes,
                               const int64_t* gt_classes, const int32_t* gt_offsets, int num_images,
                               int max_boxes_per_image, float iou_threshold, int64_t background_label,
                               int32_t* matched_idx, float* matched_iou, int32_t* matched_label,
                               int64_t* matched_class, void* stream) {
  osr::DeviceGuard device_guard(matched_idx);
  if (num_images < 0 || max_boxes_per_image < 0) return osr::fail_arg(OSR_E_ARG, "match_label: negative size");
  if (num_images == 0 || max_boxes_per_image == 0) return 0;
  if (!boxes || !box_offsets || !gt_offsets || !matched_idx || !matched_iou || !matched_label || !matched_class)
    return osr::fail_arg(OSR_E_ARG, "match_label: null pointer argument");
  if ((reinterpret_cast<uintptr_t>(boxes) & 15) || (gt_boxes && (reinterpret_cast<uintptr_t>(gt_boxes) & 15)))
    return osr::fail_arg(OSR_E_ARG, "match_label: box arrays must be 16-byte aligned");
  MatchParams p{boxes, box_offsets, box_counts, box_counts_stride, gt_boxes, gt_classes, gt_offsets, iou_threshold, background_label,
                matched_idx, matched_iou, matched_label, matched_class};
  dim3 grid(osr::ceil_div(max_boxes_per_image, kThreads), num_images);
  match_label_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  OSR_LAUNCH_CHECK();
  return 0;
}
