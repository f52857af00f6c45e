// RoI geometry shared by the ROIAlign forward and backward kernels: FPN level assignment
// (detectron2 assign_boxes_to_levels) and the per-RoI *separable* interpolation tables.
//
// torchvision's roi_align (aligned=True) computes, per channel,
//     out[ph][pw] = (1/count) * sum_{iy,ix} bilinear(F, y(ph,iy), x(pw,ix))
// Bilinear weights are products (hy|ly) * (hx|lx), the sample grid is a Cartesian product and the
// out-of-range test is a product of a y- and an x-test, so this regroups EXACTLY (in real arithmetic) to
//     out[ph][pw] = (1/count) * sum_y sum_x Wy[ph][y] * Wx[pw][x] * F[y][x]
// with Wy[ph][y] = sum over valid samples iy of (hy at y_low, ly at y_high).  Each feature pixel is then
// read once per output row instead of up to 4*grid^2 times, and the backward is the transposed operator
// built from the same tables (so it is the exact adjoint of the forward).
#pragma once

#include "osr_common.cuh"

namespace osr {

constexpr int kP = 7;          // pooler resolution handled by the fast kernels
constexpr int kRB = 48;        // max feature rows (cols) one output bin may span in the fast path

struct LevelDesc {
  float* data;
  int64_t sN, sC, sH, sW;
  int H, W;
  float scale;
};

struct RoiLevels {
  LevelDesc lv[OSR_MAX_LEVELS];
  int num_levels;
  int num_images;
  int C;
  int sampling_ratio;
  float inv_canonical_size;  // 1.0f / canonical_box_size (ATen CUDA divides by a scalar as a multiply)
  int canonical_level, min_level, max_level;
};

// detectron2 assign_boxes_to_levels, with the fp32 op sequence torch executes on CUDA:
// sqrt(area) * (1/224) + 1e-8 -> log2 -> + canonical_level -> floor -> clamp -> - min_level.
// Returns -1 for NaN sizes (negative area): torch's NaN -> int64 cast matches no level, the RoI pools to zeros.
__device__ __forceinline__ int assign_level(float x1, float y1, float x2, float y2, const RoiLevels& P) {
  const float area = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
  const float s = __fsqrt_rn(area);
  const float t = __fadd_rn(__fmul_rn(s, P.inv_canonical_size), 1e-8f);
  float lv = floorf(__fadd_rn((float)P.canonical_level, log2f(t)));
  if (lv != lv) return -1;
  lv = fminf(fmaxf(lv, (float)P.min_level), (float)P.max_level);
  return (int)lv - P.min_level;
}

struct RoiGeom {
  float start_w, start_h, bin_w, bin_h;
  int grid_w, grid_h;
  float count;
};

__device__ __forceinline__ RoiGeom roi_geometry(float x1, float y1, float x2, float y2, float scale, int sampling_ratio) {
  RoiGeom g;
  g.start_w = x1 * scale - 0.5f;
  g.start_h = y1 * scale - 0.5f;
  const float end_w = x2 * scale - 0.5f;
  const float end_h = y2 * scale - 0.5f;
  const float roi_w = end_w - g.start_w;
  const float roi_h = end_h - g.start_h;
  g.bin_w = roi_w / (float)kP;
  g.bin_h = roi_h / (float)kP;
  g.grid_h = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_h / (float)kP);
  g.grid_w = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(roi_w / (float)kP);
  const int c = g.grid_h * g.grid_w;
  g.count = (float)(c > 1 ? c : 1);
  return g;
}

// One axis, one output bin p: accumulate the separable weights of its `grid` samples into wtab[0..kRB)
// (pre-zeroed), relative to the first touched row `base`.  Returns rows touched (0 = no valid sample),
// or -1 if the bin spans more than kRB rows (caller falls back to the generic path).
__device__ __forceinline__ int build_bin_weights(float start, float bin, int grid, int L, int p, float* wtab, int* base) {
  int b = 0, maxr = -1;
  bool any = false;
  for (int i = 0; i < grid; ++i) {
    float c = start + p * bin + (i + 0.5f) * bin / (float)grid;
    if (c < -1.0f || c > (float)L) continue;
    if (c <= 0.f) c = 0.f;
    int lo = (int)c, hi;
    if (lo >= L - 1) {
      hi = lo = L - 1;
      c = (float)lo;
    } else {
      hi = lo + 1;
    }
    const float l = c - (float)lo;
    const float h = 1.f - l;
    if (!any) {
      any = true;
      b = lo;
    }
    const int r = lo - b;
    if (r + (hi - lo) >= kRB) return -1;
    wtab[r] += h;
    wtab[r + (hi - lo)] += l;
    maxr = r + (hi - lo);
  }
  *base = b;
  return maxr + 1;
}

// torchvision bilinear_interpolate (generic fallback path only)
__device__ __forceinline__ float bilinear_sample(const float* plane, int64_t sH, int64_t sW, int H, int W, float y, float x) {
  if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) return 0.f;
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  int y_low = (int)y, x_low = (int)x, y_high, x_high;
  if (y_low >= H - 1) {
    y_high = y_low = H - 1;
    y = (float)y_low;
  } else {
    y_high = y_low + 1;
  }
  if (x_low >= W - 1) {
    x_high = x_low = W - 1;
    x = (float)x_low;
  } else {
    x_high = x_low + 1;
  }
  const float ly = y - y_low, lx = x - x_low, hy = 1.f - ly, hx = 1.f - lx;
  const float v1 = __ldg(plane + y_low * sH + x_low * sW), v2 = __ldg(plane + y_low * sH + x_high * sW);
  const float v3 = __ldg(plane + y_high * sH + x_low * sW), v4 = __ldg(plane + y_high * sH + x_high * sW);
  return hy * hx * v1 + hy * lx * v2 + ly * hx * v3 + ly * lx * v4;
}

}  // namespace osr
