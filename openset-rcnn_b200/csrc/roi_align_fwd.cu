// FPN level assignment + multi-level ROIAlignV2 forward for sm_100a.
// Replaces detectron2 ROIPooler.forward (osrcnn_roi_heads.py:306): ~8 elementwise launches + 4 x
// (nonzero sync, index, torchvision roi_align, index_put) become ONE launch, no host sync.
//
// One CTA per RoI.  Prologue: level assignment (bit-exact op sequence) + the separable weight tables
// Wy[ph][row], Wx[pw][col] of the RoI (roi_geometry.cuh).  Main loop, NCHW maps (the reference layout):
//   stage 1  lanes run along x (coalesced row segments of the RoI footprint); each lane folds the footprint
//            rows of its column into 7 per-ph accumulators for 4 channels at a time:
//            U[c][ph][x] = sum_y Wy[ph][y] * F[c][y][x]          (each feature value is loaded once per ph)
//   stage 2  per warp, through a warp-private shared-memory tile:
//            out[c][ph][pw] = (1/count) * sum_x Wx[pw][x] * U[c][ph][x]
//            consecutive lanes write consecutive (c, ph, pw) => 128-byte coalesced stores of the C-major
//            (M, C, 7, 7) output the box head expects.
// Wide footprints (33..146 columns) use the same scheme one channel at a time; anything larger, or an
// output bin spanning more than kRB rows/cols, takes a generic per-sample path (correctness only).
#include "roi_geometry.cuh"

namespace {

using namespace osr;

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kWarpTile = 1024;  // floats of warp-private staging
constexpr int kKC = 4;           // channels per lane in the fast path

struct FwdParams {
  RoiLevels L;
  const float* rois;
  int M;
  float* out;
  int32_t* out_level;
};

struct Tables {
  float wy[kP * kRB];
  float wx[kP * kRB];
  int yb[kP], ny[kP], xb[kP], nx[kP];
};

template <int LX>
__device__ __forceinline__ void fwd_fast(const LevelDesc& lv, const Tables& T, float* Us, const float* img_base, int C,
                                         int xmin, int wf, float count, float* out_roi) {
  constexpr int G = 32 / LX;
  constexpr int CPW = G * kKC;   // channels per warp iteration
  constexpr int LXP = LX + 1;    // padded row stride of the staging tile
  static_assert(CPW * kP * LXP <= kWarpTile, "warp tile too small");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cg = lane / LX, x = lane % LX;
  const bool xin = x < wf;
  const float* col = img_base + (int64_t)(xmin + (xin ? x : 0)) * lv.sW;

  for (int cbase = warp * CPW; cbase < C; cbase += kWarps * CPW) {
    float U[kKC][kP];
#pragma unroll
    for (int k = 0; k < kKC; ++k)
#pragma unroll
      for (int ph = 0; ph < kP; ++ph) U[k][ph] = 0.f;
    const int c0 = cbase + cg * kKC;
    const float* chan = col + (int64_t)c0 * lv.sC;
    bool cin[kKC];
#pragma unroll
    for (int k = 0; k < kKC; ++k) cin[k] = xin && (c0 + k < C);
#pragma unroll
    for (int ph = 0; ph < kP; ++ph) {
      const int nr = T.ny[ph];
      const float* rowp = chan + (int64_t)T.yb[ph] * lv.sH;
      for (int r = 0; r < nr; ++r) {
        const float w = T.wy[ph * kRB + r];
        float v[kKC];
#pragma unroll
        for (int k = 0; k < kKC; ++k) v[k] = cin[k] ? __ldg(rowp + (int64_t)k * lv.sC) : 0.f;
#pragma unroll
        for (int k = 0; k < kKC; ++k) U[k][ph] = fmaf(w, v[k], U[k][ph]);
        rowp += lv.sH;
      }
    }
#pragma unroll
    for (int k = 0; k < kKC; ++k)
#pragma unroll
      for (int ph = 0; ph < kP; ++ph) Us[((cg * kKC + k) * kP + ph) * LXP + x] = U[k][ph];
    __syncwarp();
    const int nout = min(CPW, C - cbase) * (kP * kP);
    for (int o = lane; o < nout; o += 32) {
      const int cl = o / (kP * kP);
      const int rem = o - cl * (kP * kP);
      const int ph = rem / kP, pw = rem - ph * kP;
      const int nq = T.nx[pw];
      const float* up = Us + (cl * kP + ph) * LXP + (T.xb[pw] - xmin);
      const float* wp = T.wx + pw * kRB;
      float s = 0.f;
      for (int q = 0; q < nq; ++q) s = fmaf(wp[q], up[q], s);
      out_roi[(int64_t)cbase * (kP * kP) + o] = s / count;
    }
    __syncwarp();
  }
}

__device__ __forceinline__ void fwd_wide(const LevelDesc& lv, const Tables& T, float* Us, const float* img_base, int C,
                                         int xmin, int wf, float count, float* out_roi) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < C; c += kWarps) {
    const float* chan = img_base + (int64_t)c * lv.sC + (int64_t)xmin * lv.sW;
    for (int xc = 0; xc < wf; xc += 32) {
      const int x = xc + lane;
      const bool xin = x < wf;
      float U[kP];
#pragma unroll
      for (int ph = 0; ph < kP; ++ph) U[ph] = 0.f;
#pragma unroll
      for (int ph = 0; ph < kP; ++ph) {
        const int nr = T.ny[ph];
        const float* rowp = chan + (int64_t)T.yb[ph] * lv.sH + (int64_t)(xin ? x : 0) * lv.sW;
        for (int r = 0; r < nr; ++r) {
          const float v = xin ? __ldg(rowp) : 0.f;
          U[ph] = fmaf(T.wy[ph * kRB + r], v, U[ph]);
          rowp += lv.sH;
        }
      }
      if (xin) {
#pragma unroll
        for (int ph = 0; ph < kP; ++ph) Us[ph * wf + x] = U[ph];
      }
    }
    __syncwarp();
    for (int o = lane; o < kP * kP; o += 32) {
      const int ph = o / kP, pw = o - ph * kP;
      const int nq = T.nx[pw];
      const float* up = Us + ph * wf + (T.xb[pw] - xmin);
      const float* wp = T.wx + pw * kRB;
      float s = 0.f;
      for (int q = 0; q < nq; ++q) s = fmaf(wp[q], up[q], s);
      out_roi[(int64_t)c * (kP * kP) + o] = s / count;
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kThreads) roi_align_fwd_kernel(const __grid_constant__ FwdParams p) {
  __shared__ Tables T;
  __shared__ __align__(16) float s_U[kWarps * kWarpTile];

  const int m = blockIdx.x;
  const int tid = threadIdx.x;
  const float* roi = p.rois + (int64_t)m * 5;
  const float fimg = __ldg(roi), x1 = __ldg(roi + 1), y1 = __ldg(roi + 2), x2 = __ldg(roi + 3), y2 = __ldg(roi + 4);
  const int img = (int)fimg;
  const int level = assign_level(x1, y1, x2, y2, p.L);
  if (tid == 0) p.out_level[m] = level;
  const int C = p.L.C;
  float* out_roi = p.out + (int64_t)m * C * (kP * kP);

  bool zero = (level < 0) || (level >= p.L.num_levels) || (img < 0) || (img >= p.L.num_images);
  RoiGeom g;
  if (!zero) {
    const LevelDesc& lv0 = p.L.lv[level];
    g = roi_geometry(x1, y1, x2, y2, lv0.scale, p.L.sampling_ratio);
    for (int i = tid; i < kP * kRB; i += kThreads) {
      T.wy[i] = 0.f;
      T.wx[i] = 0.f;
    }
    __syncthreads();
    if (tid < kP) {
      T.ny[tid] = build_bin_weights(g.start_h, g.bin_h, g.grid_h, lv0.H, tid, T.wy + tid * kRB, &T.yb[tid]);
    } else if (tid >= 32 && tid < 32 + kP) {
      const int pw = tid - 32;
      T.nx[pw] = build_bin_weights(g.start_w, g.bin_w, g.grid_w, lv0.W, pw, T.wx + pw * kRB, &T.xb[pw]);
    }
    __syncthreads();
  }
  int path = 3;  // 0 fast, 1 wide, 2 generic, 3 zeros
  int xmin = 0, wf = 0;
  if (!zero) {
    bool overflow = false, anyx = false, anyy = false;
    int xmax = -1;
    xmin = 1 << 30;
#pragma unroll
    for (int i = 0; i < kP; ++i) {
      const int nx = T.nx[i], ny = T.ny[i];
      overflow |= (nx < 0) | (ny < 0);
      if (nx > 0) {
        anyx = true;
        xmin = min(xmin, T.xb[i]);
        xmax = max(xmax, T.xb[i] + nx - 1);
      }
      anyy |= ny > 0;
    }
    wf = xmax - xmin + 1;
    if (overflow) path = 2;
    else if (!anyx || !anyy) path = 3;
    else if (wf <= 32) path = 0;
    else if (wf * kP <= kWarpTile) path = 1;
    else path = 2;
  }

  if (path == 3) {
    for (int o = tid; o < C * kP * kP; o += kThreads) out_roi[o] = 0.f;
    return;
  }
  const LevelDesc& lv = p.L.lv[level];
  const float* img_base = lv.data + (int64_t)img * lv.sN;
  float* Us = s_U + (tid >> 5) * kWarpTile;
  if (path == 0) {
    if (wf <= 8) fwd_fast<8>(lv, T, Us, img_base, C, xmin, wf, g.count, out_roi);
    else if (wf <= 16) fwd_fast<16>(lv, T, Us, img_base, C, xmin, wf, g.count, out_roi);
    else fwd_fast<32>(lv, T, Us, img_base, C, xmin, wf, g.count, out_roi);
  } else if (path == 1) {
    fwd_wide(lv, T, Us, img_base, C, xmin, wf, g.count, out_roi);
  } else {
    // generic: torchvision's per-sample loop, one output element per thread iteration
    for (int o = tid; o < C * kP * kP; o += kThreads) {
      const int c = o / (kP * kP);
      const int rem = o - c * (kP * kP);
      const int ph = rem / kP, pw = rem - ph * kP;
      const float* plane = img_base + (int64_t)c * lv.sC;
      float s = 0.f;
      for (int iy = 0; iy < g.grid_h; ++iy) {
        const float y = g.start_h + ph * g.bin_h + (iy + 0.5f) * g.bin_h / (float)g.grid_h;
        for (int ix = 0; ix < g.grid_w; ++ix) {
          const float x = g.start_w + pw * g.bin_w + (ix + 0.5f) * g.bin_w / (float)g.grid_w;
          s += bilinear_sample(plane, lv.sH, lv.sW, lv.H, lv.W, y, x);
        }
      }
      out_roi[o] = s / g.count;
    }
  }
}

}  // namespace

namespace osr {
int fill_roi_levels(RoiLevels& L, const osr_feat_level_t* h_levels, int num_levels, int num_images, int C, int P,
                    int sampling_ratio, int aligned, int canonical_box_size, int canonical_level, int min_level) {
  if (!h_levels || num_levels <= 0 || num_levels > OSR_MAX_LEVELS)
    return fail_arg(OSR_E_ARG, "roi_align: num_levels=%d outside [1,%d]", num_levels, OSR_MAX_LEVELS);
  if (P != kP) return fail_arg(OSR_E_SHAPE, "roi_align: pooler resolution %d unsupported (kernels are built for 7x7)", P);
  if (!aligned) return fail_arg(OSR_E_ARG, "roi_align: only aligned=1 (ROIAlignV2) is supported");
  if (C <= 0 || num_images < 0 || sampling_ratio < 0 || canonical_box_size <= 0)
    return fail_arg(OSR_E_ARG, "roi_align: bad C / num_images / sampling_ratio / canonical_box_size");
  L.num_levels = num_levels;
  L.num_images = num_images;
  L.C = C;
  L.sampling_ratio = sampling_ratio;
  L.inv_canonical_size = 1.0f / (float)canonical_box_size;
  L.canonical_level = canonical_level;
  L.min_level = min_level;
  L.max_level = min_level + num_levels - 1;
  for (int l = 0; l < num_levels; ++l) {
    const osr_feat_level_t& h = h_levels[l];
    if (!h.data || h.H <= 0 || h.W <= 0) return fail_arg(OSR_E_ARG, "roi_align: level %d null data or empty map", l);
    L.lv[l].data = h.data;
    L.lv[l].sN = h.sN; L.lv[l].sC = h.sC; L.lv[l].sH = h.sH; L.lv[l].sW = h.sW;
    L.lv[l].H = h.H; L.lv[l].W = h.W;
    L.lv[l].scale = h.scale;
  }
  return 0;
}
}  // namespace osr

extern "C" int osr_roi_align_fwd(const osr_feat_level_t* h_levels, int num_levels, int num_images, int C,
                                 const float* rois, int M, int P, int sampling_ratio, int aligned,
                                 int canonical_box_size, int canonical_level, int min_level, float* out,
                                 int32_t* out_level, void* stream) {
  FwdParams p;
  int rc = osr::fill_roi_levels(p.L, h_levels, num_levels, num_images, C, P, sampling_ratio, aligned,
                                canonical_box_size, canonical_level, min_level);
  if (rc) return rc;
  if (M < 0) return osr::fail_arg(OSR_E_ARG, "roi_align_fwd: M < 0");
  if (M == 0) return 0;
  if (!rois || !out || !out_level) return osr::fail_arg(OSR_E_ARG, "roi_align_fwd: null pointer argument");
  p.rois = rois;
  p.M = M;
  p.out = out;
  p.out_level = out_level;
  roi_align_fwd_kernel<<<M, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  OSR_LAUNCH_CHECK();
  return 0;
}
