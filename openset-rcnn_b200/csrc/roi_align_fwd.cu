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
#include <cuda.h>
#include <stdlib.h>
#include <string.h>  // CUtensorMap types only; cuTensorMapEncodeTiled is resolved at run time (no libcuda link dependency)

#include <cuda_bf16.h>

#include "roi_geometry.cuh"

namespace {

using namespace osr;

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kWarpTile = 1024;  // floats of warp-private staging
constexpr int kKC = 4;           // channels per lane in the fast path

constexpr int kMaxRows = kP * kRB;  // footprint rows the shared-row table can describe

struct FwdParams {
  RoiLevels L;
  const float* rois;
  int M;
  float* out;
  unsigned short* out_bf16;   // non-null: write the pooled tile as bf16 (round to nearest even) INSTEAD of fp32 (NHWC kernel)
  int32_t* out_level;
  const int32_t* order;  // (M) processing order (RoIs sorted by image, level, y band) or nullptr
  float* rec;            // (M, kRecFloats) per-RoI table records written by roi_fwd_prep_kernel, or nullptr
  int tma_ok[OSR_MAX_LEVELS];  // level's map satisfies the TMA rules AND the TMA path is enabled for this call
  int ring_floats;             // floats of the staging ring at the start of dynamic shared memory
  int two_rows;                // NHWC kernel: 2 footprint rows per row-loop iteration
  int max_stages;              // NHWC kernel: most row stages the ring is cut into (<= kNhwcMaxStages)
  int* counter;                // persistent NHWC kernel: next unclaimed record (set to the grid size by roi_fwd_prep_kernel)
  int pers_grid;               // grid size of the persistent NHWC kernel (0: not used)
};

// TMA staging geometry: a stage holds the WHOLE footprint (<= kTRmax rows x 32 columns) of kTC channels of one image;
// it is filled by (rows/8) x kTC boxes of 32 x 8 floats that complete on one mbarrier.  2-stage ring.
constexpr int kTX = 32, kTR = 8, kTC = 8;
constexpr int kTRmax = 24;                       // footprints taller than this use the LDG path
constexpr int kChanFloats = kTRmax * kTX;        // 768 floats per channel plane of a stage
constexpr int kStageFloats = kTC * kChanFloats;  // 24 KB
constexpr int kNS = 2;
constexpr int kRingFloats = kNS * kStageFloats;  // 48 KB (aliases the LDG paths' 32 KB of warp tiles)
static_assert(kTC == kWarps, "one warp per channel of a stage");
static_assert(kRingFloats >= kWarps * kWarpTile, "ring must cover the LDG staging tiles");

struct alignas(64) FwdTma {
  CUtensorMap map[OSR_MAX_LEVELS];
};

struct Tables {
  float wy[kP * kRB];
  float wx[kP * kRB];
  int yb[kP], ny[kP], xb[kP], nx[kP];
  // shared-row form of the y tables: footprint row r (0 = ymin) is loaded ONCE and folded into the (up to 3)
  // consecutive output bins that contain it.  own_b/own_e[ph] = rows whose FIRST bin is ph.
  float4 rw[kMaxRows];  // (w for bin ph0, ph0+1, ph0+2, unused)
  int own_b[kP], own_e[kP];
  int ymin, hf, shared_ok;
  float4 wt[kP][2];   // x-tap weights of each bin padded to 8 taps (NHWC kernel)
  // the same weights for the packed-fp32x2 row loop: bins are processed as PAIRS (0,1) (2,3) (4,5) (6,-);
  // wt2[pp][t2] = (w[2pp][2 t2], w[2pp+1][2 t2], w[2pp][2 t2 + 1], w[2pp+1][2 t2 + 1]); the missing 8th bin has zero weights
  float4 wt2[4][4];
};

// fold one loaded value into bins PH, PH+1, PH+2 (indices are compile-time after unrolling; guards keep them in range)
#define OSR_FOLD(K, PH, W, V)                                                                       \
  do {                                                                                              \
    U[K][PH] = fmaf((W).x, (V), U[K][PH]);                                                          \
    if ((PH) + 1 < kP) U[K][(PH) + 1 < kP ? (PH) + 1 : 0] = fmaf((W).y, (V), U[K][(PH) + 1 < kP ? (PH) + 1 : 0]); \
    if ((PH) + 2 < kP) U[K][(PH) + 2 < kP ? (PH) + 2 : 0] = fmaf((W).z, (V), U[K][(PH) + 2 < kP ? (PH) + 2 : 0]); \
  } while (0)

// Fast path (footprint at most 32 columns wide).  Instruction-lean on purpose: per-image element offsets are
// 32-bit (a (C,H,W) pyramid level has < 2^31 elements), loads are unpredicated (out-of-footprint lanes re-read a
// clamped in-footprint address and their results are never consumed), and every lane owns two fixed (ph, pw)
// outputs for the whole RoI so stage 2 has no divisions and its x-weights live in registers.
template <int LX>
__device__ __forceinline__ void fwd_fast(const LevelDesc& lv, const Tables& T, float* Us, const float* img_base, int C,
                                         int xmin, int wf, float inv_count, float* out_roi) {
  constexpr int G = 32 / LX;
  constexpr int CPW = G * kKC;   // channels per warp iteration
  constexpr int LXP = LX + 1;    // padded row stride of the staging tile
  static_assert(CPW * kP * LXP <= kWarpTile, "warp tile too small");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cg = lane / LX, x = lane % LX;
  const uint32_t sC = (uint32_t)lv.sC, sH = (uint32_t)lv.sH;
  const uint32_t off_x = (uint32_t)(xmin + min(x, wf - 1)) * (uint32_t)lv.sW + (uint32_t)T.ymin * sH;

  // stage-2 ownership: outputs o = lane and o = lane + 32 (< 49) of every channel
  int so[2], snq[2];
  float sw[2][4];
  bool sslow = false;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int o = lane + 32 * t;
    const int oo = o < kP * kP ? o : 0;
    const int ph = oo / kP, pw = oo - ph * kP;
    const int nq = (o < kP * kP) ? T.nx[pw] : 0;
    snq[t] = nq;
    so[t] = nq > 0 ? ph * LXP + (T.xb[pw] - xmin) : 0;   // taps past nq carry weight 0 and read finite staging data
#pragma unroll
    for (int q = 0; q < 4; ++q) sw[t][q] = (q < nq) ? T.wx[pw * kRB + q] : 0.f;
    sslow |= nq > 4;
  }
  sslow = __any_sync(0xffffffffu, sslow);

  for (int cbase = warp * CPW; cbase < C; cbase += kWarps * CPW) {
    float U[kKC][kP];
#pragma unroll
    for (int k = 0; k < kKC; ++k)
#pragma unroll
      for (int ph = 0; ph < kP; ++ph) U[k][ph] = 0.f;
    uint32_t off_c[kKC];
#pragma unroll
    for (int k = 0; k < kKC; ++k) off_c[k] = (uint32_t)min(cbase + cg * kKC + k, C - 1) * sC + off_x;
    if (T.shared_ok) {
      // every footprint row is loaded once; rows owned by bin ph also feed bins ph+1, ph+2 (static indices)
#pragma unroll
      for (int ph = 0; ph < kP; ++ph) {
        const int rb = T.own_b[ph], re = T.own_e[ph];
        int r = rb;
        uint32_t off_r = (uint32_t)rb * sH;
        for (; r + 1 < re; r += 2) {   // two rows (8 loads) in flight per lane
          const float4 w0 = T.rw[r], w1 = T.rw[r + 1];
          float v0[kKC], v1[kKC];
#pragma unroll
          for (int k = 0; k < kKC; ++k) v0[k] = __ldg(img_base + (off_c[k] + off_r));
#pragma unroll
          for (int k = 0; k < kKC; ++k) v1[k] = __ldg(img_base + (off_c[k] + off_r + sH));
          off_r += 2 * sH;
#pragma unroll
          for (int k = 0; k < kKC; ++k) {
            OSR_FOLD(k, ph, w0, v0[k]);
            OSR_FOLD(k, ph, w1, v1[k]);
          }
        }
        if (r < re) {
          const float4 w0 = T.rw[r];
          float v0[kKC];
#pragma unroll
          for (int k = 0; k < kKC; ++k) v0[k] = __ldg(img_base + (off_c[k] + off_r));
#pragma unroll
          for (int k = 0; k < kKC; ++k) OSR_FOLD(k, ph, w0, v0[k]);
        }
      }
    } else {
#pragma unroll
      for (int ph = 0; ph < kP; ++ph) {
        const int nr = T.ny[ph];
        uint32_t off_r = (uint32_t)(T.yb[ph] - T.ymin) * sH;
        for (int r = 0; r < nr; ++r) {
          const float w = T.wy[ph * kRB + r];
          float v[kKC];
#pragma unroll
          for (int k = 0; k < kKC; ++k) v[k] = __ldg(img_base + (off_c[k] + off_r));
#pragma unroll
          for (int k = 0; k < kKC; ++k) U[k][ph] = fmaf(w, v[k], U[k][ph]);
          off_r += sH;
        }
      }
    }
    {
      float* ub = Us + (cg * kKC) * (kP * LXP) + x;
#pragma unroll
      for (int k = 0; k < kKC; ++k)
#pragma unroll
        for (int ph = 0; ph < kP; ++ph) ub[(k * kP + ph) * LXP] = U[k][ph];
    }
    __syncwarp();
    const int nch = min(CPW, C - cbase);
    float* outc = out_roi + (int64_t)cbase * (kP * kP) + lane;
    if (!sslow) {
#pragma unroll 4
      for (int cl = 0; cl < nch; ++cl) {
        const float* uc = Us + cl * (kP * LXP);
        {
          const float* up = uc + so[0];
          float a = sw[0][0] * up[0];
          a = fmaf(sw[0][1], up[1], a);
          a = fmaf(sw[0][2], up[2], a);
          a = fmaf(sw[0][3], up[3], a);
          outc[cl * (kP * kP)] = a * inv_count;
        }
        if (lane + 32 < kP * kP) {
          const float* up = uc + so[1];
          float a = sw[1][0] * up[0];
          a = fmaf(sw[1][1], up[1], a);
          a = fmaf(sw[1][2], up[2], a);
          a = fmaf(sw[1][3], up[3], a);
          outc[cl * (kP * kP) + 32] = a * inv_count;
        }
      }
    } else {
      for (int cl = 0; cl < nch; ++cl) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int o = lane + 32 * t;
          if (o < kP * kP) {
            const int pw = o % kP;
            const float* up = Us + cl * (kP * LXP) + so[t];
            const float* wp = T.wx + pw * kRB;
            float a = 0.f;
            for (int q = 0; q < snq[t]; ++q) a = fmaf(wp[q], up[q], a);
            outc[cl * (kP * kP) + 32 * t] = a * inv_count;
          }
        }
      }
    }
    __syncwarp();
  }
}

// ---- TMA path ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ bool elect_one() {  // one lane of a CONVERGED warp (TMA is issued from uniform control flow)
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
// One stage = kTC channel planes x nrb row blocks.  The level is addressed as a 2-D tensor (W, N*C*H): row index of
// (n, c, y) = (n*C + c)*H + y.  Rows past the plane's end belong to the next plane and are never consumed (r < hf);
// columns past W are zero-filled by the TMA unit.  Box starts must be 16-byte aligned in x (x0 % 4 == 0).
__device__ __forceinline__ void tma_load_stage(float* dst, const CUtensorMap* map, uint64_t* bar, int x0, int y0, int nrb,
                                               int c0, int img, int C, int H) {
  mbar_expect_tx(bar, (uint32_t)(kTC * nrb * kTR * kTX * 4));
#pragma unroll
  for (int k = 0; k < kTC; ++k) {
    const int c = min(c0 + k, C - 1);  // C % kTC != 0: the extra planes re-read the last channel (results discarded)
    const int row0 = (img * C + c) * H + y0;
    for (int rb = 0; rb < nrb; ++rb) tma_load_2d(dst + k * kChanFloats + rb * (kTR * kTX), map, bar, x0, row0 + rb * kTR);
  }
}

// Footprint (<= 32 columns from the 4-column-aligned origin x0, <= kTRmax rows) streamed through shared memory by the
// TMA unit: one elected lane of warp 0 keeps both stages in flight; warp w consumes channel w of every stage with
// conflict-free LDS (lane = column), folding the rows into the 7 per-bin accumulators, then finishes its channel
// (x contraction through a warp-private tile) and stores its 49 outputs.
__device__ __forceinline__ void fwd_tma(const CUtensorMap* map, const Tables& T, float* stages, uint64_t* full_bar,
                                        uint64_t* empty_bar, float* Us, int C, int H, int img, int x0, int hf,
                                        float inv_count, float* out_roi) {
  constexpr int LXP = kTX + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nrb = ceil_div(hf, kTR), ncb = ceil_div(C, kTC);
  if (tid == 0) {
    for (int i = 0; i < kNS; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the prologue zero-filled the ring with generic-proxy stores; order them before the async-proxy (TMA) writes
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    if (elect_one()) {
      for (int t = 0; t < min(kNS, ncb); ++t)
        tma_load_stage(stages + t * kStageFloats, map, &full_bar[t], x0, T.ymin, nrb, t * kTC, img, C, H);
    }
    __syncwarp();
  }
  // stage-2 ownership (same scheme as the fast path, staging row stride LXP)
  int so[2];
  float sw[2][4];
  bool sslow = false;
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int o = lane + 32 * t;
    const int oo = o < kP * kP ? o : 0;
    const int ph = oo / kP, pw = oo - ph * kP;
    const int nq = (o < kP * kP) ? T.nx[pw] : 0;
    so[t] = nq > 0 ? ph * LXP + (T.xb[pw] - x0) : 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) sw[t][q] = (q < nq) ? T.wx[pw * kRB + q] : 0.f;
    sslow |= nq > 4;
  }
  sslow = __any_sync(0xffffffffu, sslow);

  for (int cb = 0; cb < ncb; ++cb) {
    const int slot = cb % kNS;
    const uint32_t parity = (uint32_t)((cb / kNS) & 1);
    float U[1][kP];
#pragma unroll
    for (int ph = 0; ph < kP; ++ph) U[0][ph] = 0.f;
    mbar_wait(&full_bar[slot], parity);
    const float* tile = stages + slot * kStageFloats + warp * kChanFloats + lane;
#pragma unroll
    for (int ph = 0; ph < kP; ++ph) {
      const int rbeg = T.own_b[ph], rend = T.own_e[ph];
      int r = rbeg;
      for (; r + 1 < rend; r += 2) {
        const float4 w0 = T.rw[r], w1 = T.rw[r + 1];
        const float v0 = tile[r * kTX], v1 = tile[(r + 1) * kTX];
        OSR_FOLD(0, ph, w0, v0);
        OSR_FOLD(0, ph, w1, v1);
      }
      if (r < rend) {
        const float4 w0 = T.rw[r];
        const float v0 = tile[r * kTX];
        OSR_FOLD(0, ph, w0, v0);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[slot]);
    // producer: refill this slot with channel block cb + kNS once every warp has released it
    if (warp == 0 && cb + kNS < ncb) {
      if (elect_one()) {
        mbar_wait(&empty_bar[slot], parity);
        tma_load_stage(stages + slot * kStageFloats, map, &full_bar[slot], x0, T.ymin, nrb, (cb + kNS) * kTC, img, C, H);
      }
      __syncwarp();
    }
    const int c = cb * kTC + warp;
    if (c < C) {
#pragma unroll
      for (int ph = 0; ph < kP; ++ph) Us[ph * LXP + lane] = U[0][ph];
      __syncwarp();
      float* outc = out_roi + (int64_t)c * (kP * kP) + lane;
      if (!sslow) {
        {
          const float* up = Us + so[0];
          float a = sw[0][0] * up[0];
          a = fmaf(sw[0][1], up[1], a);
          a = fmaf(sw[0][2], up[2], a);
          a = fmaf(sw[0][3], up[3], a);
          outc[0] = a * inv_count;
        }
        if (lane + 32 < kP * kP) {
          const float* up = Us + so[1];
          float a = sw[1][0] * up[0];
          a = fmaf(sw[1][1], up[1], a);
          a = fmaf(sw[1][2], up[2], a);
          a = fmaf(sw[1][3], up[3], a);
          outc[32] = a * inv_count;
        }
      } else {
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
          const int o = lane + 32 * tt;
          if (o < kP * kP) {
            const int pw = o % kP;
            const float* up = Us + so[tt];
            const float* wp = T.wx + pw * kRB;
            float a = 0.f;
            for (int q = 0; q < T.nx[pw]; ++q) a = fmaf(wp[q], up[q], a);
            outc[32 * tt] = a * inv_count;
          }
        }
      }
      __syncwarp();
    }
  }
}

__device__ __forceinline__ void fwd_wide(const LevelDesc& lv, const Tables& T, float* Us, const float* img_base, int C,
                                         int xmin, int wf, float inv_count, float* out_roi) {
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
      out_roi[(int64_t)c * (kP * kP) + o] = s * inv_count;
    }
    __syncwarp();
  }
}

#ifndef OSR_FWD_MINB
#define OSR_FWD_MINB 3
#endif
template <bool kTma>
__global__ void __launch_bounds__(kThreads, OSR_FWD_MINB) roi_align_fwd_kernel(const __grid_constant__ FwdParams p,
                                                                               const __grid_constant__ FwdTma tm) {
  // dynamic shared memory: [ ring (48 KB; LDG paths use its first 32 KB as warp tiles) | Us_tma | barriers | T ]
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* s_U = reinterpret_cast<float*>(smem_raw);
  float* s_Ut = s_U + p.ring_floats;                                        // kWarps x 7 x 33 floats
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_Ut + kWarps * kP * (kTX + 1) + 8);
  Tables& T = *reinterpret_cast<Tables*>(s_bar + 2 * kNS);

  const int m = p.order ? p.order[blockIdx.x] : blockIdx.x;
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
    // the staging tiles' pad columns are read (with weight 0) by the fixed 4-tap stage 2: keep them finite
    for (int i = tid; i < p.ring_floats + kWarps * kP * (kTX + 1) + 8; i += kThreads) s_U[i] = 0.f;
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
  int xmin = 0, wf = 0, ymin = 0, hf = 0;
  if (!zero) {
    bool overflow = false, anyx = false, anyy = false;
    int xmax = -1, ymax = -1;
    xmin = 1 << 30;
    ymin = 1 << 30;
#pragma unroll
    for (int i = 0; i < kP; ++i) {
      const int nx = T.nx[i], ny = T.ny[i];
      overflow |= (nx < 0) | (ny < 0);
      if (nx > 0) {
        anyx = true;
        xmin = min(xmin, T.xb[i]);
        xmax = max(xmax, T.xb[i] + nx - 1);
      }
      if (ny > 0) {
        anyy = true;
        ymin = min(ymin, T.yb[i]);
        ymax = max(ymax, T.yb[i] + ny - 1);
      }
    }
    wf = xmax - xmin + 1;
    hf = ymax - ymin + 1;
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
  if (path == 0) {
    // shared-row tables: row r of the footprint -> (first bin containing it, weights for that bin and the next two)
    int bad = 0;
    for (int r = tid; r < hf; r += kThreads) {
      const int y = ymin + r;
      int ph0 = -1, cnt = 0;
      float w0 = 0.f, w1 = 0.f, w2 = 0.f;
#pragma unroll
      for (int ph = 0; ph < kP; ++ph) {
        const int rr = y - T.yb[ph];
        if (rr >= 0 && rr < T.ny[ph]) {
          if (ph0 < 0) ph0 = ph;
          const int d = ph - ph0;
          const float wv = T.wy[ph * kRB + rr];
          if (d == 0) w0 = wv;
          else if (d == 1) w1 = wv;
          else if (d == 2) w2 = wv;
          else bad = 1;
          ++cnt;
        }
      }
      // rows inside [ymin, ymax] that no bin touches cannot exist (bins tile the RoI), but stay safe:
      T.rw[r] = make_float4(w0, w1, w2, __int_as_float(ph0 < 0 ? kP - 1 : ph0));
    }
    bad = __syncthreads_or(bad);
    if (tid < kP) {
      // own range of bin ph = rows whose first bin is ph (rows are ordered, so this is a contiguous range)
      int b = hf, e = 0;
      for (int r = 0; r < hf; ++r)
        if (__float_as_int(T.rw[r].w) == tid) {
          b = min(b, r);
          e = r + 1;
        }
      T.own_b[tid] = b < e ? b : 0;
      T.own_e[tid] = b < e ? e : 0;
    }
    if (tid == 0) {
      T.ymin = ymin;
      T.hf = hf;
      T.shared_ok = !bad;
    }
    __syncthreads();
  }
  const LevelDesc& lv = p.L.lv[level];
  const float* img_base = lv.data + (int64_t)img * lv.sN;
  float* Us = s_U + (tid >> 5) * kWarpTile;
  const float inv_count = 1.0f / g.count;
  // TMA box starts must be 16-byte aligned in x: start at x0 = xmin rounded down to 4 columns
  if (kTma && path == 0 && p.tma_ok[level] && T.shared_ok && ((xmin & 3) + wf <= kTX) && hf <= kTRmax) {
    fwd_tma(&tm.map[level], T, s_U, s_bar, s_bar + kNS, s_Ut + (tid >> 5) * (kP * (kTX + 1)), C, lv.H, img, xmin & ~3, hf,
            inv_count, out_roi);
  } else if (path == 0) {
    if (wf <= 8) fwd_fast<8>(lv, T, Us, img_base, C, xmin, wf, inv_count, out_roi);
    else if (wf <= 16) fwd_fast<16>(lv, T, Us, img_base, C, xmin, wf, inv_count, out_roi);
    else fwd_fast<32>(lv, T, Us, img_base, C, xmin, wf, inv_count, out_roi);
  } else if (path == 1) {
    fwd_wide(lv, T, Us, img_base, C, xmin, wf, inv_count, out_roi);
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

// =========================================================================================================
// channels_last (NHWC) forward: the layout the north star asks for ("coalesced NHWC reads").  A footprint row of a
// channels_last map is ONE contiguous run of wf*C floats (22 KB for a 22-pixel-wide RoI at C = 256), so it is streamed
// with a single cp.async.bulk (UBLKCP) per row into a shared-memory ring guarded by mbarriers (2..6 stages, sized to the
// RoI); no address arithmetic, no sector waste, tens of KB in flight per CTA.  Thread t owns channel t and keeps all 49
// outputs of the RoI in registers:
//   per row  U[pw] = sum_x Wx[pw][x] * row[x][c]   (conflict-free LDS: consecutive threads, consecutive floats; the taps
//                                                   of every bin padded with zero weights to 4 / 6 / 8 - chosen per RoI -
//                                                   and the tap window pulled inside the row, so the loop is branch-free
//                                                   and never reads outside its own stage)
//   then     out[ph][pw] += Wy[ph][y] * U[pw]      for the <= 3 bins containing the row (uniform switch).
// The per-RoI tables come from roi_fwd_prep_kernel's records in the common case (in-CTA derivation otherwise).
// The 49 x C tile is transposed through shared memory and stored with 16-byte coalesced writes (C-major output).
constexpr int kNhwcRingCols = 96;   // ring capacity in pixel columns (x C floats); split per RoI into 2..6 row stages (p.max_stages)
constexpr int kNhwcMaxStages = 12;
constexpr int kNhwcWide = 48;       // widest footprint staged as whole rows; wider ones go in 32-column chunks

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// packed fp32x2 arithmetic (SASS FFMA2 / FMUL2): both halves of a register pair in one instruction
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                     rc = *reinterpret_cast<unsigned long long*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b) {
  unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b), rd;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 f2_fma_s(float2 a, float s, float2 c) { return f2_fma(a, make_float2(s, s), c); }

// ---------------------------------------------------------------------------------------------------------
// Per-RoI table records for the channels_last kernel, written once by a warp-per-RoI prep kernel so that the 256-thread
// RoI CTAs do not spend the first quarter of their life deriving tables with 14 active threads.  A record is only
// marked FAST for the common case the fully unrolled row loop handles (footprint <= 48 x 64 pixels, bins <= 8 pixels
// wide, every row in <= 3 bins); every other RoI keeps the in-CTA prologue.  Layout (floats):
//   [0..3]   int4   flags(bit 0 = FAST) | level << 8,  xmin,  ymin,  wf | hf << 16
//   [4..7]   float4 1/count, tmax (int bits), image (int bits), RoI index m (int bits)
//   [8..15]  int    first column of each bin relative to xmin (0 for empty bins)
//   [16..71] float  x-tap weights of each bin padded to 8
//   [72..327] float4 per footprint row: (w(ph0), w(ph0+1), w(ph0+2), ph0)
//   [328..391] float the x-tap weights again, pair-interleaved for the packed row loop (Tables::wt2)
constexpr int kRecRows = 64;
constexpr int kRecW2 = 72 + 4 * kRecRows;
constexpr int kRecFloats = kRecW2 + 64;
static_assert((kRecFloats * 4) % 16 == 0, "records are fetched with one 16-byte-granular bulk copy");
// taps the row loop executes per bin for a RoI whose widest bin spans `tmax` pixels (one unrolled loop per count)
__host__ __device__ __forceinline__ int nhwc_tap_count(int tmax) { return tmax <= 2 ? 2 : (tmax <= 6 ? tmax : 8); }
constexpr int kPrepWarpsF = 4;

struct PrepScratch {
  float wy[kP * kRB];
  float wx[kP * kRB];
  int yb[kP], ny[kP], xb[kP], nx[kP];
};

// one warp per RoI m (no block-level barrier inside)
__device__ __forceinline__ void roi_fwd_prep_body(const FwdParams& p, int m, PrepScratch& T) {
  const int lane = threadIdx.x & 31;
  if (m == 0 && lane == 0 && p.counter != nullptr) *p.counter = p.pers_grid;   // work counter of the persistent kernel
  if (m >= p.M) return;   // warp-uniform
  float* rec = p.rec + (int64_t)m * kRecFloats;
  const float* roi = p.rois + (int64_t)m * 5;
  const float fimg = __ldg(roi), x1 = __ldg(roi + 1), y1 = __ldg(roi + 2), x2 = __ldg(roi + 3), y2 = __ldg(roi + 4);
  const int img = (int)fimg;
  const int level = assign_level(x1, y1, x2, y2, p.L);
  const bool zero = (level < 0) || (level >= p.L.num_levels) || (img < 0) || (img >= p.L.num_images);
  if (zero) {
    if (lane == 0) {
      reinterpret_cast<int4*>(rec)[0] = make_int4(0, 0, 0, 0);
      reinterpret_cast<float4*>(rec)[1] = make_float4(0.f, 0.f, 0.f, __int_as_float(m));
    }
    return;
  }
  const LevelDesc& lv = p.L.lv[level];
  const RoiGeom g = roi_geometry(x1, y1, x2, y2, lv.scale, p.L.sampling_ratio);
  for (int i = lane; i < kP * kRB; i += 32) {
    T.wy[i] = 0.f;
    T.wx[i] = 0.f;
  }
  __syncwarp();
  if (lane < kP) T.ny[lane] = build_bin_weights(g.start_h, g.bin_h, g.grid_h, lv.H, lane, T.wy + lane * kRB, &T.yb[lane]);
  else if (lane < 2 * kP) T.nx[lane - kP] = build_bin_weights(g.start_w, g.bin_w, g.grid_w, lv.W, lane - kP, T.wx + (lane - kP) * kRB, &T.xb[lane - kP]);
  __syncwarp();
  int xmin = 1 << 30, xmax = -1, ymin = 1 << 30, ymax = -1, tmax = 0;
  bool bad = false;
#pragma unroll
  for (int i = 0; i < kP; ++i) {
    const int nx = T.nx[i], ny = T.ny[i];
    bad |= (nx < 0) | (ny < 0);
    tmax = max(tmax, nx);
    if (nx > 0) { xmin = min(xmin, T.xb[i]); xmax = max(xmax, T.xb[i] + nx - 1); }
    if (ny > 0) { ymin = min(ymin, T.yb[i]); ymax = max(ymax, T.yb[i] + ny - 1); }
  }
  const int wf = xmax - xmin + 1, hf = ymax - ymin + 1;
  bool fast = !bad && xmax >= 0 && ymax >= 0 && wf <= kNhwcWide && hf <= kRecRows && tmax <= 8;
  if (fast) {
    for (int r = lane; r < hf; r += 32) {   // same packing as the in-CTA prologue
      const int y = ymin + r;
      int ph0 = -1, cnt = 0;
      float w0 = 0.f, w1 = 0.f, w2 = 0.f;
#pragma unroll
      for (int ph = 0; ph < kP; ++ph) {
        const int rr = y - T.yb[ph];
        if (rr >= 0 && rr < T.ny[ph]) {
          if (ph0 < 0) ph0 = ph;
          const int d = ph - ph0;
          const float wv = T.wy[ph * kRB + rr];
          if (d == 0) w0 = wv;
          else if (d == 1) w1 = wv;
          else if (d == 2) w2 = wv;
          else cnt = 99;
        }
      }
      if (cnt == 99) bad = true;
      reinterpret_cast<float4*>(rec + 72)[r] = make_float4(w0, w1, w2, __int_as_float(ph0 < 0 ? kP : ph0));
    }
    fast = !__any_sync(0xffffffffu, bad);
  }
  if (fast) {
    // Tap window of each bin: TT = taps the row loop will execute (4 / 6 / 8 by the widest bin).  The window start is
    // pulled left so that all TT taps stay inside the row (columns [0, wf)) and the weights are shifted right by the
    // same amount: zero-weight taps then re-read real data of the SAME row, never uninitialised shared memory, and the
    // CTA does not have to zero-fill its 96 KB ring (footprints narrower than TT still do).
    const int TT = nhwc_tap_count(tmax);
    for (int i = lane; i < kP * 8; i += 32) {
      const int pw = i >> 3, q = i & 7;
      const int nx = T.nx[pw];
      const int first = nx > 0 ? T.xb[pw] - xmin : 0;
      const int start = max(0, min(first, wf - TT));
      const int qq = q - (first - start);
      const float wv = (qq >= 0 && qq < nx) ? T.wx[pw * kRB + qq] : 0.f;
      rec[16 + i] = wv;
      rec[kRecW2 + (pw >> 1) * 16 + (q >> 1) * 4 + (q & 1) * 2 + (pw & 1)] = wv;   // element (pp, t2, k): k = (tap & 1) * 2 + (bin & 1)
    }
    if (lane < 8) rec[kRecW2 + 3 * 16 + (lane >> 1) * 4 + (lane & 1) * 2 + 1] = 0.f;   // the missing 8th bin
    if (lane < 8) {
      const int nx = lane < kP ? T.nx[lane] : 0;
      const int first = nx > 0 ? T.xb[lane] - xmin : 0;
      reinterpret_cast<int*>(rec)[8 + lane] = max(0, min(first, wf - TT));
    }
  }
  if (lane == 0) {
    reinterpret_cast<float4*>(rec)[1] = make_float4(1.0f / g.count, __int_as_float(tmax), __int_as_float(img), __int_as_float(m));
    reinterpret_cast<int4*>(rec)[0] = make_int4((fast ? 1 : 0) | (level << 8), xmin, ymin, wf | (hf << 16));
  }
}

__global__ void __launch_bounds__(kPrepWarpsF * 32) roi_fwd_prep_kernel(const __grid_constant__ FwdParams p) {
  __shared__ PrepScratch S4[kPrepWarpsF];
  const int warp = threadIdx.x >> 5;
  roi_fwd_prep_body(p, blockIdx.x * kPrepWarpsF + warp, S4[warp]);
}

// One RoI (record j) by one 256-thread CTA.  kUseRec = false ignores the prep record (the persistent kernel calls this form
// for the records that are not FAST; the barriers must then be fresh, i.e. invalidated by the caller).
template <int kC, bool kUseRec>   // kC > 0: compile-time channel count (immediate LDS offsets); 0: run-time C
__device__ __forceinline__ void nhwc_roi(const FwdParams& p, int j, float* ring, uint64_t* s_bar, Tables& T) {
  const int C = kC > 0 ? kC : p.L.C;
  uint64_t* full_bar = s_bar;
  uint64_t* empty_bar = s_bar + kNhwcMaxStages;

  const int m = p.order ? p.order[j] : j;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* out_roi = p.out + (int64_t)m * C * (kP * kP);
  unsigned short* out_bf = p.out_bf16 ? p.out_bf16 + (int64_t)m * C * (kP * kP) : nullptr;
  auto store_out = [&](int o, float v) {   // rare paths (empty / oversized RoIs): scalar stores in the requested type
    if (out_bf) out_bf[o] = __bfloat16_as_ushort(__float2bfloat16_rn(v));
    else out_roi[o] = v;
  };
  // ---- prologue: tables either from the prep kernel's record (common case) or derived here ----------------------
  int level, img, xmin, ymin, wf, hf, tmax = 0;
  float inv_count;
  int toff[kP];
  bool pre = false, early = false;
  int4 h0 = make_int4(0, 0, 0, 0);
  const float* rec = nullptr;
  if (kUseRec && p.rec != nullptr) {
    rec = p.rec + (int64_t)m * kRecFloats;
    h0 = __ldg(reinterpret_cast<const int4*>(rec));
    pre = (h0.x & 1) != 0;
  }
  if (pre) {
    level = h0.x >> 8;
    xmin = h0.y; ymin = h0.z; wf = h0.w & 0xffff; hf = h0.w >> 16;
    const float4 h1 = __ldg(reinterpret_cast<const float4*>(rec) + 1);
    inv_count = h1.x;
    tmax = __float_as_int(h1.y);
    img = __float_as_int(h1.z);
    if (tid == 0) p.out_level[m] = level;
    // Early issue: the first rows only need the header, so one thread arms the barriers and starts the bulk copies NOW;
    // the table loads below (x taps, row weights: ~1 us of dependent global loads) overlap with the rows' flight time
    // instead of preceding it.  (Footprints narrower than 8 columns zero-fill the ring first and keep the late issue.)
    early = (wf >= 8);
    if (early && tid == 0) {
      const LevelDesc& lve = p.L.lv[level];
      const float* ibase = lve.data + (int64_t)img * lve.sN;
      for (int i = 0; i < kNhwcMaxStages; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], kWarps);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const int scols_e = wf <= kNhwcWide ? max(8, (wf + 3) & ~3) : 32;
      const int nst_e = min(p.max_stages, kNhwcRingCols / scols_e);
      const int total_e = ceil_div(wf, scols_e) * hf;
      for (int t = 0; t < min(nst_e, total_e); ++t) {
        const int xc = t / hf, r = t - xc * hf;
        const int ncols = min(scols_e, wf - xc * scols_e);
        const uint32_t bytes = (uint32_t)(ncols * C * 4);
        mbar_expect_tx(&full_bar[t], bytes);
        bulk_load_1d(ring + t * (scols_e * C), ibase + ((int64_t)(ymin + r) * lve.sH + (int64_t)(xmin + xc * scols_e) * lve.sW), bytes, &full_bar[t]);
      }
    }
    const int4 o0 = __ldg(reinterpret_cast<const int4*>(rec) + 2), o1 = __ldg(reinterpret_cast<const int4*>(rec) + 3);
    toff[0] = o0.x * C; toff[1] = o0.y * C; toff[2] = o0.z * C; toff[3] = o0.w * C;
    toff[4] = o1.x * C; toff[5] = o1.y * C; toff[6] = o1.z * C;
    if (tid < 2 * kP) reinterpret_cast<float4*>(&T.wt[0][0])[tid] = __ldg(reinterpret_cast<const float4*>(rec + 16) + tid);
    if (tid >= 64 && tid < 80) reinterpret_cast<float4*>(&T.wt2[0][0])[tid - 64] = __ldg(reinterpret_cast<const float4*>(rec + kRecW2) + (tid - 64));
    for (int r = tid; r < hf; r += kThreads) T.rw[r] = __ldg(reinterpret_cast<const float4*>(rec + 72) + r);
  } else {
  const float* roi = p.rois + (int64_t)m * 5;
  const float fimg = __ldg(roi), x1 = __ldg(roi + 1), y1 = __ldg(roi + 2), x2 = __ldg(roi + 3), y2 = __ldg(roi + 4);
  img = (int)fimg;
  level = assign_level(x1, y1, x2, y2, p.L);
  if (tid == 0) p.out_level[m] = level;

  const bool zero = (level < 0) || (level >= p.L.num_levels) || (img < 0) || (img >= p.L.num_images);
  RoiGeom g;
  int xmax = -1, ymax = -1;
  xmin = 1 << 30; ymin = 1 << 30;
  bool overflow = false;
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
#pragma unroll
    for (int i = 0; i < kP; ++i) {
      const int nx = T.nx[i], ny = T.ny[i];
      overflow |= (nx < 0) | (ny < 0);
      if (nx > 0) { xmin = min(xmin, T.xb[i]); xmax = max(xmax, T.xb[i] + nx - 1); }
      if (ny > 0) { ymin = min(ymin, T.yb[i]); ymax = max(ymax, T.yb[i] + ny - 1); }
    }
  }
  wf = xmax - xmin + 1; hf = ymax - ymin + 1;
  if (zero || (!overflow && (xmax < 0 || ymax < 0))) {
    for (int o = tid; o < C * kP * kP; o += kThreads) store_out(o, 0.f);
    return;
  }
  if (overflow || hf > kMaxRows) {
    const LevelDesc& lv = p.L.lv[level];
    const float* img_base = lv.data + (int64_t)img * lv.sN;
    // generic: torchvision's per-sample loop (any strides)
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
      store_out(o, s / g.count);
    }
    return;
  }
  // row -> bins table (same as the NCHW shared-row form): rw[r] = (w(ph0), w(ph0+1), w(ph0+2), ph0); rows that sit in
  // more than 3 bins (bins narrower than half a pixel) carry ph0 = -1 and are folded densely from wy.
  for (int r = tid; r < hf; r += kThreads) {
    const int y = ymin + r;
    int ph0 = -1, cnt = 0;
    float w0 = 0.f, w1 = 0.f, w2 = 0.f;
#pragma unroll
    for (int ph = 0; ph < kP; ++ph) {
      const int rr = y - T.yb[ph];
      if (rr >= 0 && rr < T.ny[ph]) {
        if (ph0 < 0) ph0 = ph;
        const int d = ph - ph0;
        const float wv = T.wy[ph * kRB + rr];
        if (d == 0) w0 = wv;
        else if (d == 1) w1 = wv;
        else if (d == 2) w2 = wv;
        else cnt = 99;
      }
    }
    T.rw[r] = make_float4(w0, w1, w2, __int_as_float(cnt == 99 ? -1 : (ph0 < 0 ? kP : ph0)));
  }
    inv_count = 1.0f / g.count;
    __syncthreads();   // T.rw complete; T.nx / T.xb / T.wx read below
#pragma unroll
    for (int pw = 0; pw < kP; ++pw) tmax = max(tmax, T.nx[pw]);
    // tap windows pulled inside the row, weights shifted by the same amount (see roi_fwd_prep_kernel)
    const int TT = nhwc_tap_count(tmax);
#pragma unroll
    for (int pw = 0; pw < kP; ++pw) {
      const int first = T.nx[pw] > 0 ? T.xb[pw] - xmin : 0;
      toff[pw] = max(0, min(first, wf - TT)) * C;
    }
    if (tid < kP) {
      const int nx = T.nx[tid];
      const int first = nx > 0 ? T.xb[tid] - xmin : 0;
      const int shift = first - max(0, min(first, wf - TT));
      float w8[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int qq = q - shift;
        w8[q] = (qq >= 0 && qq < nx && qq < kRB) ? T.wx[tid * kRB + qq] : 0.f;
      }
      T.wt[tid][0] = make_float4(w8[0], w8[1], w8[2], w8[3]);
      T.wt[tid][1] = make_float4(w8[4], w8[5], w8[6], w8[7]);
#pragma unroll
      for (int q = 0; q < 8; ++q) reinterpret_cast<float*>(&T.wt2[0][0])[(tid >> 1) * 16 + (q >> 1) * 4 + (q & 1) * 2 + (tid & 1)] = w8[q];
      if (tid == kP - 1) {
#pragma unroll
        for (int q = 0; q < 8; ++q) reinterpret_cast<float*>(&T.wt2[0][0])[3 * 16 + (q >> 1) * 4 + (q & 1) * 2 + 1] = 0.f;
      }
    }
  }
  const LevelDesc& lv = p.L.lv[level];
  const float* img_base = lv.data + (int64_t)img * lv.sN;
  if (tid == 0 && !early) {
    for (int i = 0; i < kNhwcMaxStages; ++i) {
      if (!kUseRec) {   // called per RoI by the persistent kernel: the barriers hold valid (idle) objects from the last call
        asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&full_bar[i])) : "memory");
        asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&empty_bar[i])) : "memory");
      }
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  // accumulators as bin PAIRS: acc2[ph][pp] = (out[ph][2 pp], out[ph][2 pp + 1]); pp = 3 carries bin 6 and a dummy
  float2 acc2[kP][4];
#pragma unroll
  for (int a = 0; a < kP; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc2[a][b] = make_float2(0.f, 0.f);

  // stage geometry: a stage holds one footprint row (chunk).  Narrow RoIs get more, smaller stages (deeper prefetch),
  // RoIs up to 48 pixels wide are still staged as whole rows, wider ones in 32-column chunks.
  const int scols = wf <= kNhwcWide ? max(8, (wf + 3) & ~3) : 32;   // >= 8: every (padded) tap stays inside its own stage
  const int nstages = min(p.max_stages, kNhwcRingCols / scols);
  const int stage_floats = scols * C;
  const int nxc = ceil_div(wf, scols);
  const int total = nxc * hf;          // (x chunk, row) tiles, chunk-major
  const int c = tid;                   // my channel (C <= 256 enforced by the host)
  const bool cin = c < C;
  // Common case (one x chunk): every bin's x taps are padded to 8 with zero weights (T.wt), so the row loop is fully
  // unrolled and predicate-free; padded taps read finite data (the ring is zero-initialised once and only ever holds
  // feature values; 8 columns of slack follow the last stage).
  const bool fast_taps = (nxc == 1) && (tmax <= 8);   // tmax = widest bin in pixels: picks the unrolled row loop
  const int ntaps = nhwc_tap_count(tmax);
  // Zero-weight padded taps must read finite data.  The tap windows stay inside the loaded row unless the footprint is
  // narrower than the tap count; only then the pad columns exist and the ring is zero-filled first.
  if (wf < 8) {
    for (int i = tid; i < (p.ring_floats >> 2); i += kThreads) reinterpret_cast<float4*>(ring)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic zero-fill before async-proxy bulk writes
  }
  __syncthreads();
  // producer prologue (already issued by thread 0 when `early`)
  if (warp == 0 && !early) {
    if (elect_one()) {
      for (int t = 0; t < min(nstages, total); ++t) {
        const int xc = t / hf, r = t - xc * hf;
        const int ncols = min(scols, wf - xc * scols);
        const uint32_t bytes = (uint32_t)(ncols * C * 4);
        mbar_expect_tx(&full_bar[t], bytes);
        bulk_load_1d(ring + t * stage_floats, img_base + ((int64_t)(ymin + r) * lv.sH + (int64_t)(xmin + xc * scols) * lv.sW), bytes, &full_bar[t]);
      }
    }
    __syncwarp();
  }
  int xc = 0, r = 0;
  int slot = 0;            // t % nstages and (t / nstages) & 1, kept incrementally (nstages is a run-time value)
  uint32_t parity = 0;
  int rxc = 0, rr = 0;     // (chunk, row) of tile t + nstages, the one the producer refills
  {
    const int t0 = min(nstages, total);
    rxc = t0 / hf;
    rr = t0 - rxc * hf;
  }
  int t = 0;
  // Two rows per iteration (whole-row stages only): the tap weights - a quarter of the row loop's shared-memory wavefronts -
  // are loaded once for both rows and the per-row loop overhead (barrier wait, release, refill, ~90 of the ~160 instructions
  // of a one-row iteration) is halved.  0.506 -> 0.471 ms at cfg 2; OSR_TUNE_FWD_VARIANT = 5 keeps the one-row loop.
  if (fast_taps && p.two_rows && nstages >= 2) {
    const int cc = min(c, C - 1);
    for (; t + 1 < total; t += 2) {
      const int slot_a = slot;
      const uint32_t par_a = parity;
      if (++slot == nstages) { slot = 0; parity ^= 1u; }
      const int slot_b = slot;
      const uint32_t par_b = parity;
      if (++slot == nstages) { slot = 0; parity ^= 1u; }
      mbar_wait(&full_bar[slot_a], par_a);
      mbar_wait(&full_bar[slot_b], par_b);
      const float* rowa = ring + slot_a * stage_floats + cc;
      const float* rowb = ring + slot_b * stage_floats + cc;
      float2 Ua[4], Ub[4];
#define OSR_NHWC_TAPS2(NT)                                                                                         \
  _Pragma("unroll") for (int pp = 0; pp < 3; ++pp) {                                                                \
    const int o0 = toff[2 * pp], o1 = toff[2 * pp + 1];                                                             \
    float2 ua = make_float2(0.f, 0.f), ub = make_float2(0.f, 0.f);                                                  \
    _Pragma("unroll") for (int t2 = 0; t2 < (NT) / 2; ++t2) {                                                       \
      const float4 w = T.wt2[pp][t2];                                                                               \
      const float2 wl = make_float2(w.x, w.y), wh = make_float2(w.z, w.w);                                          \
      const float2 a0 = make_float2(rowa[o0 + (2 * t2) * C], rowa[o1 + (2 * t2) * C]);                              \
      const float2 a1 = make_float2(rowa[o0 + (2 * t2 + 1) * C], rowa[o1 + (2 * t2 + 1) * C]);                      \
      const float2 b0 = make_float2(rowb[o0 + (2 * t2) * C], rowb[o1 + (2 * t2) * C]);                              \
      const float2 b1 = make_float2(rowb[o0 + (2 * t2 + 1) * C], rowb[o1 + (2 * t2 + 1) * C]);                      \
      ua = (t2 == 0) ? f2_mul(a0, wl) : f2_fma(a0, wl, ua);                                                         \
      ub = (t2 == 0) ? f2_mul(b0, wl) : f2_fma(b0, wl, ub);                                                         \
      ua = f2_fma(a1, wh, ua);                                                                                      \
      ub = f2_fma(b1, wh, ub);                                                                                      \
    }                                                                                                               \
    if ((NT) & 1) {   /* odd tap count: the last tap uses half a weight vector */                                   \
      const float4 w = T.wt2[pp][(NT) / 2];                                                                         \
      const float2 wl = make_float2(w.x, w.y);                                                                      \
      ua = f2_fma(make_float2(rowa[o0 + ((NT) - 1) * C], rowa[o1 + ((NT) - 1) * C]), wl, ua);                       \
      ub = f2_fma(make_float2(rowb[o0 + ((NT) - 1) * C], rowb[o1 + ((NT) - 1) * C]), wl, ub);                       \
    }                                                                                                               \
    Ua[pp] = ua;                                                                                                    \
    Ub[pp] = ub;                                                                                                    \
  }                                                                                                                 \
  {                                                                                                                 \
    const int o0 = toff[kP - 1];                                                                                    \
    float2 u = make_float2(0.f, 0.f);   /* (row a, row b) of the unpaired bin */                                     \
    _Pragma("unroll") for (int t2 = 0; t2 < (NT) / 2; ++t2) {                                                       \
      const float4 w = T.wt2[3][t2];                                                                                \
      u = f2_fma_s(make_float2(rowa[o0 + (2 * t2) * C], rowb[o0 + (2 * t2) * C]), w.x, u);                          \
      u = f2_fma_s(make_float2(rowa[o0 + (2 * t2 + 1) * C], rowb[o0 + (2 * t2 + 1) * C]), w.z, u);                  \
    }                                                                                                               \
    if ((NT) & 1) u = f2_fma_s(make_float2(rowa[o0 + ((NT) - 1) * C], rowb[o0 + ((NT) - 1) * C]), T.wt2[3][(NT) / 2].x, u); \
    Ua[3] = make_float2(u.x, 0.f);                                                                                  \
    Ub[3] = make_float2(u.y, 0.f);                                                                                  \
  }
      switch (ntaps) {   // CTA-uniform
        case 2: OSR_NHWC_TAPS2(2) break;
        case 3: OSR_NHWC_TAPS2(3) break;
        case 4: OSR_NHWC_TAPS2(4) break;
        case 5: OSR_NHWC_TAPS2(5) break;
        case 6: OSR_NHWC_TAPS2(6) break;
        default: OSR_NHWC_TAPS2(8) break;
      }
#undef OSR_NHWC_TAPS2
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&empty_bar[slot_a]);
        mbar_arrive(&empty_bar[slot_b]);
      }
      if (warp == 0 && t + nstages < total) {
        if (elect_one()) {
          const uint32_t bytes = (uint32_t)(wf * C * 4);
          mbar_wait(&empty_bar[slot_a], par_a);
          mbar_expect_tx(&full_bar[slot_a], bytes);
          bulk_load_1d(ring + slot_a * stage_floats, img_base + ((int64_t)(ymin + t + nstages) * lv.sH + (int64_t)xmin * lv.sW), bytes, &full_bar[slot_a]);
          if (t + 1 + nstages < total) {
            mbar_wait(&empty_bar[slot_b], par_b);
            mbar_expect_tx(&full_bar[slot_b], bytes);
            bulk_load_1d(ring + slot_b * stage_floats, img_base + ((int64_t)(ymin + t + 1 + nstages) * lv.sH + (int64_t)xmin * lv.sW), bytes, &full_bar[slot_b]);
          }
        }
        __syncwarp();
      }
#define OSR_NHWC_CASE(PH, UU, W)                                                                            \
  case PH:                                                                                                  \
    _Pragma("unroll") for (int pp = 0; pp < 4; ++pp) {                                                      \
      acc2[PH][pp] = f2_fma_s(UU[pp], (W).x, acc2[PH][pp]);                                                 \
      if (PH + 1 < kP) acc2[PH + 1 < kP ? PH + 1 : 0][pp] = f2_fma_s(UU[pp], (W).y, acc2[PH + 1 < kP ? PH + 1 : 0][pp]); \
      if (PH + 2 < kP) acc2[PH + 2 < kP ? PH + 2 : 0][pp] = f2_fma_s(UU[pp], (W).z, acc2[PH + 2 < kP ? PH + 2 : 0][pp]); \
    }                                                                                                       \
    break;
#define OSR_NHWC_FOLD(UU, ROWIDX)                                                                           \
  {                                                                                                         \
    const float4 w = T.rw[ROWIDX];                                                                          \
    switch (__float_as_int(w.w)) {                                                                          \
      OSR_NHWC_CASE(0, UU, w) OSR_NHWC_CASE(1, UU, w) OSR_NHWC_CASE(2, UU, w) OSR_NHWC_CASE(3, UU, w)       \
      OSR_NHWC_CASE(4, UU, w) OSR_NHWC_CASE(5, UU, w) OSR_NHWC_CASE(6, UU, w)                               \
      case -1: {                                                                                            \
        const int y = ymin + (ROWIDX);                                                                      \
        _Pragma("unroll") for (int ph = 0; ph < kP; ++ph) {                                                 \
          const int rr2 = y - T.yb[ph];                                                                     \
          const float wy = (rr2 >= 0 && rr2 < T.ny[ph]) ? T.wy[ph * kRB + rr2] : 0.f;                       \
          _Pragma("unroll") for (int pp = 0; pp < 4; ++pp) acc2[ph][pp] = f2_fma_s(UU[pp], wy, acc2[ph][pp]); \
        }                                                                                                   \
        break;                                                                                              \
      }                                                                                                     \
      default: break;                                                                                       \
    }                                                                                                       \
  }
      OSR_NHWC_FOLD(Ua, t)
      OSR_NHWC_FOLD(Ub, t + 1)
#undef OSR_NHWC_FOLD
#undef OSR_NHWC_CASE
    }
    // (whole-row stages: chunk 0, row t) bookkeeping of the one-row loop that finishes an odd row count
    r = t;
    {
      const int tn = min(t + nstages, total);
      rxc = tn / hf;
      rr = tn - rxc * hf;
    }
  }
  for (; t < total; ++t) {
    mbar_wait(&full_bar[slot], parity);
    const float* row = ring + slot * stage_floats + c;
    const int x_lo = xmin + xc * scols, x_hi = min(xmin + wf, x_lo + scols);
    // x contraction of this row (chunk): U[pw] = sum_x Wx[pw][x] * row[x][c], two bins per packed FMA
    float2 U2[4];
    if (fast_taps) {
      const float* rowc = ring + slot * stage_floats + min(c, C - 1);
#define OSR_NHWC_TAPS(NT)                                                                                          \
  _Pragma("unroll") for (int pp = 0; pp < 3; ++pp) {                                                                \
    const float* ra = rowc + toff[2 * pp];                                                                          \
    const float* rb = rowc + toff[2 * pp + 1];                                                                      \
    float2 u = make_float2(0.f, 0.f);                                                                               \
    _Pragma("unroll") for (int t2 = 0; t2 < (NT) / 2; ++t2) {                                                       \
      const float4 w = T.wt2[pp][t2];                                                                               \
      const float2 d0 = make_float2(ra[(2 * t2) * C], rb[(2 * t2) * C]);                                            \
      const float2 d1 = make_float2(ra[(2 * t2 + 1) * C], rb[(2 * t2 + 1) * C]);                                    \
      u = (t2 == 0) ? f2_mul(d0, make_float2(w.x, w.y)) : f2_fma(d0, make_float2(w.x, w.y), u);                     \
      u = f2_fma(d1, make_float2(w.z, w.w), u);                                                                     \
    }                                                                                                               \
    if ((NT) & 1) {                                                                                                 \
      const float4 w = T.wt2[pp][(NT) / 2];                                                                         \
      u = f2_fma(make_float2(ra[((NT) - 1) * C], rb[((NT) - 1) * C]), make_float2(w.x, w.y), u);                    \
    }                                                                                                               \
    U2[pp] = u;                                                                                                     \
  }                                                                                                                 \
  {                                                                                                                 \
    const float* ra = rowc + toff[kP - 1];                                                                          \
    float u = 0.f;                                                                                                  \
    _Pragma("unroll") for (int t2 = 0; t2 < (NT) / 2; ++t2) {                                                       \
      const float4 w = T.wt2[3][t2];                                                                                \
      u = fmaf(w.x, ra[(2 * t2) * C], u);                                                                           \
      u = fmaf(w.z, ra[(2 * t2 + 1) * C], u);                                                                       \
    }                                                                                                               \
    if ((NT) & 1) u = fmaf(T.wt2[3][(NT) / 2].x, ra[((NT) - 1) * C], u);                                            \
    U2[3] = make_float2(u, 0.f);                                                                                    \
  }
      switch (ntaps) {   // CTA-uniform: taps per bin of this RoI
        case 2: OSR_NHWC_TAPS(2) break;
        case 3: OSR_NHWC_TAPS(3) break;
        case 4: OSR_NHWC_TAPS(4) break;
        case 5: OSR_NHWC_TAPS(5) break;
        case 6: OSR_NHWC_TAPS(6) break;
        default: OSR_NHWC_TAPS(8) break;
      }
#undef OSR_NHWC_TAPS
    } else {
      float U[kP + 1];
      U[kP] = 0.f;
#pragma unroll
      for (int pw = 0; pw < kP; ++pw) {
        float u = 0.f;
        const int xb = T.xb[pw];
        const int q0 = max(0, x_lo - xb), q1 = min(T.nx[pw], x_hi - xb);
        for (int q = q0; q < q1; ++q) u = fmaf(T.wx[pw * kRB + q], cin ? row[(xb + q - x_lo) * C] : 0.f, u);
        U[pw] = u;
      }
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) U2[pp] = make_float2(U[2 * pp], U[2 * pp + 1]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[slot]);
    // refill the slot with tile t + stages once every warp has released it
    if (warp == 0 && t + nstages < total) {
      if (elect_one()) {
        mbar_wait(&empty_bar[slot], parity);
        const int nxc2 = rxc, nr2 = rr;
        const int ncols = min(scols, wf - nxc2 * scols);
        const uint32_t bytes = (uint32_t)(ncols * C * 4);
        mbar_expect_tx(&full_bar[slot], bytes);
        bulk_load_1d(ring + slot * stage_floats, img_base + ((int64_t)(ymin + nr2) * lv.sH + (int64_t)(xmin + nxc2 * scols) * lv.sW), bytes, &full_bar[slot]);
      }
      __syncwarp();
    }
    // y fold: the row feeds the <= 3 consecutive bins that contain it (bin index is CTA-uniform)
    const float4 w = T.rw[r];
    switch (__float_as_int(w.w)) {
#define OSR_NHWC_CASE(PH)                                                                                   \
  case PH:                                                                                                  \
    _Pragma("unroll") for (int pp = 0; pp < 4; ++pp) {                                                      \
      acc2[PH][pp] = f2_fma_s(U2[pp], w.x, acc2[PH][pp]);                                                   \
      if (PH + 1 < kP) acc2[PH + 1 < kP ? PH + 1 : 0][pp] = f2_fma_s(U2[pp], w.y, acc2[PH + 1 < kP ? PH + 1 : 0][pp]); \
      if (PH + 2 < kP) acc2[PH + 2 < kP ? PH + 2 : 0][pp] = f2_fma_s(U2[pp], w.z, acc2[PH + 2 < kP ? PH + 2 : 0][pp]); \
    }                                                                                                       \
    break;
      OSR_NHWC_CASE(0) OSR_NHWC_CASE(1) OSR_NHWC_CASE(2) OSR_NHWC_CASE(3) OSR_NHWC_CASE(4) OSR_NHWC_CASE(5) OSR_NHWC_CASE(6)
#undef OSR_NHWC_CASE
      case -1: {  // row inside more than 3 bins: dense fold from the per-bin tables
        const int y = ymin + r;
#pragma unroll
        for (int ph = 0; ph < kP; ++ph) {
          const int rr = y - T.yb[ph];
          const float wy = (rr >= 0 && rr < T.ny[ph]) ? T.wy[ph * kRB + rr] : 0.f;
#pragma unroll
          for (int pp = 0; pp < 4; ++pp) acc2[ph][pp] = f2_fma_s(U2[pp], wy, acc2[ph][pp]);
        }
        break;
      }
      default: break;
    }
    if (++r == hf) { r = 0; ++xc; }
    if (++rr == hf) { rr = 0; ++rxc; }
    if (++slot == nstages) { slot = 0; parity ^= 1u; }
  }
  // epilogue: 49 x C tile -> shared memory as [c][49] (stride 49 is odd: conflict-free) -> coalesced 16-byte stores
  __syncthreads();   // every warp is done with the ring (all bulk loads have landed and been consumed)
  if (cin) {
    float* o = ring + c * (kP * kP);
#pragma unroll
    for (int a = 0; a < kP; ++a)
#pragma unroll
      for (int b = 0; b < kP; ++b) o[a * kP + b] = ((b & 1) ? acc2[a][b >> 1].y : acc2[a][b >> 1].x) * inv_count;
  }
  __syncthreads();
  const int n4 = (C * kP * kP) >> 2;
  if (out_bf) {   // bf16 pooled output for the tensor-core box head: 8 values per 16-byte store (C % 8 == 0 host-checked)
    for (int i = tid; i < (n4 >> 1); i += kThreads) {
      const float4 a = reinterpret_cast<const float4*>(ring)[2 * i], b = reinterpret_cast<const float4*>(ring)[2 * i + 1];
      __nv_bfloat162 b0 = __floats2bfloat162_rn(a.x, a.y), b1 = __floats2bfloat162_rn(a.z, a.w);
      __nv_bfloat162 b2 = __floats2bfloat162_rn(b.x, b.y), b3 = __floats2bfloat162_rn(b.z, b.w);
      uint4 o;
      o.x = *reinterpret_cast<uint32_t*>(&b0); o.y = *reinterpret_cast<uint32_t*>(&b1);
      o.z = *reinterpret_cast<uint32_t*>(&b2); o.w = *reinterpret_cast<uint32_t*>(&b3);
      reinterpret_cast<uint4*>(out_bf)[i] = o;
    }
  } else if ((((uintptr_t)out_roi) & 15) == 0 && ((C * kP * kP) & 3) == 0) {
    for (int i = tid; i < n4; i += kThreads) reinterpret_cast<float4*>(out_roi)[i] = reinterpret_cast<const float4*>(ring)[i];
  } else {
    for (int i = tid; i < C * kP * kP; i += kThreads) out_roi[i] = ring[i];
  }
}

// One CTA per RoI (the shipped form): dynamic smem = [ ring | barriers | Tables ]
template <int kC>
__global__ void __launch_bounds__(kThreads, 2) roi_align_fwd_nhwc_kernel(const __grid_constant__ FwdParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* ring = reinterpret_cast<float*>(smem_raw);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(ring + p.ring_floats);
  nhwc_roi<kC, true>(p, blockIdx.x, ring, s_bar, *reinterpret_cast<Tables*>(s_bar + 2 * kNhwcMaxStages));
}

template <int kC>
__device__ __noinline__ void nhwc_roi_slow(const FwdParams& p, int j, float* ring, uint64_t* s_bar, Tables& T) {
  nhwc_roi<kC, false>(p, j, ring, s_bar, T);
}

// ---------------------------------------------------------------------------------------------------------
// Persistent form of the channels_last forward (opt-in, OSR_TUNE_FWD_VARIANT = 4): 2 CTAs per SM walk the records in
// processing order (a global counter hands out the next record), so the per-RoI fixed costs of the one-CTA-per-RoI kernel
// leave the critical path.  MEASURED RESULT (B200, cfg 2): 0.506 ms against 0.499 ms for one CTA per RoI with identical
// output - hiding the prologue buys nothing, because the kernel is not bound by per-RoI latency but by the shared-memory
// data pipe (row-loop LDS + bulk-copy fills + epilogue transposition = ~70 % of its cycles; ablations in DESIGN.md 4).
// Kept as the documented experiment and as a second implementation the tests cross-check.  What it overlaps:
//   * the record of the NEXT RoI (1.5 KB: header, tap offsets / weights, row weights) is fetched into shared memory by one
//     bulk copy while the current RoI is processed - the row loop reads its tables straight from that copy, no dependent
//     global loads, no table-building prologue;
//   * the first rows of the next RoI are requested BEFORE the current RoI's epilogue (the 49 x C output tile is staged in
//     the upper 49 columns of the ring, the early rows go to the stages below it), so the row loop of the next RoI starts
//     on data that is already in flight for the whole epilogue;
//   * the ring's mbarriers live for the whole kernel: a slot's phase parity is one bit of `cpar` (flipped per consumed
//     tile; fills and consumptions of a slot alternate, so the same bit serves the full and the empty barrier).
// Records that are not FAST (footprint > 48 x 64, bins > 8 pixels, rows in > 3 bins, off-pyramid) are rare and go through
// the general per-RoI body (nhwc_roi_slow) on a private barrier set.
// dynamic smem: [ ring 96 x C | full[6] empty[6] | slow-body barriers[12] | Tables (slow body) | 2 records | rec barriers[2] | ctl[2] ]
constexpr int kTileCols = kP * kP;   // the output tile of a RoI takes 49 columns (x C floats) of the ring

template <int kC>
__global__ void __launch_bounds__(kThreads, 2) roi_align_fwd_nhwc_pers_kernel(const __grid_constant__ FwdParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* ring = reinterpret_cast<float*>(smem_raw);
  const int C = kC > 0 ? kC : p.L.C;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + p.ring_floats);
  uint64_t* empty_bar = full_bar + kNhwcMaxStages;
  uint64_t* slow_bar = empty_bar + kNhwcMaxStages;
  Tables& T = *reinterpret_cast<Tables*>(slow_bar + 2 * kNhwcMaxStages);
  float* recbuf = reinterpret_cast<float*>(&T + 1);
  uint64_t* rec_bar = reinterpret_cast<uint64_t*>(recbuf + 2 * kRecFloats);
  volatile int* s_ctl = reinterpret_cast<volatile int*>(rec_bar + 2);   // record index held by each record buffer (-1: none)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = tid;
  const bool cin = c < C;
  const int tile_off = p.ring_floats - kTileCols * C;   // floats; the ring is 96 columns, the tile its upper 49
  const int fit_cols = tile_off / C;                    // early stages of the next RoI must end below the tile
  const uint32_t rec_bytes = kRecFloats * 4;
  int jn = 0, mn = -1;   // (thread 0) the position in processing order this CTA handles next, and its RoI
  auto roi_at = [&](int j) { return j < p.M ? (p.order ? __ldg(p.order + j) : j) : -1; };
  if (tid == 0) {
    for (int i = 0; i < kNhwcMaxStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kWarps);
      mbar_init(&slow_bar[i], 1);
      mbar_init(&slow_bar[kNhwcMaxStages + i], kWarps);
    }
    mbar_init(&rec_bar[0], 1);
    mbar_init(&rec_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_ctl[0] = blockIdx.x;   // the host launches at most M CTAs
    s_ctl[1] = -1;
    mbar_expect_tx(&rec_bar[0], rec_bytes);
    bulk_load_1d(recbuf, p.rec + (int64_t)roi_at(blockIdx.x) * kRecFloats, rec_bytes, &rec_bar[0]);
    jn = atomicAdd(p.counter, 1);
    mn = roi_at(jn);
  }
  // Zero-weight padded taps read whatever the ring holds: zero it once, afterwards it only ever holds feature values and
  // output tiles (finite whenever the inputs are).
  for (int i = tid; i < (p.ring_floats >> 2); i += kThreads) reinterpret_cast<float4*>(ring)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();

  int buf = 0;
  uint32_t rpar = 0;       // bit b: phase parity of rec_bar[b]
  uint32_t cpar = 0;       // bit s: phase parity of ring slot s
  bool after_fast = false; // the previous RoI went through the fast path (and may have requested my first rows)
  while (true) {
    const int j = s_ctl[buf];
    if (j < 0) break;
    mbar_wait(&rec_bar[buf], (rpar >> buf) & 1u);
    const float* rs = recbuf + buf * kRecFloats;
    const int4 h0 = *reinterpret_cast<const int4*>(rs);
    const float4 h1 = *reinterpret_cast<const float4*>(rs + 4);
    if (tid == 0) {   // fetch the record after this one into the other buffer (its last readers passed the previous RoI's barriers)
      const int nb = buf ^ 1;
      if (jn < p.M) {
        s_ctl[nb] = jn;
        mbar_expect_tx(&rec_bar[nb], rec_bytes);
        bulk_load_1d(recbuf + nb * kRecFloats, p.rec + (int64_t)mn * kRecFloats, rec_bytes, &rec_bar[nb]);
        jn = atomicAdd(p.counter, 1);
        mn = roi_at(jn);   // consumed one RoI later
      } else {
        s_ctl[nb] = -1;
      }
    }
    if ((h0.x & 1) == 0) {   // not FAST: general body on its own barriers
      __syncthreads();       // the previous RoI's tile has been copied out
      nhwc_roi_slow<kC>(p, j, ring, slow_bar, T);
      __syncthreads();
      rpar ^= 1u << buf;
      buf ^= 1;
      after_fast = false;
      continue;
    }
    const int level = h0.x >> 8, xmin = h0.y, ymin = h0.z, wf = h0.w & 0xffff, hf = h0.w >> 16;
    const float inv_count = h1.x;
    const int tmax = __float_as_int(h1.y), img = __float_as_int(h1.z), m = __float_as_int(h1.w);
    if (tid == 0) p.out_level[m] = level;
    int toff[kP];
    {
      const int4 o0 = *reinterpret_cast<const int4*>(rs + 8), o1 = *reinterpret_cast<const int4*>(rs + 12);
      toff[0] = o0.x * C; toff[1] = o0.y * C; toff[2] = o0.z * C; toff[3] = o0.w * C;
      toff[4] = o1.x * C; toff[5] = o1.y * C; toff[6] = o1.z * C;
    }
    const float4* wt2 = reinterpret_cast<const float4*>(rs + kRecW2);   // [pp * 4 + t2]
    const float4* rw = reinterpret_cast<const float4*>(rs + 72);
    const LevelDesc& lv = p.L.lv[level];
    const float* img_base = lv.data + (int64_t)img * lv.sN + ((int64_t)ymin * lv.sH + (int64_t)xmin * lv.sW);
    const int scols = max(8, (wf + 3) & ~3);   // FAST records are at most 48 columns wide: a stage is one whole footprint row
    const int nstages = min(p.max_stages, kNhwcRingCols / scols);
    const int stage_floats = scols * C;
    const uint32_t row_bytes = (uint32_t)(wf * C * 4);
    const int total = hf;
    const int n_first = min(nstages, total);
    const int ntaps = nhwc_tap_count(tmax);
    const int n_early = after_fast ? min(n_first, fit_cols / scols) : 0;   // already requested by the previous RoI's epilogue

    float2 acc2[kP][4];
#pragma unroll
    for (int a = 0; a < kP; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc2[a][b] = make_float2(0.f, 0.f);

    __syncthreads();   // (C) the previous RoI's tile has been copied out: the whole ring is free
    if (warp == 0 && n_early < n_first) {
      if (elect_one()) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy tile traffic before async-proxy refills
        for (int t = n_early; t < n_first; ++t) {
          mbar_expect_tx(&full_bar[t], row_bytes);
          bulk_load_1d(ring + t * stage_floats, img_base + (int64_t)t * lv.sH, row_bytes, &full_bar[t]);
        }
      }
      __syncwarp();
    }
    int slot = 0;
    const float* rowc0 = ring + min(c, C - 1);
    for (int t = 0; t < total; ++t) {
      const uint32_t parity = (cpar >> slot) & 1u;
      mbar_wait(&full_bar[slot], parity);
      const float* rowc = rowc0 + slot * stage_floats;
      float2 U2[4];
#define OSR_NHWC_TAPS(NT)                                                                                          \
  _Pragma("unroll") for (int pp = 0; pp < 3; ++pp) {                                                                \
    const float* ra = rowc + toff[2 * pp];                                                                          \
    const float* rb = rowc + toff[2 * pp + 1];                                                                      \
    float2 u = make_float2(0.f, 0.f);                                                                               \
    _Pragma("unroll") for (int t2 = 0; t2 < (NT) / 2; ++t2) {                                                       \
      const float4 w = wt2[pp * 4 + t2];                                                                            \
      const float2 d0 = make_float2(ra[(2 * t2) * C], rb[(2 * t2) * C]);                                            \
      const float2 d1 = make_float2(ra[(2 * t2 + 1) * C], rb[(2 * t2 + 1) * C]);                                    \
      u = (t2 == 0) ? f2_mul(d0, make_float2(w.x, w.y)) : f2_fma(d0, make_float2(w.x, w.y), u);                     \
      u = f2_fma(d1, make_float2(w.z, w.w), u);                                                                     \
    }                                                                                                               \
    if ((NT) & 1) {                                                                                                 \
      const float4 w = wt2[pp * 4 + (NT) / 2];                                                                      \
      u = f2_fma(make_float2(ra[((NT) - 1) * C], rb[((NT) - 1) * C]), make_float2(w.x, w.y), u);                    \
    }                                                                                                               \
    U2[pp] = u;                                                                                                     \
  }                                                                                                                 \
  {                                                                                                                 \
    const float* ra = rowc + toff[kP - 1];                                                                          \
    float u = 0.f;                                                                                                  \
    _Pragma("unroll") for (int t2 = 0; t2 < (NT) / 2; ++t2) {                                                       \
      const float4 w = wt2[12 + t2];                                                                                \
      u = fmaf(w.x, ra[(2 * t2) * C], u);                                                                           \
      u = fmaf(w.z, ra[(2 * t2 + 1) * C], u);                                                                       \
    }                                                                                                               \
    if ((NT) & 1) u = fmaf(wt2[12 + (NT) / 2].x, ra[((NT) - 1) * C], u);                                            \
    U2[3] = make_float2(u, 0.f);                                                                                    \
  }
      switch (ntaps) {   // CTA-uniform: taps per bin of this RoI
        case 2: OSR_NHWC_TAPS(2) break;
        case 3: OSR_NHWC_TAPS(3) break;
        case 4: OSR_NHWC_TAPS(4) break;
        case 5: OSR_NHWC_TAPS(5) break;
        case 6: OSR_NHWC_TAPS(6) break;
        default: OSR_NHWC_TAPS(8) break;
      }
#undef OSR_NHWC_TAPS
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[slot]);
      // refill the slot with row t + stages once every warp has released it
      if (warp == 0 && t + nstages < total) {
        if (elect_one()) {
          mbar_wait(&empty_bar[slot], parity);
          mbar_expect_tx(&full_bar[slot], row_bytes);
          bulk_load_1d(ring + slot * stage_floats, img_base + (int64_t)(t + nstages) * lv.sH, row_bytes, &full_bar[slot]);
        }
        __syncwarp();
      }
      cpar ^= 1u << slot;
      // y fold: the row feeds the <= 3 consecutive bins that contain it (bin index is CTA-uniform)
      const float4 w = rw[t];
      switch (__float_as_int(w.w)) {
#define OSR_NHWC_CASE(PH)                                                                                   \
  case PH:                                                                                                  \
    _Pragma("unroll") for (int pp = 0; pp < 4; ++pp) {                                                      \
      acc2[PH][pp] = f2_fma_s(U2[pp], w.x, acc2[PH][pp]);                                                   \
      if (PH + 1 < kP) acc2[PH + 1 < kP ? PH + 1 : 0][pp] = f2_fma_s(U2[pp], w.y, acc2[PH + 1 < kP ? PH + 1 : 0][pp]); \
      if (PH + 2 < kP) acc2[PH + 2 < kP ? PH + 2 : 0][pp] = f2_fma_s(U2[pp], w.z, acc2[PH + 2 < kP ? PH + 2 : 0][pp]); \
    }                                                                                                       \
    break;
        OSR_NHWC_CASE(0) OSR_NHWC_CASE(1) OSR_NHWC_CASE(2) OSR_NHWC_CASE(3) OSR_NHWC_CASE(4) OSR_NHWC_CASE(5) OSR_NHWC_CASE(6)
#undef OSR_NHWC_CASE
        default: break;
      }
      if (++slot == nstages) slot = 0;
    }
    __syncthreads();   // (A) every warp has consumed every row: all ring slots are free, their barriers idle
    // Request the first rows of the NEXT RoI now (stages that end below the output tile), so they fly during the epilogue.
    if (tid == 0) {
      const int nb = buf ^ 1;
      if (s_ctl[nb] >= 0) {
        mbar_wait(&rec_bar[nb], (rpar >> nb) & 1u);   // landed long ago
        const float* rn = recbuf + nb * kRecFloats;
        const int4 g0 = *reinterpret_cast<const int4*>(rn);
        if (g0.x & 1) {
          const int wfn = g0.w & 0xffff, hfn = g0.w >> 16;
          const int scn = max(8, (wfn + 3) & ~3);
          const int ne = min(min(min(p.max_stages, kNhwcRingCols / scn), hfn), fit_cols / scn);
          const LevelDesc& ln = p.L.lv[g0.x >> 8];
          const float* nbase = ln.data + (int64_t)__float_as_int(rn[6]) * ln.sN + ((int64_t)g0.z * ln.sH + (int64_t)g0.y * ln.sW);
          const uint32_t nbytes = (uint32_t)(wfn * C * 4);
          for (int t = 0; t < ne; ++t) {
            mbar_expect_tx(&full_bar[t], nbytes);
            bulk_load_1d(ring + t * (scn * C), nbase + (int64_t)t * ln.sH, nbytes, &full_bar[t]);
          }
        }
      }
    }
    // epilogue: 49 x C tile -> shared memory as [c][49] (stride 49 is odd: conflict-free) -> coalesced 16-byte stores
    float* tile = ring + tile_off;
    if (cin) {
      float* o = tile + c * (kP * kP);
#pragma unroll
      for (int a = 0; a < kP; ++a)
#pragma unroll
        for (int b = 0; b < kP; ++b) o[a * kP + b] = ((b & 1) ? acc2[a][b >> 1].y : acc2[a][b >> 1].x) * inv_count;
    }
    __syncthreads();   // (B)
    const int n4 = (C * kP * kP) >> 2;
    if (p.out_bf16) {   // bf16 pooled output for the tensor-core box head: 8 values per 16-byte store (C % 8 == 0 host-checked)
      unsigned short* out_bf = p.out_bf16 + (int64_t)m * C * (kP * kP);
      for (int i = tid; i < (n4 >> 1); i += kThreads) {
        const float4 a = reinterpret_cast<const float4*>(tile)[2 * i], b = reinterpret_cast<const float4*>(tile)[2 * i + 1];
        __nv_bfloat162 b0 = __floats2bfloat162_rn(a.x, a.y), b1 = __floats2bfloat162_rn(a.z, a.w);
        __nv_bfloat162 b2 = __floats2bfloat162_rn(b.x, b.y), b3 = __floats2bfloat162_rn(b.z, b.w);
        uint4 o;
        o.x = *reinterpret_cast<uint32_t*>(&b0); o.y = *reinterpret_cast<uint32_t*>(&b1);
        o.z = *reinterpret_cast<uint32_t*>(&b2); o.w = *reinterpret_cast<uint32_t*>(&b3);
        reinterpret_cast<uint4*>(out_bf)[i] = o;
      }
    } else {
      float* out_roi = p.out + (int64_t)m * C * (kP * kP);   // 16-byte aligned: host-checked for this kernel
      for (int i = tid; i < n4; i += kThreads) reinterpret_cast<float4*>(out_roi)[i] = reinterpret_cast<const float4*>(tile)[i];
    }
    rpar ^= 1u << buf;
    buf ^= 1;
    after_fast = true;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Processing order: RoIs bucketed by (image, level, y band) with a single-CTA counting sort, so that CTAs that
// run concurrently read neighbouring feature rows (L2-resident working set instead of a whole image's pyramid).
// The order only changes scheduling: every RoI still writes its own output rows, results are order-independent.
constexpr int kSortThreads = 1024;
constexpr int kYBands = 16;
constexpr int kMaxBuckets = 8192;

__device__ __forceinline__ void roi_order_body(const FwdParams& p, int32_t* order, int32_t* keys, int ybands, int* hist) {
  const int tid = threadIdx.x;
  const int nb = p.L.num_images * p.L.num_levels * ybands;
  for (int i = tid; i < nb; i += kSortThreads) hist[i] = 0;
  __syncthreads();
  for (int m = tid; m < p.M; m += kSortThreads) {
    const float* roi = p.rois + (int64_t)m * 5;
    const float x1 = __ldg(roi + 1), y1 = __ldg(roi + 2), x2 = __ldg(roi + 3), y2 = __ldg(roi + 4);
    int img = (int)__ldg(roi);
    int level = assign_level(x1, y1, x2, y2, p.L);
    img = min(max(img, 0), p.L.num_images - 1);
    level = min(max(level, 0), p.L.num_levels - 1);
    const LevelDesc& lv = p.L.lv[level];
    const float yc = 0.5f * (y1 + y2) * lv.scale;
    int band = (int)(yc * (float)ybands / (float)lv.H);
    band = min(max(band, 0), ybands - 1);
    const int key = (img * p.L.num_levels + level) * ybands + band;
    keys[m] = key;
    atomicAdd(&hist[key], 1);
  }
  __syncthreads();
  // exclusive scan of the buckets (single warp, nb <= 8192)
  if (tid < 32) {
    int run = 0;
    for (int base = 0; base < nb; base += 32) {
      const int i = base + tid;
      const int v = i < nb ? hist[i] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (tid >= o) inc += t;
      }
      if (i < nb) hist[i] = run + inc - v;
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  __syncthreads();
  for (int m = tid; m < p.M; m += kSortThreads) order[atomicAdd(&hist[keys[m]], 1)] = m;
}

__global__ void __launch_bounds__(kSortThreads) roi_order_kernel(const __grid_constant__ FwdParams p, int32_t* order,
                                                                int32_t* keys, int ybands) {
  __shared__ int hist[kMaxBuckets];
  roi_order_body(p, order, keys, ybands, hist);
}

// Ordering and table records in ONE launch: CTA 0 runs the counting sort, every other CTA (32 warps) writes the records of
// 32 RoIs.  The two are independent (records are indexed by RoI, the order only says who goes first) and both are
// latency-bound single passes: 13 + 14 us back to back, ~14 us side by side.
constexpr int kFusedPrepWarps = kSortThreads / 32;
__global__ void __launch_bounds__(kSortThreads) roi_order_prep_kernel(const __grid_constant__ FwdParams p, int32_t* order,
                                                                     int32_t* keys, int ybands) {
  extern __shared__ __align__(16) unsigned char prep_smem[];
  if (blockIdx.x == 0) {
    roi_order_body(p, order, keys, ybands, reinterpret_cast<int*>(prep_smem));
    return;
  }
  const int warp = threadIdx.x >> 5;
  roi_fwd_prep_body(p, (blockIdx.x - 1) * kFusedPrepWarps + warp, reinterpret_cast<PrepScratch*>(prep_smem)[warp]);
}

size_t fwd_smem_bytes(int ring_floats) {
  return (size_t)(ring_floats + kWarps * kP * (kTX + 1) + 8) * 4 + 2 * kNS * 8 + sizeof(Tables) + 16;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// 2-D tensor map (x, rows = N*C*H) of one contiguous NCHW pyramid level; returns 1 if the level can be streamed by TMA
int encode_level_map(CUtensorMap* map, const LevelDesc& lv, int num_images, int C) {
  if (lv.sW != 1 || lv.sH != lv.W) return 0;                                  // rows contiguous
  if (lv.sC != (int64_t)lv.H * lv.W || (num_images > 1 && lv.sN != (int64_t)C * lv.H * lv.W)) return 0;  // planes contiguous
  if ((reinterpret_cast<uintptr_t>(lv.data) & 15) || ((lv.sH * 4) & 15)) return 0;  // 16-byte base and row stride
  const int64_t rows = (int64_t)(num_images > 0 ? num_images : 1) * C * lv.H;
  if (rows >= ((int64_t)1 << 31)) return 0;
  EncodeTiledFn fn = encode_fn();
  if (!fn) return 0;
  cuuint64_t dims[2] = {(cuuint64_t)lv.W, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)lv.sH * 4};
  cuuint32_t box[2] = {kTX, kTR};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, lv.data, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 1 : 0;
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

extern "C" {

size_t osr_roi_align_fwd_workspace(int M) {
  const size_t m = (size_t)(M > 0 ? M : 1);
  return osr::align256(m * 4) * 2 + osr::align256(m * kRecFloats * 4) + 256;   // order, keys, per-RoI table records, work counter
}

static int roi_align_fwd_impl(const osr_feat_level_t* h_levels, int num_levels, int num_images, int C, const float* rois,
                              int M, int P, int sampling_ratio, int aligned, int canonical_box_size, int canonical_level,
                              int min_level, float* out, unsigned short* out_bf16, int32_t* out_level, void* workspace,
                              size_t workspace_bytes, void* stream) {
  osr::DeviceGuard device_guard(out_level);
  FwdParams p;
  p.counter = nullptr;
  p.pers_grid = 0;
  // ring cut into at most 6 row stages (shipped); 6: up to 12 for narrow footprints - measured 0.5 % SLOWER (0.468 vs 0.4657 ms
  // against the 6-stage build on the same box): the rows in flight are not what bounds narrow RoIs
  p.max_stages = osr::tuning(osr::kTuneFwdVariant) == 6 ? kNhwcMaxStages : 6;
  p.two_rows = osr::tuning(osr::kTuneFwdVariant) == 5 ? 0 : 1;   // 5: one row per iteration (A/B: 0.506 vs 0.471 ms at cfg 2)
  int rc = osr::fill_roi_levels(p.L, h_levels, num_levels, num_images, C, P, sampling_ratio, aligned,
                                canonical_box_size, canonical_level, min_level);
  if (rc) return rc;
  if (M < 0) return osr::fail_arg(OSR_E_ARG, "roi_align_fwd: M < 0");
  if (M == 0) return 0;
  if (!rois || (!out && !out_bf16) || !out_level) return osr::fail_arg(OSR_E_ARG, "roi_align_fwd: null pointer argument");
  p.rois = rois;
  p.M = M;
  p.out = out;
  p.out_bf16 = out_bf16;
  p.out_level = out_level;
  p.order = nullptr;
  p.rec = nullptr;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // locality ordering (optional: skipped without a workspace, or for tiny problems where it cannot pay); launched below,
  // fused with the table records when the channels_last kernel runs
  int ybands = 0;
  int32_t* order = nullptr;
  int32_t* keys = nullptr;
  if (workspace && workspace_bytes >= osr_roi_align_fwd_workspace(M) && M >= 256) {
    ybands = kYBands;
    while (ybands > 1 && num_images * num_levels * ybands > kMaxBuckets) ybands >>= 1;
    if (num_images * num_levels * ybands <= kMaxBuckets) {
      order = static_cast<int32_t*>(workspace);
      keys = reinterpret_cast<int32_t*>(static_cast<unsigned char*>(workspace) + osr::align256((size_t)M * 4));
    }
  }
  auto launch_order_alone = [&]() -> int {
    if (order) {
      roi_order_kernel<<<1, kSortThreads, 0, s>>>(p, order, keys, ybands);
      OSR_LAUNCH_CHECK();
      p.order = order;
    }
    return 0;
  };
  // TMA staging is opt-in (OSR_TUNE_FWD_VARIANT=1): on NCHW maps a footprint row is only ~50-130 bytes, and the TMA unit's
  // per-row request rate makes it slower than the LDG path (2.20 ms vs 1.71 ms at cfg2 on B200).
  // channels_last maps (sC == 1, a pixel's C channels contiguous, 16-byte aligned) take the bulk-copy NHWC kernel
  bool nhwc = (C <= kThreads) && (C % 4 == 0);
  for (int l = 0; l < num_levels && nhwc; ++l) {
    const LevelDesc& lv = p.L.lv[l];
    nhwc = lv.sC == 1 && lv.sW == C && lv.sH == (int64_t)lv.W * C && (lv.sN % 4 == 0) && ((reinterpret_cast<uintptr_t>(lv.data) & 15) == 0);
  }
  if (out_bf16 && !(nhwc && C % 8 == 0))
    return osr::fail_arg(OSR_E_SHAPE, "roi_align_fwd_bf16: needs dense channels_last feature maps with C %% 8 == 0 and C <= 256");
  if (nhwc) {
    const int variant = osr::tuning(osr::kTuneFwdVariant);
    const bool have_ws = workspace && workspace_bytes >= osr_roi_align_fwd_workspace(M) && variant != 2;
    // opt-in (OSR_TUNE_FWD_VARIANT = 4): persistent CTAs (2 per SM) over the prep records; needs the workspace and 16-byte
    // aligned output rows.  Measured equal to one CTA per RoI (0.506 vs 0.499 ms at cfg 2): see the kernel's header.
    const bool pers = have_ws && variant == 4 &&
                      ((reinterpret_cast<uintptr_t>(out_bf16 ? (const void*)out_bf16 : (const void*)out) & 15) == 0);
    for (int l = 0; l < num_levels; ++l) p.tma_ok[l] = 0;
    if (have_ws) {
      unsigned char* wsb = static_cast<unsigned char*>(workspace);
      p.rec = reinterpret_cast<float*>(wsb + 2 * osr::align256((size_t)M * 4));
      if (pers) {
        int dev = 0, sms = 0;
        OSR_CUDA_CHECK(cudaGetDevice(&dev));
        OSR_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        p.pers_grid = std::min(M, 2 * sms);
        p.counter = reinterpret_cast<int*>(wsb + 2 * osr::align256((size_t)M * 4) + osr::align256((size_t)M * kRecFloats * 4));
      }
      if (order) {   // counting sort (CTA 0) and records (the other CTAs) side by side in one launch
        const size_t psm = std::max(sizeof(PrepScratch) * (size_t)kFusedPrepWarps, sizeof(int) * (size_t)kMaxBuckets);
        OSR_CUDA_CHECK(cudaFuncSetAttribute(roi_order_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
        roi_order_prep_kernel<<<1 + osr::ceil_div(M, kFusedPrepWarps), kSortThreads, psm, s>>>(p, order, keys, ybands);
        OSR_LAUNCH_CHECK();
        p.order = order;   // (the prep CTAs do not read it)
      } else {
        roi_fwd_prep_kernel<<<osr::ceil_div(M, kPrepWarpsF), kPrepWarpsF * 32, 0, s>>>(p);
        OSR_LAUNCH_CHECK();
      }
    } else if ((rc = launch_order_alone())) {
      return rc;
    }
    if (pers) {
      p.ring_floats = kNhwcRingCols * C;   // 96 columns; the upper 49 double as the output tile
      const size_t smem = (size_t)p.ring_floats * 4 + 4 * kNhwcMaxStages * 8 + sizeof(Tables) + 2 * kRecFloats * 4 + 2 * 8 + 16;
      if (C == 256) {
        OSR_CUDA_CHECK(cudaFuncSetAttribute(roi_align_fwd_nhwc_pers_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        roi_align_fwd_nhwc_pers_kernel<256><<<p.pers_grid, kThreads, smem, s>>>(p);
      } else {
        OSR_CUDA_CHECK(cudaFuncSetAttribute(roi_align_fwd_nhwc_pers_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        roi_align_fwd_nhwc_pers_kernel<0><<<p.pers_grid, kThreads, smem, s>>>(p);
      }
    } else {
      int ring = (kNhwcRingCols + 8) * C;              // + 8 columns of slack for the zero-weight padded taps
      if (ring < kP * kP * C) ring = kP * kP * C;                 // the ring doubles as the 49 x C output tile
      p.ring_floats = (ring + 31) & ~31;
      const size_t smem = (size_t)p.ring_floats * 4 + 2 * kNhwcMaxStages * 8 + sizeof(Tables) + 16;
      if (C == 256) {
        OSR_CUDA_CHECK(cudaFuncSetAttribute(roi_align_fwd_nhwc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        roi_align_fwd_nhwc_kernel<256><<<M, kThreads, smem, s>>>(p);
      } else {
        OSR_CUDA_CHECK(cudaFuncSetAttribute(roi_align_fwd_nhwc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        roi_align_fwd_nhwc_kernel<0><<<M, kThreads, smem, s>>>(p);
      }
    }
    OSR_LAUNCH_CHECK();
    return 0;
  }
  if ((rc = launch_order_alone())) return rc;
  FwdTma tm;
  memset(&tm, 0, sizeof(tm));
  const bool use_tma = osr::tuning(osr::kTuneFwdVariant) == 1;
  for (int l = 0; l < num_levels; ++l)
    p.tma_ok[l] = use_tma ? encode_level_map(&tm.map[l], p.L.lv[l], num_images, C) : 0;
  p.ring_floats = use_tma ? kRingFloats : kWarps * kWarpTile;
  const size_t smem = fwd_smem_bytes(p.ring_floats);
  if (use_tma) {
    OSR_CUDA_CHECK(cudaFuncSetAttribute(roi_align_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roi_align_fwd_kernel<true><<<M, kThreads, smem, s>>>(p, tm);
  } else {
    OSR_CUDA_CHECK(cudaFuncSetAttribute(roi_align_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    roi_align_fwd_kernel<false><<<M, kThreads, smem, s>>>(p, tm);
  }
  OSR_LAUNCH_CHECK();
  return 0;
}


int osr_roi_align_fwd(const osr_feat_level_t* h_levels, int num_levels, int num_images, int C, const float* rois,
                      int M, int P, int sampling_ratio, int aligned, int canonical_box_size, int canonical_level,
                      int min_level, float* out, int32_t* out_level, void* workspace, size_t workspace_bytes,
                      void* stream) {
  return roi_align_fwd_impl(h_levels, num_levels, num_images, C, rois, M, P, sampling_ratio, aligned, canonical_box_size,
                            canonical_level, min_level, out, nullptr, out_level, workspace, workspace_bytes, stream);
}

int osr_roi_align_fwd_bf16(const osr_feat_level_t* h_levels, int num_levels, int num_images, int C, const float* rois,
                           int M, int P, int sampling_ratio, int aligned, int canonical_box_size, int canonical_level,
                           int min_level, void* out_bf16, int32_t* out_level, void* workspace, size_t workspace_bytes,
                           void* stream) {
  if (M > 0 && (!out_bf16 || (reinterpret_cast<uintptr_t>(out_bf16) & 15)))
    return osr::fail_arg(OSR_E_ARG, "roi_align_fwd_bf16: null or misaligned output");
  return roi_align_fwd_impl(h_levels, num_levels, num_images, C, rois, M, P, sampling_ratio, aligned, canonical_box_size,
                            canonical_level, min_level, nullptr, static_cast<unsigned short*>(out_bf16), out_level, workspace,
                            workspace_bytes, stream);
}

}  // extern "C"
