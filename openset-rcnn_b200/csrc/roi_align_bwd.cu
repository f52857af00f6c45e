// Multi-level ROIAlignV2 backward for sm_100a: deterministic, atomic-free GATHER formulation.
// Replaces autograd's torchvision roi_align_backward (fp32 atomicAdd scatter, non-deterministic, plus a
// separate zero-fill of every dense gradient map) at train.py:145.
//
//   prep kernel   one thread per RoI: FPN level + conservative pixel bounding box of its footprint.
//   gather kernel one CTA per (image, level, 16x16-pixel tile).  The CTA scans the RoIs of its image in index
//                 order, keeps those whose box meets the tile (ordered => fixed accumulation order), builds
//                 their separable weight tables restricted to the tile (same sample arithmetic as the
//                 forward => exact adjoint), and then, 16 channels at a time, every thread (= one pixel)
//                 accumulates   g[c][y][x] += Wy[ph][y] * Wx[pw][x] / count * grad_out[roi][c][ph][pw]
//                 in registers over the RoI list and writes its pixel ONCE with a plain store - which also
//                 provides the zero fill of untouched pixels.  No atomics, no memset, run-to-run bit-identical.
#include "roi_geometry.cuh"

namespace osr {
int fill_roi_levels(RoiLevels& L, const osr_feat_level_t* h_levels, int num_levels, int num_images, int C, int P,
                    int sampling_ratio, int aligned, int canonical_box_size, int canonical_level, int min_level);
}

namespace {

using namespace osr;

constexpr int kTH = 16, kTW = 16;       // tile
constexpr int kThreads = kTH * kTW;     // one thread per tile pixel
constexpr int kCC = 32;                 // channels accumulated in registers per pass
constexpr int kNB = 32;                 // RoIs whose tables are resident in shared memory at once
constexpr int kG = 4;                   // RoIs whose grad_out chunk is staged per pipeline stage
constexpr int kBlk = kCC * kP * kP;     // floats of one (RoI, channel-chunk) block of grad_out = 784 (16-byte multiple)
constexpr int kWin = 4 * kThreads;      // RoI indices scanned per batch (4 per thread, in order)

struct RoiInfo {
  int x0, x1, y0, y1;  // inclusive pixel bounds of the footprint (conservative); x1 < x0 => empty
  int level;
};

struct BwdParams {
  RoiLevels L;
  const float* grad_out;
  const float* rois;
  const int32_t* roi_off;  // (N+1)
  int M;
  RoiInfo* info;           // (M) workspace
  int tile_base[OSR_MAX_LEVELS + 1];  // first tile id of each level (tiles ordered level, image, ty, tx)
  int tiles_x[OSR_MAX_LEVELS], tiles_y[OSR_MAX_LEVELS];
};

__device__ __forceinline__ void axis_bounds(float start, float bin, int grid, int L, int* lo, int* hi) {
  // valid samples lie in [start, start + 7*bin]; rows touched = floor(clamped sample) and +1.
  if (grid <= 0) {
    *lo = 1; *hi = 0;
    return;
  }
  const float cmin = start, cmax = start + (float)kP * bin;
  if (cmax < -2.0f || cmin > (float)L + 1.0f || !(cmax >= cmin)) {
    *lo = 1; *hi = 0;
    return;
  }
  int a = (int)floorf(fmaxf(cmin, 0.f)) - 1;
  int b = (int)floorf(fminf(fmaxf(cmax, 0.f), (float)L)) + 2;
  *lo = max(a, 0);
  *hi = min(b, L - 1);
}

__global__ void __launch_bounds__(256) roi_bwd_prep_kernel(const __grid_constant__ BwdParams p) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= p.M) return;
  const float* roi = p.rois + (int64_t)m * 5;
  const float x1 = __ldg(roi + 1), y1 = __ldg(roi + 2), x2 = __ldg(roi + 3), y2 = __ldg(roi + 4);
  RoiInfo r;
  r.level = assign_level(x1, y1, x2, y2, p.L);
  r.x0 = 1; r.x1 = 0; r.y0 = 1; r.y1 = 0;
  if (r.level >= 0 && r.level < p.L.num_levels) {
    const LevelDesc& lv = p.L.lv[r.level];
    const RoiGeom g = roi_geometry(x1, y1, x2, y2, lv.scale, p.L.sampling_ratio);
    axis_bounds(g.start_w, g.bin_w, g.grid_w, lv.W, &r.x0, &r.x1);
    axis_bounds(g.start_h, g.bin_h, g.grid_h, lv.H, &r.y0, &r.y1);
    if (r.y1 < r.y0) { r.x0 = 1; r.x1 = 0; }
  }
  p.info[m] = r;
}

// Accumulate the separable weights of output bin `p` that land on rows [t0, t0+tn) into w[(row - t0) * kP + p].
// Sample arithmetic is identical to build_bin_weights (forward) => the backward is the exact adjoint.
__device__ __forceinline__ void tile_bin_weights(float start, float bin, int grid, int L, int p, int t0, int tn, float* w) {
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
    const int rl = lo - t0, rh = hi - t0;
    if (rl >= 0 && rl < tn) w[rl * kP + p] += h;
    if (rh >= 0 && rh < tn) w[rh * kP + p] += l;
  }
}

struct BatchEntry {
  int m;
  float inv_count;
};

struct __align__(16) BwdSmem {
  float sg[2][kG][kBlk];  // double-buffered staging of grad_out[(roi, c0 .. c0+kCC), 7, 7]
  float wy[kNB][kTH * kP];
  float wx[kNB][kTW * kP];
  uchar2 yi[kNB][kTH];  // (first bin, number of bins) with non-zero weight for each tile row
  uchar2 xi[kNB][kTW];
  BatchEntry e[kNB];
  int warp_cnt[kThreads / 32 + 1];
  int nb, next_pos;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Scan RoI indices [pos, min(pos + kWin, r1)) in order, keep up to kNB that hit the tile; returns count in S.nb and
// the next scan position in S.next_pos.
__device__ __forceinline__ void collect_batch(const BwdParams& p, BwdSmem& S, int level, int tx0, int ty0, int pos, int r1) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int hit[4], cnt = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int m = pos + tid * 4 + k;
    hit[k] = 0;
    if (m < r1) {
      const RoiInfo r = p.info[m];
      hit[k] = (r.level == level) && (r.x0 <= tx0 + kTW - 1) && (r.x1 >= tx0) && (r.y0 <= ty0 + kTH - 1) && (r.y1 >= ty0);
    }
    cnt += hit[k];
  }
  int inc = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) S.warp_cnt[warp] = inc;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int w = 0; w < kThreads / 32; ++w) {
      int c = S.warp_cnt[w];
      S.warp_cnt[w] = run;
      run += c;
    }
    S.warp_cnt[kThreads / 32] = run;
    S.nb = min(run, kNB);
    S.next_pos = min(pos + kWin, r1);  // overwritten below if more than kNB hits
  }
  __syncthreads();
  int rank = S.warp_cnt[warp] + inc - cnt;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (hit[k]) {
      const int m = pos + tid * 4 + k;
      if (rank < kNB) S.e[rank].m = m;
      else if (rank == kNB) S.next_pos = m;  // first hit that did not fit: resume here
      ++rank;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void build_tables(const BwdParams& p, BwdSmem& S, const LevelDesc& lv, int tx0, int ty0) {
  const int tid = threadIdx.x;
  const int nb = S.nb;
  float* wyf = &S.wy[0][0];
  float* wxf = &S.wx[0][0];
  for (int i = tid; i < nb * kTH * kP; i += kThreads) wyf[i] = 0.f;
  for (int i = tid; i < nb * kTW * kP; i += kThreads) wxf[i] = 0.f;
  __syncthreads();
  for (int q = tid; q < nb * 2 * kP; q += kThreads) {
    const int j = q / (2 * kP);
    const int ab = q - j * (2 * kP);
    const float* roi = p.rois + (int64_t)S.e[j].m * 5;
    const RoiGeom g = roi_geometry(__ldg(roi + 1), __ldg(roi + 2), __ldg(roi + 3), __ldg(roi + 4), lv.scale, p.L.sampling_ratio);
    if (ab < kP) {
      tile_bin_weights(g.start_h, g.bin_h, g.grid_h, lv.H, ab, ty0, kTH, S.wy[j]);
      if (ab == 0) S.e[j].inv_count = 1.0f / g.count;
    } else {
      tile_bin_weights(g.start_w, g.bin_w, g.grid_w, lv.W, ab - kP, tx0, kTW, S.wx[j]);
    }
  }
  __syncthreads();
  for (int q = tid; q < nb * (kTH + kTW); q += kThreads) {
    const int j = q / (kTH + kTW);
    const int r = q - j * (kTH + kTW);
    const float* w = (r < kTH) ? &S.wy[j][r * kP] : &S.wx[j][(r - kTH) * kP];
    int lo = kP, hi = -1;
#pragma unroll
    for (int b = 0; b < kP; ++b)
      if (w[b] != 0.f) {
        lo = min(lo, b);
        hi = b;
      }
    const uchar2 v = make_uchar2((unsigned char)(hi >= 0 ? lo : 0), (unsigned char)(hi >= 0 ? hi - lo + 1 : 0));
    if (r < kTH) S.yi[j][r] = v;
    else S.xi[j][r - kTH] = v;
  }
  __syncthreads();
}

// stage grad_out blocks of RoIs [j0, min(j0 + kG, nb)) for channels [c0, c0 + kCC) into S.sg[buf] (async)
__device__ __forceinline__ void stage_group(const BwdParams& p, BwdSmem& S, int buf, int j0, int nb, int c0, int C,
                                            bool vec_ok) {
  const int tid = threadIdx.x;
  const int nj = min(kG, nb - j0);
  const int cc = min(kCC, C - c0);
  if (vec_ok && cc == kCC) {
    for (int q = tid; q < nj * (kBlk / 4); q += kThreads) {
      const int jj = q / (kBlk / 4), v = q - jj * (kBlk / 4);
      const float* src = p.grad_out + ((int64_t)S.e[j0 + jj].m * C + c0) * (kP * kP) + v * 4;
      cp_async16(&S.sg[buf][jj][v * 4], src);
    }
  } else {
    for (int q = tid; q < nj * kBlk; q += kThreads) {
      const int jj = q / kBlk, v = q - jj * kBlk;
      S.sg[buf][jj][v] = (v < cc * kP * kP) ? __ldg(p.grad_out + ((int64_t)S.e[j0 + jj].m * C + c0) * (kP * kP) + v) : 0.f;
    }
  }
  cp_async_commit();
}

// accumulate the contributions of staged RoIs [j0, j0 + nj) to this thread's pixel, kCC channels
__device__ __forceinline__ void accumulate_group(const BwdSmem& S, int buf, int j0, int nj, int ty, int tx, float* acc) {
  for (int jj = 0; jj < nj; ++jj) {
    const int j = j0 + jj;
    const uchar2 yi = S.yi[j][ty];
    const uchar2 xi = S.xi[j][tx];
    if (yi.y == 0 || xi.y == 0) continue;
    const float ic = S.e[j].inv_count;
    const float* g = S.sg[buf][jj];
    for (int a = 0; a < yi.y; ++a) {
      const int ph = yi.x + a;
      const float wy = S.wy[j][ty * kP + ph] * ic;
      for (int b = 0; b < xi.y; ++b) {
        const int pw = xi.x + b;
        const float w = wy * S.wx[j][tx * kP + pw];
        const float* gp = g + ph * kP + pw;
#pragma unroll
        for (int c = 0; c < kCC; ++c) acc[c] = fmaf(w, gp[c * (kP * kP)], acc[c]);
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads, 2) roi_align_bwd_kernel(const __grid_constant__ BwdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdSmem& S = *reinterpret_cast<BwdSmem*>(smem_raw);

  // tile decode
  int t = blockIdx.x, level = 0;
  while (level + 1 < p.L.num_levels && t >= p.tile_base[level + 1]) ++level;
  t -= p.tile_base[level];
  const int per_img = p.tiles_x[level] * p.tiles_y[level];
  const int n = t / per_img;
  t -= n * per_img;
  const int ty0 = (t / p.tiles_x[level]) * kTH, tx0 = (t % p.tiles_x[level]) * kTW;
  const LevelDesc& lv = p.L.lv[level];
  const int tid = threadIdx.x;
  const int ty = tid / kTW, tx = tid % kTW;
  const int y = ty0 + ty, x = tx0 + tx;
  const bool inside = (y < lv.H) && (x < lv.W);
  float* gpix = lv.data + (int64_t)n * lv.sN + (int64_t)y * lv.sH + (int64_t)x * lv.sW;

  const int r0 = p.roi_off[n], r1 = p.roi_off[n + 1];
  const int C = p.L.C;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(p.grad_out) & 15) == 0) && ((C * kP * kP) % 4 == 0);
  const bool nhwc_vec = (lv.sC == 1) && ((lv.sW & 3) == 0) && ((lv.sH & 3) == 0) && ((lv.sN & 3) == 0) &&
                        ((reinterpret_cast<uintptr_t>(lv.data) & 15) == 0);

  // first batch; if it covers the whole image's RoI range the tables are built once and reused by every channel pass
  collect_batch(p, S, level, tx0, ty0, r0, r1);
  const bool single = (S.next_pos >= r1);
  if (S.nb > 0) build_tables(p, S, lv, tx0, ty0);

  if (single) {
    const int nb = S.nb;
    const int nchunks = ceil_div(C, kCC);
    if (nb == 0) {  // untouched tile: zero fill
      if (inside) {
        if (nhwc_vec && (C & 3) == 0) {
          for (int c = 0; c < C; c += 4) *reinterpret_cast<float4*>(gpix + c) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          for (int c = 0; c < C; ++c) gpix[(int64_t)c * lv.sC] = 0.f;
        }
      }
      return;
    }
    const int ngroups = ceil_div(nb, kG);
    const int total = nchunks * ngroups;
    float acc[kCC];
    stage_group(p, S, 0, 0, nb, 0, C, vec_ok);
    int chunk = 0, grp = 0;
    for (int st = 0; st < total; ++st) {
      if (st + 1 < total) {  // prefetch the next (chunk, group) while this one is consumed
        int nc = chunk, ng = grp + 1;
        if (ng == ngroups) { ng = 0; ++nc; }
        stage_group(p, S, (st + 1) & 1, ng * kG, nb, nc * kCC, C, vec_ok);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      if (grp == 0) {
#pragma unroll
        for (int c = 0; c < kCC; ++c) acc[c] = 0.f;
      }
      accumulate_group(S, st & 1, grp * kG, min(kG, nb - grp * kG), ty, tx, acc);
      if (grp == ngroups - 1 && inside) {
        const int c0 = chunk * kCC;
        if (nhwc_vec && c0 + kCC <= C) {   // channels_last gradient: 4 x 16-byte stores per pixel and pass
#pragma unroll
          for (int c = 0; c < kCC; c += 4)
            *reinterpret_cast<float4*>(gpix + c0 + c) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
        } else {
#pragma unroll
          for (int c = 0; c < kCC; ++c)
            if (c0 + c < C) gpix[(int64_t)(c0 + c) * lv.sC] = acc[c];
        }
      }
      __syncthreads();
      if (++grp == ngroups) { grp = 0; ++chunk; }
    }
    return;
  }

  // dense tile (more RoIs than the resident tables hold, or > kWin RoIs in the image): batches are re-collected for
  // every channel pass; staging is not overlapped.  Correct for any density, fast enough for the rare case.
  for (int c0 = 0; c0 < C; c0 += kCC) {
    float acc[kCC];
#pragma unroll
    for (int c = 0; c < kCC; ++c) acc[c] = 0.f;
    int pos = r0;
    bool first = true;
    while (true) {
      if (!(first && c0 == 0)) {
        collect_batch(p, S, level, tx0, ty0, pos, r1);
        if (S.nb > 0) build_tables(p, S, lv, tx0, ty0);
      }
      first = false;
      const int nb = S.nb;
      const int next = S.next_pos;
      for (int j0 = 0; j0 < nb; j0 += kG) {
        stage_group(p, S, 0, j0, nb, c0, C, vec_ok);
        cp_async_wait<0>();
        __syncthreads();
        accumulate_group(S, 0, j0, min(kG, nb - j0), ty, tx, acc);
        __syncthreads();
      }
      if (next >= r1) break;
      pos = next;
      __syncthreads();
    }
    if (inside) {
#pragma unroll
      for (int c = 0; c < kCC; ++c)
        if (c0 + c < C) gpix[(int64_t)(c0 + c) * lv.sC] = acc[c];
    }
    __syncthreads();
  }
}

int fill_bwd(BwdParams& p, const osr_feat_level_t* h_levels, int num_levels, int num_images, int C, int P,
             int sampling_ratio, int aligned, int canonical_box_size, int canonical_level, int min_level) {
  int rc = osr::fill_roi_levels(p.L, h_levels, num_levels, num_images, C, P, sampling_ratio, aligned,
                                canonical_box_size, canonical_level, min_level);
  if (rc) return rc;
  int base = 0;
  for (int l = 0; l < num_levels; ++l) {
    p.tiles_x[l] = osr::ceil_div(p.L.lv[l].W, kTW);
    p.tiles_y[l] = osr::ceil_div(p.L.lv[l].H, kTH);
    p.tile_base[l] = base;
    base += p.tiles_x[l] * p.tiles_y[l] * num_images;
  }
  p.tile_base[num_levels] = base;
  return 0;
}

}  // namespace

extern "C" {

size_t osr_roi_align_bwd_workspace(const osr_feat_level_t* h_levels, int num_levels, int num_images, int C, int M) {
  (void)h_levels; (void)num_levels; (void)num_images; (void)C;
  return osr::align256((size_t)(M > 0 ? M : 1) * sizeof(RoiInfo));
}

int osr_roi_align_bwd(const osr_feat_level_t* h_grad_levels, int num_levels, int num_images, int C,
                      const float* grad_out, const float* rois, const int32_t* roi_batch_offsets, int M, int P,
                      int sampling_ratio, int aligned, int canonical_box_size, int canonical_level, int min_level,
                      void* workspace, size_t workspace_bytes, void* stream) {
  BwdParams p;
  int rc = fill_bwd(p, h_grad_levels, num_levels, num_images, C, P, sampling_ratio, aligned, canonical_box_size,
                    canonical_level, min_level);
  if (rc) return rc;
  if (M < 0) return osr::fail_arg(OSR_E_ARG, "roi_align_bwd: M < 0");
  if (num_images == 0) return 0;
  if (!roi_batch_offsets || !workspace || (M > 0 && (!grad_out || !rois)))
    return osr::fail_arg(OSR_E_ARG, "roi_align_bwd: null pointer argument");
  if (workspace_bytes < osr_roi_align_bwd_workspace(h_grad_levels, num_levels, num_images, C, M))
    return osr::fail_arg(OSR_E_WORKSPACE, "roi_align_bwd: workspace too small");
  p.grad_out = grad_out;
  p.rois = rois;
  p.roi_off = roi_batch_offsets;
  p.M = M;
  p.info = static_cast<RoiInfo*>(workspace);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (M > 0) {
    roi_bwd_prep_kernel<<<osr::ceil_div(M, 256), 256, 0, s>>>(p);
    OSR_LAUNCH_CHECK();
  }
  const int tiles = p.tile_base[num_levels];
  const size_t smem = sizeof(BwdSmem);
  OSR_CUDA_CHECK(cudaFuncSetAttribute(roi_align_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  roi_align_bwd_kernel<<<tiles, kThreads, smem, s>>>(p);
  OSR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
