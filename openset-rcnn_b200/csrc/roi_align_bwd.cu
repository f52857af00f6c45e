// Multi-level ROIAlignV2 backward for sm_100a: deterministic, atomic-free GATHER formulation.
// Replaces autograd's torchvision roi_align_backward (fp32 atomicAdd scatter, non-deterministic, plus a
// separate zero-fill of every dense gradient map) at train.py:145.
//
//   prep kernel    one WARP per RoI: FPN level, exact pixel box of the footprint and (per axis, extents <= 64 px) the
//                  dense weight rows Wy[row][bin] / count, Wx[col][bin] written once to the workspace.
//   gather kernels one CTA per (image, level, 16x16-pixel tile).  The CTA scans the RoIs of its image in index order,
//                  keeps those whose box meets the tile (ordered => fixed accumulation order), loads / derives their
//                  separable weight tables restricted to the tile (same sample arithmetic as the forward => exact
//                  adjoint) and accumulates
//                      g[c][y][x] += Wy[ph][y] * Wx[pw][x] / count * grad_out[roi][c][ph][pw]
//                  writing every pixel ONCE with a plain store - which is also the zero fill of untouched pixels.
//                  No atomics, no memset, run-to-run bit-identical.
//                  * roi_align_bwd_cl_kernel (channels_last gradients, C % 32 == 0): thread = channel, accumulators in
//                    shared memory, separable in both directions - see the comment above that kernel;
//                  * roi_align_bwd_kernel (any strides): thread = pixel, 32 channels at a time in registers.
#include <cstdlib>

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
constexpr int kFW = 64;                 // footprint rows / columns per RoI kept in the workspace weight table
constexpr int kWRoi = 2 * kFW * kP;     // floats per RoI in that table: [y rows | x columns] x 7 bins

struct RoiInfo {
  int x0, x1, y0, y1;  // inclusive pixel bounds of the footprint (exact); x1 < x0 => empty
  int level;
  int flags;           // bit 0 / bit 1: the RoI's y / x weight rows are in the workspace table (extent <= kFW)
};
struct __align__(16) RoiInfoPacked {   // 16 bytes: the per-tile scans read it with one coalesced 16-byte load per RoI
  short x0, x1, y0, y1;
  int level, pad;
};
__device__ __forceinline__ RoiInfo unpack_info(const int4 v) {
  RoiInfo r;
  r.x0 = (short)(v.x & 0xffff); r.x1 = v.x >> 16;
  r.y0 = (short)(v.y & 0xffff); r.y1 = v.y >> 16;
  r.level = v.z;
  r.flags = v.w;
  return r;
}
__device__ __forceinline__ RoiInfo load_info(const RoiInfoPacked* q) { return unpack_info(__ldg(reinterpret_cast<const int4*>(q))); }

struct BwdParams {
  RoiLevels L;
  const float* grad_out;
  const float* rois;
  const int32_t* roi_off;  // (N+1)
  int M;
  RoiInfoPacked* info;     // (M) workspace
  float* wfull;            // (M, 2, kFW, 7) workspace: per RoI, per footprint row / column, the 7 bin weights
  int tile_base[OSR_MAX_LEVELS + 1];  // first tile id of each level (tiles ordered level, image, ty, tx)
  int tiles_x[OSR_MAX_LEVELS], tiles_y[OSR_MAX_LEVELS];
  int cl_tile_base[OSR_MAX_LEVELS + 1];  // same for the channels_last kernel's 16x16 tiles
  int cl_tiles_x[OSR_MAX_LEVELS], cl_tiles_y[OSR_MAX_LEVELS];
  // register-accumulator kernel: a tile of level l is worked by ps[l] * pt[l] CTAs - ps slab groups x pt sub-tile groups
  // (coarse levels hold few, heavily covered tiles: one CTA per tile made them the critical path of the launch)
  int clr_cta_base[OSR_MAX_LEVELS + 1];
  int clr_ps[OSR_MAX_LEVELS], clr_pt[OSR_MAX_LEVELS];
  int clr_dyn_min;   // > 0: a tile is only worked by several CTAs when its first batch holds at least this many RoIs
};

// one bin of one axis: visit its valid samples as (low row, high row, weight at low, weight at high) - the sample
// arithmetic of torchvision's bilinear_interpolate, shared by the forward tables (=> exact adjoint)
template <class F>
__device__ __forceinline__ void for_bin_samples(float start, float bin, int grid, int L, int p, F&& f) {
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
    f(lo, hi, 1.f - l, l);
  }
}

// Prep: one WARP per RoI.  Lanes 0..6 own the y bins, lanes 7..13 the x bins.  Pass 1 finds the exact footprint box,
// pass 2 (footprints up to kFW x kFW) accumulates the dense weight rows Wy[row][bin] / count and Wx[col][bin] in
// shared memory and writes them to the workspace, so the gather CTAs (one RoI is seen by ~7 tiles x 2 slabs) only
// load and pack them instead of re-deriving them from the samples.
constexpr int kPrepWarps = 4;
__global__ void __launch_bounds__(kPrepWarps * 32) roi_bwd_prep_kernel(const __grid_constant__ BwdParams p) {
  __shared__ float sw[kPrepWarps][kWRoi];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int m = blockIdx.x * kPrepWarps + warp;
  if (m >= p.M) return;   // warp-uniform; no block-level barrier below
  const float* roi = p.rois + (int64_t)m * 5;
  const float x1 = __ldg(roi + 1), y1 = __ldg(roi + 2), x2 = __ldg(roi + 3), y2 = __ldg(roi + 4);
  const int level = assign_level(x1, y1, x2, y2, p.L);
  const bool lvl_ok = (level >= 0 && level < p.L.num_levels);
  float* w = sw[warp];
  for (int i = lane; i < kWRoi; i += 32) w[i] = 0.f;
  RoiGeom g{};
  int L = 1;
  const bool isy = lane < kP;
  const int bin = isy ? lane : lane - kP;
  int lo = 1 << 30, hi = -1;
  if (lvl_ok) {
    const LevelDesc& lv = p.L.lv[level];
    g = roi_geometry(x1, y1, x2, y2, lv.scale, p.L.sampling_ratio);
    L = isy ? lv.H : lv.W;
    if (lane < 2 * kP)
      for_bin_samples(isy ? g.start_h : g.start_w, isy ? g.bin_h : g.bin_w, isy ? g.grid_h : g.grid_w, L, bin,
                      [&](int l, int h, float, float) { lo = min(lo, l); hi = max(hi, h); });
  }
  int ylo = isy ? lo : (1 << 30), yhi = isy ? hi : -1;
  int xlo = (!isy && lane < 2 * kP) ? lo : (1 << 30), xhi = (!isy && lane < 2 * kP) ? hi : -1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ylo = min(ylo, __shfl_xor_sync(0xffffffffu, ylo, o));
    yhi = max(yhi, __shfl_xor_sync(0xffffffffu, yhi, o));
    xlo = min(xlo, __shfl_xor_sync(0xffffffffu, xlo, o));
    xhi = max(xhi, __shfl_xor_sync(0xffffffffu, xhi, o));
  }
  const bool nonempty = (yhi >= 0) && (xhi >= 0);
  const bool prey = nonempty && (yhi - ylo < kFW), prex = nonempty && (xhi - xlo < kFW);
  __syncwarp();
  if (lane < 2 * kP && (isy ? prey : prex)) {
    float* wa = w + (isy ? 0 : kFW * kP);
    const int base = isy ? ylo : xlo;
    for_bin_samples(isy ? g.start_h : g.start_w, isy ? g.bin_h : g.bin_w, isy ? g.grid_h : g.grid_w, L, bin,
                    [&](int l, int h, float wl, float wh) {
                      wa[(l - base) * kP + bin] += wl;
                      wa[(h - base) * kP + bin] += wh;
                    });
  }
  __syncwarp();
  float* dst = p.wfull + (int64_t)m * kWRoi;
  if (prey) {
    const float ic = 1.0f / g.count;
    const int ny = (yhi - ylo + 1) * kP;
    for (int i = lane; i < ny; i += 32) dst[i] = w[i] * ic;
  }
  if (prex) {
    const int nx = (xhi - xlo + 1) * kP;
    for (int i = lane; i < nx; i += 32) dst[kFW * kP + i] = w[kFW * kP + i];
  }
  const int pre = (prey ? 1 : 0) | (prex ? 2 : 0);
  if (lane == 0) {
    RoiInfoPacked q;
    q.x0 = (short)(nonempty ? xlo : 1); q.x1 = (short)(nonempty ? xhi : 0);
    q.y0 = (short)(nonempty ? ylo : 1); q.y1 = (short)(nonempty ? yhi : 0);
    q.level = level; q.pad = pre;
    p.info[m] = q;
  }
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
      const RoiInfo r = load_info(p.info + m);
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

// =========================================================================================================
// channels_last gather kernel: thread = CHANNEL (the mirror image of roi_align_fwd_nhwc_kernel).
// One CTA per (image, level, 16x16-pixel tile), worked through 128 channels at a time (same RoI scan and tables for
// every slab) as four 4x16 sub-tiles; warp w owns channels [32w, 32w+32) of the slab and a private [pixel][lane] fp32
// accumulator sub-tile in shared memory (bank = lane: conflict-free).  Per (sub-tile, RoI) pair, in RoI index order
// (=> fixed summation order), the warp
//   1. copies its 32 x 49 block of grad_out (6272 contiguous bytes) into a private staging buffer with cp.async
//      (lane = channel reads it at stride 49: conflict-free);
//   2. folds the y weights of the 4 rows:  rg[r][pw] = sum_ph Wy[ph][y_r]/count * g[ph][pw]; the rows share one window
//      of 4 consecutive bins, so 4 rows of g are read once and each row is a branch-free 4-term combination (dense
//      7-bin fold when bins are narrower than a pixel); the staging buffer is then refilled with the next pair's block
//      while the columns are walked;
//   3. walks the tile columns the RoI touches:  acc[y][x] += sum_pw Wx[pw][x] * rg[y][pw]  (<= 3 bins per column; the
//      first bin never decreases with x, so the columns split into <= 5 runs of equal first bin, each a branch-free
//      loop with compile-time register indices; columns in > 3 bins take a dense 7-bin loop).
// This uses the separability in BOTH directions (the pixel-per-thread kernel above cannot: it pays one FMA per
// (pixel, bin pair, channel)), every lane is busy, there is no CTA-wide barrier while a batch of RoIs is processed,
// and the finished tile leaves with 16-byte coalesced stores (which is also the zero fill).  Arithmetic: same
// sample -> weight code as the forward (tile_bin_weights), fp32 FMAs, deterministic.
constexpr int kCT = 4, kCW = 16;            // sub-tile rows x cols: the unit whose accumulators live in shared memory
constexpr int kCS = 4;                      // sub-tiles (stacked vertically) per CTA: they share one RoI scan + tables
constexpr int kCH = kCT * kCS;              // CTA tile rows
constexpr int kCWarps = 4;                  // warps per CTA = 32-channel groups per slab
constexpr int kCThreads = kCWarps * 32;
constexpr int kCNB = 32;                    // RoIs whose tables are resident at once
constexpr int kGBlk = 32 * kP * kP;         // floats of one (RoI, 32-channel) block of grad_out = 1568

struct __align__(16) ClSmem {
  float acc[kCWarps][kCT * kCW * 32];       // [warp][pixel][lane]
  float sg[kCWarps][kGBlk];                 // per-warp staging of grad_out[(roi, 32 channels), 7, 7]
  float4 yrow[kCNB][kCH];                   // y weights of the row at bins pm .. pm+3, pm = first bin of its SUB-TILE (below)
  float4 xcol[kCNB][kCW];                   // (w(p0), w(p0+1), w(p0+2), code): code 0..4 = p0, 5 = dense, 6 = empty
  uchar2 lohi[kCNB][kCH];                   // scratch: (first, last) bin with weight of each tile row; first > last = none
  unsigned char pm[kCNB][kCS];              // first bin of the 4-bin window shared by the sub-tile's rows; 254 = no weight,
                                            // 255 = the rows span more than 4 bins (dense 7-bin fold)
  __align__(16) uchar2 run[kCNB][8];        // [k] = (first, last + 1) tile column whose code is k (k = 0..5)
  BatchEntry e[kCNB];
  int2 org[kCNB];                           // (y0, x0) of the RoI's footprint: origin of its rows in the workspace table
  unsigned stmask[kCS];                     // bit j: RoI j of the batch puts weight on sub-tile st
  int warp_cnt[kCWarps + 1];
  int nb, next_pos;
  static constexpr bool kDenseCols = false;
  static constexpr bool kHasInfo = false;
  static constexpr int kNB = kCNB;
  __device__ __forceinline__ float* scratch() { return &sg[0][0]; }
  __device__ __forceinline__ const float* scratch() const { return &sg[0][0]; }
};

static_assert((kCNB * kCH + kCThreads) * 8 <= kCWarps * kGBlk, "table scratch must fit in the staging buffers");


// Scan the image's RoI indices from `pos` in order and keep the first kCNB that hit the tile (level + footprint box):
// S.nb = how many, S.next_pos = where the next batch resumes (r1 when the image is exhausted).  The scan advances in
// windows of 4 x kCThreads indices and keeps going until the batch is full - an image with more RoIs than one window
// (1024 per image at cfg 5) must not split the few RoIs of a tile over several batches.
template <class SM>
__device__ __forceinline__ void cl_collect(const BwdParams& p, SM& S, int level, int tx0, int ty0, int pos, int r1) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int filled = 0, cur = pos;
  constexpr int kNBm = SM::kNB;   // RoIs per batch of this kernel's table set
  for (; cur < r1 && filled < kNBm; cur += 4 * kCThreads) {
    int hit[4], cnt = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int m = cur + tid * 4 + k;
      hit[k] = 0;
      if (m < r1) {
        const RoiInfo r = load_info(p.info + m);
        hit[k] = (r.level == level) && (r.x0 <= tx0 + kCW - 1) && (r.x1 >= tx0) && (r.y0 <= ty0 + kCH - 1) && (r.y1 >= ty0);
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
      for (int w = 0; w < kCWarps; ++w) {
        int c = S.warp_cnt[w];
        S.warp_cnt[w] = run;
        run += c;
      }
      S.warp_cnt[kCWarps] = run;
    }
    __syncthreads();
    int rank = filled + S.warp_cnt[warp] + inc - cnt;
    filled += S.warp_cnt[kCWarps];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (hit[k]) {
        const int m = cur + tid * 4 + k;
        if (rank < kNBm) {
          S.e[rank].m = m;
          if constexpr (SM::kHasInfo) S.einfo[rank] = __ldg(reinterpret_cast<const int4*>(p.info + m));   // (L1 hit) spares the table pass a dependent global load
        } else if (rank == kNBm) S.next_pos = m;   // the first hit that did not fit: the next batch resumes here
        ++rank;
      }
    }
    __syncthreads();   // warp_cnt is rewritten by the next window
  }
  if (tid == 0) {
    S.nb = min(filled, kNBm);
    if (filled <= kNBm) S.next_pos = min(cur, r1);   // no hit was left over: resume at the first index not scanned yet
  }
  __syncthreads();
}

// (w(p0), w(p0+1), w(p0+2), code) of one tile row / column from its dense 7-bin weight vector
__device__ __forceinline__ float4 cl_pack3(const float* w) {
  int lo = kP, hi = -1;
#pragma unroll
  for (int b = 0; b < kP; ++b)
    if (w[b] != 0.f) {
      lo = min(lo, b);
      hi = b;
    }
  if (hi < 0) return make_float4(0.f, 0.f, 0.f, __int_as_float(6));
  if (hi - lo + 1 > 3) return make_float4(0.f, 0.f, 0.f, __int_as_float(5));
  const int p0 = min(lo, kP - 3);
  return make_float4(w[p0], w[p0 + 1], w[p0 + 2], __int_as_float(p0));
}

template <class SM>
__device__ __forceinline__ void cl_build_tables(const BwdParams& p, SM& S, const LevelDesc& lv, int tx0, int ty0) {
  const int tid = threadIdx.x;
  const int nb = S.nb;
  // one thread per (RoI, tile row | tile column): fetch its 7 bin weights (precomputed per RoI by the prep kernel;
  // extents wider than kFW re-derive them from the samples) and pack them.  Rows / columns in more than 3 bins (code 5;
  // only possible for extents of a few pixels, which are always in the table) re-read the table when they are used.
  for (int q = tid; q < nb * (kCH + kCW); q += kCThreads) {
    const int j = q / (kCH + kCW);
    const int r = q - j * (kCH + kCW);
    const bool isy = r < kCH;
    const int m = S.e[j].m;
    RoiInfo info;
    if constexpr (SM::kHasInfo) info = unpack_info(S.einfo[j]);
    else info = load_info(p.info + m);
    const int pos = isy ? ty0 + r : tx0 + (r - kCH);
    const int base = isy ? info.y0 : info.x0, last = isy ? info.y1 : info.x1;
    // scratch in the idle staging buffers: y rows keep their 7 weights until the packing pass below ((j, row) slot),
    // columns only need them inside this iteration (per-thread slot behind the row slots)
    float* wd = S.scratch() + (isy ? (j * kCH + r) * 8 : SM::kNB * kCH * 8 + tid * 8);
    if (r == 0) S.org[j] = make_int2(info.y0, info.x0);
#pragma unroll
    for (int b = 0; b < kP; ++b) wd[b] = 0.f;
    if (pos >= base && pos <= last) {
      if (info.flags & (isy ? 1 : 2)) {
        const float* src = p.wfull + (int64_t)m * kWRoi + (isy ? 0 : kFW * kP) + (pos - base) * kP;
#pragma unroll
        for (int b = 0; b < kP; ++b) wd[b] = __ldg(src + b);
      } else {
        const float* roi = p.rois + (int64_t)m * 5;
        const RoiGeom g = roi_geometry(__ldg(roi + 1), __ldg(roi + 2), __ldg(roi + 3), __ldg(roi + 4), lv.scale, p.L.sampling_ratio);
        const float sc = isy ? 1.0f / g.count : 1.0f;
#pragma unroll 1
        for (int b = 0; b < kP; ++b) {
          float a = 0.f;
          for_bin_samples(isy ? g.start_h : g.start_w, isy ? g.bin_h : g.bin_w, isy ? g.grid_h : g.grid_w,
                          isy ? lv.H : lv.W, b, [&](int l, int h, float wl, float wh) {
                            if (l == pos) a += wl;
                            if (h == pos) a += wh;
                          });
          wd[b] = a * sc;
        }
      }
    }
    if (isy) {
      int lo = kP, hi = 0;
#pragma unroll
      for (int b = 0; b < kP; ++b)
        if (wd[b] != 0.f) {
          lo = min(lo, b);
          hi = b;
        }
      S.lohi[j][r] = make_uchar2((unsigned char)lo, (unsigned char)hi);
    } else {
      if constexpr (SM::kDenseCols) {   // register-accumulator kernel: all 7 bin weights of the column stay resident
        S.xd0[j][r - kCH] = make_float4(wd[0], wd[1], wd[2], wd[3]);
        S.xd1[j][r - kCH] = make_float4(wd[4], wd[5], wd[6], 0.f);
        unsigned bits = 0;
#pragma unroll
        for (int b = 0; b < kP; ++b) bits |= (wd[b] != 0.f) ? (1u << b) : 0u;
        S.xne[j][r - kCH] = (unsigned char)bits;
      } else {
        S.xcol[j][r - kCH] = cl_pack3(wd);
      }
    }
  }
  __syncthreads();
  // y rows: the rows of a sub-tile share ONE window of 4 consecutive bins (pm .. pm+3) whenever their bins fit in it
  // (always, unless bins are narrower than a pixel), so the fold reads 4 rows of g once for all of them.
  for (int q = tid; q < nb * kCH; q += kCThreads) {
    const int j = q / kCH;
    const int r = q - j * kCH;
    const int st = r / kCT;
    int lo = kP, hi = -1;
#pragma unroll
    for (int i = 0; i < kCT; ++i) {
      const uchar2 lh = S.lohi[j][st * kCT + i];
      if (lh.x <= lh.y) {
        lo = min(lo, (int)lh.x);
        hi = max(hi, (int)lh.y);
      }
    }
    const int pmv = min(lo, kP - 4);
    const bool fits = (hi < 0) || (hi <= pmv + 3);
    const float* wd = S.scratch() + (j * kCH + r) * 8;
    const float4 yw = (hi >= 0 && fits) ? make_float4(wd[pmv], wd[pmv + 1], wd[pmv + 2], wd[pmv + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (SM::kDenseCols) {   // row pairs, window weights interleaved for the packed FMAs
      float* d = reinterpret_cast<float*>(&S.yrow2[j][r >> 1][0]) + (r & 1);
      d[0] = yw.x; d[2] = yw.y; d[4] = yw.z; d[6] = yw.w;
    } else {
      S.yrow[j][r] = yw;
    }
    if (r == st * kCT) S.pm[j][st] = (unsigned char)(hi < 0 ? 254 : (fits ? pmv : 255));
  }
  __syncthreads();
  if constexpr (SM::kDenseCols) {
    // [first, last + 1) tile column the RoI puts weight on
    for (int j = tid; j < nb; j += kCThreads) {
      int lo = kCW, hi = 0;
#pragma unroll
      for (int x = 0; x < kCW; ++x)
        if (S.xne[j][x]) {
          lo = min(lo, x);
          hi = x + 1;
        }
      S.xr[j] = make_uchar2((unsigned char)lo, (unsigned char)hi);
#pragma unroll
      for (int g4 = 0; g4 < kCW / 4; ++g4)
        S.xgm[j][g4] = S.xne[j][4 * g4] | S.xne[j][4 * g4 + 1] | S.xne[j][4 * g4 + 2] | S.xne[j][4 * g4 + 3];
    }
    __syncthreads();
    if (tid >= kCThreads - kCS) {   // which RoIs put weight on each sub-tile
      const int st = tid - (kCThreads - kCS);
      unsigned m = 0;
      for (int j = 0; j < nb; ++j)
        if (S.pm[j][st] != 254 && S.xr[j].x < S.xr[j].y) m |= 1u << j;
      S.stmask[st] = m;
    }
  } else {
    // column runs: the first bin of a column never decreases with x, so columns of equal code are contiguous
    for (int q = tid; q < nb * 6; q += kCThreads) {
      const int j = q / 6, k = q - j * 6;
      int lo = kCW, hi = 0;
#pragma unroll
      for (int x = 0; x < kCW; ++x)
        if (__float_as_int(S.xcol[j][x].w) == k) {
          lo = min(lo, x);
          hi = x + 1;
        }
      S.run[j][k] = make_uchar2((unsigned char)lo, (unsigned char)hi);
    }
    if (tid >= kCThreads - kCS) {   // which RoIs put weight on each sub-tile
      const int st = tid - (kCThreads - kCS);
      unsigned m = 0;
      for (int j = 0; j < nb; ++j) {
        const bool anyr = S.pm[j][st] != 254;
        bool anyc = false;
#pragma unroll
        for (int x = 0; x < kCW; ++x) anyc |= (__float_as_int(S.xcol[j][x].w) != 6);
        if (anyr && anyc) m |= 1u << j;
      }
      S.stmask[st] = m;
    }
  }
  __syncthreads();
}

// rg[r][pw] = sum_ph Wy[ph][row r] / count * g[ph][pw] for the kCT rows of a sub-tile; gl = this lane's channel in the
// staged block (49 floats, stride-49 across lanes: conflict-free).  The rows share one window of 4 bins, so 4 rows of
// g are read once (28 LDS, dynamic base) and every row is a branch-free 4-term combination; sub-tiles whose rows span
// more bins (bins narrower than a pixel) loop over all 7 bins with the dense weights from the workspace table.
template <class SM>
__device__ __forceinline__ void cl_fold_subtile(const float* gl, const SM& S, int j, int st, const float* wdense,
                                                float (&rg)[kCT][kP]) {
  const int pmv = S.pm[j][st];
  if (pmv < 254) {
    const float* gp = gl + pmv * kP;
    float g4[4][kP];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int b = 0; b < kP; ++b) g4[k][b] = gp[k * kP + b];
#pragma unroll
    for (int r = 0; r < kCT; ++r) {
      const float4 w = S.yrow[j][st * kCT + r];
#pragma unroll
      for (int b = 0; b < kP; ++b) rg[r][b] = fmaf(w.w, g4[3][b], fmaf(w.z, g4[2][b], fmaf(w.y, g4[1][b], w.x * g4[0][b])));
    }
  } else {
#pragma unroll
    for (int r = 0; r < kCT; ++r)
#pragma unroll
      for (int b = 0; b < kP; ++b) rg[r][b] = 0.f;
#pragma unroll 1
    for (int a = 0; a < kP; ++a) {
      float ga[kP];
#pragma unroll
      for (int b = 0; b < kP; ++b) ga[b] = gl[a * kP + b];
#pragma unroll
      for (int r = 0; r < kCT; ++r) {
        const uchar2 lh = S.lohi[j][st * kCT + r];   // rows outside the footprint have no table entry
        const float wa = (lh.x <= lh.y) ? __ldg(wdense + r * kP + a) : 0.f;
#pragma unroll
        for (int b = 0; b < kP; ++b) rg[r][b] = fmaf(wa, ga[b], rg[r][b]);
      }
    }
  }
}

// columns [x0, x1) of the tile all start at bin K: acc[r][x] += w0 rg[r][K] + w1 rg[r][K+1] + w2 rg[r][K+2]
template <int K>
__device__ __forceinline__ void cl_run(const float4* xcol, uchar2 rn, float* accl, const float (&rg)[kCT][kP]) {
#pragma unroll 1
  for (int x = rn.x; x < rn.y; ++x) {
    const float4 w = xcol[x];
    float* ap = accl + x * 32;
    float v[kCT];
#pragma unroll
    for (int r = 0; r < kCT; ++r) v[r] = ap[r * kCW * 32];
#pragma unroll
    for (int r = 0; r < kCT; ++r) v[r] = fmaf(w.z, rg[r][K + 2], fmaf(w.y, rg[r][K + 1], fmaf(w.x, rg[r][K], v[r])));
#pragma unroll
    for (int r = 0; r < kCT; ++r) ap[r * kCW * 32] = v[r];
  }
}

// issue the async copy of the warp's 32 x 49 block of grad_out for RoI m
__device__ __forceinline__ void cl_stage(const BwdParams& p, float* sg, int m, int c0w, int C, int lane) {
  const float* src = p.grad_out + ((int64_t)m * C + c0w) * (kP * kP);
#pragma unroll
  for (int i = 0; i < (kGBlk / 4 + 31) / 32; ++i) {
    const int v = i * 32 + lane;
    if (v < kGBlk / 4) cp_async16(sg + v * 4, src + v * 4);
  }
  cp_async_commit();
}

// next (sub-tile, RoI) pair after sub-tile st with remaining mask m; returns the RoI slot or -1
template <class SM>
__device__ __forceinline__ int cl_next_pair(const SM& S, int st, unsigned m, int st_end = kCS) {
  while (m == 0) {
    if (++st >= st_end) return -1;
    m = S.stmask[st];
  }
  return __ffs(m) - 1;
}

__global__ void __launch_bounds__(kCThreads, 3) roi_align_bwd_cl_kernel(const __grid_constant__ BwdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ClSmem& S = *reinterpret_cast<ClSmem*>(smem_raw);

  // coarsest level first: its few tiles see the most (and largest) RoIs, so the longest CTAs start early
  int t = (int)(gridDim.x - 1 - blockIdx.x), level = 0;
  while (level + 1 < p.L.num_levels && t >= p.cl_tile_base[level + 1]) ++level;
  t -= p.cl_tile_base[level];
  const int per_img = p.cl_tiles_x[level] * p.cl_tiles_y[level];
  const int n = t / per_img;
  t -= n * per_img;
  const int ty0 = (t / p.cl_tiles_x[level]) * kCH, tx0 = (t % p.cl_tiles_x[level]) * kCW;
  const LevelDesc& lv = p.L.lv[level];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.L.C;
  const int nslab = ceil_div(C, kCWarps * 32);   // 128-channel slabs, walked one after the other with the same tables
  const int r0 = p.roi_off[n], r1 = p.roi_off[n + 1];
  float* acc = S.acc[warp];
  float* accl = acc + lane;
  float* sg = S.sg[warp];
  const bool vec = ((lv.sW & 3) == 0) && ((lv.sH & 3) == 0) && ((lv.sN & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(lv.data) & 15) == 0);
  // write-out slot: lane -> (pixel column xs + 4 q, channels quad .. quad + 4): four 128-byte runs per 16-byte store
  const int quad = (lane & 7) * 4, xs = lane >> 3;

  int pos = r0;
  bool first = true;   // first batch of RoIs: the write-out stores; later batches (dense tiles) add to what is there
  while (true) {
    cl_collect(p, S, level, tx0, ty0, pos, r1);
    const int nb = S.nb, next = S.next_pos;
    if (nb > 0) {
      cl_build_tables(p, S, lv, tx0, ty0);
    } else {
      if (tid < kCS) S.stmask[tid] = 0;
      __syncthreads();
    }
    for (int slab = 0; slab < nslab; ++slab) {
      const int c0w = slab * (kCWarps * 32) + warp * 32;   // first channel of this warp; C % 32 == 0 (host-checked)
      if (c0w >= C) break;
      float* gimg = lv.data + (int64_t)n * lv.sN + c0w;
      int nj = cl_next_pair(S, -1, 0);
      if (nj >= 0) cl_stage(p, sg, S.e[nj].m, c0w, C, lane);
      for (int st = 0; st < kCS; ++st) {
        unsigned m = S.stmask[st];
        const bool any = (m != 0);
        if (any) {
          float4* a4 = reinterpret_cast<float4*>(acc);
#pragma unroll 4
          for (int i = lane; i < kCT * kCW * 8; i += 32) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          __syncwarp();
        }
        while (m != 0) {
          const int j = __ffs(m) - 1;
          m &= m - 1;
          cp_async_wait<0>();
          __syncwarp();
          float rg[kCT][kP];
          const int2 org = S.org[j];
          const float* wtab = p.wfull + (int64_t)S.e[j].m * kWRoi;   // dense rows, only dereferenced for code-5 rows / columns
          cl_fold_subtile(sg + lane * (kP * kP), S, j, st, wtab + (ty0 + st * kCT - org.x) * kP, rg);
          __syncwarp();   // every lane is done with the staged block: refill it while the columns are walked
          nj = cl_next_pair(S, st, m);
          if (nj >= 0) cl_stage(p, sg, S.e[nj].m, c0w, C, lane);
          const float4* xcol = S.xcol[j];
          const uint4 rb = *reinterpret_cast<const uint4*>(S.run[j]);   // all run bounds of this RoI in one load
          cl_run<0>(xcol, make_uchar2(rb.x & 0xff, (rb.x >> 8) & 0xff), accl, rg);
          cl_run<1>(xcol, make_uchar2((rb.x >> 16) & 0xff, rb.x >> 24), accl, rg);
          cl_run<2>(xcol, make_uchar2(rb.y & 0xff, (rb.y >> 8) & 0xff), accl, rg);
          cl_run<3>(xcol, make_uchar2((rb.y >> 16) & 0xff, rb.y >> 24), accl, rg);
          cl_run<4>(xcol, make_uchar2(rb.z & 0xff, (rb.z >> 8) & 0xff), accl, rg);
          const uchar2 dn = make_uchar2((rb.z >> 16) & 0xff, rb.z >> 24);   // columns in more than 3 bins (bins narrower than half a pixel)
          for (int x = dn.x; x < dn.y; ++x) {
            if (__float_as_int(xcol[x].w) != 5) continue;
            const float* wd = wtab + kFW * kP + (tx0 + x - org.y) * kP;
            float wv[kP];
#pragma unroll
            for (int b = 0; b < kP; ++b) wv[b] = __ldg(wd + b);
#pragma unroll
            for (int r = 0; r < kCT; ++r) {
              float v = accl[(r * kCW + x) * 32];
#pragma unroll
              for (int b = 0; b < kP; ++b) v = fmaf(wv[b], rg[r][b], v);
              accl[(r * kCW + x) * 32] = v;
            }
          }
        }
        // write-out of sub-tile st
        if (!first && !any) continue;
        __syncwarp();
        const int ys = ty0 + st * kCT;
        if (vec) {
          const int ny = min(kCT, lv.H - ys);
          float* gp = gimg + (int64_t)ys * lv.sH + (int64_t)(tx0 + xs) * lv.sW + quad;
          const float* ap = acc + xs * 32 + quad;
          const int64_t step = 4 * lv.sW;
          for (int yy = 0; yy < ny; ++yy) {
#pragma unroll
            for (int q = 0; q < kCW / 4; ++q) {
              if (tx0 + xs + 4 * q < lv.W) {
                float4 v = any ? *reinterpret_cast<const float4*>(ap + (yy * kCW + 4 * q) * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4* dst = reinterpret_cast<float4*>(gp + q * step);
                if (!first) {
                  const float4 o = *dst;
                  v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                }
                *dst = v;
              }
            }
            gp += lv.sH;
          }
        } else {
          for (int pix = 0; pix < kCT * kCW; ++pix) {
            const int y = ys + pix / kCW, x = tx0 + pix % kCW;
            if (y < lv.H && x < lv.W) {
              float* dst = gimg + (int64_t)y * lv.sH + (int64_t)x * lv.sW + lane;
              const float v = any ? acc[pix * 32 + lane] : 0.f;
              *dst = first ? v : (*dst + v);
            }
          }
        }
        __syncwarp();
      }
    }
    first = false;
    pos = next;
    if (pos >= r1) break;
    __syncthreads();   // tables are rebuilt by the next batch
  }
}


// =========================================================================================================
// channels_last gather kernel, REGISTER accumulators + packed fp32x2 FMAs (round 2).  Same decomposition as
// roi_align_bwd_cl_kernel (CTA = (image, level, 16x16 tile); warp = 32 channels of a 128-channel slab; four 4x16
// sub-tiles; per (sub-tile, RoI): staged grad_out block -> y fold -> x expansion, RoIs in index order), but
//   * the 4x16 accumulator sub-tile of a lane lives in 64 REGISTERS, not in shared memory: the x expansion is unrolled
//     over the 16 tile columns so every accumulator has a compile-time index.  That removes the accumulator
//     read-modify-write (8 of the 9 shared-memory operations per 4 pixels of the shared-memory kernel, whose limiter
//     was the L1/shared pipe: 142 M wavefronts per launch) and the write-out staging;
//   * rows are held as PAIRS (float2 = rows 2q, 2q+1 of a column) and all arithmetic is Blackwell's packed
//     fma.rn.f32x2 (SASS FFMA2 with a scalar-broadcast operand): half the FMA instructions of the scalar form - the
//     kernel is issue-bound, not FP32-pipe-bound;
//   * the x expansion is dense over the 7 bins (14 FFMA2 per column, branch-free, weights in two broadcast LDS.128);
//     4-column groups the RoI does not touch are skipped.  [A 3-bin form behind a per-column switch on the first bin
//     was measured at 1.36 ms: 16 indirect branches per (sub-tile, RoI) and 5.6 k SASS instructions - instruction-
//     fetch bound; DESIGN.md section 4.]
//   * a finished sub-tile leaves straight from registers: lane = channel, one 128-byte row segment per store.
// Bit-identical summation order to nothing else - but deterministic (fixed RoI order, fixed FMA order) and the exact
// adjoint arithmetic of the forward tables.
// kNBuf staging buffers per warp.  2 = the block of the NEXT (sub-tile, RoI) pair is requested before the current pair is
// folded (a whole pair of lead instead of the x expansion only); the table set shrinks to 30 RoIs so that three CTAs
// still fit one SM's shared memory.
template <int kNBuf, bool kInfo>
struct __align__(16) RegSmemT {
  static constexpr int kNB = kNBuf == 2 ? 30 : kCNB;
  float sg[kCWarps][kNBuf][kGBlk];          // per-warp staging of grad_out[(roi, 32 channels), 7, 7]
  float4 yrow2[kNB][kCH / 2][2];            // row PAIR (2q, 2q+1): window weights interleaved (w0a,w0b,w1a,w1b | w2a,w2b,w3a,w3b)
  float4 xd0[kNB][kCW], xd1[kNB][kCW];      // all 7 bin weights of the column (w0..w3 | w4..w6, 0)
  int4 einfo[kInfo ? kNB : 1];              // the RoI's packed footprint record, kept by the scan for the table pass
  uchar2 lohi[kNB][kCH];
  unsigned char pm[kNB][kCS];
  unsigned char xne[kNB][kCW];              // column has weight: bit b = bin b
  unsigned char xgm[kNB][kCW / 4];          // OR of xne over the 4 columns of a group: bins the group's expansion must visit
  unsigned long long mbar[kCWarps][kNBuf];  // per-warp transaction barriers of the bulk copies that fill sg[warp][b]
  uchar2 xr[kNB];                           // [first, last + 1) tile column with weight
  BatchEntry e[kNB];
  int2 org[kNB];
  unsigned stmask[kCS];
  int warp_cnt[kCWarps + 1];
  int nb, next_pos;
  static constexpr bool kDenseCols = true;
  static constexpr bool kHasInfo = kInfo;
  __device__ __forceinline__ float* scratch() { return &sg[0][0][0]; }
  __device__ __forceinline__ const float* scratch() const { return &sg[0][0][0]; }
};
static_assert(sizeof(RegSmemT<2, true>) + 1024 <= 233472 / 3, "three CTAs of the double-buffered kernel must fit one SM");

// v * s + c on both halves (SASS: FFMA2 Rd, Rv.F32x2, Rs.F32, Rc.F32x2)
__device__ __forceinline__ float2 ffma2(float2 v, float s, float2 c) {
  unsigned long long rv, rs, rc, rd;
  const float2 sb = make_float2(s, s);
  rv = *reinterpret_cast<const unsigned long long*>(&v);
  rs = *reinterpret_cast<const unsigned long long*>(&sb);
  rc = *reinterpret_cast<const unsigned long long*>(&c);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(rv), "l"(rs), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fmul2(float2 v, float s) {
  unsigned long long rv, rs, rd;
  const float2 sb = make_float2(s, s);
  rv = *reinterpret_cast<const unsigned long long*>(&v);
  rs = *reinterpret_cast<const unsigned long long*>(&sb);
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(rv), "l"(rs));
  return *reinterpret_cast<float2*>(&rd);
}

// rg2[q][b] = (rg[2q][b], rg[2q+1][b]) with rg[r][b] = sum_ph Wy[ph][row r] / count * g[ph][b]
template <class SM>
__device__ __forceinline__ void clr_fold_subtile(const float* gl, const SM& S, int j, int st, const BwdParams& p, int y0,
                                                 float2 (&rg2)[kCT / 2][kP]) {
  const int pmv = S.pm[j][st];
  if (pmv < 254) {   // the 4 rows share one window of 4 bins: 28 LDS, 56 FFMA2
    const float* gp = gl + pmv * kP;
    float g4[4][kP];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int b = 0; b < kP; ++b) g4[k][b] = gp[k * kP + b];
#pragma unroll
    for (int q = 0; q < kCT / 2; ++q) {
      const float4 wa = S.yrow2[j][st * (kCT / 2) + q][0], wb = S.yrow2[j][st * (kCT / 2) + q][1];
      const float2 w0 = make_float2(wa.x, wa.y), w1 = make_float2(wa.z, wa.w), w2 = make_float2(wb.x, wb.y), w3 = make_float2(wb.z, wb.w);
#pragma unroll
      for (int b = 0; b < kP; ++b) rg2[q][b] = ffma2(w3, g4[3][b], ffma2(w2, g4[2][b], ffma2(w1, g4[1][b], fmul2(w0, g4[0][b]))));
    }
  } else {           // bins narrower than a pixel: dense 7-bin fold with the weights from the workspace table (rare)
    const float* wdense = p.wfull + (int64_t)S.e[j].m * kWRoi + (y0 - S.org[j].x) * kP;
#pragma unroll
    for (int q = 0; q < kCT / 2; ++q)
#pragma unroll
      for (int b = 0; b < kP; ++b) rg2[q][b] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int a = 0; a < kP; ++a) {
      float ga[kP];
#pragma unroll
      for (int b = 0; b < kP; ++b) ga[b] = gl[a * kP + b];
#pragma unroll
      for (int q = 0; q < kCT / 2; ++q) {
        const uchar2 l0 = S.lohi[j][st * kCT + 2 * q], l1 = S.lohi[j][st * kCT + 2 * q + 1];
        const float2 wa = make_float2((l0.x <= l0.y) ? __ldg(wdense + (2 * q) * kP + a) : 0.f,
                                      (l1.x <= l1.y) ? __ldg(wdense + (2 * q + 1) * kP + a) : 0.f);
#pragma unroll
        for (int b = 0; b < kP; ++b) rg2[q][b] = ffma2(wa, ga[b], rg2[q][b]);
      }
    }
  }
}

// One bulk copy (cp.async.bulk, TMA 1-D: SASS UBLKCP) brings the warp's whole 32 x 49 block of grad_out (6272 contiguous,
// 16-byte aligned bytes) into its staging buffer: one instruction from one lane instead of 13 LDGSTS from every lane, and
// completion is an mbarrier phase instead of a cp.async group.
__device__ __forceinline__ void clr_stage_bulk(const BwdParams& p, float* sg, unsigned long long* bar, int m, int c0w, int C, int lane) {
  if (lane == 0) {
    const float* src = p.grad_out + ((int64_t)m * C + c0w) * (kP * kP);
    const unsigned ba = (unsigned)__cvta_generic_to_shared(bar), da = (unsigned)__cvta_generic_to_shared(sg);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"((unsigned)(kGBlk * 4)) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(da), "l"(src), "r"((unsigned)(kGBlk * 4)), "r"(ba) : "memory");
  }
}
__device__ __forceinline__ void clr_wait_bulk(unsigned long long* bar, unsigned parity) {
  const unsigned ba = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(ba), "r"(parity) : "memory");
}

// kSW > 0: compile-time pixel stride of the gradient maps in floats (= C for dense channels_last); 0: run time.  kMinB: CTAs / SM
// the register allocation targets.  kNBuf: staging buffers per warp (RegSmemT), kInfo: footprint records kept in shared memory.
// kDyn: tiles are shared by several CTAs only when dense (p.clr_dyn_min; fill_bwd).
template <int kSW, int kMinB, int kNBuf, bool kInfo, bool kDyn>
__global__ void __launch_bounds__(kCThreads, kMinB) roi_align_bwd_clr_kernel(const __grid_constant__ BwdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  using SM = RegSmemT<kNBuf, kInfo>;
  SM& S = *reinterpret_cast<SM*>(smem_raw);

  int t = (int)(gridDim.x - 1 - blockIdx.x), level = 0;   // coarsest level first (longest CTAs start early)
  while (level + 1 < p.L.num_levels && t >= p.clr_cta_base[level + 1]) ++level;
  t -= p.clr_cta_base[level];
  const int ps = p.clr_ps[level], pt = p.clr_pt[level];
  const int part = t % (ps * pt);     // which share of the tile this CTA works: slab group x sub-tile group
  t /= ps * pt;
  const int per_img = p.cl_tiles_x[level] * p.cl_tiles_y[level];
  const int n = t / per_img;
  t -= n * per_img;
  const int ty0 = (t / p.cl_tiles_x[level]) * kCH, tx0 = (t % p.cl_tiles_x[level]) * kCW;
  const LevelDesc& lv = p.L.lv[level];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.L.C;
  const int nslab = ceil_div(C, kCWarps * 32);
  int slab0 = (part % ps) * nslab / ps, slab1 = ((part % ps) + 1) * nslab / ps;     // my slabs     (only ever changed
  int st0 = (part / ps) * (kCS / pt), st1 = st0 + kCS / pt;                         // my sub-tiles  when kDyn)
  const int r0 = p.roi_off[n], r1 = p.roi_off[n + 1];
  float* const sg = S.sg[warp][0];               // buffer b at sg + b * kGBlk
  unsigned long long* const bar = S.mbar[warp];  // barrier b at bar + b
  unsigned bar_phase = 0;                        // bit b: parity the next wait on barrier b expects
  int cur = 0;                                   // buffer the next pair to be folded arrives in
  if (lane == 0) {
#pragma unroll
    for (int b = 0; b < kNBuf; ++b) {
      const unsigned ba = (unsigned)__cvta_generic_to_shared(bar + b);
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ba));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int64_t sW = kSW > 0 ? (int64_t)kSW : lv.sW;
  const int64_t sH = lv.sH;
  const bool vec = ((sW & 3) == 0) && ((sH & 3) == 0) && ((lv.sN & 3) == 0) && ((reinterpret_cast<uintptr_t>(lv.data) & 15) == 0);
  const int quad = (lane & 7) * 4, xs = lane >> 3;   // zero-fill slot: lane -> (pixel column xs + 4 q, 4 channels)
  const int ncol = min(kCW, lv.W - tx0);             // tile columns inside the map

  int pos = r0;
  bool first = true;
  while (true) {
    cl_collect(p, S, level, tx0, ty0, pos, r1);
    const int nb = S.nb, next = S.next_pos;
    if constexpr (kDyn) {
      // dynamic split: only DENSE tiles (RoIs clustered on a ground-truth box: 100+ RoIs on one coarse-level tile are one
      // warp's 0.3 ms serial chain) are shared by the tile's CTAs; on every other tile CTA 0 does all the work and its
      // siblings leave after this one scan (block-uniform: nb is shared).  A separate instantiation: the same lines as
      // run-time code in the one-CTA-per-tile kernel cost it 4 - 6 % (0.582 -> 0.606 / 0.616 ms at cfg 2: register
      // allocation and code size of a kernel that is instruction-fetch sensitive).
      if (first && nb < p.clr_dyn_min) {
        if (part != 0) return;
        slab0 = 0; slab1 = nslab; st0 = 0; st1 = kCS;
      }
    }
    if (nb > 0) {
      cl_build_tables(p, S, lv, tx0, ty0);   // (uses the staging buffers as scratch: generic-proxy accesses ...)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // ... ordered before the bulk copies that refill them
    } else {
      if (tid < kCS) S.stmask[tid] = 0;
      __syncthreads();
    }
    for (int slab = slab0; slab < slab1; ++slab) {
      const int c0w = slab * (kCWarps * 32) + warp * 32;
      if (c0w >= C) break;
      float* gimg = lv.data + (int64_t)n * lv.sN + c0w;
      int nj = cl_next_pair(S, st0 - 1, 0, st1);
      if (nj >= 0) clr_stage_bulk(p, sg + cur * kGBlk, bar + cur, S.e[nj].m, c0w, C, lane);
      for (int st = st0; st < st1; ++st) {
        unsigned m = S.stmask[st];
        const int ys = ty0 + st * kCT;
        const int ny = min(kCT, lv.H - ys);
        if (m == 0) {   // nothing lands on this sub-tile: zero fill (first batch only)
          if (!first) continue;
          if (vec) {
            float* gp = gimg + (int64_t)ys * sH + (int64_t)(tx0 + xs) * sW + quad;
            const int64_t step = 4 * sW;
            for (int yy = 0; yy < ny; ++yy) {
#pragma unroll
              for (int q = 0; q < kCW / 4; ++q)
                if (xs + 4 * q < ncol) *reinterpret_cast<float4*>(gp + q * step) = make_float4(0.f, 0.f, 0.f, 0.f);
              gp += sH;
            }
          } else {
            for (int yy = 0; yy < ny; ++yy)
              for (int x = 0; x < ncol; ++x) gimg[(int64_t)(ys + yy) * sH + (int64_t)(tx0 + x) * sW + lane] = 0.f;
          }
          continue;
        }
        float2 acc[kCT / 2][kCW];   // acc[q][x] = rows (2q, 2q+1) of column x
        float* gp = gimg + (int64_t)ys * sH + (int64_t)tx0 * sW + lane;
        const bool interior = (ncol == kCW) && (ny == kCT);
        if (first) {
#pragma unroll
          for (int q = 0; q < kCT / 2; ++q)
#pragma unroll
            for (int x = 0; x < kCW; ++x) acc[q][x] = make_float2(0.f, 0.f);
        } else {
          // a later batch of a dense tile CONTINUES the sums of the earlier batches: the 64 partial sums are loaded up
          // front (independent loads, one latency) instead of a load-add-store chain per pixel at write-out - and the
          // result is what one batch holding all the RoIs would have produced
#pragma unroll
          for (int q = 0; q < kCT / 2; ++q) {
#pragma unroll
            for (int x = 0; x < kCW; ++x) {
              const bool okx = interior || (x < ncol);
              const float* g0 = gp + (int64_t)(2 * q) * sH + (int64_t)x * sW;
              acc[q][x].x = (okx && (interior || 2 * q < ny)) ? __ldcg(g0) : 0.f;
              acc[q][x].y = (okx && (interior || 2 * q + 1 < ny)) ? __ldcg(g0 + sH) : 0.f;
            }
          }
        }
        while (m != 0) {
          const int j = __ffs(m) - 1;
          m &= m - 1;
          if constexpr (kNBuf == 2) {   // the other buffer was folded one pair ago (and every lane passed the __syncwarp behind that fold): request the next pair's block now
            nj = cl_next_pair(S, st, m, st1);
            if (nj >= 0) clr_stage_bulk(p, sg + (cur ^ 1) * kGBlk, bar + (cur ^ 1), S.e[nj].m, c0w, C, lane);
          }
          clr_wait_bulk(bar + cur, (bar_phase >> cur) & 1u);
          bar_phase ^= 1u << cur;
          float2 rg2[kCT / 2][kP];
          clr_fold_subtile(sg + cur * kGBlk + lane * (kP * kP), S, j, st, p, ty0 + st * kCT, rg2);
          __syncwarp();   // every lane is done with the staged block
          if constexpr (kNBuf == 2) {
            cur ^= 1;
          } else {        // one buffer: refill it while the columns are expanded
            nj = cl_next_pair(S, st, m, st1);
            if (nj >= 0) clr_stage_bulk(p, sg, bar, S.e[nj].m, c0w, C, lane);
          }
          const unsigned gm = *reinterpret_cast<const unsigned*>(S.xgm[j]);   // 4 x 7-bit bin masks, one per 4-column group
          const float4* xd0 = S.xd0[j];
          const float4* xd1 = S.xd1[j];
#pragma unroll
          for (int gq = 0; gq < kCW / 4; ++gq) {
            const unsigned bm = (gm >> (8 * gq)) & 0x7fu;
            if (bm != 0) {   // warp-uniform: the RoI touches this 4-column group
              float wv[4][8];
#pragma unroll
              for (int xx = 0; xx < 4; ++xx) {
                const float4 a = xd0[4 * gq + xx], b = xd1[4 * gq + xx];
                wv[xx][0] = a.x; wv[xx][1] = a.y; wv[xx][2] = a.z; wv[xx][3] = a.w;
                wv[xx][4] = b.x; wv[xx][5] = b.y; wv[xx][6] = b.z; wv[xx][7] = 0.f;
              }
              // taps are skipped in two halves (bins 0-3 / 4-6): measured faster than bin-by-bin skipping (0.577 vs
              // 0.597 ms) although it executes ~5.6 instead of ~3.5 taps per group - every extra warp-uniform branch
              // costs more than the 8 packed FMAs it saves
              if (bm & 0x0fu) {
#pragma unroll
                for (int b = 0; b < 4; ++b)
#pragma unroll
                  for (int xx = 0; xx < 4; ++xx)
#pragma unroll
                    for (int q = 0; q < kCT / 2; ++q) acc[q][4 * gq + xx] = ffma2(rg2[q][b], wv[xx][b], acc[q][4 * gq + xx]);
              }
              if (bm & 0x70u) {
#pragma unroll
                for (int b = 4; b < kP; ++b)
#pragma unroll
                  for (int xx = 0; xx < 4; ++xx)
#pragma unroll
                    for (int q = 0; q < kCT / 2; ++q) acc[q][4 * gq + xx] = ffma2(rg2[q][b], wv[xx][b], acc[q][4 * gq + xx]);
              }
            }
          }
        }
        // write-out straight from registers: one 128-byte row segment (32 channels of one pixel) per store instruction
        if (interior) {   // interior tile: 64 unconditional stores
#pragma unroll
          for (int q = 0; q < kCT / 2; ++q) {
            float* g0 = gp + (int64_t)(2 * q) * sH;
            float* g1 = g0 + sH;
#pragma unroll
            for (int x = 0; x < kCW; ++x) {
              g0[(int64_t)x * sW] = acc[q][x].x;
              g1[(int64_t)x * sW] = acc[q][x].y;
            }
          }
        } else {
#pragma unroll
          for (int q = 0; q < kCT / 2; ++q) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int r = 2 * q + h;
              if (r < ny) {
#pragma unroll
                for (int x = 0; x < kCW; ++x) {
                  if (x < ncol) gp[(int64_t)r * sH + (int64_t)x * sW] = h ? acc[q][x].y : acc[q][x].x;
                }
              }
            }
          }
        }
      }
    }
    first = false;
    pos = next;
    if (pos >= r1) break;
    __syncthreads();   // tables are rebuilt by the next batch
  }
}

template <int kNBuf, bool kInfo, bool kDyn>
int launch_clr(const BwdParams& p, dim3 grid, bool sw256, cudaStream_t s) {
  const size_t smem = sizeof(RegSmemT<kNBuf, kInfo>);
  if (sw256) {
    OSR_CUDA_CHECK(cudaFuncSetAttribute(roi_align_bwd_clr_kernel<256, 3, kNBuf, kInfo, kDyn>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (kNBuf == 2) OSR_CUDA_CHECK(cudaFuncSetAttribute(roi_align_bwd_clr_kernel<256, 3, kNBuf, kInfo, kDyn>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    roi_align_bwd_clr_kernel<256, 3, kNBuf, kInfo, kDyn><<<grid, kCThreads, smem, s>>>(p);
  } else {
    OSR_CUDA_CHECK(cudaFuncSetAttribute(roi_align_bwd_clr_kernel<0, 3, kNBuf, kInfo, kDyn>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (kNBuf == 2) OSR_CUDA_CHECK(cudaFuncSetAttribute(roi_align_bwd_clr_kernel<0, 3, kNBuf, kInfo, kDyn>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    roi_align_bwd_clr_kernel<0, 3, kNBuf, kInfo, kDyn><<<grid, kCThreads, smem, s>>>(p);
  }
  return 0;
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
  base = 0;
  for (int l = 0; l < num_levels; ++l) {
    p.cl_tiles_x[l] = osr::ceil_div(p.L.lv[l].W, kCW);
    p.cl_tiles_y[l] = osr::ceil_div(p.L.lv[l].H, kCH);
    p.cl_tile_base[l] = base;
    base += p.cl_tiles_x[l] * p.cl_tiles_y[l] * num_images;
  }
  p.cl_tile_base[num_levels] = base;
  // CTAs per tile of the register-accumulator kernel.  The labelled sampler clusters RoIs on the ground-truth boxes, and a
  // coarse-level 16x16 tile (a quarter of the image on P5) then collects 100+ RoIs: one CTA's serial chain of 0.2 - 0.35 ms,
  // which IS the launch time when the batch is small (measured, backward only: 8 images 0.43 ms, 1 image 0.36 ms) - the
  // chain's length follows the RoIs per image, the launch's total work the RoIs per batch.  Shipped for batches of up to 12
  // images: the two coarsest levels launch TWO CTAs per tile (one 128-channel slab each - no pixel is visited twice) and the
  // split is DYNAMIC: the siblings only share a tile whose first batch is full (>= 30 RoIs); everywhere else CTA 0 does all
  // the work and its sibling leaves after the RoI scan.  cfg 2 shapes, backward ms, one CTA -> two: 12 images 0.488 -> 0.468,
  // 8: 0.430 -> 0.326, 4: 0.366 -> 0.248, 1: 0.356 -> 0.255; cfg 3 (8 images) step 0.766 -> 0.685 ms.  At 16 images the dense
  // tiles start first and finish inside the launch, and the siblings' repeated scan / table builds only cost (bench cfg 2:
  // 0.584 -> 0.611, cfg 5: 1.081 -> 1.137), so larger batches keep one CTA per tile.  Every STATIC split (all tiles of a
  // level, up to 8 CTAs) was slower at 16 images too (0x2110 0.607, 0x3310 0.662 against 0.583).
  // OSR_TUNE_BWD_SPLIT (A/B): -1 = one CTA per tile; > 0: hex digit l (finest level first) = log2 of level l's CTAs per tile
  // (slabs first, then sub-tile groups), bits 20-27 = the dynamic threshold in RoIs (0 = static split).
  const int nslab = osr::ceil_div(C, kCWarps * 32);
  int code = osr::tuning(osr::kTuneBwdSplit);
  if (code == 0 && num_images <= 12) {
    code = 30 << 20;
    int c0 = -1, c1 = -1;   // the two coarsest levels (smallest scale), whatever order the caller lists them in
    for (int l = 0; l < num_levels && l < 5; ++l) {
      if (c0 < 0 || p.L.lv[l].scale < p.L.lv[c0].scale) { c1 = c0; c0 = l; }
      else if (c1 < 0 || p.L.lv[l].scale < p.L.lv[c1].scale) c1 = l;
    }
    if (c0 >= 0) code |= 1 << (4 * c0);
    if (c1 >= 0) code |= 1 << (4 * c1);
  }
  p.clr_dyn_min = code > 0 ? ((code >> 20) & 0xff) : 0;
  base = 0;
  for (int l = 0; l < num_levels; ++l) {
    int lg = (code > 0 && l < 5) ? ((code >> (4 * l)) & 0xf) : 0;
    int ps = 1, pt = 1;
    while (lg > 0 && ps * 2 <= nslab && nslab % (ps * 2) == 0 && ps < 2) { ps *= 2; --lg; }   // slabs first: no table work is repeated for pixels a CTA does not own
    while (lg > 0 && pt * 2 <= kCS) { pt *= 2; --lg; }
    p.clr_ps[l] = ps;
    p.clr_pt[l] = pt;
    p.clr_cta_base[l] = base;
    base += p.cl_tiles_x[l] * p.cl_tiles_y[l] * num_images * ps * pt;
  }
  p.clr_cta_base[num_levels] = base;
  return 0;
}

}  // namespace

extern "C" {

size_t osr_roi_align_bwd_workspace(const osr_feat_level_t* h_levels, int num_levels, int num_images, int C, int M) {
  (void)h_levels; (void)num_levels; (void)num_images; (void)C;
  const size_t m = (size_t)(M > 0 ? M : 1);
  return osr::align256(m * sizeof(RoiInfoPacked)) + osr::align256(m * kWRoi * sizeof(float));
}

// mode 0: prepare + gather; 1: prepare only (grad_out / level data pointers unused); 2: gather only (workspace prepared)
static int roi_align_bwd_impl(int mode, const osr_feat_level_t* h_grad_levels, int num_levels, int num_images, int C,
                              const float* grad_out, const float* rois, const int32_t* roi_batch_offsets, int M, int P,
                              int sampling_ratio, int aligned, int canonical_box_size, int canonical_level, int min_level,
                              void* workspace, size_t workspace_bytes, void* stream) {
  osr::DeviceGuard device_guard(workspace);
  BwdParams p;
  int rc = fill_bwd(p, h_grad_levels, num_levels, num_images, C, P, sampling_ratio, aligned, canonical_box_size,
                    canonical_level, min_level);
  if (rc) return rc;
  if (M < 0) return osr::fail_arg(OSR_E_ARG, "roi_align_bwd: M < 0");
  if (num_images == 0) return 0;
  if (!roi_batch_offsets || !workspace || (M > 0 && ((mode != 1 && !grad_out) || !rois)))
    return osr::fail_arg(OSR_E_ARG, "roi_align_bwd: null pointer argument");
  if (workspace_bytes < osr_roi_align_bwd_workspace(h_grad_levels, num_levels, num_images, C, M))
    return osr::fail_arg(OSR_E_WORKSPACE, "roi_align_bwd: workspace too small");
  p.grad_out = grad_out;
  p.rois = rois;
  p.roi_off = roi_batch_offsets;
  p.M = M;
  p.info = static_cast<RoiInfoPacked*>(workspace);
  p.wfull = reinterpret_cast<float*>(static_cast<char*>(workspace) + osr::align256((size_t)(M > 0 ? M : 1) * sizeof(RoiInfoPacked)));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (M > 0 && mode != 2) {
    roi_bwd_prep_kernel<<<osr::ceil_div(M, kPrepWarps), kPrepWarps * 32, 0, s>>>(p);
    OSR_LAUNCH_CHECK();
  }
  if (mode == 1) return 0;
  // channels_last gradient maps with whole 32-channel groups: thread-per-channel kernel
  bool cl = (C % 32 == 0) && ((reinterpret_cast<uintptr_t>(grad_out) & 15) == 0) && osr::tuning(osr::kTuneBwdVariant) != 3;
  for (int l = 0; l < num_levels; ++l) cl = cl && (p.L.lv[l].sC == 1);
  if (cl) {
    const int variant = osr::tuning(osr::kTuneBwdVariant);
    dim3 grid(variant == 2 ? p.cl_tile_base[num_levels] : p.clr_cta_base[num_levels], 1);
    if (variant == 2) {          // shared-memory accumulators (round-1 kernel)
      const size_t smem = sizeof(ClSmem);
      OSR_CUDA_CHECK(cudaFuncSetAttribute(roi_align_bwd_cl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      roi_align_bwd_cl_kernel<<<grid, kCThreads, smem, s>>>(p);
    } else {                     // register accumulators + packed fp32x2 FMAs (shipped)
      bool sw256 = true;
      for (int l = 0; l < num_levels; ++l) sw256 = sw256 && (p.L.lv[l].sW == 256);
      // 0 (shipped): two staging buffers per warp - the next pair's block is requested a whole pair ahead - and the footprint
      // records kept in shared memory by the scan (cfg 2: 0.570 -> 0.562 ms) | 1: one buffer, refilled behind the fold
      const bool dyn = p.clr_dyn_min > 0;
      if (variant == 1) rc = dyn ? launch_clr<1, false, true>(p, grid, sw256, s) : launch_clr<1, false, false>(p, grid, sw256, s);
      else rc = dyn ? launch_clr<2, true, true>(p, grid, sw256, s) : launch_clr<2, true, false>(p, grid, sw256, s);
      if (rc) return rc;
    }
    OSR_LAUNCH_CHECK();
    return 0;
  }
  const int tiles = p.tile_base[num_levels];
  const size_t smem = sizeof(BwdSmem);
  OSR_CUDA_CHECK(cudaFuncSetAttribute(roi_align_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  roi_align_bwd_kernel<<<tiles, kThreads, smem, s>>>(p);
  OSR_LAUNCH_CHECK();
  return 0;
}

int osr_roi_align_bwd(const osr_feat_level_t* h_grad_levels, int num_levels, int num_images, int C,
                      const float* grad_out, const float* rois, const int32_t* roi_batch_offsets, int M, int P,
                      int sampling_ratio, int aligned, int canonical_box_size, int canonical_level, int min_level,
                      void* workspace, size_t workspace_bytes, void* stream) {
  return roi_align_bwd_impl(0, h_grad_levels, num_levels, num_images, C, grad_out, rois, roi_batch_offsets, M, P, sampling_ratio,
                            aligned, canonical_box_size, canonical_level, min_level, workspace, workspace_bytes, stream);
}

int osr_roi_align_bwd_prepare(const osr_feat_level_t* h_levels, int num_levels, int num_images, int C, const float* rois,
                              const int32_t* roi_batch_offsets, int M, int P, int sampling_ratio, int aligned,
                              int canonical_box_size, int canonical_level, int min_level, void* workspace,
                              size_t workspace_bytes, void* stream) {
  return roi_align_bwd_impl(1, h_levels, num_levels, num_images, C, nullptr, rois, roi_batch_offsets, M, P, sampling_ratio,
                            aligned, canonical_box_size, canonical_level, min_level, workspace, workspace_bytes, stream);
}

int osr_roi_align_bwd_prepared(const osr_feat_level_t* h_grad_levels, int num_levels, int num_images, int C,
                               const float* grad_out, const float* rois, const int32_t* roi_batch_offsets, int M, int P,
                               int sampling_ratio, int aligned, int canonical_box_size, int canonical_level, int min_level,
                               void* workspace, size_t workspace_bytes, void* stream) {
  return roi_align_bwd_impl(2, h_grad_levels, num_levels, num_images, C, grad_out, rois, roi_batch_offsets, M, P, sampling_ratio,
                            aligned, canonical_box_size, canonical_level, min_level, workspace, workspace_bytes, stream);
}

}  // extern "C"
