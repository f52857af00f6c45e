// PLN prototype loss (MODEL.PLN.DISTANCE_TYPE = COS | L1 | L2) forward + closed-form backward for sm_100a.
// Replaces prototype_learning_network.py:134,137-187 (F.normalize x2, nonzero (host sync), index, mm, reshape/min,
// two index_puts, K-iteration Python loop :179-180, three relu/sum chains) and its autograd graph.
//
// forward  pln_rows_kernel   : one warp per RoI row; non-foreground rows are skipped without reading emb.
//                              Prototypes are normalised into shared memory once per CTA (K*D*4 <= 28 KB).
//                              Per-CTA partial hinge sums go to the workspace (fixed order => deterministic); the
//                              prototype-separation term (K x K) is spread over the same grid, one warp per prototype.
//          pln_final_kernel  : ordered sum of the partials and the separation hinges (one CTA, fixed tree).
// backward pln_grad_emb_kernel   : one warp per row, d loss/d emb through the normalisation (zeros for inactive rows)
//          pln_grad_reps_partial : grid (rep, row-segment): ordered compaction of the rows whose active hinge points
//                                  at this prototype, then a fixed-order sum of their unit embeddings
//          pln_grad_reps_final   : one CTA per prototype: ordered sum over segments + separation-term gradient + projection.
// No atomics anywhere: results are run-to-run bit-identical.
#include "osr_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxD = 256;   // embedding dim handled: D == 256 fast layout (8 floats per lane), checked on host
constexpr int kMaxReps = 160; // K * reps_per_class (28 classes x 5 representatives = 140: 140 KB of unit prototypes in shared memory)
constexpr int kSeg = 8;      // row segments of the grad_reps partial reduction
constexpr float kEps = 1e-12f;

struct PlnParams {
  const float* emb;
  const float* reps;
  const int64_t* labels;
  const float* ious;
  int R, D, K, rpc, Kr;
  float alpha, beta, loss_weight, iou_thr, r_norm, center_weight;
  // outputs / saved
  float* loss_terms;
  float* emb_inv_norm;
  float* rep_inv_norm;
  int32_t* intra_rep;
  int32_t* inter_rep;
  int32_t* center_rep;
  // workspace
  float* partial;      // fwd: (num_ctas, 2) + (Kr)
  float* bpartial;     // bwd: (kSeg, Kr, D) - a separate region, so the fused call can reduce the loss after the backward partials
  int num_ctas;
  int dist_type;       // OSR_PLN_DIST_COS / _L1 / _L2 (prototype_learning_network.py:156-161,171-176)
  float* saved_dist;   // (2 R + Kr): intra distance, inter distance of every foreground row, separation distance of every
                       // prototype - read by the L2 backward (d dist / d x = (x - y) / dist); may be null for COS / L1
  // backward only
  const float* grad_loss;
  float* grad_emb;
  float* grad_reps;
};

// ---- the three distances between unit vectors and their derivatives (torch.cdist semantics: sign(0) = 0, and the L2
// gradient is 0 where the distance is 0)
enum { kCos = OSR_PLN_DIST_COS, kL1 = OSR_PLN_DIST_L1, kL2 = OSR_PLN_DIST_L2 };
template <int kDist>
__device__ __forceinline__ float dist_acc(float a, float b, float acc) {   // one dimension's contribution to the lane partial
  if (kDist == kCos) return fmaf(a, b, acc);
  const float d = a - b;
  return kDist == kL1 ? acc + fabsf(d) : fmaf(d, d, acc);
}
template <int kDist>
__device__ __forceinline__ float dist_finish(float s) {   // the reduced sum -> the distance
  return kDist == kCos ? 1.0f - s : (kDist == kL1 ? s : sqrtf(s));
}
template <int kDist>
__device__ __forceinline__ float dist_d_first(float a, float b, float dist) {   // d dist(a, b) / d a, one dimension
  if (kDist == kCos) return -b;
  const float d = a - b;
  if (kDist == kL1) return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  return dist > 0.f ? d / dist : 0.f;
}
template <int kDist>
__device__ __forceinline__ float dist_d_second(float a, float b, float dist) {  // d dist(a, b) / d b
  if (kDist == kCos) return -a;
  return -dist_d_first<kDist>(a, b, dist);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// four independent butterfly reductions, interleaved (the distance loops are bound by the latency of these chains)
__device__ __forceinline__ void warp_sum4(float (&v)[4]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] += __shfl_xor_sync(0xffffffffu, v[u], o);
  }
}

// Normalise the prototypes into shared memory: rh[j][d] = r[j][d] / max(||r_j||, eps).  One warp per prototype, FOUR
// prototypes of a warp in flight at a time (their loads and their reductions overlap): every CTA of the loss kernels runs
// this prologue, and as a chain of one prototype after the other it cost 3.9 us of a 16 us kernel (%globaltimer stamps).
__device__ __forceinline__ void load_unit_reps(const PlnParams& p, float* rh, float* inv_norm_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int j0 = warp; j0 < p.Kr; j0 += 4 * kWarps) {
    float v[4][kMaxD / 32], ss[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * kWarps;
      ss[u] = 0.f;
#pragma unroll
      for (int q = 0; q < kMaxD / 32; ++q) {
        const int d = q * 32 + lane;
        v[u][q] = (j < p.Kr && d < p.D) ? __ldg(p.reps + (int64_t)j * p.D + d) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int q = 0; q < kMaxD / 32; ++q) ss[u] = fmaf(v[u][q], v[u][q], ss[u]);   // same order as a strided loop over d
    warp_sum4(ss);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * kWarps;
      if (j < p.Kr) {
        const float denom = fmaxf(sqrtf(ss[u]), kEps);
#pragma unroll
        for (int q = 0; q < kMaxD / 32; ++q) {
          const int d = q * 32 + lane;
          if (d < p.D) rh[j * p.D + d] = v[u][q] / denom;
        }
        if (inv_norm_out && lane == 0) inv_norm_out[j] = 1.0f / denom;
      }
    }
  }
}

// kGrad: also write d loss / d emb of every row (the closed form of pln_grad_emb_kernel, same arithmetic, fused here so the
// forward + backward of the loss is one pass over the embeddings: osr_pln_loss_fwd_bwd)
template <bool kGrad, int kDist>
__global__ void __launch_bounds__(kThreads) pln_rows_kernel(const __grid_constant__ PlnParams p) {
  extern __shared__ __align__(16) float rh[];  // (Kr, D)
  __shared__ float s_part[kWarps][2];
  load_unit_reps(p, rh, blockIdx.x == 0 ? p.rep_inv_norm : nullptr);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (blockIdx.x >= (unsigned)p.num_ctas) {
    // Prototype-separation term (:171-181) in CTAs of its own behind the row CTAs: one warp per prototype k; hinge value to
    // the workspace, arg-min prototype saved for the backward.  (On the first row CTAs it sat on the critical path: 6 us.)
    for (int k = (blockIdx.x - p.num_ctas) * kWarps + warp; k < p.Kr; k += (gridDim.x - p.num_ctas) * kWarps) {
      float rk[kMaxD / 32];
#pragma unroll
      for (int q = 0; q < kMaxD / 32; ++q) rk[q] = (q * 32 + lane < p.D) ? rh[k * p.D + q * 32 + lane] : 0.f;
      float best = 1000.f;
      int best_j = -1;
      for (int j0 = 0; j0 < p.Kr; j0 += 4) {
        float dd[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = min(j0 + u, p.Kr - 1);
          float dot = 0.f;
#pragma unroll
          for (int q = 0; q < kMaxD / 32; ++q) {
            const int d = q * 32 + lane;   // same order over d as the strided loop it replaces
            if (d < p.D) dot = dist_acc<kDist>(rk[q], rh[j * p.D + d], dot);
          }
          dd[u] = dot;
        }
        warp_sum4(dd);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + u;
          if (j >= p.Kr || j / p.rpc == k / p.rpc) continue;  // own-class block masked with 1000 (:179-180)
          const float dist = dist_finish<kDist>(dd[u]);
          if (dist < best) {
            best = dist;
            best_j = j;
          }
        }
      }
      const float h = (p.beta + p.alpha) - best;
      if (lane == 0) {
        p.partial[2 * p.num_ctas + k] = h > 0.f ? h : 0.f;
        p.center_rep[k] = (h > 0.f) ? best_j : -1;
        if (p.saved_dist) p.saved_dist[2 * p.R + k] = best;
      }
    }
    return;
  }
  float intra_sum = 0.f, inter_sum = 0.f;
  const int rows_per_cta = osr::ceil_div(p.R, p.num_ctas);
  const int row_begin = blockIdx.x * rows_per_cta;
  const int row_end = min(p.R, row_begin + rows_per_cta);
  for (int i = row_begin + warp; i < row_end; i += kWarps) {
    const int64_t y = __ldg(p.labels + i);
    const float iou = __ldg(p.ious + i);
    const bool fg = (y >= 0) && (y < p.K) && (iou > p.iou_thr);   // :149-151 (strict >)
    if (!fg) {
      if (lane == 0) {
        p.emb_inv_norm[i] = 0.f;
        p.intra_rep[i] = -1;
        p.inter_rep[i] = -1;
      }
      if (kGrad) {
        float* go = p.grad_emb + (int64_t)i * p.D;
        for (int d = lane; d < p.D; d += 32) go[d] = 0.f;
      }
      continue;
    }
    // unit embedding, 8 dims per lane (D == 256) or strided in general
    float e[kMaxD / 32], eraw[kMaxD / 32];
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < kMaxD / 32; ++q) {
      const int d = q * 32 + lane;
      e[q] = (d < p.D) ? __ldg(p.emb + (int64_t)i * p.D + d) : 0.f;
      eraw[q] = e[q];
      ss = fmaf(e[q], e[q], ss);
    }
    ss = warp_sum(ss);
    const float denom = fmaxf(sqrtf(ss), kEps);
#pragma unroll
    for (int q = 0; q < kMaxD / 32; ++q) e[q] = e[q] / denom;
    // distances to all prototypes; min over the reps of each class (:164)
    float intra = 0.f, inter = 1000.f;   // sentinel 1000 as in :168
    int intra_j = -1, inter_j = -1;
    {
      // prototypes in index order, four at a time (their reductions overlap); per class the first minimum wins (:164), the
      // own class feeds intra, the first minimum over the other classes inter (:166-169)
      float best = 0.f;
      int best_j = -1, r = 0, c = 0;
      for (int j0 = 0; j0 < p.Kr; j0 += 4) {
        float dd[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = min(j0 + u, p.Kr - 1);
          float dot = 0.f;
#pragma unroll
          for (int q = 0; q < kMaxD / 32; ++q) {
            const int d = q * 32 + lane;
            if (d < p.D) dot = dist_acc<kDist>(e[q], rh[j * p.D + d], dot);
          }
          dd[u] = dot;
        }
        warp_sum4(dd);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = j0 + u;
          if (j < p.Kr) {
            const float dist = dist_finish<kDist>(dd[u]);
            if (r == 0 || dist < best) {
              best = dist;
              best_j = j;
            }
            if (++r == p.rpc) {   // class c complete
              if (c == (int)y) {
                intra = best;
                intra_j = best_j;
              } else if (best < inter) {
                inter = best;
                inter_j = best_j;
              }
              r = 0;
              ++c;
            }
          }
        }
      }
    }
    const bool intra_on = (intra - p.alpha) > 0.f;
    const bool inter_on = (p.beta - inter) > 0.f;
    if (intra_on) intra_sum += intra - p.alpha;
    if (inter_on) inter_sum += p.beta - inter;
    if (lane == 0) {
      p.emb_inv_norm[i] = 1.0f / denom;
      p.intra_rep[i] = intra_on ? intra_j : -1;
      p.inter_rep[i] = inter_on ? inter_j : -1;
      if (p.saved_dist) {
        p.saved_dist[i] = intra;
        p.saved_dist[p.R + i] = inter;
      }
    }
    if (kGrad) {   // pln_grad_emb_kernel's row, with the values already in registers
      const int ja = intra_on ? intra_j : -1, jb = inter_on ? inter_j : -1;
      float* go = p.grad_emb + (int64_t)i * p.D;
      if (ja < 0 && jb < 0) {
        for (int d = lane; d < p.D; d += 32) go[d] = 0.f;
      } else {
        const float inv = 1.0f / denom;
        const float sc = __ldg(p.grad_loss) * p.loss_weight / fmaxf(p.r_norm, 1.0f);
        float eh[kMaxD / 32], g[kMaxD / 32];
        float dot = 0.f;
#pragma unroll
        for (int q = 0; q < kMaxD / 32; ++q) {
          const int d = q * 32 + lane;
          eh[q] = 0.f; g[q] = 0.f;
          if (d < p.D) {
            eh[q] = eraw[q] * inv;
            float t = 0.f;   // d (relu(d_y - alpha) + relu(beta - d_c*)) / d e_hat
            if (ja >= 0) t += dist_d_first<kDist>(eh[q], rh[ja * p.D + d], intra);
            if (jb >= 0) t -= dist_d_first<kDist>(eh[q], rh[jb * p.D + d], inter);
            g[q] = t;
            dot = fmaf(eh[q], t, dot);
          }
        }
        dot = warp_sum(dot);
        const float f = sc * inv;
#pragma unroll
        for (int q = 0; q < kMaxD / 32; ++q) {
          const int d = q * 32 + lane;
          if (d < p.D) go[d] = f * (g[q] - eh[q] * dot);
        }
      }
    }
  }
  if (lane == 0) {
    s_part[warp][0] = intra_sum;
    s_part[warp][1] = inter_sum;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < kWarps; ++w) {
      a += s_part[w][0];
      b += s_part[w][1];
    }
    p.partial[2 * blockIdx.x] = a;
    p.partial[2 * blockIdx.x + 1] = b;
  }
}

// final ordered reduction (:183-187) of the per-CTA partial sums and the per-prototype separation hinges (one CTA)
__device__ __forceinline__ void pln_final_body(const PlnParams& p) {
  __shared__ float s_c[kMaxReps];
  if (threadIdx.x < p.Kr) s_c[threadIdx.x] = p.partial[2 * p.num_ctas + threadIdx.x];
  // ordered sum of the per-CTA partials: thread t adds partials t, t + 256, ... in that order, then a fixed tree
  __shared__ float s_a[kThreads], s_b[kThreads];
  {
    float a = 0.f, b = 0.f;
    for (int i = threadIdx.x; i < p.num_ctas; i += kThreads) {
      a += p.partial[2 * i];
      b += p.partial[2 * i + 1];
    }
    s_a[threadIdx.x] = a;
    s_b[threadIdx.x] = b;
  }
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_a[threadIdx.x] += s_a[threadIdx.x + o];
      s_b[threadIdx.x] += s_b[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float a = s_a[0], b = s_b[0];
    float c = 0.f;
    for (int k = 0; k < p.Kr; ++k) c += s_c[k];
    const float denom = fmaxf(p.r_norm, 1.0f);
    p.loss_terms[0] = (a + b + p.center_weight * c) * p.loss_weight / denom;
    p.loss_terms[1] = a;
    p.loss_terms[2] = b;
    p.loss_terms[3] = c;
  }
}

__global__ void __launch_bounds__(kThreads) pln_final_kernel(const __grid_constant__ PlnParams p) { pln_final_body(p); }

// ------------------------------------------------------------------------------------------ backward
template <int kDist>
__global__ void __launch_bounds__(kThreads) pln_grad_emb_kernel(const __grid_constant__ PlnParams p) {
  extern __shared__ __align__(16) float rh[];
  load_unit_reps(p, rh, nullptr);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float s = __ldg(p.grad_loss) * p.loss_weight / fmaxf(p.r_norm, 1.0f);
  for (int i = blockIdx.x * kWarps + warp; i < p.R; i += gridDim.x * kWarps) {
    const int ja = p.intra_rep[i], jb = p.inter_rep[i];
    float* go = p.grad_emb + (int64_t)i * p.D;
    if (ja < 0 && jb < 0) {
      for (int d = lane; d < p.D; d += 32) go[d] = 0.f;
      continue;
    }
    const float inv = p.emb_inv_norm[i];
    const float da = kDist == kL2 ? p.saved_dist[i] : 0.f, db = kDist == kL2 ? p.saved_dist[p.R + i] : 0.f;
    float eh[kMaxD / 32], g[kMaxD / 32];
    float dot = 0.f;
#pragma unroll
    for (int q = 0; q < kMaxD / 32; ++q) {
      const int d = q * 32 + lane;
      eh[q] = 0.f; g[q] = 0.f;
      if (d < p.D) {
        eh[q] = __ldg(p.emb + (int64_t)i * p.D + d) * inv;
        float t = 0.f;
        if (ja >= 0) t += dist_d_first<kDist>(eh[q], rh[ja * p.D + d], da);   // relu(d_y - alpha): +d dist / d e_hat (COS: -r_hat)
        if (jb >= 0) t -= dist_d_first<kDist>(eh[q], rh[jb * p.D + d], db);   // relu(beta - d_c*): -d dist / d e_hat
        g[q] = t;
        dot = fmaf(eh[q], t, dot);
      }
    }
    dot = warp_sum(dot);
    const float f = s * inv;
#pragma unroll
    for (int q = 0; q < kMaxD / 32; ++q) {
      const int d = q * 32 + lane;
      if (d < p.D) go[d] = f * (g[q] - eh[q] * dot);
    }
  }
}

// grid (Kr, kSeg): partial[seg][j][:] = sum over rows of segment seg (in row order) of G[i][j] * d dist(e_hat_i, r_hat_j) / d r_hat_j
// (COS: -e_hat_i)
template <int kDist>
__global__ void __launch_bounds__(kThreads) pln_grad_reps_partial(const __grid_constant__ PlnParams p) {
  __shared__ int s_rows[2048];       // |row+1| with sign = sign of G; segment length <= 2048 enforced by the host loop
  __shared__ int s_wcnt[kWarps + 1];
  const int j = blockIdx.x, seg = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int seg_len = osr::ceil_div(p.R, kSeg);
  const int r0 = seg * seg_len, r1 = min(p.R, r0 + seg_len);
  float acc[kMaxD / kThreads > 0 ? kMaxD / kThreads : 1] = {0.f};
  const float rj = (kDist != kCos && tid < p.D) ? __ldg(p.reps + (int64_t)j * p.D + tid) * p.rep_inv_norm[j] : 0.f;   // r_hat_j[tid]
  for (int base = r0; base < r1; base += 2048) {
    // ordered compaction of matching rows of [base, base+2048): each thread owns 8 consecutive rows
    int code[8], cnt = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = base + tid * 8 + k;
      code[k] = 0;
      if (i < r1) {
        const int ja = p.intra_rep[i], jb = p.inter_rep[i];
        // a row can point at j through only one of the two hinges (own class vs other class)
        if (ja == j) code[k] = -(i + 1);
        else if (jb == j) code[k] = (i + 1);
      }
      cnt += code[k] != 0;
    }
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_wcnt[warp] = inc;
    __syncthreads();
    if (tid == 0) {
      int run = 0;
      for (int w = 0; w < kWarps; ++w) {
        int c = s_wcnt[w];
        s_wcnt[w] = run;
        run += c;
      }
      s_wcnt[kWarps] = run;
    }
    __syncthreads();
    int pos = s_wcnt[warp] + inc - cnt;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (code[k] != 0) s_rows[pos++] = code[k];
    __syncthreads();
    const int n = s_wcnt[kWarps];
    // four rows per iteration: their loads are independent and issued together (the loop is latency-bound otherwise);
    // the FMAs keep the row order, so the sum is bit-identical to the one-row-at-a-time loop
    int q = 0;
    if (kDist != kCos) {   // general form: the derivative depends on (e_hat_i - r_hat_j) and, for L2, on the saved distance
      for (; q < n; ++q) {
        const int c = s_rows[q];
        const int i = (c < 0 ? -c : c) - 1;
        const float coef = c < 0 ? 1.f : -1.f;   // intra hinge: +d dist, inter hinge: -d dist
        const float dist = kDist == kL2 ? p.saved_dist[(c < 0 ? 0 : p.R) + i] : 0.f;
        if (tid < p.D) {
          const float eh = __ldg(p.emb + (int64_t)i * p.D + tid) * p.emb_inv_norm[i];
          acc[0] = fmaf(coef, dist_d_second<kDist>(eh, rj, dist), acc[0]);
        }
      }
    }
    for (; q + 4 <= n; q += 4) {
      float x[4], w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = s_rows[q + u];
        const int i = (c < 0 ? -c : c) - 1;
        w[u] = p.emb_inv_norm[i] * (c < 0 ? -1.f : 1.f);
        x[u] = (tid < p.D) ? __ldg(p.emb + (int64_t)i * p.D + tid) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[0] = fmaf(x[u], w[u], acc[0]);
    }
    for (; q < n; ++q) {
      const int c = s_rows[q];
      const int i = (c < 0 ? -c : c) - 1;
      const float sg = c < 0 ? -1.f : 1.f;
      const float inv = p.emb_inv_norm[i] * sg;
      if (tid < p.D) acc[0] = fmaf(__ldg(p.emb + (int64_t)i * p.D + tid), inv, acc[0]);
    }
    __syncthreads();
  }
  if (tid < p.D) p.bpartial[((int64_t)seg * p.Kr + j) * p.D + tid] = acc[0];
}

// kWithLoss: the grid has one extra CTA (blockIdx.x == Kr) that performs pln_final_kernel's reduction - the fused
// forward + backward call saves a launch
template <int kDist, bool kWithLoss>
__global__ void __launch_bounds__(kThreads) pln_grad_reps_final(const __grid_constant__ PlnParams p) {
  extern __shared__ __align__(16) float rh[];
  __shared__ float s_red[kWarps];
  if (kWithLoss && blockIdx.x == (unsigned)p.Kr) {   // CTA-uniform
    pln_final_body(p);
    return;
  }
  load_unit_reps(p, rh, nullptr);
  __syncthreads();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float s = __ldg(p.grad_loss) * p.loss_weight / fmaxf(p.r_norm, 1.0f);
  {
    const int j = blockIdx.x;   // one CTA per prototype
    float g = 0.f;
    if (tid < p.D) {
      for (int seg = 0; seg < kSeg; ++seg) g += p.bpartial[((int64_t)seg * p.Kr + j) * p.D + tid];
      // separation term: (Gamma + Gamma^T) r_hat, Gamma[k][j*_k] = center_weight * [hinge active]
      // relu(beta + alpha - c_dist_k): -center_weight * d dist(r_hat_k, r_hat_j*k) / d (either argument)
      const float* cd = p.saved_dist ? p.saved_dist + 2 * p.R : nullptr;
      const int js = p.center_rep[j];
      if (js >= 0) g = fmaf(-p.center_weight, dist_d_first<kDist>(rh[j * p.D + tid], rh[js * p.D + tid], kDist == kL2 ? cd[j] : 0.f), g);
      for (int k = 0; k < p.Kr; ++k)
        if (p.center_rep[k] == j)
          g = fmaf(-p.center_weight, dist_d_second<kDist>(rh[k * p.D + tid], rh[j * p.D + tid], kDist == kL2 ? cd[k] : 0.f), g);
    }
    float dot = (tid < p.D) ? rh[j * p.D + tid] * g : 0.f;
    dot = warp_sum(dot);
    if (lane == 0) s_red[warp] = dot;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < kWarps; ++w) tot += s_red[w];
    if (tid < p.D) p.grad_reps[(int64_t)j * p.D + tid] = s * p.rep_inv_norm[j] * (g - rh[j * p.D + tid] * tot);
  }
}

struct NearestParams {
  PlnParams base;
  float unk_thr;
  int64_t unknown_id;
  const int64_t* class_id_map;
  int64_t* pred;
  float* min_dist;
};

// PLN.inference: nearest prototype + unknown threshold, one warp per row
template <int kDist>
__global__ void __launch_bounds__(kThreads) pln_nearest_kernel(const __grid_constant__ NearestParams q) {
  extern __shared__ __align__(16) float rh[];
  const PlnParams& p = q.base;
  load_unit_reps(p, rh, nullptr);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = blockIdx.x * kWarps + warp; i < p.R; i += gridDim.x * kWarps) {
    float e[kMaxD / 32];
    float ss = 0.f;
#pragma unroll
    for (int t = 0; t < kMaxD / 32; ++t) {
      const int d = t * 32 + lane;
      e[t] = (d < p.D) ? __ldg(p.emb + (int64_t)i * p.D + d) : 0.f;
      ss = fmaf(e[t], e[t], ss);
    }
    ss = warp_sum(ss);
    const float denom = fmaxf(sqrtf(ss), kEps);
#pragma unroll
    for (int t = 0; t < kMaxD / 32; ++t) e[t] = e[t] / denom;
    float best = 0.f;
    int best_c = -1;
    for (int j0 = 0; j0 < p.Kr; j0 += 4) {
      float dd[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = min(j0 + u, p.Kr - 1);
        float dot = 0.f;
#pragma unroll
        for (int t = 0; t < kMaxD / 32; ++t) {
          const int d = t * 32 + lane;
          if (d < p.D) dot = dist_acc<kDist>(e[t], rh[j * p.D + d], dot);
        }
        dd[u] = dot;
      }
      warp_sum4(dd);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u;
        if (j < p.Kr) {
          const float dist = dist_finish<kDist>(dd[u]);
          if (best_c < 0 || dist < best) {   // first minimum over all prototypes = first minimum over the class minima (:217-218)
            best = dist;
            best_c = j / p.rpc;
          }
        }
      }
    }
    if (lane == 0) {
      int64_t cls = q.class_id_map ? q.class_id_map[best_c] : (int64_t)best_c;
      if (best > q.unk_thr) cls = q.unknown_id;
      q.pred[i] = cls;
      q.min_dist[i] = best;
    }
  }
}

int check_shape(int R, int D, int K, int rpc) {
  if (R < 0 || D <= 0 || K <= 0 || rpc <= 0) return osr::fail_arg(OSR_E_ARG, "pln: bad R/D/K/reps_per_class");
  if (D > kMaxD) return osr::fail_arg(OSR_E_SHAPE, "pln: embedding dim %d > %d unsupported", D, kMaxD);
  if (K * rpc > kMaxReps) return osr::fail_arg(OSR_E_SHAPE, "pln: K*reps_per_class=%d > %d unsupported", K * rpc, kMaxReps);
  return 0;
}

int check_dist(int distance_type, const void* saved_dist, bool need_saved) {
  if (distance_type != kCos && distance_type != kL1 && distance_type != kL2)
    return osr::fail_arg(OSR_E_ARG, "pln: distance_type %d is not OSR_PLN_DIST_COS / _L1 / _L2", distance_type);
  if (need_saved && distance_type == kL2 && !saved_dist)
    return osr::fail_arg(OSR_E_ARG, "pln: the L2 distance needs the saved_dist buffer (2 R + K * reps_per_class floats)");
  return 0;
}

// run STMT with kD bound to the compile-time distance type
#define OSR_PLN_DISPATCH(DT, ...)                                  \
  do {                                                             \
    if ((DT) == kL1) { constexpr int kD = kL1; __VA_ARGS__; }      \
    else if ((DT) == kL2) { constexpr int kD = kL2; __VA_ARGS__; } \
    else { constexpr int kD = kCos; __VA_ARGS__; }                 \
  } while (0)

// two rows per warp: enough CTAs to fill the GPU in one wave at R = 8192 (the stage is latency-bound)
int fwd_ctas(int R) { return R < 2 * kWarps ? 1 : (R / (2 * kWarps) > 592 ? 592 : R / (2 * kWarps)); }
size_t fwd_partial_bytes(int R, int Kr) { return ((size_t)fwd_ctas(R > 0 ? R : 1) * 2 + (size_t)Kr) * sizeof(float); }

}  // namespace

extern "C" {

size_t osr_pln_workspace(int R, int D, int K, int reps_per_class) {
  return osr::align256(fwd_partial_bytes(R, K * reps_per_class)) + osr::align256((size_t)kSeg * K * reps_per_class * D * sizeof(float));
}

int osr_pln_loss_fwd(const float* emb, const float* reps, const int64_t* labels, const float* ious, int R, int D,
                     int K, int reps_per_class, int distance_type, float alpha, float beta, float loss_weight,
                     float iou_threshold, float r_norm, float center_weight, float* loss_terms, float* emb_inv_norm,
                     float* rep_inv_norm, int32_t* intra_rep, int32_t* inter_rep, int32_t* center_rep, float* saved_dist,
                     void* workspace, size_t workspace_bytes, void* stream) {
  osr::DeviceGuard device_guard(emb);
  int rc = check_shape(R, D, K, reps_per_class);
  if (rc) return rc;
  if ((rc = check_dist(distance_type, saved_dist, false))) return rc;
  if (!reps || !loss_terms || !rep_inv_norm || !center_rep || !workspace ||
      (R > 0 && (!emb || !labels || !ious || !emb_inv_norm || !intra_rep || !inter_rep)))
    return osr::fail_arg(OSR_E_ARG, "pln_loss_fwd: null pointer argument");
  if (workspace_bytes < osr_pln_workspace(R, D, K, reps_per_class))
    return osr::fail_arg(OSR_E_WORKSPACE, "pln_loss_fwd: workspace too small");
  PlnParams p{};
  p.emb = emb; p.reps = reps; p.labels = labels; p.ious = ious;
  p.R = R; p.D = D; p.K = K; p.rpc = reps_per_class; p.Kr = K * reps_per_class;
  p.alpha = alpha; p.beta = beta; p.loss_weight = loss_weight; p.iou_thr = iou_threshold;
  p.r_norm = r_norm; p.center_weight = center_weight;
  p.loss_terms = loss_terms; p.emb_inv_norm = emb_inv_norm; p.rep_inv_norm = rep_inv_norm;
  p.intra_rep = intra_rep; p.inter_rep = inter_rep; p.center_rep = center_rep;
  p.dist_type = distance_type; p.saved_dist = saved_dist;
  p.partial = static_cast<float*>(workspace);
  p.bpartial = reinterpret_cast<float*>(static_cast<unsigned char*>(workspace) + osr::align256(fwd_partial_bytes(R, p.Kr)));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t smem = (size_t)p.Kr * D * sizeof(float);
  p.num_ctas = fwd_ctas(R > 0 ? R : 1);   // also launched for R == 0: the separation term does not depend on the rows
  OSR_PLN_DISPATCH(distance_type, {
    OSR_CUDA_CHECK(cudaFuncSetAttribute(pln_rows_kernel<false, kD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pln_rows_kernel<false, kD><<<p.num_ctas + osr::ceil_div(p.Kr, kWarps), kThreads, smem, s>>>(p);
  });
  OSR_LAUNCH_CHECK();
  pln_final_kernel<<<1, kThreads, 0, s>>>(p);
  OSR_LAUNCH_CHECK();
  return 0;
}

// Loss forward AND closed-form backward in three launches (rows + d loss / d emb fused, prototype-gradient
// partials, prototype-gradient final + loss reduction in an extra CTA) instead of five plus the autograd plumbing: the training step's S5.
// phase 0: everything; 1: the row launch only (loss partials, saved state, d loss / d emb); 2: the two prototype-gradient
// launches + loss reduction on what phase 1 left in the buffers
static int pln_loss_fwd_bwd_impl(int phase, const float* emb, const float* reps, const int64_t* labels, const float* ious,
                                 const float* grad_loss, int R, int D, int K, int reps_per_class, int distance_type, float alpha,
                                 float beta, float loss_weight, float iou_threshold, float r_norm, float center_weight,
                                 float* loss_terms, float* emb_inv_norm, float* rep_inv_norm, int32_t* intra_rep,
                                 int32_t* inter_rep, int32_t* center_rep, float* saved_dist, float* grad_emb, float* grad_reps,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  osr::DeviceGuard device_guard(grad_reps);
  int rc = check_shape(R, D, K, reps_per_class);
  if (rc) return rc;
  if ((rc = check_dist(distance_type, saved_dist, true))) return rc;
  if (!reps || !loss_terms || !rep_inv_norm || !center_rep || !workspace || !grad_loss || !grad_reps ||
      (R > 0 && (!emb || !labels || !ious || !emb_inv_norm || !intra_rep || !inter_rep || !grad_emb)))
    return osr::fail_arg(OSR_E_ARG, "pln_loss_fwd_bwd: null pointer argument");
  if (workspace_bytes < osr_pln_workspace(R, D, K, reps_per_class))
    return osr::fail_arg(OSR_E_WORKSPACE, "pln_loss_fwd_bwd: workspace too small");
  PlnParams p{};
  p.emb = emb; p.reps = reps; p.labels = labels; p.ious = ious;
  p.R = R; p.D = D; p.K = K; p.rpc = reps_per_class; p.Kr = K * reps_per_class;
  p.alpha = alpha; p.beta = beta; p.loss_weight = loss_weight; p.iou_thr = iou_threshold;
  p.r_norm = r_norm; p.center_weight = center_weight;
  p.loss_terms = loss_terms; p.emb_inv_norm = emb_inv_norm; p.rep_inv_norm = rep_inv_norm;
  p.intra_rep = intra_rep; p.inter_rep = inter_rep; p.center_rep = center_rep;
  p.grad_loss = grad_loss; p.grad_emb = grad_emb; p.grad_reps = grad_reps;
  p.dist_type = distance_type; p.saved_dist = saved_dist;
  p.partial = static_cast<float*>(workspace);
  p.bpartial = reinterpret_cast<float*>(static_cast<unsigned char*>(workspace) + osr::align256(fwd_partial_bytes(R, p.Kr)));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t smem = (size_t)p.Kr * D * sizeof(float);
  p.num_ctas = fwd_ctas(R > 0 ? R : 1);
  OSR_PLN_DISPATCH(distance_type, {
    OSR_CUDA_CHECK(cudaFuncSetAttribute(pln_rows_kernel<true, kD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OSR_CUDA_CHECK(cudaFuncSetAttribute(pln_grad_reps_final<kD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (phase != 2) {
      pln_rows_kernel<true, kD><<<p.num_ctas + osr::ceil_div(p.Kr, kWarps), kThreads, smem, s>>>(p);
      OSR_LAUNCH_CHECK();
    }
    if (phase != 1) {
      pln_grad_reps_partial<kD><<<dim3(p.Kr, kSeg), kThreads, 0, s>>>(p);
      OSR_LAUNCH_CHECK();
      pln_grad_reps_final<kD, true><<<p.Kr + 1, kThreads, smem, s>>>(p);   // CTA Kr reduces the loss (pln_final_kernel's job)
      OSR_LAUNCH_CHECK();
    }
  });
  return 0;
}

int osr_pln_loss_fwd_bwd(const float* emb, const float* reps, const int64_t* labels, const float* ious, const float* grad_loss,
                         int R, int D, int K, int reps_per_class, int distance_type, float alpha, float beta,
                         float loss_weight, float iou_threshold, float r_norm, float center_weight, float* loss_terms,
                         float* emb_inv_norm, float* rep_inv_norm, int32_t* intra_rep, int32_t* inter_rep,
                         int32_t* center_rep, float* saved_dist, float* grad_emb, float* grad_reps, void* workspace,
                         size_t workspace_bytes, void* stream) {
  return pln_loss_fwd_bwd_impl(0, emb, reps, labels, ious, grad_loss, R, D, K, reps_per_class, distance_type, alpha, beta,
                               loss_weight, iou_threshold, r_norm, center_weight, loss_terms, emb_inv_norm, rep_inv_norm,
                               intra_rep, inter_rep, center_rep, saved_dist, grad_emb, grad_reps, workspace, workspace_bytes,
                               stream);
}

int osr_pln_loss_fwd_bwd_phase(int phase, const float* emb, const float* reps, const int64_t* labels, const float* ious,
                               const float* grad_loss, int R, int D, int K, int reps_per_class, int distance_type, float alpha,
                               float beta, float loss_weight, float iou_threshold, float r_norm, float center_weight,
                               float* loss_terms, float* emb_inv_norm, float* rep_inv_norm, int32_t* intra_rep,
                               int32_t* inter_rep, int32_t* center_rep, float* saved_dist, float* grad_emb, float* grad_reps,
                               void* workspace, size_t workspace_bytes, void* stream) {
  if (phase != 1 && phase != 2) return osr::fail_arg(OSR_E_ARG, "pln_loss_fwd_bwd_phase: phase must be 1 (rows) or 2 (prototype gradient + loss)");
  return pln_loss_fwd_bwd_impl(phase, emb, reps, labels, ious, grad_loss, R, D, K, reps_per_class, distance_type, alpha, beta,
                               loss_weight, iou_threshold, r_norm, center_weight, loss_terms, emb_inv_norm, rep_inv_norm,
                               intra_rep, inter_rep, center_rep, saved_dist, grad_emb, grad_reps, workspace, workspace_bytes,
                               stream);
}

int osr_pln_loss_bwd(const float* emb, const float* reps, const int64_t* labels, const float* emb_inv_norm,
                     const float* rep_inv_norm, const int32_t* intra_rep, const int32_t* inter_rep,
                     const int32_t* center_rep, const float* saved_dist, const float* grad_loss, int R, int D, int K,
                     int reps_per_class, int distance_type, float loss_weight, float r_norm, float center_weight,
                     float* grad_emb, float* grad_reps, void* workspace, size_t workspace_bytes, void* stream) {
  osr::DeviceGuard device_guard(emb);
  (void)labels;
  int rc = check_shape(R, D, K, reps_per_class);
  if (rc) return rc;
  if ((rc = check_dist(distance_type, saved_dist, true))) return rc;
  if (!reps || !rep_inv_norm || !center_rep || !grad_loss || !grad_reps || !workspace ||
      (R > 0 && (!emb || !emb_inv_norm || !intra_rep || !inter_rep || !grad_emb)))
    return osr::fail_arg(OSR_E_ARG, "pln_loss_bwd: null pointer argument");
  if (workspace_bytes < osr_pln_workspace(R, D, K, reps_per_class))
    return osr::fail_arg(OSR_E_WORKSPACE, "pln_loss_bwd: workspace too small");
  PlnParams p{};
  p.emb = emb; p.reps = reps;
  p.R = R; p.D = D; p.K = K; p.rpc = reps_per_class; p.Kr = K * reps_per_class;
  p.loss_weight = loss_weight; p.r_norm = r_norm; p.center_weight = center_weight;
  p.emb_inv_norm = const_cast<float*>(emb_inv_norm); p.rep_inv_norm = const_cast<float*>(rep_inv_norm);
  p.intra_rep = const_cast<int32_t*>(intra_rep); p.inter_rep = const_cast<int32_t*>(inter_rep);
  p.center_rep = const_cast<int32_t*>(center_rep);
  p.grad_loss = grad_loss; p.grad_emb = grad_emb; p.grad_reps = grad_reps;
  p.dist_type = distance_type; p.saved_dist = const_cast<float*>(saved_dist);
  p.partial = static_cast<float*>(workspace);
  p.bpartial = reinterpret_cast<float*>(static_cast<unsigned char*>(workspace) + osr::align256(fwd_partial_bytes(R, p.Kr)));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t smem = (size_t)p.Kr * D * sizeof(float);
  OSR_PLN_DISPATCH(distance_type, {
    OSR_CUDA_CHECK(cudaFuncSetAttribute(pln_grad_emb_kernel<kD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    OSR_CUDA_CHECK(cudaFuncSetAttribute(pln_grad_reps_final<kD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (R > 0) {
      int ctas = osr::ceil_div(R, kWarps * 4);
      if (ctas > 592) ctas = 592;
      pln_grad_emb_kernel<kD><<<ctas, kThreads, smem, s>>>(p);
      OSR_LAUNCH_CHECK();
    }
    pln_grad_reps_partial<kD><<<dim3(p.Kr, kSeg), kThreads, 0, s>>>(p);
    OSR_LAUNCH_CHECK();
    pln_grad_reps_final<kD, false><<<p.Kr, kThreads, smem, s>>>(p);
    OSR_LAUNCH_CHECK();
  });
  return 0;
}

int osr_pln_nearest(const float* emb, const float* reps, int R, int D, int K, int reps_per_class, int distance_type,
                    float unk_thr, int64_t unknown_id, const int64_t* class_id_map, int64_t* pred, float* min_dist,
                    void* stream) {
  osr::DeviceGuard device_guard(emb);
  int rc = check_shape(R, D, K, reps_per_class);
  if (rc) return rc;
  if ((rc = check_dist(distance_type, nullptr, false))) return rc;
  if (R == 0) return 0;
  if (!emb || !reps || !pred || !min_dist) return osr::fail_arg(OSR_E_ARG, "pln_nearest: null pointer argument");
  NearestParams q{};
  q.base.emb = emb; q.base.reps = reps;
  q.base.R = R; q.base.D = D; q.base.K = K; q.base.rpc = reps_per_class; q.base.Kr = K * reps_per_class;
  q.unk_thr = unk_thr; q.unknown_id = unknown_id; q.class_id_map = class_id_map; q.pred = pred; q.min_dist = min_dist;
  const size_t smem = (size_t)q.base.Kr * D * sizeof(float);
  int ctas = osr::ceil_div(R, kWarps * 4);
  if (ctas > 592) ctas = 592;
  OSR_PLN_DISPATCH(distance_type, {
    OSR_CUDA_CHECK(cudaFuncSetAttribute(pln_nearest_kernel<kD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pln_nearest_kernel<kD><<<ctas, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(q);
  });
  OSR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
