// CF-RPN proposal stage for sm_100a: per-(image, level) segmented top-k (radix select over a thread-block
// cluster, scores read from HBM exactly once into distributed shared memory), then - only for the k
// survivors - gather deltas, Box2BoxTransformLinear decode, finite check, clip, small-box filter, ordered
// compaction.  Replaces classification_free_rpn.py:558-610 + find_top_proposals.py:63-110.
//
// Work decomposition: one cluster of 8 CTAs per (image, level) segment.  Each CTA owns a contiguous 1/8
// slice of the segment's scores, converted once to order-preserving uint32 keys in its shared memory.
//   pass 0..3 : 8-bit MSD radix select; per-CTA histograms are merged through DSMEM (one cluster.sync per
//               pass, histograms double-buffered), every CTA redundantly finds the digit of the k-th key.
//   collect   : keys > T plus the lowest-index ties (key == T) are pushed into rank 0's candidate buffer
//               through DSMEM at deterministic positions (ordered block scan + cross-rank prefix).
//   rank 0    : bitonic sort of (key << 32 | ~index) => score descending, ties by lower anchor index; decode
//               the survivors with separately rounded fp32 ops (bit-exact with the torch elementwise chain),
//               clip to the image, drop non-finite / empty boxes, compact in order, write counts + flags.
// A second tiny kernel concatenates the per-level runs of every image (needs all levels' counts).
#include <cooperative_groups.h>

#include "osr_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kCluster = 8;
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kBins = 256;

struct RpnParams {
  osr_rpn_level_t lv[OSR_MAX_LEVELS];
  int koff[OSR_MAX_LEVELS + 1];
  int ktop[OSR_MAX_LEVELS];
  int num_levels, num_images, kmax;
  int slice_cap;  // keys per CTA the shared-memory carve-up can hold
  // cluster packing: a level whose slice fits in fewer CTAs shares its 8-CTA cluster between 8 / split images, so small
  // levels do not burn whole clusters (800x1333: 26 clusters = 208 CTAs = ONE wave on 148 SMs, instead of 80 = 640 = three)
  int split[OSR_MAX_LEVELS];          // CTAs that share one (image, level) segment: 8, 4, 2 or 1
  int cl_base[OSR_MAX_LEVELS + 1];    // first cluster id of each level
  int kpad;       // candidate-buffer capacity (power of two >= max k)
  float min_box_size;
  const int32_t* image_hw;
  float* st_boxes;    // staging (N, Kmax, 4): level l of image n compacted at koff[l]
  float* st_scores;   // (N, Kmax)
  int32_t* st_index;  // (N, Kmax)
  int32_t* st_flags;  // (N, L)
  int32_t* counts;    // (N, L+2)
};

__device__ __forceinline__ uint32_t score_to_key(float f) {
  uint32_t b = __float_as_uint(f);
  if ((b & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;  // NaN ranks first, as in torch.topk / torch.sort
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_score(uint32_t k) {
  uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(b);
}

__device__ __forceinline__ float relu_keep_nan(float x) { return (x <= 0.f) ? 0.f : x; }

// exclusive scan of one uint32 per thread over the CTA; returns exclusive prefix, *total = CTA sum.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sums, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = (lane < kWarps) ? warp_sums[lane] : 0u;
    uint32_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < kWarps) warp_sums[lane] = winc - w;  // exclusive warp offsets
    if (lane == kWarps - 1) warp_sums[kWarps] = winc;
  }
  __syncthreads();
  uint32_t res = warp_sums[warp] + inc - v;
  *total = warp_sums[kWarps];
  __syncthreads();  // warp_sums reusable
  return res;
}

__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads, 2)
    rpn_select_decode_kernel(const __grid_constant__ RpnParams p) {
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int cid = blockIdx.x / kCluster;
  int level = 0;   // big (fine) levels first
  while (level + 1 < p.num_levels && cid >= p.cl_base[level + 1]) ++level;
  const int split = p.split[level];             // CTAs sharing this segment
  const int seg = crank / split;                // segment of this cluster the CTA works on
  const int rank = crank - seg * split;         // rank inside the segment's CTA group
  const int base_rank = seg * split;            // cluster rank of the group's first CTA (owner of the candidate buffer)
  const int n_raw = (cid - p.cl_base[level]) * (kCluster / split) + seg;
  const bool active = n_raw < p.num_images;     // the last cluster of a level may have idle groups: they join every
  const int n = active ? n_raw : 0;             //   cluster.sync with an empty slice and leave before the decode
  const osr_rpn_level_t& L = p.lv[level];
  const int S = (int)L.num_anchors;
  const int k = p.ktop[level];
  const int tid = threadIdx.x, lane = tid & 31;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* cand = reinterpret_cast<unsigned long long*>(smem_raw);
  uint32_t* keys = reinterpret_cast<uint32_t*>(cand + p.kpad);
  uint32_t* hist = keys + p.slice_cap;  // [2][256]
  uint32_t* total = hist + 2 * kBins;   // [256]
  uint32_t* misc = total + kBins;       // [64]

  // ---- slice of this CTA -------------------------------------------------------------------------
  const int per = osr::round_up(osr::ceil_div(S, split), 4);
  const int begin = min(S, rank * per);
  const int len = active ? min(S, begin + per) - begin : 0;

  {
    const float* sp = L.scores + (int64_t)n * L.score_stride_n;
    if (L.score_stride_a == 1 && ((reinterpret_cast<uintptr_t>(sp + begin) & 15) == 0)) {
      const int nv = len >> 2;
      const float4* v = reinterpret_cast<const float4*>(sp + begin);
      uint4* kv = reinterpret_cast<uint4*>(keys);
      for (int i = tid; i < nv; i += kThreads) {
        float4 f = __ldg(v + i);
        kv[i] = make_uint4(score_to_key(f.x), score_to_key(f.y), score_to_key(f.z), score_to_key(f.w));
      }
      for (int i = (nv << 2) + tid; i < len; i += kThreads) keys[i] = score_to_key(__ldg(sp + begin + i));
    } else {
      for (int i = tid; i < len; i += kThreads)
        keys[i] = score_to_key(__ldg(sp + (int64_t)(begin + i) * L.score_stride_a));
    }
  }
  __syncthreads();

  // ---- radix select of the k-th largest key ------------------------------------------------------------
  uint32_t T = 0;          // threshold key: selected = key > T, plus `remaining` lowest-index keys == T
  uint32_t remaining = 0;  // ties to take
  if (k < S) {
    uint32_t prefix = 0, mask = 0;
    remaining = (uint32_t)k;
    const int len_pad = osr::round_up(len, 32);
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      uint32_t* h = hist + (pass & 1) * kBins;
      for (int i = tid; i < kBins; i += kThreads) h[i] = 0;
      __syncthreads();
      for (int i = tid; i < len_pad; i += kThreads) {
        uint32_t bin = 0xffffffffu;
        if (i < len) {
          uint32_t key = keys[i];
          if ((key & mask) == prefix) bin = (key >> shift) & 0xffu;
        }
        if (pass == 0) {
          // warp-aggregated shared-memory histogram: scores in (0,1) put most keys in 1-2 top-byte bins.  Only for the top
          // byte - the lower bytes are spread over many bins, where match_any serialises (9 us vs 2.7 us for the pass)
          uint32_t peers = __match_any_sync(0xffffffffu, bin);
          if (bin != 0xffffffffu && lane == (__ffs(peers) - 1)) atomicAdd(&h[bin], (uint32_t)__popc(peers));
        } else if (bin != 0xffffffffu) {
          atomicAdd(&h[bin], 1u);
        }
      }
      cluster.sync();
      if (tid < kBins) {
        uint32_t s = 0;
        for (int r = 0; r < split; ++r) s += cluster.map_shared_rank(h, base_rank + r)[tid];
        total[tid] = s;
      }
      __syncthreads();
      if (tid < 32) {
        // lane j owns bins [255-8j-7, 255-8j], scanned from the top
        uint32_t c[8], blk = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          c[j] = total[255 - (lane * 8 + j)];
          blk += c[j];
        }
        uint32_t inc = blk;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        const uint32_t exc = inc - blk;
        if (exc < remaining && remaining <= inc) {  // exactly one lane
          uint32_t run = exc;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (run < remaining && remaining <= run + c[j]) {
              misc[0] = 255 - (lane * 8 + j);
              misc[1] = remaining - run;
            }
            run += c[j];
          }
        }
      }
      __syncthreads();
      prefix |= misc[0] << shift;
      mask |= 0xffu << shift;
      remaining = misc[1];
      __syncthreads();
    }
    T = prefix;
  }

  // ---- collect: ordered counts of (key > T) and (key == T) in this slice ---------------------------------
  int chunk = osr::ceil_div(len, kThreads) | 1;  // odd stride => conflict-free strided smem reads
  const int i0 = min(len, tid * chunk), i1 = min(len, i0 + chunk);
  uint32_t my = 0;
  for (int i = i0; i < i1; ++i) {
    uint32_t key = keys[i];
    my += (key > T) ? 1u : 0u;
    my += (key == T) ? 0x10000u : 0u;
  }
  uint32_t cta_total;
  uint32_t pre = block_exclusive_scan(my, misc + 8, &cta_total);
  if (tid == 0) {
    misc[2] = cta_total & 0xffffu;
    misc[3] = cta_total >> 16;
  }
  cluster.sync();
  uint32_t base_gt = 0, base_eq = 0, total_gt = 0;
  for (int r = 0; r < split; ++r) {
    const uint32_t* rm = cluster.map_shared_rank(misc, base_rank + r);
    uint32_t g = rm[2], e = rm[3];
    if (r < rank) {
      base_gt += g;
      base_eq += e;
    }
    total_gt += g;
  }
  {
    unsigned long long* cand0 = cluster.map_shared_rank(cand, base_rank);
    uint32_t pg = base_gt + (pre & 0xffffu), pe = base_eq + (pre >> 16);
    for (int i = i0; i < i1; ++i) {
      uint32_t key = keys[i];
      unsigned long long e = ((unsigned long long)key << 32) | (unsigned long long)(0xffffffffu - (uint32_t)(begin + i));
      if (key > T) {
        cand0[pg++] = e;
      } else if (key == T) {
        if (pe < remaining) cand0[total_gt + pe] = e;
        ++pe;
      }
    }
  }
  cluster.sync();
  if (rank != 0 || !active) return;

  // ---- rank 0 of the group: sort the k survivors -----------------------------------------------------------------
  int kp = 32;
  while (kp < k) kp <<= 1;
  for (int i = k + tid; i < kp; i += kThreads) cand[i] = 0ull;
  // Bitonic network, two steps per pass: the steps with strides 2h and h exchange inside groups {b, b+h, b+2h, b+3h}, so a
  // thread that owns such a group does both in registers - 33 passes over shared memory (4 loads + 4 stores per thread, one
  // barrier each) instead of 66 (25 us -> 19 us for 2048 candidates, measured with %globaltimer stamps).  A group lies inside one `size`-aligned block, so
  // its four elements share the sort direction.
  auto cx = [](unsigned long long& a, unsigned long long& b, bool desc) {
    if (desc ? (a < b) : (a > b)) {
      const unsigned long long t = a;
      a = b;
      b = t;
    }
  };
  for (int size = 2; size <= kp; size <<= 1) {
    int stride = size >> 1;
    for (; stride >= 2; stride >>= 2) {
      const int h = stride >> 1;
      __syncthreads();
      for (int q = tid; q < (kp >> 2); q += kThreads) {
        const int b0 = ((q & ~(h - 1)) << 2) | (q & (h - 1));
        const bool desc = ((b0 & size) == 0);
        unsigned long long e0, e1, e2, e3;
        if (h == 1) {   // four consecutive elements: 16-byte accesses (8-byte ones at a 32-byte lane pitch conflict 8-way)
          const ulonglong2 v0 = reinterpret_cast<const ulonglong2*>(cand)[2 * q], v1 = reinterpret_cast<const ulonglong2*>(cand)[2 * q + 1];
          e0 = v0.x; e1 = v0.y; e2 = v1.x; e3 = v1.y;
        } else {
          e0 = cand[b0]; e1 = cand[b0 + h]; e2 = cand[b0 + 2 * h]; e3 = cand[b0 + 3 * h];
        }
        cx(e0, e2, desc);
        cx(e1, e3, desc);
        cx(e0, e1, desc);
        cx(e2, e3, desc);
        if (h == 1) {
          reinterpret_cast<ulonglong2*>(cand)[2 * q] = make_ulonglong2(e0, e1);
          reinterpret_cast<ulonglong2*>(cand)[2 * q + 1] = make_ulonglong2(e2, e3);
        } else {
          cand[b0] = e0;
          cand[b0 + h] = e1;
          cand[b0 + 2 * h] = e2;
          cand[b0 + 3 * h] = e3;
        }
      }
    }
    if (stride == 1) {   // odd number of steps for this size: the last one on its own
      __syncthreads();
      for (int t = tid; t < (kp >> 1); t += kThreads) {
        const int lo = 2 * t;
        const bool desc = ((lo & size) == 0);
        ulonglong2 v = reinterpret_cast<const ulonglong2*>(cand)[t];
        if (desc ? (v.x < v.y) : (v.x > v.y)) reinterpret_cast<ulonglong2*>(cand)[t] = make_ulonglong2(v.y, v.x);
      }
    }
  }
  __syncthreads();

  // ---- decode, clip, filter, ordered compaction ---------------------------------------------------
  const float img_h = (float)p.image_hw[2 * n], img_w = (float)p.image_hw[2 * n + 1];
  const int64_t out_base = (int64_t)n * p.kmax + p.koff[level];
  const bool vec_delta = (L.delta_stride_c == 1) && ((L.delta_stride_a & 3) == 0) && ((L.delta_stride_n & 3) == 0) &&
                         ((reinterpret_cast<uintptr_t>(L.deltas) & 15) == 0);
  uint32_t kept = 0;
  uint32_t bad = 0;
  for (int j0 = 0; j0 < k; j0 += kThreads) {
    const int j = j0 + tid;
    bool valid = false;
    float x1 = 0, y1 = 0, x2 = 0, y2 = 0, score = 0;
    uint32_t idx = 0;
    if (j < k) {
      unsigned long long e = cand[j];
      idx = 0xffffffffu - (uint32_t)(e & 0xffffffffull);
      score = key_to_score((uint32_t)(e >> 32));
      const float* dp = L.deltas + (int64_t)n * L.delta_stride_n + (int64_t)idx * L.delta_stride_a;
      float d0, d1, d2, d3;
      if (vec_delta) {
        float4 d = __ldg(reinterpret_cast<const float4*>(dp));
        d0 = d.x; d1 = d.y; d2 = d.z; d3 = d.w;
      } else {
        d0 = __ldg(dp);
        d1 = __ldg(dp + L.delta_stride_c);
        d2 = __ldg(dp + 2 * L.delta_stride_c);
        d3 = __ldg(dp + 3 * L.delta_stride_c);
      }
      if (L.anchors != nullptr) {
        // Box2BoxTransformLinear(normalize_by_size=True).apply_deltas: every op rounded separately (no FMA)
        const float4 a = __ldg(reinterpret_cast<const float4*>(L.anchors) + idx);
        d0 = relu_keep_nan(d0); d1 = relu_keep_nan(d1); d2 = relu_keep_nan(d2); d3 = relu_keep_nan(d3);
        const float ctr_x = __fmul_rn(0.5f, __fadd_rn(a.x, a.z));
        const float ctr_y = __fmul_rn(0.5f, __fadd_rn(a.y, a.w));
        const float sw = __fsub_rn(a.z, a.x), sh = __fsub_rn(a.w, a.y);
        x1 = __fsub_rn(ctr_x, __fmul_rn(d0, sw));
        y1 = __fsub_rn(ctr_y, __fmul_rn(d1, sh));
        x2 = __fadd_rn(ctr_x, __fmul_rn(d2, sw));
        y2 = __fadd_rn(ctr_y, __fmul_rn(d3, sh));
      } else {  // `deltas` already holds decoded xyxy proposals (find_top_rpn_proposals signature)
        x1 = d0; y1 = d1; x2 = d2; y2 = d3;
      }
      const bool finite = isfinite(x1) && isfinite(y1) && isfinite(x2) && isfinite(y2) && isfinite(score);
      if (!finite) bad = 1;
      // Boxes.clip((h, w)) then Boxes.nonempty(threshold=min_box_size)   (find_top_proposals.py:105-110)
      x1 = fminf(fmaxf(x1, 0.f), img_w);
      y1 = fminf(fmaxf(y1, 0.f), img_h);
      x2 = fminf(fmaxf(x2, 0.f), img_w);
      y2 = fminf(fmaxf(y2, 0.f), img_h);
      valid = finite && (__fsub_rn(x2, x1) > p.min_box_size) && (__fsub_rn(y2, y1) > p.min_box_size);
    }
    uint32_t round_total;
    uint32_t pos = kept + block_exclusive_scan(valid ? 1u : 0u, misc + 8, &round_total);
    if (valid) {
      reinterpret_cast<float4*>(p.st_boxes)[out_base + pos] = make_float4(x1, y1, x2, y2);
      p.st_scores[out_base + pos] = score;
      p.st_index[out_base + pos] = (int32_t)idx;
    }
    kept += round_total;
  }
  bad = __syncthreads_or((int)bad);
  if (tid == 0) {
    p.counts[n * (p.num_levels + 2) + level] = (int32_t)kept;
    p.st_flags[n * p.num_levels + level] = bad ? 1 : 0;
  }
}

// Concatenate the per-level runs of each image (find_top_proposals.py:85-87 cat over levels + :108-110 filter).
// grid (N, kConcatY): every block derives the level prefix from the counts and moves a strided share of the rows.
constexpr int kConcatY = 16;
__global__ void __launch_bounds__(256) rpn_concat_kernel(const __grid_constant__ RpnParams p, float* out_boxes,
                                                          float* out_scores, int32_t* out_level,
                                                          int32_t* out_index) {
  const int n = blockIdx.x;
  const int L = p.num_levels;
  int32_t* cnt = p.counts + n * (L + 2);
  int off[OSR_MAX_LEVELS + 1];
  int flags = 0;
  off[0] = 0;
#pragma unroll
  for (int l = 0; l < OSR_MAX_LEVELS; ++l) {
    const int c = l < L ? cnt[l] : 0;
    off[l + 1] = off[l] + c;
    if (l < L) flags |= p.st_flags[n * L + l];
  }
  const int total = off[L];
  const int64_t base = (int64_t)n * p.kmax;
  for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < total; i += gridDim.y * blockDim.x) {
    int l = 0;
#pragma unroll
    for (int k = 1; k < OSR_MAX_LEVELS; ++k) l += (k < L && i >= off[k]) ? 1 : 0;
    const int64_t src = base + p.koff[l] + (i - off[l]);
    reinterpret_cast<float4*>(out_boxes)[base + i] = reinterpret_cast<const float4*>(p.st_boxes)[src];
    out_scores[base + i] = p.st_scores[src];
    out_index[base + i] = p.st_index[src];
    out_level[base + i] = l;
  }
  if (blockIdx.y == 0 && threadIdx.x == 0) {
    // written last in program order by this thread only; other blocks never read cnt[L], cnt[L+1]
    cnt[L] = total;
    cnt[L + 1] = flags;
  }
}

int fill_params(RpnParams& p, const osr_rpn_level_t* h_levels, int num_levels, int num_images, int pre_nms_topk) {
  if (!h_levels || num_levels <= 0 || num_levels > OSR_MAX_LEVELS)
    return osr::fail_arg(OSR_E_ARG, "rpn: num_levels=%d outside [1,%d]", num_levels, OSR_MAX_LEVELS);
  if (num_images < 0 || pre_nms_topk <= 0) return osr::fail_arg(OSR_E_ARG, "rpn: bad num_images / pre_nms_topk");
  p.num_levels = num_levels;
  p.num_images = num_images;
  int off = 0, kmaxlvl = 1, cap = 4;
  for (int l = 0; l < num_levels; ++l) {
    p.lv[l] = h_levels[l];
    const int64_t S = h_levels[l].num_anchors;
    if (S <= 0 || S > (int64_t)1 << 24) return osr::fail_arg(OSR_E_SHAPE, "rpn: level %d num_anchors=%lld unsupported", l, (long long)S);
    p.ktop[l] = (int)(S < pre_nms_topk ? S : pre_nms_topk);
    p.koff[l] = off;
    off += p.ktop[l];
    kmaxlvl = p.ktop[l] > kmaxlvl ? p.ktop[l] : kmaxlvl;
    const int per = osr::round_up(osr::ceil_div((int)S, kCluster), 4);
    cap = per > cap ? per : cap;
  }
  p.koff[num_levels] = off;
  p.kmax = off;
  p.kpad = osr::next_pow2(kmaxlvl < 32 ? 32 : kmaxlvl);
  p.slice_cap = cap;
  // cluster packing: the fewest CTAs per segment whose slice still fits the carve-up sized for the largest level
  int base = 0;
  for (int l = 0; l < num_levels; ++l) {
    const int S = (int)h_levels[l].num_anchors;
    int split = kCluster;
    while (split > 1 && osr::round_up(osr::ceil_div(S, split / 2), 4) <= cap) split /= 2;
    p.split[l] = split;
    p.cl_base[l] = base;
    base += osr::ceil_div(num_images, kCluster / split);
  }
  p.cl_base[num_levels] = base;
  return 0;
}

size_t select_smem_bytes(const RpnParams& p) {
  return (size_t)p.kpad * 8 + (size_t)p.slice_cap * 4 + (2 * kBins + kBins + 64) * 4;
}

}  // namespace

extern "C" {

int64_t osr_rpn_kmax(const osr_rpn_level_t* h_levels, int num_levels, int pre_nms_topk) {
  RpnParams p;
  if (fill_params(p, h_levels, num_levels, 1, pre_nms_topk)) return -1;
  return p.kmax;
}

size_t osr_rpn_select_decode_workspace(const osr_rpn_level_t* h_levels, int num_levels, int num_images,
                                       int pre_nms_topk) {
  RpnParams p;
  if (fill_params(p, h_levels, num_levels, num_images, pre_nms_topk)) return 0;
  const size_t nk = (size_t)num_images * p.kmax;
  return osr::align256(nk * 16) + osr::align256(nk * 4) + osr::align256(nk * 4) +
         osr::align256((size_t)num_images * num_levels * 4) + 256;
}

int osr_rpn_select_decode(const osr_rpn_level_t* h_levels, int num_levels, int num_images, int pre_nms_topk,
                          float min_box_size, const int32_t* image_hw, float* out_boxes, float* out_scores,
                          int32_t* out_level, int32_t* out_index, int32_t* out_counts, void* workspace,
                          size_t workspace_bytes, void* stream) {
  osr::DeviceGuard device_guard(out_counts);
  RpnParams p;
  int rc = fill_params(p, h_levels, num_levels, num_images, pre_nms_topk);
  if (rc) return rc;
  if (num_images == 0) return 0;
  if (!image_hw || !out_boxes || !out_scores || !out_level || !out_index || !out_counts || !workspace)
    return osr::fail_arg(OSR_E_ARG, "rpn: null pointer argument");
  for (int l = 0; l < num_levels; ++l)
    if (!p.lv[l].deltas || !p.lv[l].scores || (reinterpret_cast<uintptr_t>(p.lv[l].anchors) & 15))
      return osr::fail_arg(OSR_E_ARG, "rpn: level %d has a null pointer or anchors not 16-byte aligned", l);
  if ((reinterpret_cast<uintptr_t>(out_boxes) & 15) || (reinterpret_cast<uintptr_t>(workspace) & 255))
    return osr::fail_arg(OSR_E_ARG, "rpn: out_boxes must be 16-byte and workspace 256-byte aligned");
  if (workspace_bytes < osr_rpn_select_decode_workspace(h_levels, num_levels, num_images, pre_nms_topk))
    return osr::fail_arg(OSR_E_WORKSPACE, "rpn: workspace too small");
  const size_t smem = select_smem_bytes(p);
  if (smem > 227 * 1024)
    return osr::fail_arg(OSR_E_SHAPE, "rpn: level too large for the shared-memory carve-up (%zu B needed)", smem);

  const size_t nk = (size_t)num_images * p.kmax;
  unsigned char* w = static_cast<unsigned char*>(workspace);
  p.st_boxes = reinterpret_cast<float*>(w);
  w += osr::align256(nk * 16);
  p.st_scores = reinterpret_cast<float*>(w);
  w += osr::align256(nk * 4);
  p.st_index = reinterpret_cast<int32_t*>(w);
  w += osr::align256(nk * 4);
  p.st_flags = reinterpret_cast<int32_t*>(w);
  p.counts = out_counts;
  p.image_hw = image_hw;
  p.min_box_size = min_box_size;

  cudaStream_t s = static_cast<cudaStream_t>(stream);
  OSR_CUDA_CHECK(cudaFuncSetAttribute(rpn_select_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  rpn_select_decode_kernel<<<p.cl_base[num_levels] * kCluster, kThreads, smem, s>>>(p);
  OSR_LAUNCH_CHECK();
  rpn_concat_kernel<<<dim3(num_images, kConcatY), 256, 0, s>>>(p, out_boxes, out_scores, out_level, out_index);
  OSR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
