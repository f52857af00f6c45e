// Labelled RoI sampling for the ROI-head glue: detectron2's subsample_labels
// (called per image from ROIHeads._sample_proposals, osrcnn_roi_heads.py:136-175 / :203-206) for ALL images in one launch,
// fused with the gather of every sampled field (osrcnn_roi_heads.py:208-226: proposal_boxes, objectness_logits,
// gt_classes, ious and the matched ground-truth row).
//
// The reference draws two torch.randperm permutations per image (positives, negatives) and keeps a prefix of each.  Here
// every row carries one random key (one torch.rand draw for the whole batch) and a kind keeps the rows with the SMALLEST
// keys, emitted in ascending (key, row) order: the subset positive[perm[:num_pos]] the reference would take if its
// permutation were the stable arg-sort of the kind's keys - a uniformly random subset in uniformly random order (parity
// tests inject exactly that permutation into the reference's subsample_labels).
//
// One CTA per image:
//   count      positives (class != -1, != background) and negatives (class == background) -> quotas
//              num_pos = min(#pos, num_pos_max), num_neg = min(#neg, num_samples - num_pos)
//   select     per kind whose quota is smaller than its population: 4-pass 8-bit MSD radix select of the quota-th smallest
//              key (shared-memory histograms), ties at the threshold taken by lowest row index
//   compact    selected rows -> (kind << 63 | key << 31 | row) in shared memory (ordered block scan for the ties)
//   sort       one bitonic network over the <= num_samples survivors: positives first, each kind by ascending (key, row)
//   gather     boxes / logits / classes / IoUs / matched ground-truth rows of the survivors, coalesced over the sample slot
#include "osr_common.cuh"

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kBins = 256;
constexpr int kMaxSamples = 4096;

struct SampleParams {
  const int64_t* labels;
  const float* keys;
  const int32_t* box_off;
  const int32_t* box_cnt;
  int cnt_stride;
  int num_samples, num_pos_max, kp, cap;
  int64_t bg;
  const float* boxes;
  const float* logits;
  const float* ious;
  const int32_t* midx;
  const int32_t* gt_off;
  int32_t* out_index;
  int32_t* out_count;
  float* out_boxes;
  float* out_logits;
  int64_t* out_classes;
  float* out_ious;
  int64_t* out_gt;
  float* out_rois;
};

// order-preserving map float -> uint32 (ascending); NaN keys sort last
__device__ __forceinline__ uint32_t key_bits(float f) {
  const uint32_t b = __float_as_uint(f);
  if ((b & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;
  if (b == 0x80000000u) return 0x80000000u;   // -0.0 == +0.0, as a float comparison (and torch.argsort) has it
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// 1 = positive, 2 = negative, 0 = ignored
__device__ __forceinline__ int row_kind(int64_t label, int64_t bg) { return label == bg ? 2 : (label == -1 ? 0 : 1); }

__device__ __forceinline__ unsigned long long block_exclusive_scan64(unsigned long long v, unsigned long long* warp_sums,
                                                                     unsigned long long* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const unsigned long long w = (lane < kWarps) ? warp_sums[lane] : 0ull;
    unsigned long long winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < kWarps) warp_sums[lane] = winc - w;
    if (lane == kWarps - 1) warp_sums[kWarps] = winc;
  }
  __syncthreads();
  const unsigned long long res = warp_sums[warp] + inc - v;
  *total = warp_sums[kWarps];
  __syncthreads();
  return res;
}

// kCache: the rows' (kind, key) pairs are read from global memory ONCE into shared memory and every later pass runs on the
// copy (an image's ~7 k rows are visited by ~7 passes; from L2 each pass is a chain of dependent ~1 us loads per thread).
// Images with more rows than the carve-up holds (host hint max_boxes_per_image) use the instantiation that re-reads them.
template <bool kCache>
__global__ void __launch_bounds__(kThreads) sample_rois_kernel(const __grid_constant__ SampleParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* cand = reinterpret_cast<unsigned long long*>(smem_raw);   // [kp]
  uint32_t* skey = reinterpret_cast<uint32_t*>(cand + p.kp);                    // [cap]   (kCache)
  unsigned char* skind = reinterpret_cast<unsigned char*>(skey + p.cap);        // [cap]
  __shared__ uint32_t hist[2][kBins];
  __shared__ unsigned long long wsum[kWarps + 1];
  __shared__ uint32_t misc[2][2];
  const int n = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b0 = p.box_off[n];
  const int P = (p.box_cnt ? p.box_cnt[n * p.cnt_stride] : p.box_off[n + 1] - b0);
  const int64_t* lab = p.labels + b0;
  const float* key = p.keys + b0;
  auto kind_of = [&](int i) -> int { return kCache ? (int)skind[i] : row_kind(__ldg(lab + i), p.bg); };
  auto key_of = [&](int i) -> uint32_t { return kCache ? skey[i] : key_bits(__ldg(key + i)); };

  // ---- populations -> quotas (and the shared-memory copy of the rows) ---------------------------------------------
  unsigned long long my = 0;
#pragma unroll 4
  for (int i = tid; i < P; i += kThreads) {
    const int kd = row_kind(__ldg(lab + i), p.bg);
    if (kCache) {
      skind[i] = (unsigned char)kd;
      skey[i] = key_bits(__ldg(key + i));
    }
    my += (kd == 1) ? 1ull : 0ull;
    my += (kd == 2) ? (1ull << 32) : 0ull;
  }
  for (int i = tid; i < p.kp; i += kThreads) cand[i] = ~0ull;   // padding sorts last
  unsigned long long tot;
  block_exclusive_scan64(my, wsum, &tot);   // (its barriers also publish skey / skind / cand)
  const int pop[2] = {(int)(tot & 0xffffffffull), (int)(tot >> 32)};
  int take[2];
  take[0] = min(pop[0], p.num_pos_max);
  take[1] = min(pop[1], p.num_samples - take[0]);
  const int taken = take[0] + take[1];

  // ---- thresholds: the take-th smallest key of each kind whose quota is below its population (both kinds per pass) -------
  // selected rows of kind k: key < T[k], plus the `rem[k]` lowest rows with key == T[k]
  uint32_t T[2] = {0xffffffffu, 0xffffffffu}, rem[2] = {0xffffffffu, 0xffffffffu};
  const bool sel0 = take[0] > 0 && take[0] < pop[0], sel1 = take[1] > 0 && take[1] < pop[1];
  if (sel0 || sel1) {   // CTA-uniform
    uint32_t prefix[2] = {0, 0}, remaining[2] = {(uint32_t)take[0], (uint32_t)take[1]}, mask = 0;
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = tid; i < 2 * kBins; i += kThreads) (&hist[0][0])[i] = 0;
      __syncthreads();
      for (int i = tid; i < P; i += kThreads) {
        const int kd = kind_of(i);
        if (kd == 0 || !(kd == 1 ? sel0 : sel1)) continue;
        const uint32_t k = key_of(i);
        if ((k & mask) == prefix[kd - 1]) atomicAdd(&hist[kd - 1][(k >> shift) & 0xffu], 1u);
      }
      __syncthreads();
      if (warp < 2 && (warp == 0 ? sel0 : sel1)) {   // warp w resolves kind w + 1; lane j owns bins [8j, 8j + 7], scanned from the bottom
        uint32_t c[8], blk = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          c[j] = hist[warp][lane * 8 + j];
          blk += c[j];
        }
        uint32_t inc = blk;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += t;
        }
        const uint32_t exc = inc - blk, want = remaining[warp];
        if (exc < want && want <= inc) {   // exactly one lane
          uint32_t run = exc;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (run < want && want <= run + c[j]) {
              misc[warp][0] = lane * 8 + j;
              misc[warp][1] = want - run;
            }
            run += c[j];
          }
        }
      }
      __syncthreads();
      if (sel0) { prefix[0] |= misc[0][0] << shift; remaining[0] = misc[0][1]; }
      if (sel1) { prefix[1] |= misc[1][0] << shift; remaining[1] = misc[1][1]; }
      mask |= 0xffu << shift;
    }
    if (sel0) { T[0] = prefix[0]; rem[0] = remaining[0]; }
    if (sel1) { T[1] = prefix[1]; rem[1] = remaining[1]; }
  }

  // ---- ordered compaction of both kinds (the block scans rank the ties by row index; rows below the threshold may land in
  //      any order - the sort fixes it) ----------------------------------------------------------------------------------
  const int chunk = osr::ceil_div(max(P, 1), kThreads);
  const int i0 = min(P, tid * chunk), i1 = min(P, i0 + chunk);
  unsigned long long mine[2] = {0ull, 0ull};
  for (int i = i0; i < i1; ++i) {
    const int kd = kind_of(i);
    if (kd == 0 || take[kd - 1] == 0) continue;
    const uint32_t k = key_of(i);
    mine[kd - 1] += (k < T[kd - 1]) ? 1ull : 0ull;
    mine[kd - 1] += (k == T[kd - 1]) ? (1ull << 32) : 0ull;
  }
  unsigned long long all[2], pre[2];
  pre[0] = block_exclusive_scan64(mine[0], wsum, &all[0]);
  pre[1] = block_exclusive_scan64(mine[1], wsum, &all[1]);
  uint32_t pl[2] = {(uint32_t)(pre[0] & 0xffffffffull), (uint32_t)(pre[1] & 0xffffffffull)};
  uint32_t pe[2] = {(uint32_t)(pre[0] >> 32), (uint32_t)(pre[1] >> 32)};
  const uint32_t total_lt[2] = {(uint32_t)(all[0] & 0xffffffffull), (uint32_t)(all[1] & 0xffffffffull)};
  const int base[2] = {0, take[0]};
  for (int i = i0; i < i1; ++i) {
    const int kd = kind_of(i);
    if (kd == 0 || take[kd - 1] == 0) continue;
    const int q = kd - 1;
    const uint32_t k = key_of(i);
    const unsigned long long e = ((unsigned long long)q << 63) | ((unsigned long long)k << 31) | (unsigned long long)i;
    if (k < T[q]) {
      cand[base[q] + pl[q]++] = e;
    } else if (k == T[q]) {
      if (pe[q] < rem[q]) cand[base[q] + total_lt[q] + pe[q]] = e;
      ++pe[q];
    }
  }
  __syncthreads();

  // ---- bitonic sort (ascending) of the survivors: positives first, each kind by (key, row).  Two steps per pass (strides
  //      2h and h exchange inside groups {b, b+h, b+2h, b+3h}: a thread does both in registers) -----------------------------
  int kp = 4;
  while (kp < taken) kp <<= 1;
  auto cx = [](unsigned long long& a, unsigned long long& b, bool asc) {
    if (asc ? (a > b) : (a < b)) {
      const unsigned long long t = a;
      a = b;
      b = t;
    }
  };
  for (int size = 2; size <= kp; size <<= 1) {
    int stride = size >> 1;
    for (; stride >= 2; stride >>= 2) {
      const int h = stride >> 1;
      for (int q = tid; q < (kp >> 2); q += kThreads) {
        const int g0 = ((q & ~(h - 1)) << 2) | (q & (h - 1));
        const bool asc = ((g0 & size) == 0);
        unsigned long long e0 = cand[g0], e1 = cand[g0 + h], e2 = cand[g0 + 2 * h], e3 = cand[g0 + 3 * h];
        cx(e0, e2, asc);
        cx(e1, e3, asc);
        cx(e0, e1, asc);
        cx(e2, e3, asc);
        cand[g0] = e0; cand[g0 + h] = e1; cand[g0 + 2 * h] = e2; cand[g0 + 3 * h] = e3;
      }
      __syncthreads();
    }
    if (stride == 1) {   // odd number of steps for this size: the last one on its own
      for (int t = tid; t < (kp >> 1); t += kThreads) {
        const int lo = 2 * t;
        const bool asc = ((lo & size) == 0);
        unsigned long long a = cand[lo], b = cand[lo + 1];
        if (asc ? (a > b) : (a < b)) {
          cand[lo] = b;
          cand[lo + 1] = a;
        }
      }
      __syncthreads();
    }
  }

  // ---- gather the sampled fields --------------------------------------------------------------------------------
  const int g0 = p.gt_off ? p.gt_off[n] : 0;
  for (int j = tid; j < p.num_samples; j += kThreads) {
    const int64_t o = (int64_t)n * p.num_samples + j;
    if (j >= taken) {
      p.out_index[o] = -1;
      continue;
    }
    const int i = (int)(cand[j] & 0x7fffffffull);
    p.out_index[o] = i;
    if (p.out_boxes) reinterpret_cast<float4*>(p.out_boxes)[o] = __ldg(reinterpret_cast<const float4*>(p.boxes) + b0 + i);
    if (p.out_logits) p.out_logits[o] = __ldg(p.logits + b0 + i);
    if (p.out_classes) p.out_classes[o] = __ldg(lab + i);
    if (p.out_ious) p.out_ious[o] = __ldg(p.ious + b0 + i);
    if (p.out_gt) p.out_gt[o] = (int64_t)g0 + (int64_t)__ldg(p.midx + b0 + i);
    if (p.out_rois) {   // (image, x1, y1, x2, y2): the row format ROIAlign consumes (osr_roi_align_fwd `rois`)
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.boxes) + b0 + i);
      float* r = p.out_rois + o * 5;
      r[0] = (float)n; r[1] = b.x; r[2] = b.y; r[3] = b.z; r[4] = b.w;
    }
  }
  if (tid == 0) {
    p.out_count[2 * n] = take[0];
    p.out_count[2 * n + 1] = taken;
  }
}

}  // namespace

extern "C" int osr_sample_rois(const int64_t* labels, const float* keys, const int32_t* box_offsets, const int32_t* box_counts,
                               int box_counts_stride, int num_images, int max_boxes_per_image, int num_samples,
                               int num_pos_max, int64_t background_label, const float* boxes, const float* logits, const float* ious,
                               const int32_t* matched_idx, const int32_t* gt_offsets, int32_t* out_index,
                               int32_t* out_count, float* out_boxes, float* out_logits, int64_t* out_classes,
                               float* out_ious, int64_t* out_gt, float* out_rois, void* stream) {
  osr::DeviceGuard device_guard(out_index);
  if (num_images < 0 || num_samples <= 0 || num_pos_max < 0 || num_pos_max > num_samples)
    return osr::fail_arg(OSR_E_ARG, "sample_rois: bad num_images / num_samples / num_pos_max");
  if (num_samples > kMaxSamples) return osr::fail_arg(OSR_E_ARG, "sample_rois: num_samples=%d > %d", num_samples, kMaxSamples);
  if (num_images == 0) return 0;
  if (!labels || !keys || !box_offsets || !out_index || !out_count)
    return osr::fail_arg(OSR_E_ARG, "sample_rois: null pointer argument");
  if (((out_boxes || out_rois) && !boxes) || (out_logits && !logits) || (out_ious && !ious) || (out_gt && !matched_idx))
    return osr::fail_arg(OSR_E_ARG, "sample_rois: an output field is requested without its source");
  if ((boxes && (reinterpret_cast<uintptr_t>(boxes) & 15)) || (out_boxes && (reinterpret_cast<uintptr_t>(out_boxes) & 15)))
    return osr::fail_arg(OSR_E_ARG, "sample_rois: box arrays must be 16-byte aligned");
  SampleParams p{labels, keys, box_offsets, box_counts, box_counts_stride, num_samples, num_pos_max, 0, 0, background_label,
                 boxes, logits, ious, matched_idx, gt_offsets, out_index, out_count, out_boxes, out_logits, out_classes,
                 out_ious, out_gt, out_rois};
  p.kp = osr::next_pow2(num_samples < 4 ? 4 : num_samples);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t smem_sort = (size_t)p.kp * sizeof(unsigned long long);
  const size_t cap = (size_t)osr::round_up(max_boxes_per_image > 0 ? max_boxes_per_image : 0, 16);
  if (max_boxes_per_image > 0 && smem_sort + cap * 5 <= 200 * 1024) {   // rows cached in shared memory
    p.cap = (int)cap;
    const size_t smem = smem_sort + cap * 5;
    OSR_CUDA_CHECK(cudaFuncSetAttribute(sample_rois_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sample_rois_kernel<true><<<num_images, kThreads, smem, s>>>(p);
  } else {
    sample_rois_kernel<false><<<num_images, kThreads, smem_sort, s>>>(p);
  }
  OSR_LAUNCH_CHECK();
  return 0;
}
