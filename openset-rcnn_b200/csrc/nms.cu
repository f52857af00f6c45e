// Segmented greedy NMS for sm_100a: all (image x category) segments of a batch in three launches, no host sync.
// Replaces detectron2.layers.batched_nms -> torchvision nms_kernel_impl + gather_keep_from_mask, called per image
// in a Python loop at osrcnn_fast_rcnn.py:135, softmax_classifier.py:93/:154 (and find_top_proposals.py:112).
//
//   nms_sort_kernel   one CTA per segment: stable descending sort of (score key << 32 | ~index) in shared memory
//                     (bitonic), gathers the boxes into sorted order.  Skipped for presorted segments.
//   nms_mask_kernel   upper-triangular 64x64 tiles: the column tile's boxes staged in shared memory, each thread owns
//                     one row box and emits a uint64 of suppressed columns.  IoU arithmetic is the exact sequence of
//                     torchvision's CUDA kernel (SASS-verified):
//                         inter / (fma(bw, bh, rn(aw*ah)) - inter) > (float)thr      a = row (higher score) box
//   nms_sweep_kernel  one CTA per segment, 64 rows at a time: one thread resolves the 64x64 diagonal word chain,
//                     then all threads OR the kept rows' mask words into the `removed` bit-vector held in shared
//                     memory (each kept row's mask is read exactly once, coalesced); kept indices are written in
//                     score order with a popcount prefix.
#include "osr_common.cuh"

namespace {

constexpr int kSortThreads = 1024;
constexpr int kMaxSortLen = 16384;
constexpr int kMaxLen = 65536;
constexpr int kSweepThreads = 512;

struct NmsParams {
  const float* boxes;
  const float* scores;
  const int32_t* seg_begin;
  const int32_t* seg_len;
  int num_segments, max_len, wpr;  // wpr = mask words per row = ceil(max_len / 64)
  float thr;
  int presorted;
  int64_t* keep_idx;
  int32_t* keep_counts;
  uint8_t* keep_mask;
  // workspace
  float4* sorted_boxes;   // (T) boxes in sorted order (segment-relative positions)   [unsorted path only]
  int32_t* order;         // (T) sorted position -> index relative to the segment start [unsorted path only]
  unsigned long long* mask;  // (S, max_len, wpr)
};

__device__ __forceinline__ uint32_t score_key(float f) {
  uint32_t b = __float_as_uint(f);
  if ((b & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;  // NaN first, as torch.sort(descending=True)
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(kSortThreads) nms_sort_kernel(const __grid_constant__ NmsParams p) {
  extern __shared__ __align__(16) unsigned long long skey[];
  const int s = blockIdx.x;
  const int begin = p.seg_begin[s], len = min(p.seg_len[s], p.max_len);
  int kp = 32;
  while (kp < len) kp <<= 1;
  const int tid = threadIdx.x;
  for (int i = tid; i < kp; i += kSortThreads) {
    unsigned long long e = 0ull;
    if (i < len) e = ((unsigned long long)score_key(__ldg(p.scores + begin + i)) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)i);
    skey[i] = e;
  }
  for (int size = 2; size <= kp; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = tid; t < (kp >> 1); t += kSortThreads) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        unsigned long long a = skey[lo], b = skey[hi];
        if (desc ? (a < b) : (a > b)) {
          skey[lo] = b;
          skey[hi] = a;
        }
      }
    }
  }
  __syncthreads();
  const float4* bx = reinterpret_cast<const float4*>(p.boxes);
  for (int i = tid; i < len; i += kSortThreads) {
    const uint32_t idx = 0xffffffffu - (uint32_t)(skey[i] & 0xffffffffull);
    p.order[begin + i] = (int32_t)idx;
    p.sorted_boxes[begin + i] = __ldg(bx + begin + idx);
  }
}

__global__ void __launch_bounds__(64) nms_mask_kernel(const __grid_constant__ NmsParams p) {
  const int s = blockIdx.z;
  const int len = min(p.seg_len[s], p.max_len);
  const int row_start = blockIdx.y, col_start = blockIdx.x;
  if (row_start > col_start) return;
  if (row_start * 64 >= len || col_start * 64 >= len) return;
  const int begin = p.seg_begin[s];
  const float4* bx = p.presorted ? reinterpret_cast<const float4*>(p.boxes) : p.sorted_boxes;
  const int row_size = min(len - row_start * 64, 64);
  const int col_size = min(len - col_start * 64, 64);
  __shared__ float4 cb[64];
  const int tid = threadIdx.x;
  if (tid < col_size) cb[tid] = __ldg(bx + begin + col_start * 64 + tid);
  __syncthreads();
  if (tid < row_size) {
    const int row = row_start * 64 + tid;
    const float4 a = __ldg(bx + begin + row);
    const float sa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    unsigned long long t = 0;
    const int start = (row_start == col_start) ? tid + 1 : 0;
    const bool neg_thr = !(p.thr >= 0.f);
    for (int i = start; i < col_size; ++i) {
      const float4 b = cb[i];
      const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
      const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
      const float w = fmaxf(__fsub_rn(right, left), 0.f), h = fmaxf(__fsub_rn(bottom, top), 0.f);
      const float inter = __fmul_rn(w, h);
      // disjoint boxes (inter == 0: most pairs): 0 / uni is +-0 or NaN, never above a threshold >= 0 - skip the IEEE division
      if (inter > 0.f || neg_thr) {
        const float uni = __fsub_rn(__fmaf_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y), sa), inter);
        if (__fdiv_rn(inter, uni) > p.thr) t |= 1ull << i;
      }
    }
    p.mask[((int64_t)s * p.max_len + row) * p.wpr + col_start] = t;
  }
}

// Suppression sweep of one segment (boxes already in score order).  64 rows are resolved per step from the DIAGONAL mask
// words (greedy NMS inside the block is a 64-bit recurrence: thread 0 walks only the KEPT rows with ffs), then the kept
// rows' mask words are OR-ed into the `removed` bit-vector of the later column blocks with a WARP-BALLOT style reduction:
// a warp owns one column word at a time, lane l loads the words of rows l and l + 32 (if kept) - 64 independent loads in
// flight per warp instead of a dependent load-OR chain per thread - and __reduce_or_sync folds them.  Same keep set and
// order as the serial sweep, bit for bit.
__global__ void __launch_bounds__(kSweepThreads) nms_sweep_kernel(const __grid_constant__ NmsParams p) {
  extern __shared__ __align__(16) unsigned long long removed[];  // [wpr]
  __shared__ unsigned long long diag[64];
  __shared__ unsigned long long s_kept;
  __shared__ int s_count;
  const int s = blockIdx.x;
  const int begin = p.seg_begin[s], len = min(p.seg_len[s], p.max_len);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kSweepWarps = kSweepThreads / 32;
  const int nblk = (len + 63) >> 6;
  for (int w = tid; w < nblk; w += kSweepThreads) removed[w] = 0ull;
  if (tid == 0) s_count = 0;
  if (p.keep_mask)
    for (int i = tid; i < len; i += kSweepThreads) p.keep_mask[begin + i] = 0;
  __syncthreads();
  const unsigned long long* mrow = p.mask + (int64_t)s * p.max_len * p.wpr;
  for (int rb = 0; rb < nblk; ++rb) {
    const int rows = min(64, len - rb * 64);
    if (tid < rows) diag[tid] = mrow[(int64_t)(rb * 64 + tid) * p.wpr + rb];
    __syncthreads();
    if (tid == 0) {
      const unsigned long long valid = rows == 64 ? ~0ull : ((1ull << rows) - 1ull);
      unsigned long long cur = removed[rb], kept = 0ull;
      unsigned long long avail = ~cur & valid;
      while (avail) {                       // visits kept rows only
        const int i = __ffsll((long long)avail) - 1;
        kept |= 1ull << i;
        cur |= diag[i];
        const unsigned long long above = (i == 63) ? 0ull : (~0ull << (i + 1));
        avail = ~cur & valid & above;
      }
      s_kept = kept;
    }
    __syncthreads();
    const unsigned long long kept = s_kept;
    const int base = s_count;
    // write kept indices (score order) : position = base + number of kept rows before it in this block
    if (tid < rows && ((kept >> tid) & 1ull)) {
      const int pos = base + __popcll(kept & ((1ull << tid) - 1ull));
      const int row = rb * 64 + tid;
      const int orig = p.presorted ? row : p.order[begin + row];
      p.keep_idx[begin + pos] = (int64_t)orig;
      if (p.keep_mask) p.keep_mask[begin + orig] = 1;
    }
    // OR the kept rows' masks into `removed` for the column blocks after rb: one warp per column word
    if (kept != 0ull) {
      const bool k0 = (kept >> lane) & 1ull, k1 = (kept >> (lane + 32)) & 1ull;
      const unsigned long long* r0 = mrow + (int64_t)(rb * 64 + lane) * p.wpr;
      const unsigned long long* r1 = r0 + (int64_t)32 * p.wpr;
      for (int w = rb + 1 + warp; w < nblk; w += kSweepWarps) {
        unsigned long long v = 0ull;
        if (k0) v = r0[w];
        if (k1) v |= r1[w];
        const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)v);
        const unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(v >> 32));
        if (lane == 0) removed[w] |= ((unsigned long long)hi << 32) | lo;
      }
    }
    __syncthreads();
    if (tid == 0) s_count = base + __popcll(kept);
    __syncthreads();
  }
  if (tid == 0) p.keep_counts[s] = s_count;
}

// iou_threshold >= 1 (the reference's shipped TEST.NMS thresholds, osrcnn_fast_rcnn.py:135 with test_nms_thresh = 1.0):
// the kernel's IoU can not exceed 1.0f - w <= min(aw, bw) and h <= min(ah, bh) after monotonic rounding, so
// inter = rn(w h) <= rn(aw ah) and rn(bw bh + rn(aw ah)) >= 2 inter (ties go to the even 2 inter), hence uni >= inter -
// and `iou > thr` never holds (nor for NaN / degenerate boxes: 0 / x and NaN compare false).  Greedy NMS then keeps every
// box: the result is the sorted order itself, no mask, no sweep.
__global__ void __launch_bounds__(256) nms_keep_all_kernel(const __grid_constant__ NmsParams p) {
  const int s = blockIdx.x;
  const int begin = p.seg_begin[s], len = min(p.seg_len[s], p.max_len);
  for (int i = threadIdx.x; i < len; i += 256) {
    p.keep_idx[begin + i] = p.presorted ? (int64_t)i : (int64_t)p.order[begin + i];
    if (p.keep_mask) p.keep_mask[begin + i] = 1;
  }
  if (threadIdx.x == 0) p.keep_counts[s] = len;
}

size_t ws_layout(int64_t T, int S, int max_len, size_t* off_sorted, size_t* off_order, size_t* off_mask) {
  const int wpr = (max_len + 63) / 64;
  size_t o = 0;
  *off_sorted = o; o += osr::align256((size_t)T * 16);
  *off_order = o;  o += osr::align256((size_t)T * 4);
  *off_mask = o;   o += osr::align256((size_t)S * max_len * wpr * 8);
  return o;
}

}  // namespace

extern "C" {

size_t osr_nms_workspace(int64_t total_boxes, int num_segments, int max_segment_len) {
  if (total_boxes < 0 || num_segments < 0 || max_segment_len < 0) return 0;
  size_t a, b, c;
  return ws_layout(total_boxes, num_segments, max_segment_len > 0 ? max_segment_len : 1, &a, &b, &c) + 256;
}

int osr_nms_segmented(const float* boxes, const float* scores, int64_t total_boxes, const int32_t* seg_begin,
                      const int32_t* seg_len, int num_segments, int max_segment_len, float iou_threshold,
                      int presorted, int64_t* keep_idx, int32_t* keep_counts, uint8_t* keep_mask, void* workspace,
                      size_t workspace_bytes, void* stream) {
  osr::DeviceGuard device_guard(boxes);
  if (num_segments < 0 || max_segment_len < 0 || total_boxes < 0) return osr::fail_arg(OSR_E_ARG, "nms: negative sizes");
  if (num_segments == 0) return 0;
  if (num_segments > 65535) return osr::fail_arg(OSR_E_SHAPE, "nms: more than 65535 segments in one call");
  if (!seg_begin || !seg_len || !keep_counts || !workspace) return osr::fail_arg(OSR_E_ARG, "nms: null pointer argument");
  if (max_segment_len > 0 && (!boxes || !scores || !keep_idx)) return osr::fail_arg(OSR_E_ARG, "nms: null pointer argument");
  if (reinterpret_cast<uintptr_t>(boxes) & 15) return osr::fail_arg(OSR_E_ARG, "nms: boxes must be 16-byte aligned");
  if (max_segment_len > kMaxLen || (!presorted && max_segment_len > kMaxSortLen))
    return osr::fail_arg(OSR_E_SHAPE, "nms: max_segment_len=%d unsupported (in-kernel sort <= %d, presorted <= %d)",
                         max_segment_len, kMaxSortLen, kMaxLen);
  NmsParams p{};
  p.boxes = boxes; p.scores = scores; p.seg_begin = seg_begin; p.seg_len = seg_len;
  p.num_segments = num_segments;
  p.max_len = max_segment_len > 0 ? max_segment_len : 1;
  p.wpr = (p.max_len + 63) / 64;
  p.thr = iou_threshold;
  p.presorted = presorted;
  p.keep_idx = keep_idx; p.keep_counts = keep_counts; p.keep_mask = keep_mask;
  size_t off_sorted, off_order, off_mask;
  const size_t need = ws_layout(total_boxes, num_segments, p.max_len, &off_sorted, &off_order, &off_mask);
  if (workspace_bytes < need) return osr::fail_arg(OSR_E_WORKSPACE, "nms: workspace too small");
  unsigned char* w = static_cast<unsigned char*>(workspace);
  p.sorted_boxes = reinterpret_cast<float4*>(w + off_sorted);
  p.order = reinterpret_cast<int32_t*>(w + off_order);
  p.mask = reinterpret_cast<unsigned long long*>(w + off_mask);

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!presorted) {
    const size_t smem = (size_t)osr::next_pow2(p.max_len < 32 ? 32 : p.max_len) * 8;
    OSR_CUDA_CHECK(cudaFuncSetAttribute(nms_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nms_sort_kernel<<<num_segments, kSortThreads, smem, st>>>(p);
    OSR_LAUNCH_CHECK();
  }
  if (!(iou_threshold < 1.0f)) {   // nothing can be suppressed (see nms_keep_all_kernel): sorted order = keep list
    nms_keep_all_kernel<<<num_segments, 256, 0, st>>>(p);
    OSR_LAUNCH_CHECK();
    return 0;
  }
  const int nb = (p.max_len + 63) / 64;
  nms_mask_kernel<<<dim3(nb, nb, num_segments), 64, 0, st>>>(p);
  OSR_LAUNCH_CHECK();
  nms_sweep_kernel<<<num_segments, kSweepThreads, (size_t)p.wpr * 8, st>>>(p);
  OSR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
