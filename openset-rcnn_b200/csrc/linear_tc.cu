// Box-head fully connected layers on 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//   out[R, N] = act(A[R, K] . W[N, K]^T + bias)          (detectron2 FastRCNNConvFCHead fc1 / fc2 behind
//                                                          osrcnn_roi_heads.py:308)
// A and W are bf16 (A = the ROIAlign output written in bf16 by osr_roi_align_fwd_bf16, or the previous layer's bf16
// output; W = the layer's weight cast once per optimizer step), accumulation fp32 in TMEM, bias + ReLU + output cast
// fused into the epilogue.  The pooled tensor therefore makes ONE trip through memory in bf16 (205 MB at cfg 2, read
// back through L2 by the TMA loads below) instead of the reference's 411 MB fp32 write + 411 MB read.
//
//   CTA tile 128 (rows) x 256 (cols), K blocks of 64 bf16 (= one 128-byte swizzle row), 4-stage TMA ring (48 KB / stage)
//   warp 0     TMA producer   cp.async.bulk.tensor 2-D boxes {64, 128} (A) and 2 x {64, 128} (W), SWIZZLE_128B
//   warp 1     MMA issuer     one elected thread: 4 x tcgen05.mma.cta_group::1.kind::f16 (M128 N256 K16) per K block,
//                             tcgen05.commit releases the stage / signals the epilogue; TMEM allocator (256 columns)
//   warps 2-5  epilogue       tcgen05.ld 32x32b.x32 (TMEM lane = output row) -> + bias -> ReLU -> bf16 / fp32 row stores
//   Tile order: the N index runs fastest, so the 4 CTAs that share an A row-tile run together and A is fetched from
//   HBM once (the other three read it from L2); W (26 MB at fc1) stays L2-resident.
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>

#include "osr_common.cuh"

namespace {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4, UMMA_K = 16;
constexpr int kThreads = 192;
constexpr uint32_t kTmemCols = 256;
constexpr int kStageBytesA = BM * BK * 2, kStageBytesB = BN * BK * 2;
constexpr size_t kSmemBytes = 1024 /*align slack*/ + (size_t)STAGES * (kStageBytesA + kStageBytesB) + 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ bool elect_one() {
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
// K-major, SWIZZLE_128B shared-memory matrix descriptor: start >> 4 | LBO = 1 | SBO = 1024 B >> 4 | version 1 | layout 2
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  const uint32_t lo = ((smem_addr >> 4) & 0x3fffu) | (1u << 16);
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, N = 256, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct LinParams {
  const float* bias;   // (N) or null
  void* out;           // (R, N) bf16 or fp32, row-major
  int R, K, N;
  int relu, out_bf16;
  int tiles_n;
};

__global__ void __launch_bounds__(kThreads, 1)
    linear_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ LinParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = base;
  unsigned char* sB = base + STAGES * kStageBytesA;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(base + STAGES * (kStageBytesA + kStageBytesB));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int m0 = (tile / p.tiles_n) * BM, n0 = (tile % p.tiles_n) * BN;   // N fastest: the CTAs sharing an A tile are neighbours
  const int num_kb = p.K / BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)((kb / STAGES) & 1);
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_expect_tx(&full_bar[s], kStageBytesA + kStageBytesB);
        tma_load_2d(sA + s * kStageBytesA, &map_a, &full_bar[s], kb * BK, m0);
        tma_load_2d(sB + s * kStageBytesB, &map_b, &full_bar[s], kb * BK, n0);                          // rows n0 .. n0+127
        tma_load_2d(sB + s * kStageBytesB + kStageBytesB / 2, &map_b, &full_bar[s], kb * BK, n0 + 128);   // rows n0+128 .. n0+255
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)((kb / STAGES) & 1);
        mbar_wait(&full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t adesc = make_smem_desc(smem_u32(sA + s * kStageBytesA));
        const uint64_t bdesc = make_smem_desc(smem_u32(sB + s * kStageBytesB));
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), kIdesc, (kb | k) ? 1u : 0u);
        umma_commit(&empty_bar[s]);
        if (kb == num_kb - 1) umma_commit(tmem_full);
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue (warps 2..5): TMEM lane quarter = warp % 4; lane = one output row =====
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float x = __uint_as_float(v[j]) + (p.bias ? __ldg(p.bias + n0 + c0 + j) : 0.f);
        f[j] = p.relu ? fmaxf(x, 0.f) : x;
      }
      if (row < p.R) {
        if (p.out_bf16) {   // 32 bf16 = 64 contiguous bytes of this thread's row
          uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + (int64_t)row * p.N + n0 + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            __nv_bfloat162 b0 = __floats2bfloat162_rn(f[8 * j], f[8 * j + 1]), b1 = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
            __nv_bfloat162 b2 = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]), b3 = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
            uint4 o;
            o.x = *reinterpret_cast<uint32_t*>(&b0); o.y = *reinterpret_cast<uint32_t*>(&b1);
            o.z = *reinterpret_cast<uint32_t*>(&b2); o.w = *reinterpret_cast<uint32_t*>(&b3);
            dst[j] = o;
          }
        } else {
          float4* dst = reinterpret_cast<float4*>(static_cast<float*>(p.out) + (int64_t)row * p.N + n0 + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) dst[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// fp32 -> bf16 (round to nearest even), n multiple of 4
__global__ void __launch_bounds__(256) cast_bf16_lin_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    reinterpret_cast<uint2*>(dst)[i] = o;
  }
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
// (rows, K) bf16 row-major -> 2-D map, box {64, 128}, 128-byte swizzle; rows past the end read as zero
int encode_map(CUtensorMap* map, const void* ptr, int rows, int K) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return 0;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {BK, 128};
  cuuint32_t es[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

extern "C" {

int osr_cast_bf16(const float* src, void* dst_bf16, int64_t n, void* stream) {
  osr::DeviceGuard device_guard(dst_bf16);
  if (n < 0 || (n & 3)) return osr::fail_arg(OSR_E_ARG, "cast_bf16: n must be a non-negative multiple of 4");
  if (n == 0) return 0;
  if (!src || !dst_bf16 || (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst_bf16) & 7))
    return osr::fail_arg(OSR_E_ARG, "cast_bf16: null or misaligned pointer");
  const int64_t n4 = n / 4;
  const int blocks = (int)((n4 + 255) / 256 < 148 * 8 ? (n4 + 255) / 256 : 148 * 8);
  cast_bf16_lin_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, static_cast<__nv_bfloat16*>(dst_bf16), n4);
  OSR_LAUNCH_CHECK();
  return 0;
}

int osr_linear_bf16_fwd(const void* a_bf16, const void* w_bf16, const float* bias, int R, int K, int N, int relu, void* out,
                        int out_is_bf16, void* stream) {
  osr::DeviceGuard device_guard(out);
  if (R < 0 || K <= 0 || N <= 0) return osr::fail_arg(OSR_E_ARG, "linear: bad R/K/N");
  if (K % BK != 0 || N % BN != 0)
    return osr::fail_arg(OSR_E_SHAPE, "linear: K=%d must be a multiple of %d and N=%d of %d", K, BK, N, BN);
  if (R == 0) return 0;
  if (!a_bf16 || !w_bf16 || !out) return osr::fail_arg(OSR_E_ARG, "linear: null pointer argument");
  if ((reinterpret_cast<uintptr_t>(a_bf16) & 15) || (reinterpret_cast<uintptr_t>(w_bf16) & 15) || (reinterpret_cast<uintptr_t>(out) & 15))
    return osr::fail_arg(OSR_E_ARG, "linear: A / W / out must be 16-byte aligned");
  CUtensorMap ma, mb;
  memset(&ma, 0, sizeof(ma));
  memset(&mb, 0, sizeof(mb));
  if (!encode_map(&ma, a_bf16, R, K) || !encode_map(&mb, w_bf16, N, K))
    return osr::fail_arg(OSR_E_ARG, "linear: cuTensorMapEncodeTiled failed");
  LinParams p;
  p.bias = bias; p.out = out; p.R = R; p.K = K; p.N = N; p.relu = relu; p.out_bf16 = out_is_bf16;
  p.tiles_n = N / BN;
  OSR_CUDA_CHECK(cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
  linear_tc_kernel<<<osr::ceil_div(R, BM) * p.tiles_n, kThreads, kSmemBytes, static_cast<cudaStream_t>(stream)>>>(ma, mb, p);
  OSR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
