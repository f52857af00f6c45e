// PLN encoder GEMM on 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//   emb[R, E] = x[R, F] . W[E, F]^T + b          (prototype_learning_network.py:133, nn.Linear(1024, 256))
// bf16 operands, fp32 accumulate in TMEM (the north star's "only kernels that use tensor cores").
//
//   cast kernel   x, W fp32 -> bf16 (vectorised; 16-byte loads, 8-byte stores) into the caller's workspace
//   GEMM kernel   CTA tile 128 (rows) x 128 (cols), K blocks of 64 (= one 128-byte swizzle row of bf16).
//                 warp 0 : TMA producer  - cp.async.bulk.tensor 2-D boxes {64, 128} with SWIZZLE_128B for A and B,
//                          4-stage ring, full/empty mbarriers
//                 warp 1 : MMA issuer    - one elected thread issues 4 x tcgen05.mma (M128 N128 K16, kind::f16) per K
//                          block from shared-memory descriptors; tcgen05.commit releases the stage / signals the epilogue
//                 warps 2-5 : epilogue   - tcgen05.ld 32x32b.x32 (TMEM lane = output row), + bias, fp32 stores
//                 TMEM: 128 columns x 128 lanes fp32 accumulator, allocated / freed by warp 1.
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>

#include "osr_common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 64, STAGES = 4, UMMA_K = 16;
constexpr int kThreads = 192;
constexpr uint32_t kTmemCols = 128;
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
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO=1 | SBO=1024B>>4 |
// version=1 (Blackwell) | layout_type=2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  const uint32_t lo = ((smem_addr >> 4) & 0x3fffu) | (1u << 16);
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return ((uint64_t)hi << 32) | lo;
}
// kind::f16 instruction descriptor: D=F32, A=B=BF16, both K-major, N=128, M=128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
// kind::tf32: D=F32, A=B=TF32 (format 2): the fp32 operands are read from shared memory as they are (the tensor core uses
// the top 19 bits), so x and W need no cast pass at all.  K per instruction = 8 (32 bytes), K block = 32 floats = 128 bytes.
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
constexpr int BK_TF32 = 32;

__device__ __forceinline__ void umma_f16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int kMaxPeers = 16;
struct EncParams {
  const float* bias;
  float* emb;              // destination 0 (local)
  int R, F, E;
  // fused all-gather: the epilogue stores every tile into `num_outs` buffers (this rank's and its peers', mapped
  // through NVLink) at row offset `row_off` - the embeddings never take a second trip through HBM or a collective.
  float* outs[kMaxPeers];
  int num_outs;
  int64_t row_off;
  float* mc;               // NVLS multicast mapping of the same buffer (NULL: one P2P store per destination)
};

template <bool kTf32>   // true: fp32 operands straight from HBM (kind::tf32); false: bf16 copies made by cast_bf16_kernel
__global__ void __launch_bounds__(kThreads, 1)
    pln_encode_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         const __grid_constant__ EncParams p) {
  extern __shared__ unsigned char smem_raw[];
  // 1024-byte alignment for the 128B-swizzled tiles
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = base;
  unsigned char* sB = base + STAGES * kStageBytesA;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(base + STAGES * (kStageBytesA + kStageBytesB));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  constexpr int BKe = kTf32 ? BK_TF32 : BK;   // elements per 128-byte K block
  const int num_kb = p.F / BKe;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation: one full warp, address written to shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)((kb / STAGES) & 1);
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_expect_tx(&full_bar[s], kStageBytesA + kStageBytesB);
        tma_load_2d(sA + s * kStageBytesA, &map_a, &full_bar[s], kb * BKe, m0);
        tma_load_2d(sB + s * kStageBytesB, &map_b, &full_bar[s], kb * BKe, n0);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)((kb / STAGES) & 1);
        mbar_wait(&full_bar[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t adesc = make_smem_desc(smem_u32(sA + s * kStageBytesA));
        const uint64_t bdesc = make_smem_desc(smem_u32(sB + s * kStageBytesB));
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 16 bf16 (or 8 tf32) = 32 bytes along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
          if (kTf32) umma_tf32(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), kIdescTf32, (kb | k) ? 1u : 0u);
          else umma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), kIdesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);                   // stage reusable once these MMAs have read it
        if (kb == num_kb - 1) umma_commit(tmem_full);  // accumulator complete
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue (warps 2..5): TMEM lane quarter = warp % 4 =====
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
      // bias, then transpose the 32 x 32 chunk through shared memory (the operand stages are idle: every MMA has
      // retired) so that each store instruction writes four 128-byte row segments - what NVLink peers want -
      // instead of 32 scattered 16-byte pieces
      float* tr = reinterpret_cast<float*>(sA) + q * (32 * 33);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        tr[lane * 33 + j] = __uint_as_float(v[j]) + (p.bias ? __ldg(p.bias + n0 + c0 + j) : 0.f);
      __syncwarp();
      const int sub = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll 1
      for (int d = 0; d < (p.mc ? 1 : p.num_outs); ++d) {
        float* base = p.mc ? p.mc : p.outs[d];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + sub;
          const int grow = m0 + q * 32 + rr;
          if (grow < p.R) {
            const float* sp = tr + rr * 33 + c4;
            const float4 o = make_float4(sp[0], sp[1], sp[2], sp[3]);
            float* out = base + (p.row_off + grow) * p.E + n0 + c0 + c4;
            if (p.mc) {
              asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out), "f"(o.x), "f"(o.y),
                           "f"(o.z), "f"(o.w)
                           : "memory");
            } else {
              *reinterpret_cast<float4*>(out) = o;
            }
          }
        }
      }
      __syncwarp();
    }
  }
  // teardown
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// fp32 -> bf16 (round to nearest even), n multiple of 4
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n4) {
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
// (rows, K) fp32 row-major -> 2-D map, box {32, 128} (= 128 bytes x 128 rows), 128-byte swizzle
int encode_f32_map(CUtensorMap* map, const void* ptr, int rows, int K) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return 0;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  cuuint32_t box[2] = {BK_TF32, BM};
  cuuint32_t es[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// (rows, K) bf16 row-major -> 2-D map, box {64, 128}, 128-byte swizzle
int encode_bf16_map(CUtensorMap* map, void* ptr, int rows, int K) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return 0;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {BK, BM};
  cuuint32_t es[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

extern "C" {

size_t osr_pln_encode_workspace(int R, int F, int E) {
  if (R < 0 || F <= 0 || E <= 0) return 0;
  return osr::align256((size_t)(R > 0 ? R : 1) * F * 2) + osr::align256((size_t)E * F * 2) + 256;
}

static int encode_launch(const float* x, const float* W, const float* bias, int R, int F, int E, float* const* outs,
                         int num_outs, int64_t row_off, float* mc, void* workspace, size_t workspace_bytes, void* stream);

int osr_pln_encode_fwd(const float* x, const float* W, const float* bias, int R, int F, int E, float* emb,
                       void* workspace, size_t workspace_bytes, void* stream) {
  float* outs[1] = {emb};
  return encode_launch(x, W, bias, R, F, E, outs, 1, 0, nullptr, workspace, workspace_bytes, stream);
}

int osr_pln_encode_gather_fwd(const float* x, const float* W, const float* bias, int R, int F, int E,
                              const uint64_t* h_peer_buffers, int world, int rank, uint64_t multicast_buffer,
                              void* workspace, size_t workspace_bytes, void* stream) {
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || !h_peer_buffers)
    return osr::fail_arg(OSR_E_ARG, "pln_encode_gather: bad world/rank (1 <= world <= %d) or null buffer list", kMaxPeers);
  float* outs[kMaxPeers];
  for (int d = 0; d < world; ++d) {   // own buffer first, then the peers in ring order (spreads the NVLink traffic)
    outs[d] = reinterpret_cast<float*>(static_cast<uintptr_t>(h_peer_buffers[(rank + d) % world]));
    if (!outs[d]) return osr::fail_arg(OSR_E_ARG, "pln_encode_gather: null peer buffer %d", d);
  }
  return encode_launch(x, W, bias, R, F, E, outs, world, (int64_t)rank * R,
                       reinterpret_cast<float*>(static_cast<uintptr_t>(multicast_buffer)), workspace, workspace_bytes, stream);
}

static int encode_launch(const float* x, const float* W, const float* bias, int R, int F, int E, float* const* outs,
                         int num_outs, int64_t row_off, float* mc, void* workspace, size_t workspace_bytes, void* stream) {
  osr::DeviceGuard device_guard(workspace);
  float* emb = outs[0];
  if (R < 0 || F <= 0 || E <= 0) return osr::fail_arg(OSR_E_ARG, "pln_encode: bad R/F/E");
  if (F % BK != 0 || E % BN != 0)
    return osr::fail_arg(OSR_E_SHAPE, "pln_encode: feature dim %d must be a multiple of %d and embedding dim %d of %d", F, BK, E, BN);
  if (R == 0) return 0;
  if (!x || !W || !emb || !workspace) return osr::fail_arg(OSR_E_ARG, "pln_encode: null pointer argument");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(W) & 15) || (reinterpret_cast<uintptr_t>(emb) & 15) ||
      (reinterpret_cast<uintptr_t>(workspace) & 255))
    return osr::fail_arg(OSR_E_ARG, "pln_encode: x / W / emb must be 16-byte and workspace 256-byte aligned");
  if (workspace_bytes < osr_pln_encode_workspace(R, F, E)) return osr::fail_arg(OSR_E_WORKSPACE, "pln_encode: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool tf32 = osr::tuning(osr::kTunePlnVariant) != 1;   // default: fp32 operands through kind::tf32, no cast passes
  CUtensorMap ma, mb;
  memset(&ma, 0, sizeof(ma));
  memset(&mb, 0, sizeof(mb));
  if (tf32) {
    if (!encode_f32_map(&ma, x, R, F) || !encode_f32_map(&mb, W, E, F))
      return osr::fail_arg(OSR_E_ARG, "pln_encode: cuTensorMapEncodeTiled failed");
  } else {
    __nv_bfloat16* xb = static_cast<__nv_bfloat16*>(workspace);
    __nv_bfloat16* wb = reinterpret_cast<__nv_bfloat16*>(static_cast<unsigned char*>(workspace) + osr::align256((size_t)R * F * 2));
    const int64_t nx4 = (int64_t)R * F / 4, nw4 = (int64_t)E * F / 4;
    cast_bf16_kernel<<<(int)((nx4 + 255) / 256 < 1184 ? (nx4 + 255) / 256 : 1184), 256, 0, s>>>(x, xb, nx4);
    OSR_LAUNCH_CHECK();
    cast_bf16_kernel<<<(int)((nw4 + 255) / 256 < 1184 ? (nw4 + 255) / 256 : 1184), 256, 0, s>>>(W, wb, nw4);
    OSR_LAUNCH_CHECK();
    if (!encode_bf16_map(&ma, xb, R, F) || !encode_bf16_map(&mb, wb, E, F))
      return osr::fail_arg(OSR_E_ARG, "pln_encode: cuTensorMapEncodeTiled failed");
  }
  EncParams p;
  p.bias = bias; p.emb = emb; p.R = R; p.F = F; p.E = E;
  p.num_outs = num_outs; p.row_off = row_off; p.mc = mc;
  for (int d = 0; d < kMaxPeers; ++d) p.outs[d] = d < num_outs ? outs[d] : nullptr;
  for (int d = 0; d < num_outs; ++d)
    if (reinterpret_cast<uintptr_t>(outs[d]) & 15) return osr::fail_arg(OSR_E_ARG, "pln_encode: output buffers must be 16-byte aligned");
  if (tf32) {
    OSR_CUDA_CHECK(cudaFuncSetAttribute(pln_encode_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    pln_encode_tc_kernel<true><<<dim3(osr::ceil_div(R, BM), E / BN), kThreads, kSmemBytes, s>>>(ma, mb, p);
  } else {
    OSR_CUDA_CHECK(cudaFuncSetAttribute(pln_encode_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    pln_encode_tc_kernel<false><<<dim3(osr::ceil_div(R, BM), E / BN), kThreads, kSmemBytes, s>>>(ma, mb, p);
  }
  OSR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
