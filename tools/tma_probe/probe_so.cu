// Minimal TMA probe: 4-D fp32 tensor (W,H,C,N), box 32x8x8x1, descriptor passed three ways.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include <stdlib.h>

struct alignas(64) Maps { CUtensorMap map[8]; };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
__device__ void body(const CUtensorMap* map, float* out, int x, int y, int c, int n) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* tile = reinterpret_cast<float*>(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    if (elect_one()) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(8192) : "memory");
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
          ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(c), "r"(n) : "memory");
    }
    __syncwarp();
  }
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = tile[i];
}
__global__ void k_2d(const __grid_constant__ CUtensorMap m, float* out, int x, int y) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* tile = reinterpret_cast<float*>(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(1024) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(&m)), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW2:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D2;\nbra W2;\nD2:\n}\n" ::"r"(smem_u32(bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < 256; i += blockDim.x) out[i] = tile[i];
}
__device__ void body2d(const CUtensorMap* m, float* out, int x, int y, int use_elect) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* tile = reinterpret_cast<float*>(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    bool go = use_elect ? elect_one() : (threadIdx.x == 0);
    if (go) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(8192) : "memory");
      for (int k = 0; k < 8; ++k)
        asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
          ::"r"(smem_u32(tile + k * 256)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(x), "r"(y + k * 80) : "memory");
    }
    __syncwarp();
  }
  asm volatile("{\n.reg .pred p;\nW4:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D4;\nbra W4;\nD4:\n}\n" ::"r"(smem_u32(bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = tile[i];
}
__global__ void k2_direct(const __grid_constant__ CUtensorMap m, float* out, int x, int y, int e) { body2d(&m, out, x, y, e); }
__global__ void k2_struct(const __grid_constant__ Maps ms, const int* idx, float* out, int x, int y, int e) { body2d(&ms.map[idx[blockIdx.x]], out, x, y, e); }
struct Dummy { char b[552]; };
__global__ void k2_struct2(const __grid_constant__ Dummy dd, const __grid_constant__ Maps ms, const int* idx, float* out, int x, int y, int e) {
  if (dd.b[3] == 77) out[0] = 1.f;
  body2d(&ms.map[idx[blockIdx.x]], out, x, y, e);
}
__global__ void k2_global(const CUtensorMap* m, float* out, int x, int y, int e) { body2d(m, out, x, y, e); }
__global__ void k_3d(const __grid_constant__ CUtensorMap m, float* out, int x, int y, int c, int tx_bytes) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* tile = reinterpret_cast<float*>(smem);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 8192);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(tx_bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(tile)), "l"(reinterpret_cast<uint64_t>(&m)), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(c) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW3:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D3;\nbra W3;\nD3:\n}\n" ::"r"(smem_u32(bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = tile[i];
}
__global__ void k_direct(const __grid_constant__ CUtensorMap m, float* out, int x, int y, int c, int n) { body(&m, out, x, y, c, n); }
__global__ void k_struct(const __grid_constant__ Maps ms, int idx, float* out, int x, int y, int c, int n) { body(&ms.map[idx], out, x, y, c, n); }
__global__ void k_global(const CUtensorMap* m, float* out, int x, int y, int c, int n) { body(m, out, x, y, c, n); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
extern "C" int probe_main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int W = getenv("PW") ? atoi(getenv("PW")) : 120, H = getenv("PH") ? atoi(getenv("PH")) : 80, C = 16, N = 2;
  std::vector<float> h((size_t)W * H * C * N);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *out;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 8192);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  printf("entry point: err=%d q=%d f=%p\n", (int)e, (int)q, f);
  Maps ms; memset(&ms, 0, sizeof(ms));
  cuuint64_t dims[4] = {W, H, C, N};
  cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
  cuuint32_t box[4] = {32, 8, 8, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = ((EncodeTiledFn)f)(&ms.map[2], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  if (variant >= 30) {
    Maps ms2; memset(&ms2, 0, sizeof(ms2));
    cuuint64_t d2[2] = {W, (cuuint64_t)H * C * N}; cuuint64_t s2[1] = {(cuuint64_t)W * 4};
    cuuint32_t b2[2] = {32, 8}, e2[2] = {1, 1};
    r = ((EncodeTiledFn)f)(&ms2.map[2], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, d2, s2, b2, e2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode2d: %d base=%p\n", (int)r, (void*)d);
    { unsigned char* b = (unsigned char*)&ms2.map[2]; for (int i = 0; i < 128; ++i) printf("%02x%s", b[i], (i % 32 == 31) ? "\n" : ""); }
    int two = 2, *didx; cudaMalloc(&didx, 4); cudaMemcpy(didx, &two, 4, cudaMemcpyHostToDevice);
    CUtensorMap* dm2; cudaMalloc(&dm2, sizeof(CUtensorMap)); cudaMemcpy(dm2, &ms2.map[2], sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    const int el = (variant % 2);
    const int mode = (variant - 30) / 2;
    if (mode == 0) k2_direct<<<1, 128, 8192 + 64>>>(ms2.map[2], out, 4, 3, el);
    if (mode == 1) k2_struct<<<1, 128, 8192 + 64>>>(ms2, didx, out, 4, 3, el);
    if (mode == 2) k2_global<<<1, 128, 8192 + 64>>>(dm2, out, 4, 3, el);
    if (mode == 3) { Dummy dd; memset(&dd, 0, sizeof(dd)); k2_struct2<<<1, 128, 8192 + 64>>>(dd, ms2, didx, out, 4, 3, el); }
    e = cudaDeviceSynchronize();
    printf("2d mode %d elect %d: %s\n", mode, el, cudaGetErrorString(e));
    std::vector<float> o2(2048); cudaMemcpy(o2.data(), out, 8192, cudaMemcpyDeviceToHost);
    int bad = 0; for (int k = 0; k < 8; ++k) for (int rr = 0; rr < 8; ++rr) for (int xx = 0; xx < 32; ++xx) if (o2[k*256 + rr*32+xx] != h[(size_t)(3+rr+k*80)*W + 4 + xx]) ++bad;
    printf("mismatches %d\n", bad);
    return 0;
  }
  if (variant >= 20) {
    CUtensorMap m3; memset(&m3, 0, sizeof(m3));
    cuuint64_t d3[3] = {W, H, (cuuint64_t)C * N}; cuuint64_t s3[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t b3[3] = {(cuuint32_t)(argc > 2 ? atoi(argv[2]) : 32), (cuuint32_t)(argc > 3 ? atoi(argv[3]) : 8), (cuuint32_t)(argc > 4 ? atoi(argv[4]) : 8)}, e3[3] = {1, 1, 1};
    const int tx_bytes = b3[0] * b3[1] * b3[2] * 4;
    r = ((EncodeTiledFn)f)(&m3, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, d3, s3, b3, e3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode3d: %d\n", (int)r);
    const int x3 = variant == 21 ? -3 : 5, y3 = variant == 21 ? 76 : 3, c3 = 1 * C + 8;
    k_3d<<<1, 128, 8192 + 64>>>(m3, out, x3, y3, c3, tx_bytes);
    e = cudaDeviceSynchronize();
    printf("3d variant %d: %s\n", variant, cudaGetErrorString(e));
    std::vector<float> o3(2048); cudaMemcpy(o3.data(), out, 8192, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int cc = 0; cc < 8; ++cc) for (int rr = 0; rr < 8; ++rr) for (int xx = 0; xx < 32; ++xx) {
      int gx = x3 + xx, gy = y3 + rr;
      float exp = (gx < 0 || gx >= W || gy < 0 || gy >= H) ? 0.f : h[(((size_t)c3 + cc) * H + gy) * W + gx];
      if (o3[(cc * 8 + rr) * 32 + xx] != exp) ++bad;
    }
    printf("3d mismatches %d\n", bad);
    return 0;
  }
  if (variant >= 10) {
    CUtensorMap m2; memset(&m2, 0, sizeof(m2));
    cuuint64_t d2[2] = {W, (cuuint64_t)H * C * N}; cuuint64_t s2[1] = {(cuuint64_t)W * 4};
    cuuint32_t b2[2] = {32, 8}, e2[2] = {1, 1};
    CUtensorMapL2promotion prom = variant == 11 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    r = ((EncodeTiledFn)f)(&m2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, d2, s2, b2, e2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode2d: %d\n", (int)r);
    k_2d<<<1, 128, 8192 + 64>>>(m2, out, 4, 3);
    e = cudaDeviceSynchronize();
    printf("2d variant %d: %s\n", variant, cudaGetErrorString(e));
    std::vector<float> o2(256); cudaMemcpy(o2.data(), out, 1024, cudaMemcpyDeviceToHost);
    int bad = 0; for (int rr = 0; rr < 8; ++rr) for (int xx = 0; xx < 32; ++xx) if (o2[rr*32+xx] != h[(size_t)(3+rr)*W + 4 + xx]) ++bad;
    printf("2d mismatches %d\n", bad);
    return 0;
  }
  CUtensorMap* dm; cudaMalloc(&dm, sizeof(CUtensorMap)); cudaMemcpy(dm, &ms.map[2], sizeof(CUtensorMap), cudaMemcpyHostToDevice);
  std::vector<float> o(2048);
  const int x = 5, y = 3, c = 8, n = 1;
  for (int mode = 0; mode < 3; ++mode) {
    cudaMemset(out, 0, 8192);
    if (mode == 0) k_direct<<<1, 128, 8192 + 64>>>(ms.map[2], out, x, y, c, n);
    if (mode == 1) k_struct<<<1, 128, 8192 + 64>>>(ms, 2, out, x, y, c, n);
    if (mode == 2) k_global<<<1, 128, 8192 + 64>>>(dm, out, x, y, c, n);
    e = cudaDeviceSynchronize();
    printf("mode %d: %s\n", mode, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(o.data(), out, 8192, cudaMemcpyDeviceToHost);
    // element (cc, rr, xx) of the box = tensor[n][c+cc][y+rr][x+xx]
    int bad = 0;
    for (int cc = 0; cc < 8; ++cc) for (int rr = 0; rr < 8; ++rr) for (int xx = 0; xx < 32; ++xx) {
      float exp = h[(((size_t)n * C + c + cc) * H + y + rr) * W + x + xx];
      if (o[(cc * 8 + rr) * 32 + xx] != exp) ++bad;
    }
    printf("mode %d mismatches: %d\n", mode, bad);
  }
  return 0;
}
