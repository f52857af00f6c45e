// Micro-benchmark: issue rate of legacy mma.sync (m16n8k8 tf32 / m16n8k16 bf16) and FFMA on sm_100a, per SM.
// Decides whether a 3xTF32 mma.sync formulation of the ROIAlign-backward x expansion can beat the FFMA form
// (DESIGN.md section 4).  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void k_tf32(float* out, int iters) {
  float c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = a0 * 5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_bf16(float* out, int iters) {
  float c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
  unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = a0 * 5;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_ffma(float* out, int iters, float w) {
  float c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = (float)i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fmaf(c[i], w, 1.0f);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_ffma2(float* out, int iters, float w) {
  unsigned long long c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = (unsigned long long)i * 0x3f8000003f800000ull;
  const float2 wb = make_float2(w, w);
  const float2 ob = make_float2(1.f, 1.f);
  const unsigned long long ww = *reinterpret_cast<const unsigned long long*>(&wb), one = *reinterpret_cast<const unsigned long long*>(&ob);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(c[i]) : "l"(ww), "l"(one));
  }
  unsigned long long s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s ^= c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)(s & 0xffff);
}

template <class F>
float time_ms(F&& launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); cudaDeviceSynchronize();
  cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  float* out; cudaMalloc(&out, sizeof(float) * sms * 8 * 1024);
  const int iters = 4096;
  const double ghz = prop.clockRate * 1e-6;
  printf("{\"sms\": %d, \"clock_ghz_nominal\": %.3f", sms, ghz);
  for (int warps = 4; warps <= 16; warps *= 2) {
    const int threads = warps * 32;
    float ms = time_ms([&] { k_tf32<8><<<sms, threads>>>(out, iters); });
    double mmas = (double)sms * warps * iters * 8;
    printf(", \"tf32_m16n8k8_w%d\": {\"ms\": %.3f, \"mma_per_ns_per_sm\": %.4f, \"tflops\": %.1f}", warps, ms,
           mmas / (ms * 1e6) / sms, mmas * 2048.0 / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { k_bf16<8><<<sms, threads>>>(out, iters); });
    printf(", \"bf16_m16n8k16_w%d\": {\"ms\": %.3f, \"mma_per_ns_per_sm\": %.4f, \"tflops\": %.1f}", warps, ms,
           mmas / (ms * 1e6) / sms, mmas * 4096.0 / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { k_ffma<16><<<sms, threads>>>(out, iters, 0.999f); });
    double ff = (double)sms * warps * iters * 16;
    printf(", \"ffma_w%d\": {\"ms\": %.3f, \"warp_ffma_per_ns_per_sm\": %.4f, \"tflops\": %.1f}", warps, ms,
           ff / (ms * 1e6) / sms, ff * 64.0 / (ms * 1e-3) / 1e12);
    ms = time_ms([&] { k_ffma2<16><<<sms, threads>>>(out, iters, 0.999f); });
    printf(", \"ffma2_w%d\": {\"ms\": %.3f, \"warp_ffma2_per_ns_per_sm\": %.4f, \"tflops\": %.1f}", warps, ms,
           ff / (ms * 1e6) / sms, ff * 128.0 / (ms * 1e-3) / 1e12);
  }
  printf("}\n");
  return 0;
}
