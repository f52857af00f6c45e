// dlopen libosr_sm100a.so and run its TMA self-test without torch in the process
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "../../include/osr.h"
typedef int (*SelfTest)(const osr_feat_level_t*, int, int, int, int, int, int, float*, void*);
int main(int argc, char** argv) {
  void* h = dlopen(argv[1], RTLD_NOW);
  if (!h) { printf("dlopen failed: %s\n", dlerror()); return 1; }
  SelfTest fn = (SelfTest)dlsym(h, "osr_debug_tma_selftest");
  const int W = 120, H = 80, C = 16;
  std::vector<float> hst((size_t)W * H * C);
  for (size_t i = 0; i < hst.size(); ++i) hst[i] = (float)i;
  float *d, *out;
  cudaMalloc(&d, hst.size() * 4); cudaMalloc(&out, 65536);
  cudaMemcpy(d, hst.data(), hst.size() * 4, cudaMemcpyHostToDevice);
  osr_feat_level_t lv; lv.data = d; lv.sN = (int64_t)C * H * W; lv.sC = (int64_t)H * W; lv.sH = W; lv.sW = 1; lv.H = H; lv.W = W; lv.scale = 0.25f;
  if (argc > 2 && atoi(argv[2]) == 9999) {
    typedef int (*MinFn)(float*, int, int, float*, void*);
    MinFn mf = (MinFn)dlsym(h, "osr_debug_tma_min");
    int r = (argc > 3) ? mf(nullptr, W, H * C, nullptr, (void*)1) : mf(d, W, H * C, out, 0);
    cudaError_t e2 = cudaDeviceSynchronize();
    printf("min rc=%d sync=%s\n", r, cudaGetErrorString(e2));
    return 0;
  }
  typedef int (*DbgMap)(const osr_feat_level_t*, int, int, int, int, unsigned char*);
  DbgMap dm = (DbgMap)dlsym(h, "osr_debug_tensormap");
  unsigned char b[128];
  int ok = dm(&lv, 1, 1, C, 0, b);
  printf("lib map ok=%d base=%p\n", ok, (void*)d);
  for (int i = 0; i < 128; ++i) printf("%02x%s", b[i], (i % 32 == 31) ? "\n" : "");
  int rc = fn(&lv, 1, 1, C, 0, argc > 2 ? atoi(argv[2]) : 5, 3, out, 0);
  cudaError_t e = cudaDeviceSynchronize();
  printf("harness rc=%d sync=%s\n", rc, cudaGetErrorString(e));
  return 0;
}
