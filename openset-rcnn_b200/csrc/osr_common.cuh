// Shared host/device helpers for libosr_sm100a.so.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/osr.h"

namespace osr {

// thread-local error string + launch counter (the only state the library keeps; per thread)
char* tls_error_buf();
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// kernel-variant switches (osr_set_tuning / OSR_TUNE_* read once at load): A/B measurement only
enum TuneKey { kTuneBwdVariant = OSR_TUNE_BWD_VARIANT, kTuneFwdVariant = OSR_TUNE_FWD_VARIANT,
               kTunePlnVariant = OSR_TUNE_PLN_VARIANT, kTuneRpnVariant = OSR_TUNE_RPN_VARIANT,
               kTuneNmsVariant = OSR_TUNE_NMS_VARIANT, kTuneBwdSplit = OSR_TUNE_BWD_SPLIT };
int tuning(int key);

inline int fail_arg(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tls_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

#define OSR_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      osr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                     \
    }                                                                                     \
  } while (0)

#define OSR_LAUNCH_CHECK()                                                                \
  do {                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      osr::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return (int)_e;                                                                     \
    }                                                                                     \
    osr::count_launch();                                                                  \
  } while (0)

// Launch on the device that OWNS the data: every entry point opens with `osr::DeviceGuard guard(<first device pointer>)`,
// so a caller whose current device differs from the tensors' device (multi-GPU in one process, autograd worker threads)
// gets correct launches and per-device function attributes; the previous current device is restored on return.
struct DeviceGuard {
  int prev = -1;
  bool changed = false;
  explicit DeviceGuard(const void* p) {
    cudaPointerAttributes a;
    if (p && cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeDevice && cudaGetDevice(&prev) == cudaSuccess &&
        a.device != prev) {
      changed = cudaSetDevice(a.device) == cudaSuccess;
    } else {
      (void)cudaGetLastError();   // an unregistered pointer is not an error here; the entry point validates its arguments
    }
  }
  ~DeviceGuard() {
    if (changed) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) {
  return (a + b - 1) / b;
}
template <typename T>
__host__ __device__ constexpr T round_up(T a, T b) {
  return ceil_div(a, b) * b;
}

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

__host__ inline int next_pow2(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

}  // namespace osr
