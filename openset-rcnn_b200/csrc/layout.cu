// NCHW <-> NHWC staging for feature maps / gradient maps (SURVEY.md F-layout; DESIGN.md section 4, "NCHW maps").
// The reference's backbone emits NCHW-contiguous FPN maps (osrcnn_roi_heads.py:306); the fast ROIAlign kernels want a
// pixel's C channels contiguous.  One tiled transpose per level in front of the forward (and one behind the backward)
// costs 2 x the map's bytes at ~HBM speed, which is cheaper than running the NCHW kernels (3.0 -> ~2.3 ms per cfg-2 step).
//   per image:  src (C, HW) row-major  ->  dst (HW, C) row-major      (or back)
// 32 x 32 tiles through padded shared memory: both the reads and the writes are 128-byte coalesced.
#include "osr_common.cuh"

namespace {

constexpr int kTile = 32, kRows = 8;

// in: (rows, cols) row-major with leading dimension cols;  out: (cols, rows).  blockIdx.z = image.
__global__ void __launch_bounds__(kTile * kRows) transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows,
                                                                  int cols, int64_t in_img_stride, int64_t out_img_stride) {
  __shared__ float tile[kTile][kTile + 1];
  const float* src = in + (int64_t)blockIdx.z * in_img_stride;
  float* dst = out + (int64_t)blockIdx.z * out_img_stride;
  const int c0 = blockIdx.x * kTile, r0 = blockIdx.y * kTile;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int j = 0; j < kTile; j += kRows) {
    const int r = r0 + ty + j, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + j][tx] = __ldg(src + (int64_t)r * cols + c);
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kTile; j += kRows) {
    const int c = c0 + ty + j, r = r0 + tx;   // output row = input column
    if (c < cols && r < rows) dst[(int64_t)c * rows + r] = tile[tx][ty + j];
  }
}

int launch(const float* in, float* out, int N, int rows, int cols, void* stream) {
  if (N < 0 || rows < 0 || cols < 0) return osr::fail_arg(OSR_E_ARG, "layout: negative size");
  if (N == 0 || rows == 0 || cols == 0) return 0;
  if (!in || !out) return osr::fail_arg(OSR_E_ARG, "layout: null pointer argument");
  if (N > 65535 || osr::ceil_div(rows, kTile) > 65535) return osr::fail_arg(OSR_E_SHAPE, "layout: too many images / rows for one launch");
  dim3 grid(osr::ceil_div(cols, kTile), osr::ceil_div(rows, kTile), N);
  transpose_kernel<<<grid, dim3(kTile, kRows), 0, static_cast<cudaStream_t>(stream)>>>(in, out, rows, cols, (int64_t)rows * cols,
                                                                                      (int64_t)rows * cols);
  OSR_LAUNCH_CHECK();
  return 0;
}

}  // namespace

extern "C" {

// (N, C, H*W) contiguous -> (N, H*W, C) contiguous, i.e. NCHW -> channels_last memory
int osr_nchw_to_nhwc(const float* src, float* dst, int N, int C, int64_t HW, void* stream) {
  osr::DeviceGuard device_guard(dst);
  if (HW > 0x7fffffff) return osr::fail_arg(OSR_E_SHAPE, "layout: H*W too large");
  return launch(src, dst, N, C, (int)HW, stream);
}

// (N, H*W, C) contiguous -> (N, C, H*W) contiguous
int osr_nhwc_to_nchw(const float* src, float* dst, int N, int C, int64_t HW, void* stream) {
  osr::DeviceGuard device_guard(dst);
  if (HW > 0x7fffffff) return osr::fail_arg(OSR_E_SHAPE, "layout: H*W too large");
  return launch(src, dst, N, (int)HW, C, stream);
}

}  // extern "C"
