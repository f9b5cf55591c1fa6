// Shared helpers for libobman_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define OBMAN_OK 0
#define OBMAN_ERR_BAD_ARG -1
#define OBMAN_ERR_CUDA -2
#define OBMAN_ERR_UNSUPPORTED -3
#define OBMAN_ERR_DRIVER -4
#define OBMAN_PREC_TF32 1
#define OBMAN_PREC_3XBF16 2
#define OBMAN_PREC_3XTF32 3

namespace obman {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA launch error: %s", what, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return OBMAN_ERR_CUDA;
  }
  return OBMAN_OK;
}

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum; result valid in thread 0. `scratch` must hold >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  v = warp_sum(v);
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    v = lane < nw ? scratch[lane] : 0.f;
    v = warp_sum(v);
  }
  return v;
}

}  // namespace obman

#define OBMAN_REQUIRE(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      obman::set_error(__VA_ARGS__);      \
      return OBMAN_ERR_BAD_ARG;           \
    }                                     \
  } while (0)
