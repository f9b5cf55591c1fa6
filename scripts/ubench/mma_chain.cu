// Micro-benchmark: issue rate of tcgen05.mma (kind::f16, A in tensor memory, B in shared memory, M = 128, K = 16)
// when consecutive MMAs accumulate into the SAME accumulator (one dependent chain) versus round-robin over several
// accumulators.  One CTA per SM on `ctas` SMs, thread 32 issues `reps` MMAs, commits, everybody waits.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I obman_train_b200/csrc scripts/ubench/mma_chain.cu -o gpurun_out/mma_chain
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "common.cuh"
#include "sm100.cuh"
using namespace obman::sm100;

template <int N>
__global__ void __launch_bounds__(128, 1) chain_kernel(int reps, int n_acc, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + N * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < N * 32; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *slot, 0);
  long long t0 = 0, t1 = 0, t2 = 0;
  if (warp == 1) {
    if ((threadIdx.x & 31) == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, N);
      const uint64_t db = umma_desc(smem_u32(smem), 16, 1024, 2);
      const uint32_t a_tmem = tmem_base + 480;   // 16 columns of (garbage) A operand, outside the accumulators
      t0 = clock64();
      int acc = 0;
      for (int i = 0; i < reps; ++i) {
        umma_f16_ts(tmem_base + acc * N, a_tmem, db, idesc, 1u);
        if (++acc == n_acc) acc = 0;
      }
      t1 = clock64();
      umma_commit(bar);
    }
    __syncwarp();
  }
  mbar_wait(bar, 0);
  t2 = clock64();
  tc_fence_after();
  if (threadIdx.x == 32) { out[blockIdx.x * 3 + 0] = t1 - t0; out[blockIdx.x * 3 + 1] = t2 - t0; }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int N>
static void run(int reps, int n_acc, int ctas, long long* d_out) {
  const int smem = N * 128 + 1024 + 64;
  cudaFuncSetAttribute(chain_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  chain_kernel<N><<<ctas, 128, smem>>>(reps, n_acc, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d acc=%d: %s\n", N, n_acc, cudaGetErrorString(e)); exit(1); }
  long long h[6];
  cudaMemcpy(h, d_out, sizeof(long long) * 3, cudaMemcpyDeviceToHost);
  printf("N=%3d accumulators=%d reps=%d : issue %.1f clk/MMA, issue+drain %.1f clk/MMA (ideal at 8192 flop/clk/SM: %.1f)\n", N,
         n_acc, reps, (double)h[0] / reps, (double)h[1] / reps, 128.0 * N * 16 * 2 / 8192.0);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * 3 * 256);
  for (int rep = 0; rep < 2; ++rep) {   // second round is warm
    run<64>(2048, 1, 1, d_out);
    run<64>(2048, 2, 1, d_out);
    run<64>(2048, 3, 1, d_out);
    run<64>(2048, 6, 1, d_out);
    run<128>(2048, 1, 1, d_out);
    run<128>(2048, 2, 1, d_out);
    run<128>(2048, 3, 1, d_out);
    run<256>(2048, 1, 1, d_out);
    run<256>(2048, 1, 148, d_out);
    run<64>(2048, 1, 148, d_out);
  }
  return 0;
}
