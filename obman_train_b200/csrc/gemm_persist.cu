// Persistent 3xBF16 GEMM for plain matrices with a short reduction: out[M,N] = epilogue(alpha * A[M,K] * W[N,K]^T).
//
// The AtlasNet point decoder (atlasutils.py:65-75) is three GEMMs over B*N_points rows (655 872 at batch 256) with
// K = 515 / 257 / 128, forward and data gradient.  gemm_tc_kernel gives every 128 x 128 tile its own CTA: with 4-17 K
// blocks per tile the tensor-memory allocation, barrier set-up, pipeline fill and the epilogue (TMEM -> registers ->
// shared memory -> global, with the ReLU mask read) cost more than the main loop (measured round 2: 14-40 % tensor
// pipe on these layers against 76-95 % on the convolutions with K >= 1152).  Here one CTA per SM walks over the tiles:
//
//   warp 0      TMA producer: A tile (128 rows x 32 fp32) + packed weight tile (128 rows x [32 hi | 32 lo] bf16) per K
//               block into a 6-stage ring that runs ACROSS tile boundaries (no fill / drain between tiles)
//   warp 1      tcgen05.mma issuer: A from tensor memory, B from shared memory, 3 products per 16-wide K step
//   warps 2-5   splitters: fp32 A row -> bf16 hi / lo -> tensor memory (one row per thread, as in gemm_tc_kernel)
//   warps 6-9   epilogue of the PREVIOUS tile, concurrently: two accumulators (2 x 128 TMEM columns) alternate
//
// Tensor memory: 2 x 128 accumulator columns + 6 x 32 operand columns.  TAIL = 1: N = 128 j + (1..4); the extra
// columns are evaluated on the CUDA cores by the splitter threads of column tile 0 (see gemm_tc_kernel's TAIL note).
#include <stdlib.h>

#include "gemm_shared.cuh"

namespace obman {

constexpr int GP_BN = 128;
constexpr int GP_S = 6;
constexpr int GP_B_TILE = GP_BN * 128;
constexpr int GP_STAGE = A_TILE_BYTES + GP_B_TILE;      // 32 KB
constexpr int GP_STAGING = 16384;                        // 4 epilogue warps x 32 rows x 128 B
constexpr int GP_THREADS = 320;
constexpr int GP_SMEM = GP_S * GP_STAGE + GP_STAGING + GP_S * 512 + 256 + 1024;
static_assert(GP_SMEM <= 232448, "shared memory budget");

template <int TAIL>
__global__ void __launch_bounds__(GP_THREADS, 1)
gemm_persist_kernel(const __grid_constant__ GemmMaps maps, const GemmProgram prog, const GemmEpilogue epi,
                    int total_tiles) {
  constexpr int S = GP_S;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + S * GP_STAGE;
  float* tailw = reinterpret_cast<float*>(staging + GP_STAGING);          // [S][4 columns][32 k]
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + GP_STAGING + S * 512);
  uint64_t* full = bars;              // [S] TMA bytes landed
  uint64_t* conv = bars + S;          // [S] A operand written to tensor memory
  uint64_t* empty = bars + 2 * S;     // [S] MMAs reading the stage retired
  uint64_t* accfull = bars + 3 * S;   // [2] accumulator of a tile complete
  uint64_t* accfree = accfull + 2;    // [2] epilogue warps have drained the accumulator
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfree + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  const int KB = prog.kblocks;
  const int n_tiles = prog.n_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&conv[i], 128); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&accfull[i], 1); mbar_init(&accfree[i], 128); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  auto stage_a = [&](int s) { return smem + s * GP_STAGE; };
  auto stage_b = [&](int s) { return smem + s * GP_STAGE + A_TILE_BYTES; };
  // wait-time accounting (obman_debug_trace + OBMAN_GEMM_PERSIST=2, scripts/trace_gemm_persist.py): cycles every warp
  // role spends blocked on each of its barriers, one 16-slot record per CTA
  const bool timed = epi.trace != nullptr;
  const long long t_start = clock64();

  if (warp == 0) {
    // ===== TMA producer (TAIL: all lanes unpack the tail-column weights of the K block, lane = k) =====
    if (TAIL || lane == 0) {
      if (lane == 0) {
        tma_prefetch_desc(&maps.a[0]);
        tma_prefetch_desc(&maps.b);
      }
      int g = 0;
      long long w_empty = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * GP_BN;
        const bool tail_on = TAIL && n0 == 0;
        for (int kb = 0; kb < KB; ++kb, ++g) {
          const int s = g % S;
          const uint32_t ph = (g / S) & 1;
          float tw[4] = {0.f, 0.f, 0.f, 0.f};
          if (TAIL && tail_on) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (e < prog.n_tail) {
                const unsigned short* rowp =
                    reinterpret_cast<const unsigned short*>(prog.tail_w + e * prog.tail_ldw + (long long)kb * 128);
                tw[e] = __uint_as_float((uint32_t)__ldg(rowp + lane) << 16) +
                        __uint_as_float((uint32_t)__ldg(rowp + 32 + lane) << 16);
              }
            }
          }
          if (lane == 0) mbar_wait_timed(&empty[s], ph ^ 1, timed, w_empty);
          if (TAIL) {
            __syncwarp();
            if (tail_on) {
#pragma unroll
              for (int e = 0; e < 4; ++e) tailw[(s * 4 + e) * 32 + lane] = tw[e];
            }
            __syncwarp();   // lane 0's arrive on full[s] publishes the whole warp's stores
            if (lane != 0) continue;
          }
          mbar_arrive_expect_tx(&full[s], A_TILE_BYTES + GP_B_TILE);
          tma_load_2d(stage_a(s), &maps.a[0], &full[s], kb * BK, m0);
          tma_load_2d(stage_b(s), &maps.b, &full[s], kb * BK, n0);
        }
      }
      if (lane == 0) trace_put(epi, 1, w_empty);
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = umma_idesc_bf16(BM, GP_BN);
    int g = 0, ti = 0;
    long long w_full = 0, w_conv = 0, w_accfree = 0;
    const long long t_loop = clock64();
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const int acc = ti & 1;
      mbar_wait_timed(&accfree[acc], ((ti >> 1) & 1) ^ 1, timed, w_accfree);
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(acc * GP_BN);
      for (int kb = 0; kb < KB; ++kb, ++g) {
        const int s = g % S;
        const uint32_t ph = (g / S) & 1;
        mbar_wait_timed(&full[s], ph, timed, w_full);
        mbar_wait_timed(&conv[s], ph, timed, w_conv);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t b = smem_u32(stage_b(s));
          const uint32_t ta = tmem_base + (uint32_t)(2 * GP_BN + 32 * s);
          // every 128-byte B row holds 32 bf16 hi then 32 bf16 lo of this K block; K = 16 per MMA
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t db = umma_desc(b + k * 32, 16, 1024, 2);
            const uint64_t dbl = umma_desc(b + 64 + k * 32, 16, 1024, 2);
            const uint32_t ta_hi = ta + k * 8, ta_lo = ta + 16 + k * 8;
            umma_f16_ts(d, ta_lo, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_f16_ts(d, ta_hi, dbl, idesc, 1u);
            umma_f16_ts(d, ta_hi, db, idesc, 1u);
          }
          umma_commit(&empty[s]);
          if (kb == KB - 1) umma_commit(&accfull[acc]);
        }
        __syncwarp();
      }
    }
    if (lane == 0) {
      trace_put(epi, 2, w_full); trace_put(epi, 3, w_conv); trace_put(epi, 4, w_accfree);
      trace_put(epi, 5, clock64() - t_loop); trace_put(epi, 12, ti);
    }
  } else if (warp < 6) {
    // ===== splitters: A row r (32 fp32 along K, 8 swizzled 16-byte chunks) -> hi / lo -> tensor memory =====
    const int q = warp & 3;
    const int r = q * 32 + lane;       // row of the tile == TMEM lane
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    int g = 0;
    long long w_full = 0, w_st = 0;
    const long long t_loop = clock64();
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * GP_BN;
      const bool tail_on = TAIL && n0 == 0;
      float ext[4] = {0.f, 0.f, 0.f, 0.f};
      for (int kb = 0; kb < KB; ++kb, ++g) {
        const int s = g % S;
        const uint32_t ph = (g / S) & 1;
        mbar_wait_timed(&full[s], ph, timed, w_full);
        const uint32_t row = smem_u32(stage_a(s)) + r * 128;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = lds_v4(row + ((j ^ (r & 7)) << 4));
          split_bf16x2(v.x, v.y, hi[2 * j], lo[2 * j]);
          split_bf16x2(v.z, v.w, hi[2 * j + 1], lo[2 * j + 1]);
          if (TAIL && tail_on) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (e < prog.n_tail) {
                const float4 w = *reinterpret_cast<const float4*>(tailw + (s * 4 + e) * 32 + 4 * j);   // broadcast
                ext[e] = fmaf(v.w, w.w, fmaf(v.z, w.z, fmaf(v.y, w.y, fmaf(v.x, w.x, ext[e]))));
              }
            }
          }
        }
        const uint32_t dst = lane_base + (uint32_t)(2 * GP_BN + 32 * s);
        const long long t_st = timed ? clock64() : 0;
        tmem_st_32x16(dst, hi);
        tmem_st_32x16(dst + 16, lo);
        tmem_st_wait();
        if (timed) w_st += clock64() - t_st;
        tc_fence_before();
        mbar_arrive(&conv[s]);
      }
      if (TAIL && tail_on && m0 + r < prog.M) {
        // tail columns: same epilogue as the tiles (alpha, bias, addend, ReLU, mask), one row per thread
        const long long row_off = (long long)(m0 + r) * epi.ld;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (e < prog.n_tail) {
            const int col = prog.N + e;
            float y = epi.alpha * ext[e];
            if (epi.bias) y += __ldg(epi.bias + col);
            if (epi.addend) y += epi.addend[row_off + col];
            if (epi.relu) y = fmaxf(y, 0.f);
            if (epi.mask_src) y = epi.mask_src[row_off + col] > 0.f ? y : 0.f;
            if (epi.accumulate) atomicAdd(epi.out + row_off + col, y);
            else epi.out[row_off + col] = y;
          }
        }
      }
    }
    if (threadIdx.x == 64) { trace_put(epi, 6, w_full); trace_put(epi, 7, clock64() - t_loop); trace_put(epi, 8, w_st); }
  } else {
    // ===== epilogue warps 6-9: TMEM lane quadrant = warp % 4 =====
    const int q = warp & 3;
    const int r = q * 32 + lane;
    int ti = 0;
    long long w_accfull = 0;
    const long long t_loop = clock64();
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const int m0 = (tile / n_tiles) * BM, n0 = (tile % n_tiles) * GP_BN;
      const int acc = ti & 1;
      if (timed) {   // diagnostics: time blocked on the accumulator (the epilogue function waits again, instantly)
        const long long t0 = clock64();
        mbar_wait(&accfull[acc], (uint32_t)((ti >> 1) & 1));
        w_accfull += clock64() - t0;
      }
      gemm_epilogue<GP_BN, 0>(staging, tmem_base + (uint32_t)(acc * GP_BN), &accfull[acc], prog, epi, m0, n0, 0, 0, 0, q,
                              lane, r, (uint32_t)((ti >> 1) & 1));
      tc_fence_before();
      mbar_arrive(&accfree[acc]);
    }
    if (warp == 6 && lane == 0) { trace_put(epi, 9, w_accfull); trace_put(epi, 10, clock64() - t_loop); }
  }
  if (threadIdx.x == 0) trace_put(epi, 0, clock64() - t_start);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <int TAIL>
static int launch_persist(const GemmMaps& maps, const GemmProgram& prog, const GemmEpilogue& epi, int total_tiles,
                          cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_persist_kernel<TAIL>, cudaFuncAttributeMaxDynamicSharedMemorySize, GP_SMEM);
    if (e != cudaSuccess) {
      set_error("gemm_persist: cudaFuncSetAttribute(%d bytes) failed: %s", GP_SMEM, cudaGetErrorString(e));
      return OBMAN_ERR_CUDA;
    }
    attr = true;
  }
  const int grid = total_tiles < num_sms() ? total_tiles : num_sms();
  gemm_persist_kernel<TAIL><<<grid, GP_THREADS, GP_SMEM, st>>>(maps, prog, epi, total_tiles);
  return check_launch("gemm_persist_kernel");
}

// Called by obman_gemm for packed-bf16 problems with 128-wide column tiles (maps.b box = 128 rows).  Returns 0 when the
// problem is left to gemm_tc_kernel, 1 when launched, < 0 on error.
int try_gemm_persist(const GemmMaps& maps, const GemmProgram& prog, const GemmEpilogue& epi, long long m_tiles,
                     cudaStream_t st) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("OBMAN_GEMM_PERSIST");   // 0: one CTA per tile (gemm_tc_kernel) everywhere; 2: also while tracing
    on = e ? atoi(e) : 1;
  }
  if (epi.trace != nullptr && on != 2) return 0;   // obman_debug_trace records gemm_tc_kernel's CTA phases by default
  const long long total = m_tiles * prog.n_tiles;
  // Worth it when every SM gets several tiles and the reduction is short (long-K tiles amortise their own set-up).
  // Measured at B = 256 (gpurun r2y/r2z, profiles/gemm_persist_trace_r2.txt): 655872 x 128 x 257 0.246 ms against
  // 0.295 ms with one CTA per tile; with two or more column tiles per row tile this kernel is SLOWER (1.04 vs 0.86 ms
  // for N = 257, K = 515): the loop is paced by tensor-memory traffic (~830 clk per K block: 16 KB of tcgen05.st, 24 KB
  // of A reads by the MMAs, the epilogue's tcgen05.ld) either way, and two co-resident CTAs per SM hide each other's
  // barrier round trips better than one persistent CTA does.  OBMAN_GEMM_PERSIST=2 forces it for every eligible shape.
  if (!on || total < 4LL * num_sms() || total > 0x7fffffffLL || prog.kblocks > 24) return 0;
  if (on != 2 && (prog.n_tiles > 1 || prog.n_tail)) return 0;
  // one or two K blocks per tile (the decoder's K = 3 data gradient): the tile is all epilogue, and two co-resident
  // CTAs with four epilogue warps each drain it faster than one persistent CTA (0.29 vs 0.21 ms at B = 256)
  if (on != 2 && prog.kblocks < 4) return 0;
  const int rc = prog.n_tail ? launch_persist<1>(maps, prog, epi, (int)total, st)
                             : launch_persist<0>(maps, prog, epi, (int)total, st);
  return rc == OBMAN_OK ? 1 : (rc < 0 ? rc : -1);
}

}  // namespace obman
