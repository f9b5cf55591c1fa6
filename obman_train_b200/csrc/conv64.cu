// Persistent tcgen05 convolution kernel for the 64-output-channel layers of ResNet-18 (the 7x7/2 stem in its
// 4-tap space-to-depth form and the four 3x3 convolutions of layer1, fprop and dgrad;
// mano_train/networks/bases/resnet.py:25-54,154-171).  These layers were the furthest from the tensor-core roofline
// (27-44 % pipe, 24-30 % of a step's tensor time): N = 64 MMAs issue at 54 % of the N >= 128 rate, the A tile was
// re-fetched from L2 once per tap (operand delivery ~53 B/clk/SM, the fabric limit), and a 128-pixel tile spends 38 %
// of its CTA lifetime in prologue and epilogue.  This kernel removes all three:
//   * ONE CTA per SM loops over output tiles (static round-robin): barrier init, TMEM allocation and the weight fetch
//     happen once per CTA, not once per tile;
//   * the WHOLE packed weight matrix (<= 18 (tap, K-block) tiles of 8 KB) stays resident in shared memory, landed in the
//     stacked layout (rows 0-63 = bf16 hi, rows 64-127 = bf16 lo of the same K block), so one N = 128 MMA yields
//     a_hi*b_hi | a_hi*b_lo at the full rate and one N = 64 MMA adds a_lo*b_hi: 124 instead of 180 cycles per K = 16;
//   * the input arrives as ONE halo box per 32 input channels (TMA, zero fill at image borders); the splitter warps
//     convert it IN PLACE to bf16 hi | lo once, then every tap is a shifted row read + tcgen05.st into a ring of A
//     stages in tensor memory: input bytes per tile drop ~6x, the per-tap work of a splitter thread is 8 LDS + 2 STTM;
//   * the accumulator is double-buffered in tensor memory (2 x 128 columns) and drained by eight dedicated epilogue warps
//     (bias / BN shift, residual add, ReLU, ReLU-mask, coalesced stores through a shared-memory transpose) while the MMA
//     warp is already working on the next tile.
// Warp roles: 0 = TMA producer, 1 = MMA issuer, then NS sets of four splitter warps, then eight epilogue warps (two per
// TMEM lane quadrant, 32 output channels each).  Tensor memory: columns 0-255 accumulators, 256-511 eight A stages of 32 columns.
#include <stdlib.h>

#include "gemm_shared.cuh"

namespace obman {

extern long long* g_trace;
extern long long g_trace_cap;

constexpr int P64_HALO_BYTES = 25600;   // up to 200 halo pixels x 32 channels fp32
constexpr int P64_RA_MAX = 6;           // halo boxes in flight (as many as fit next to the resident weights, >= 2)
constexpr int P64_WT_TILE = 8192;       // 128 rows (64 hi + 64 lo) x 64 B
constexpr int P64_MAX_WT = 18;          // (tap, K-block) weight tiles resident in shared memory
constexpr int P64_C = 8;                // A stages in tensor memory
constexpr int P64_SMEM_LIMIT = 232448;  // 227 KB of dynamic shared memory per CTA
template <int NS, int EW> struct P64Threads { static constexpr int value = 32 * (2 + 4 * NS + EW); };
// shared memory: [weights: n_iters x 8 KB][halo ring: RA x halo_stride][staging: EW x 2 KB][barriers], RA and
// halo_stride (the box rounded up to 1 KB: swizzle phase) chosen on the host
__host__ __device__ inline int p64_smem_bytes(int n_wt, int ra, int halo_stride, int ew) {
  return n_wt * P64_WT_TILE + ra * halo_stride + ew * 2048 + 1024 /*align*/ + 512 /*barriers*/;
}

__device__ __forceinline__ void named_barrier_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

__device__ __forceinline__ void tile_coords(const GemmProgram& prog, int tile, int& n_img0, int& h0, int& w0) {
  const int tw_i = tile % prog.tiles_w; tile /= prog.tiles_w;
  const int th_i = tile % prog.tiles_h; tile /= prog.tiles_h;
  n_img0 = tile * prog.TN;
  h0 = th_i * prog.TH;
  w0 = tw_i * prog.TW;
}

// NS = sets of four splitter warps; iteration gi (= one tap of one K block) belongs to set gi % NS, so NS tcgen05.st
// round trips are in flight per TMEM lane quadrant (one set alone is latency bound: LDS -> STTM -> wait::st -> arrive).
// EW = epilogue warps: 8 (two per TMEM lane quadrant, 32 output channels each) or 4 (64 channels each).
template <int NS, int EW>
__global__ void __launch_bounds__(P64Threads<NS, EW>::value, 1)
conv64_persistent_kernel(const __grid_constant__ GemmMaps maps, const GemmProgram prog, const GemmEpilogue epi,
                         int total_tiles) {
  constexpr int C = P64_C;
  const int RA = prog.halo_ring;
  const int T = prog.num_taps, KB = prog.kblocks;
  const int n_iters = T * KB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wt = smem;
  uint8_t* halo_base = smem + n_iters * P64_WT_TILE;
  uint8_t* staging = halo_base + RA * prog.halo_stride;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + EW * 2048);
  uint64_t* wfull = bars;                    // weights landed
  uint64_t* afull = bars + 1;                // [RA] halo box landed
  uint64_t* afree = afull + P64_RA_MAX;      // [RA] splitter warps are done with the halo box
  uint64_t* conv = afree + P64_RA_MAX;       // [C]  A stage written to tensor memory
  uint64_t* empty = conv + C;                // [C]  MMAs reading the A stage retired
  uint64_t* accfull = empty + C;             // [2]  accumulator of a tile complete
  uint64_t* accfree = accfull + 2;           // [2]  epilogue warps have drained the accumulator
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfree + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform by construction
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(wfull, 1);
    for (int i = 0; i < RA; ++i) { mbar_init(&afull[i], 1); mbar_init(&afree[i], 128 * NS); }
    for (int i = 0; i < C; ++i) { mbar_init(&conv[i], 128); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&accfull[i], 1); mbar_init(&accfree[i], 32 * EW); }
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
  auto halo = [&](int i) { return halo_base + i * prog.halo_stride; };
  const bool timed = epi.trace != nullptr;
  const long long t_start = clock64();

  if (warp == 0) {
    // ===== TMA producer: the weights once, then one halo box per (tile, K block) =====
    if (lane == 0) {
      tma_prefetch_desc(&maps.a[0]);
      tma_prefetch_desc(&maps.b);
      mbar_arrive_expect_tx(wfull, (uint32_t)n_iters * P64_WT_TILE);
      for (int kb = 0; kb < KB; ++kb)
        for (int tap = 0; tap < T; ++tap) {
          uint8_t* dst = wt + (kb * T + tap) * P64_WT_TILE;
          const int kc = prog.tap_bk[tap] + kb * BK;
          tma_load_2d(dst, &maps.b, wfull, kc, 0);             // 64 rows x 64 B of bf16 hi
          tma_load_2d(dst + 4096, &maps.b, wfull, kc + 16, 0); // 64 rows x 64 B of bf16 lo
        }
      int g = 0;
      long long w_afree = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int n_img0, h0, w0;
        tile_coords(prog, tile, n_img0, h0, w0);
        for (int kb = 0; kb < KB; ++kb, ++g) {
          const int a = g % RA;
          mbar_wait_timed(&afree[a], ((g / RA) & 1) ^ 1, timed, w_afree);
          mbar_arrive_expect_tx(&afull[a], (uint32_t)prog.halo_bytes);
          tma_load_4d(halo(a), &maps.a[0], &afull[a], kb * BK, w0 + prog.halo_dw0, h0 + prog.halo_dh0, n_img0);
        }
      }
      trace_put(epi, 1, w_afree);
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = umma_idesc_bf16(BM, 64);
    const uint32_t idesc_wide = umma_idesc_bf16(BM, 128);
    long long w_conv = 0, w_accfree = 0, w_wfull = 0;
    mbar_wait_timed(wfull, 0, timed, w_wfull);
    const long long t_loop = clock64();
    int gi = 0, ti = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const int acc = ti & 1;
      mbar_wait_timed(&accfree[acc], ((ti >> 1) & 1) ^ 1, timed, w_accfree);
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(acc * 128);
      for (int it = 0; it < n_iters; ++it, ++gi) {
        const int c = gi % C;
        mbar_wait_timed(&conv[c], (gi / C) & 1, timed, w_conv);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t b = smem_u32(wt + it * P64_WT_TILE);
          const uint32_t ta = tmem_base + (uint32_t)(256 + 32 * c);
          if (!(prog.debug_skip & 1) || it == 0) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t db = umma_desc(b + k * 32, 16, 512, 4);   // 128 rows x 64 B, SW64: 8-row groups 512 B apart
            umma_f16_ts(d, ta + k * 8, db, idesc_wide, (it > 0 || k > 0) ? 1u : 0u);   // [a_hi*b_hi | a_hi*b_lo]
            if (!(prog.debug_skip & 8)) umma_f16_ts(d, ta + 16 + k * 8, db, idesc, 1u);   // columns 0-63 += a_lo*b_hi
          }
          }
          umma_commit(&empty[c]);
          if (it == n_iters - 1) umma_commit(&accfull[acc]);
        }
        __syncwarp();
      }
    }
    if (lane == 0) {
      trace_put(epi, 13, w_conv); trace_put(epi, 7, w_accfree); trace_put(epi, 8, clock64() - t_loop);
      trace_put(epi, 11, w_wfull); trace_put(epi, 12, ti);
    }
  } else if (warp < 2 + 4 * NS) {
    // ===== splitters: halo box -> bf16 hi | lo in place, then one shifted row per tap -> tensor memory =====
    const int q = warp & 3;
    const int set = (warp - 2) >> 2;
    const int r = q * 32 + lane;                 // tile row == TMEM lane
    const int sid = threadIdx.x - 64;            // 0 .. 128 NS - 1
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const int tw = r % prog.TW, th = (r / prog.TW) % prog.TH, tn = r / (prog.TW * prog.TH);
    const int hp0 = (tn * prog.halo_h + th) * prog.halo_w + tw;
    // the packed row of halo pixel hp: 16 words of bf16 hi pairs (chunks 0-3), 16 words of lo pairs (chunks 4-7)
    auto load_row = [&](uint32_t box, int hp, uint32_t* hi, uint32_t* lo) {
      const uint32_t row = box + (uint32_t)hp * 128u;
      const int sw = hp & 7;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 h4 = lds_v4(row + ((j ^ sw) << 4));
        const float4 l4 = lds_v4(row + (((j + 4) ^ sw) << 4));
        hi[4 * j] = __float_as_uint(h4.x); hi[4 * j + 1] = __float_as_uint(h4.y);
        hi[4 * j + 2] = __float_as_uint(h4.z); hi[4 * j + 3] = __float_as_uint(h4.w);
        lo[4 * j] = __float_as_uint(l4.x); lo[4 * j + 1] = __float_as_uint(l4.y);
        lo[4 * j + 2] = __float_as_uint(l4.z); lo[4 * j + 3] = __float_as_uint(l4.w);
      }
    };
    int g = 0;
    long long w_afull = 0, w_empty = 0, t_split = 0;
    const long long t_loop = clock64();
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < KB; ++kb, ++g) {
        const int a = g % RA;
        mbar_wait_timed(&afull[a], (g / RA) & 1, timed, w_afull);
        const long long t_s0 = timed ? clock64() : 0;
        const uint32_t box = smem_u32(halo(a));
        for (int p = sid; p < ((prog.debug_skip & 4) ? 0 : prog.halo_pix); p += 128 * NS) {
          const uint32_t row = box + (uint32_t)p * 128u;
          const int sw = p & 7;
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 v = lds_v4(row + ((j ^ sw) << 4));
            split_bf16x2(v.x, v.y, hi[2 * j], lo[2 * j]);
            split_bf16x2(v.z, v.w, hi[2 * j + 1], lo[2 * j + 1]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            sts_v4(row + ((j ^ sw) << 4), hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
            sts_v4(row + (((j + 4) ^ sw) << 4), lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
          }
        }
        named_barrier_sync(1, 128 * NS);   // every pixel of the box is converted before any tap row is read
        if (timed) t_split += clock64() - t_s0;
        const int gi0 = g * T;             // iteration index of tap 0 of this box
        int tap = (set - gi0 % NS + NS) % NS;   // first tap of this box that belongs to this set
        uint32_t hi[16], lo[16];
        if (tap < T && !(prog.debug_skip & 2)) load_row(box, hp0 + prog.tap_delta[tap], hi, lo);
        for (; tap < T; tap += NS) {
          const int gi = gi0 + tap;
          const int c = gi % C;
          mbar_wait_timed(&empty[c], ((gi / C) & 1) ^ 1, timed, w_empty);
          tc_fence_after();
          const uint32_t dst = lane_base + (uint32_t)(256 + 32 * c);
          tmem_st_32x16(dst, hi);
          tmem_st_32x16(dst + 16, lo);
          // the next row's shared-memory reads are issued underneath the tensor-memory store round trip
          uint32_t nhi[16], nlo[16];
          const bool more = tap + NS < T;
          if (more && !(prog.debug_skip & 2)) load_row(box, hp0 + prog.tap_delta[tap + NS], nhi, nlo);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&conv[c]);
          if (more) {
#pragma unroll
            for (int j = 0; j < 16; ++j) { hi[j] = nhi[j]; lo[j] = nlo[j]; }
          }
        }
        fence_proxy_async_smem();   // the box was rewritten through the generic proxy; the next TMA write follows it
        mbar_arrive(&afree[a]);
      }
    }
    if (sid == 0) {
      trace_put(epi, 2, w_afull); trace_put(epi, 3, w_empty); trace_put(epi, 4, t_split);
      trace_put(epi, 5, clock64() - t_loop);
    }
  } else {
    // ===== epilogue: the last EW warps, TMEM lane quadrant = warp % 4, 64 * 4 / EW output channels per warp =====
    constexpr int COLS = 64 * 4 / EW;
    const int q = warp & 3;
    const int part = (warp - (2 + 4 * NS)) >> 2;
    const int r = q * 32 + lane;
    int ti = 0;
    const long long t_loop = clock64();
    long long w_accfull = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      int n_img0, h0, w0;
      tile_coords(prog, tile, n_img0, h0, w0);
      const int acc = ti & 1;
      if (timed) {   // diagnostics: time blocked on the accumulator (the epilogue function waits again, instantly)
        const long long t0 = clock64();
        mbar_wait(&accfull[acc], (uint32_t)((ti >> 1) & 1));
        w_accfull += clock64() - t0;
      }
      gemm_epilogue_stacked<64>(staging + part * 8192, tmem_base + (uint32_t)(acc * 128), &accfull[acc], prog, epi, 0, 0,
                                n_img0, h0, w0, q, lane, r, (uint32_t)((ti >> 1) & 1), part * COLS, part * COLS + COLS);
      tc_fence_before();
      mbar_arrive(&accfree[acc]);
    }
    if (warp == 2 + 4 * NS && lane == 0) { trace_put(epi, 9, w_accfull); trace_put(epi, 10, clock64() - t_loop); }
  }
  if (threadIdx.x == 0) trace_put(epi, 0, clock64() - t_start);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- second generation: operand roles swapped ----------------------------------------------------------------------
// Measured on the kernel above (scripts/trace_conv64.py, profiles/conv64_trace_r2.txt): with the A operand in tensor
// memory every tap costs 16 KB of tcgen05.st plus 16 KB of A reads by the MMAs, and that tensor-memory traffic - not the
// MMA rate - paces the loop (~650 clk per (tap, K block) against 248 clk of MMA work).  Here NOTHING is written per tap:
//   * A operand (M = 128) = the resident stacked weight tile of a (tap, K block): rows 0-63 = bf16 hi, 64-127 = bf16 lo
//     of the 64 output channels, read from shared memory (K-major, SW64) - the same bytes the first generation used as B;
//   * B operand (N = a run of consecutive halo pixels) = the activations, read from shared memory through a NON-swizzled
//     K-major descriptor.  The halo box is converted once, in place, from TMA's fp32 [pixel][32 ch] rows into eight
//     "planes" [16-byte K chunk][pixel][8 bf16] (4 planes of hi, 4 of lo).  In that layout the 8-row x 16-byte core
//     matrices of consecutive pixels are contiguous (stride-between-row-groups = 128 B), so ANY pixel offset is a valid
//     descriptor start address: a tap is a different start address, nothing else.  The run covers the TH x TW output
//     tile as flattened halo positions ((TH-1) * halo_w + TW columns, rounded up to 16); the halo columns in between are
//     computed and discarded by the epilogue (89 % of the columns are used for the 8 x 16 tile of a 3x3 convolution).
//   * two MMAs per K = 16 step, both with A = [w_hi; w_lo]: B = a_hi, then B = a_lo.  Accumulator lanes 0-63 hold
//     w_hi * (a_hi + a_lo), lanes 64-127 hold w_lo * (a_hi + a_lo); the epilogue adds the two halves (all four bf16
//     products, one more than the 3xBF16 scheme needs).
// Warp roles (448 threads): 0 = TMA producer, 1 = MMA issuer, 2-5 = converters, 6-13 = epilogue (two groups of four).  Tensor memory: two
// accumulators of N <= 256 columns.
constexpr int P64V2_THREADS = 448;

__global__ void __launch_bounds__(P64V2_THREADS, 1)
conv64_v2_kernel(const __grid_constant__ GemmMaps maps, const GemmProgram prog, const GemmEpilogue epi, int total_tiles) {
  const int RA = prog.halo_ring;
  const int T = prog.num_taps, KB = prog.kblocks;
  const int n_iters = T * KB;
  const int NRUN = prog.run_cols;            // MMA N: columns of the accumulator (multiple of 16, <= 256)
  const int P = prog.halo_pix;               // pixels of the halo box = rows of every plane
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wt = smem;
  uint8_t* halo_base = smem + n_iters * P64_WT_TILE;
  uint8_t* staging = halo_base + RA * prog.halo_stride;      // 2 groups x [2 parts][32 pixels][64 channels] fp32 + column table
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + 32768 + 2048);
  uint64_t* wfull = bars;                    // weights landed
  uint64_t* afull = bars + 1;                // [RA] halo box landed (TMA)
  uint64_t* bconv = afull + P64_RA_MAX;      // [RA] halo box converted to bf16 planes
  uint64_t* afree = bconv + P64_RA_MAX;      // [RA] MMAs reading the box retired
  uint64_t* accfull = afree + P64_RA_MAX;    // [2]  accumulator of a tile complete
  uint64_t* accfree = accfull + 2;           // [2]  epilogue warps have drained the accumulator
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accfree + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(wfull, 1);
    for (int i = 0; i < RA; ++i) { mbar_init(&afull[i], 1); mbar_init(&bconv[i], 128); mbar_init(&afree[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&accfull[i], 1); mbar_init(&accfree[i], 256); }
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
  auto halo = [&](int i) { return halo_base + i * prog.halo_stride; };
  const bool timed = epi.trace != nullptr;
  const long long t_start = clock64();

  if (warp == 0) {
    // ===== TMA producer: the weights once, then one halo box per (tile, K block) =====
    if (lane == 0) {
      tma_prefetch_desc(&maps.a[0]);
      tma_prefetch_desc(&maps.b);
      mbar_arrive_expect_tx(wfull, (uint32_t)n_iters * P64_WT_TILE);
      for (int kb = 0; kb < KB; ++kb)
        for (int tap = 0; tap < T; ++tap) {
          uint8_t* dst = wt + (kb * T + tap) * P64_WT_TILE;
          const int kc = prog.tap_bk[tap] + kb * BK;
          tma_load_2d(dst, &maps.b, wfull, kc, 0);             // 64 rows x 64 B of bf16 hi
          tma_load_2d(dst + 4096, &maps.b, wfull, kc + 16, 0); // 64 rows x 64 B of bf16 lo
        }
      int g = 0;
      long long w_afree = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        int n_img0, h0, w0;
        tile_coords(prog, tile, n_img0, h0, w0);
        for (int kb = 0; kb < KB; ++kb, ++g) {
          const int a = g % RA;
          mbar_wait_timed(&afree[a], ((g / RA) & 1) ^ 1, timed, w_afree);
          mbar_arrive_expect_tx(&afull[a], (uint32_t)prog.halo_bytes);
          tma_load_4d(halo(a), &maps.a[0], &afull[a], kb * BK, w0 + prog.halo_dw0, h0 + prog.halo_dh0, n_img0);
        }
      }
      trace_put(epi, 1, w_afree);
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = umma_idesc_bf16(BM, NRUN);
    const uint32_t plane = (uint32_t)P * 16u;            // bytes between consecutive 16-byte K chunks of a pixel row
    mbar_wait(wfull, 0);
    int g = 0, ti = 0;
    long long w_bconv = 0, w_accfree = 0;
    const long long t_loop = clock64();
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      const int acc = ti & 1;
      mbar_wait_timed(&accfree[acc], ((ti >> 1) & 1) ^ 1, timed, w_accfree);
      tc_fence_after();
      const uint32_t d = tmem_base + (uint32_t)(acc * 256);
      for (int kb = 0; kb < KB; ++kb, ++g) {
        const int a = g % RA;
        mbar_wait_timed(&bconv[a], (g / RA) & 1, timed, w_bconv);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t box = smem_u32(halo(a));
          for (int tap = 0; tap < T; ++tap) {
            const uint32_t w = smem_u32(wt + (kb * T + tap) * P64_WT_TILE);
            const uint32_t brow = box + (uint32_t)prog.tap_delta[tap] * 16u;   // first pixel of the run for this tap
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              // A: 128 rows x 64 B, SW64, K-major: 8-row groups 512 B apart; k-th 32-byte K slice
              const uint64_t da = umma_desc(w + k * 32, 16, 512, 4);
              // B: no swizzle, K-major: core matrix = 8 pixels x 16 B contiguous; next 8 pixels +128 B (SBO),
              // next 16-byte K chunk +plane (LBO); hi planes 0-3, lo planes 4-7; K = 16 -> chunks 2k, 2k+1
              const uint64_t db_hi = umma_desc(brow + (uint32_t)(2 * k) * plane, plane, 128, 0);
              const uint64_t db_lo = umma_desc(brow + (uint32_t)(4 + 2 * k) * plane, plane, 128, 0);
              umma_f16_ss(d, da, db_hi, idesc, (kb > 0 || tap > 0 || k > 0) ? 1u : 0u);
              umma_f16_ss(d, da, db_lo, idesc, 1u);
            }
          }
          umma_commit(&afree[a]);                       // the box may be overwritten once these MMAs have read it
          if (kb == KB - 1) umma_commit(&accfull[acc]);
        }
        __syncwarp();
      }
    }
    if (lane == 0) {
      trace_put(epi, 13, w_bconv); trace_put(epi, 7, w_accfree); trace_put(epi, 8, clock64() - t_loop); trace_put(epi, 12, ti);
    }
  } else if (warp < 6) {
    // ===== converters: fp32 [pixel][32 ch] (SW128 rows as TMA wrote them) -> eight bf16 planes, in place =====
    const int sid = threadIdx.x - 64;            // 0 .. 127
    int g = 0;
    long long w_afull = 0;
    const long long t_loop = clock64();
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < KB; ++kb, ++g) {
        const int a = g % RA;
        mbar_wait_timed(&afull[a], (g / RA) & 1, timed, w_afull);
        const uint32_t box = smem_u32(halo(a));
        uint32_t hi[2][16], lo[2][16];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int p = sid + 128 * e;
          if (p < P) {
            const uint32_t row = box + (uint32_t)p * 128u;
            const int sw = p & 7;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 v = lds_v4(row + ((j ^ sw) << 4));
              split_bf16x2(v.x, v.y, hi[e][2 * j], lo[e][2 * j]);
              split_bf16x2(v.z, v.w, hi[e][2 * j + 1], lo[e][2 * j + 1]);
            }
          }
        }
        named_barrier_sync(1, 128);   // every row has been read before the planes overwrite the box
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int p = sid + 128 * e;
          if (p < P) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              sts_v4(box + (uint32_t)(j * P + p) * 16u, hi[e][4 * j], hi[e][4 * j + 1], hi[e][4 * j + 2], hi[e][4 * j + 3]);
              sts_v4(box + (uint32_t)((4 + j) * P + p) * 16u, lo[e][4 * j], lo[e][4 * j + 1], lo[e][4 * j + 2], lo[e][4 * j + 3]);
            }
          }
        }
        fence_proxy_async_smem();     // the planes are read by the tensor core (async proxy)
        mbar_arrive(&bconv[a]);
      }
    }
    if (sid == 0) { trace_put(epi, 2, w_afull); trace_put(epi, 5, clock64() - t_loop); }
  } else {
    // ===== epilogue: warps 6..13 = two groups of four (TMEM lane quadrant = warp % 4), alternating 32-column rounds =====
    // lanes 0-63 of the accumulator = w_hi part of channels 0-63, lanes 64-127 = w_lo part.  In a round every warp
    // stores its 32 channels x 32 pixel columns into the group's staging tile ([part][pixel][64 channels] fp32), one
    // barrier later the four warps walk the pixels with 16 lanes x float4 per pixel, adding the two parts on the way
    // (256 contiguous bytes per output row and store instruction; four independent pixels per thread).  The column ->
    // pixel mapping does not depend on the tile: it is tabulated once in shared memory, so a round costs one table
    // read per pixel instead of a chain of integer divisions (a lone warp per scheduler runs dependent ALU chains at
    // ~4 clk per instruction: the divisions alone were 800 clk per round, scripts/trace_conv64.py).
    constexpr int RW2 = 32, PE = RW2 / 8;
    const int q = warp & 3;
    const int group = (warp - 6) >> 2;           // 0 | 1
    const int et = (threadIdx.x - 192) & 127;    // 0 .. 127 inside the group
    const int part = q >> 1;                     // 0: lanes 0-63 (hi), 1: lanes 64-127 (lo)
    const int ch = (q & 1) * 32 + lane;          // channel of this TMEM lane
    const int px_sub = et >> 4;                  // pixel (of 8) this thread stores in each eighth of a round
    const int c4 = (et & 15) * 4;                // its four channels
    float* S = reinterpret_cast<float*>(staging + group * 16384);   // [2 parts][32 pixels][64 channels]
    int2* tab = reinterpret_cast<int2*>(staging + 32768);           // [256] column -> {relative offset, th | tw<<8 | tn<<16}
    for (int j = threadIdx.x - 192; j < 256; j += 256) {
      const int hr = j / prog.halo_w, tw = j - hr * prog.halo_w;    // halo row (over all images of the box), column
      const int tn = hr / prog.halo_h, th = hr - tn * prog.halo_h;
      const bool ok = j < NRUN && tw < prog.TW && th < prog.TH && tn < prog.TN;
      tab[j] = make_int2((int)(tn * epi.sN + th * epi.sH + tw * epi.sW), ok ? (th | (tw << 8) | (tn << 16)) : -1);
    }
    named_barrier_sync(4, 256);                  // both groups: the table is complete
    const bool ptr_ok = ((reinterpret_cast<uintptr_t>(epi.out) | reinterpret_cast<uintptr_t>(epi.addend) |
                          reinterpret_cast<uintptr_t>(epi.mask_src)) & 15) == 0;
    // vector path: 16-byte aligned rows (pointers and strides) and four whole channels for this thread
    const bool fast = ptr_ok && ((epi.sN | epi.sH | epi.sW) & 3) == 0 && c4 + 3 < prog.N;
    const float relu_floor = epi.relu ? 0.f : -3.0e38f;
    float bias4[4] = {0.f, 0.f, 0.f, 0.f};
    if (epi.bias) {
#pragma unroll
      for (int e = 0; e < 4; ++e) if (c4 + e < prog.N) bias4[e] = __ldg(epi.bias + c4 + e);
    }
    const int rounds = (NRUN + RW2 - 1) / RW2;
    int ti = 0;
    long long w_accfull = 0;
    const long long t_loop = clock64();
    long long t_phase[5] = {0, 0, 0, 0, 0};
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++ti) {
      int n_img0, h0, w0;
      tile_coords(prog, tile, n_img0, h0, w0);
      const long long base = n_img0 * epi.sN + h0 * epi.sH + w0 * epi.sW;
      const int lim_n = prog.n_img - n_img0, lim_h = prog.h_out - h0, lim_w = prog.w_out - w0;
      const int acc = ti & 1;
      // operands of this group's first round: independent of the accumulator, fetched before waiting for it
      const int rr0 = (group + ti) & 1;
      long long roff[PE];
      bool okp[PE];
      float4 add4[PE], msk4[PE];
      // branch-free per pixel (the four pixels of a thread are independent chains the scheduler can interleave):
      // validity from the tabulated tile coordinates, operand loads predicated
      auto fetch = [&](int rr) {
#pragma unroll
        for (int e = 0; e < PE; ++e) {
          const int2 t = tab[rr * RW2 + px_sub + 8 * e];
          okp[e] = (t.y >= 0) & ((t.y & 255) < lim_h) & (((t.y >> 8) & 255) < lim_w) & ((t.y >> 16) < lim_n) & (c4 < prog.N);
          roff[e] = base + t.x;
        }
        if (fast) {
#pragma unroll
          for (int e = 0; e < PE; ++e) {
            add4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
            msk4[e] = make_float4(1.f, 1.f, 1.f, 1.f);
            if (epi.addend && okp[e]) add4[e] = __ldg(reinterpret_cast<const float4*>(epi.addend + roff[e] + c4));
            if (epi.mask_src && okp[e]) msk4[e] = __ldg(reinterpret_cast<const float4*>(epi.mask_src + roff[e] + c4));
          }
        }
      };
      fetch(rr0);
      mbar_wait_timed(&accfull[acc], (uint32_t)((ti >> 1) & 1), timed, w_accfull);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256);
      // the groups take alternate rounds; which one starts alternates with the tile (odd round counts stay balanced)
      for (int rr = rr0; rr < rounds; rr += 2) {
        const long long tp0 = timed ? clock64() : 0;
        if (rr != rr0) fetch(rr);
        const long long tp1 = timed ? clock64() : 0;
        uint32_t v[RW2];
        tmem_ld_32x32(lane_addr + (uint32_t)(rr * RW2), v);
        tmem_ld_wait();
        if (fast) {
#pragma unroll
          for (int e = 0; e < PE; ++e)
            asm volatile("" : "+f"(msk4[e].x), "+f"(msk4[e].y), "+f"(msk4[e].z), "+f"(msk4[e].w));
        }
#pragma unroll
        for (int i = 0; i < RW2; ++i) S[(part * RW2 + i) * 64 + ch] = __uint_as_float(v[i]);
        const long long tp2 = timed ? clock64() : 0;
        named_barrier_sync(2 + group, 128);
        const long long tp3 = timed ? clock64() : 0;
        if (fast) {
          float4 s_hi[PE], s_lo[PE], o4[PE];
#pragma unroll
          for (int e = 0; e < PE; ++e) {
            s_hi[e] = *reinterpret_cast<const float4*>(S + (px_sub + 8 * e) * 64 + c4);
            s_lo[e] = *reinterpret_cast<const float4*>(S + (RW2 + px_sub + 8 * e) * 64 + c4);
          }
#pragma unroll
          for (int e = 0; e < PE; ++e) {
            float x[4] = {epi.alpha * (s_hi[e].x + s_lo[e].x) + bias4[0] + add4[e].x,
                          epi.alpha * (s_hi[e].y + s_lo[e].y) + bias4[1] + add4[e].y,
                          epi.alpha * (s_hi[e].z + s_lo[e].z) + bias4[2] + add4[e].z,
                          epi.alpha * (s_hi[e].w + s_lo[e].w) + bias4[3] + add4[e].w};
#pragma unroll
            for (int c = 0; c < 4; ++c) x[c] = fmaxf(x[c], relu_floor);
            o4[e].x = msk4[e].x > 0.f ? x[0] : 0.f; o4[e].y = msk4[e].y > 0.f ? x[1] : 0.f;
            o4[e].z = msk4[e].z > 0.f ? x[2] : 0.f; o4[e].w = msk4[e].w > 0.f ? x[3] : 0.f;
          }
#pragma unroll
          for (int e = 0; e < PE; ++e)
            if (okp[e]) *reinterpret_cast<float4*>(epi.out + roff[e] + c4) = o4[e];
        } else {
#pragma unroll 1
          for (int e = 0; e < PE; ++e) {
            if (!okp[e]) continue;
            const long long ro = roff[e];
            const int px = px_sub + 8 * e;
            for (int c = 0; c < 4; ++c) {
              if (c4 + c >= prog.N) break;
              float y = epi.alpha * (S[px * 64 + c4 + c] + S[(RW2 + px) * 64 + c4 + c]) + bias4[c];
              if (epi.addend) y += epi.addend[ro + c4 + c];
              if (epi.relu) y = fmaxf(y, 0.f);
              if (epi.mask_src) y = epi.mask_src[ro + c4 + c] > 0.f ? y : 0.f;
              epi.out[ro + c4 + c] = y;
            }
          }
        }
        const long long tp4 = timed ? clock64() : 0;
        named_barrier_sync(2 + group, 128);   // the staging tile is rewritten by the group's next round
        if (timed) {
          t_phase[0] += tp1 - tp0; t_phase[1] += tp2 - tp1; t_phase[2] += tp3 - tp2; t_phase[3] += tp4 - tp3;
          t_phase[4] += clock64() - tp4;
        }
      }
      tc_fence_before();
      mbar_arrive(&accfree[acc]);
    }
    if (et == 0) { trace_put(epi, 9 + 5 * group, w_accfull); trace_put(epi, 10 + 5 * group, clock64() - t_loop); }
    if (et == 0 && group == 0) {
      trace_put(epi, 3, t_phase[0]); trace_put(epi, 4, t_phase[1]); trace_put(epi, 6, t_phase[2]);
      trace_put(epi, 11, t_phase[3] + t_phase[4]);
    }
  }
  if (threadIdx.x == 0) trace_put(epi, 0, clock64() - t_start);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

static bool conv64_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OBMAN_CONV64");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// Returns 1 when the convolution was launched on the persistent kernel, 0 when the shape does not qualify (the caller
// then takes the generic shifted-box path), a negative error code on failure.
int try_conv64(const float* x, int n_img, int h_in, int w_in, int c_in, long long x_sN, long long x_sH, long long x_sW,
               const float* w, int c_out, int w_slots, int num_taps, const int* tap_dh, const int* tap_dw,
               const int* tap_wslot, float* out, int h_out, int w_out, long long o_sN, long long o_sH, long long o_sW,
               const float* bias, const float* addend, const float* mask_src, int relu, cudaStream_t st) {
  if (!conv64_enabled()) return 0;
  if (c_out > 64 || c_out % 4 != 0 || c_in % 32 != 0 || w_out < 8 || h_out < 1) return 0;
  const int KB = c_in / 32;
  if (num_taps * KB > P64_MAX_WT) return 0;
  int dh0 = tap_dh[0], dh1 = tap_dh[0], dw0 = tap_dw[0], dw1 = tap_dw[0];
  for (int t = 1; t < num_taps; ++t) {
    dh0 = min(dh0, tap_dh[t]); dh1 = max(dh1, tap_dh[t]);
    dw0 = min(dw0, tap_dw[t]); dw1 = max(dw1, tap_dw[t]);
  }
  int TW = 1;
  while (TW * 2 <= w_out && TW * 2 <= 16) TW *= 2;
  int TH = 1;
  while (TH * 2 <= h_out && TW * TH * 2 <= 128) TH *= 2;
  const int TN = 128 / (TW * TH);
  const int HW = TW + dw1 - dw0, HH = TH + dh1 - dh0;
  const int halo_pix = HW * HH * TN;
  if (halo_pix * 128 > P64_HALO_BYTES || HW > 256 || HH > 256 || TN > 256) return 0;
  static int cfg = -1;   // OBMAN_CONV64_CFG = <splitter sets><epilogue warps>: 18, 24 (default), 28, 34
  if (cfg < 0) {
    const char* e = getenv("OBMAN_CONV64_CFG");
    cfg = e ? atoi(e) : 24;
    if (cfg != 18 && cfg != 24 && cfg != 28 && cfg != 34) cfg = 24;
  }
  const int ew = cfg % 10;
  const int halo_stride = (halo_pix * 128 + 1023) / 1024 * 1024;
  int ring = P64_RA_MAX;
  {
    static int ring_cap = -1;
    if (ring_cap < 0) {
      const char* e = getenv("OBMAN_CONV64_RING");   // experiment knob: cap on the halo ring depth
      ring_cap = e ? atoi(e) : P64_RA_MAX;
      if (ring_cap < 2 || ring_cap > P64_RA_MAX) ring_cap = P64_RA_MAX;
    }
    ring = ring_cap;
  }
  while (ring > 2 && p64_smem_bytes(num_taps * KB, ring, halo_stride, ew) > P64_SMEM_LIMIT) --ring;
  const int smem_bytes = p64_smem_bytes(num_taps * KB, ring, halo_stride, ew);
  if (smem_bytes > P64_SMEM_LIMIT) return 0;
  GemmProgram prog;
  memset(&prog, 0, sizeof(prog));
  prog.spatial = 1;
  prog.num_taps = num_taps;
  prog.kblocks = KB;
  prog.N = c_out;
  prog.TN = TN; prog.TH = TH; prog.TW = TW;
  prog.tiles_h = (h_out + TH - 1) / TH;
  prog.tiles_w = (w_out + TW - 1) / TW;
  prog.n_img = n_img; prog.h_out = h_out; prog.w_out = w_out;
  prog.halo_w = HW; prog.halo_h = HH; prog.halo_dw0 = dw0; prog.halo_dh0 = dh0; prog.halo_pix = halo_pix;
  prog.halo_bytes = halo_pix * 128;
  prog.halo_stride = halo_stride;
  prog.halo_ring = ring;
  for (int t = 0; t < num_taps; ++t) {
    prog.tap_bk[t] = (tap_wslot ? tap_wslot[t] : t) * c_in;
    prog.tap_delta[t] = (tap_dh[t] - dh0) * HW + (tap_dw[t] - dw0);
  }
  {
    // diagnostics only (results are wrong when set): bit 0 skip the MMAs, 1 skip the tap row reads, 2 skip the in-place
    // split, 3 skip the narrow MMA - which stage bounds the loop
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("OBMAN_CONV64_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    prog.debug_skip = dbg;
  }
  const long long total = (long long)((n_img + TN - 1) / TN) * prog.tiles_h * prog.tiles_w;
  if (total > 0x7fffffff) return 0;
  GemmMaps maps;
  memset(&maps, 0, sizeof(maps));
  {
    uint64_t dims[4] = {(uint64_t)c_in, (uint64_t)w_in, (uint64_t)h_in, (uint64_t)n_img};
    uint64_t strides[3] = {(uint64_t)x_sW * 4, (uint64_t)x_sH * 4, (uint64_t)x_sN * 4};
    uint32_t box[4] = {BK, (uint32_t)HW, (uint32_t)HH, (uint32_t)TN};
    int rc = make_tensor_map(&maps.a[0], x, 4, dims, strides, box);
    if (rc) return rc;
    uint64_t dimsb[2] = {(uint64_t)w_slots * (uint64_t)c_in, (uint64_t)c_out};
    uint64_t stridesb[1] = {dimsb[0] * 4};
    uint32_t boxb[2] = {16, 64};   // 64-byte-wide boxes: the hi half and the lo half of a packed K block separately
    rc = make_tensor_map(&maps.b, w, 2, dimsb, stridesb, boxb, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }
  GemmEpilogue epi;
  memset(&epi, 0, sizeof(epi));
  epi.out = out; epi.bias = bias; epi.addend = addend; epi.mask_src = mask_src;
  epi.alpha = 1.f; epi.relu = relu; epi.accumulate = 0;
  epi.sN = o_sN; epi.sH = o_sH; epi.sW = o_sW;
  epi.trace = g_trace; epi.trace_cap = g_trace_cap;
  const int grid = (int)min((long long)num_sms(), total);
  {
    // second generation (operands swapped, no per-tap tensor-memory traffic) whenever the flattened run fits
    static int gen = -1;
    if (gen < 0) {
      const char* e = getenv("OBMAN_CONV64_GEN");
      gen = (e && e[0] == '1') ? 1 : 2;
    }
    const int run = (((TN - 1) * HH + TH - 1) * HW + TW + 15) / 16 * 16;
    auto smem2 = [&](int r) { return num_taps * KB * P64_WT_TILE + r * halo_stride + 32768 + 2048 + 1024 + 512; };
    int ring2 = P64_RA_MAX;
    while (ring2 > 2 && smem2(ring2) > P64_SMEM_LIMIT) --ring2;
    if (gen == 2 && run >= 32 && run <= 256 && halo_pix <= 256 && smem2(ring2) <= P64_SMEM_LIMIT) {   // >= 2 epilogue rounds
      prog.run_cols = run;
      prog.halo_ring = ring2;
      static bool attr2 = false;
      if (!attr2) {
        cudaError_t err = cudaFuncSetAttribute(conv64_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P64_SMEM_LIMIT);
        if (err != cudaSuccess) {
          set_error("conv64: cudaFuncSetAttribute(%d bytes) failed: %s", P64_SMEM_LIMIT, cudaGetErrorString(err));
          return OBMAN_ERR_CUDA;
        }
        attr2 = true;
      }
      conv64_v2_kernel<<<grid, P64V2_THREADS, smem2(ring2), st>>>(maps, prog, epi, (int)total);
      int rc2 = check_launch("conv64_v2_kernel");
      return rc2 ? rc2 : 1;
    }
  }
#define OBMAN_P64_CASE(id, NS, EW)                                                                                    \
  if (cfg == id) {                                                                                                    \
    static int configured = 0;                                                                                        \
    if (smem_bytes > configured) {                                                                                    \
      cudaError_t err = cudaFuncSetAttribute(conv64_persistent_kernel<NS, EW>,                                        \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, P64_SMEM_LIMIT);            \
      if (err != cudaSuccess) {                                                                                       \
        set_error("conv64: cudaFuncSetAttribute(%d bytes) failed: %s", P64_SMEM_LIMIT, cudaGetErrorString(err));      \
        return OBMAN_ERR_CUDA;                                                                                        \
      }                                                                                                               \
      configured = P64_SMEM_LIMIT;                                                                                    \
    }                                                                                                                 \
    conv64_persistent_kernel<NS, EW><<<grid, P64Threads<NS, EW>::value, smem_bytes, st>>>(maps, prog, epi, (int)total); \
  }
  OBMAN_P64_CASE(18, 1, 8)
  OBMAN_P64_CASE(24, 2, 4)
  OBMAN_P64_CASE(28, 2, 8)
  OBMAN_P64_CASE(34, 3, 4)
#undef OBMAN_P64_CASE
  int rc = check_launch("conv64_persistent_kernel");
  return rc ? rc : 1;
}

}  // namespace obman
