// tcgen05 / TMA implicit-GEMM kernel for every dense contraction of the hot path:
//   * ResNet-18 convolutions, fprop and dgrad   (mano_train/networks/bases/resnet.py:25-54,154-188)
//   * AtlasNet point-MLP 1x1 convs              (mano_train/networks/branches/atlasutils.py:65-75)
//   * ManoBranch / AtlasBranch linear layers    (manobranch.py:124-147, atlasbranch.py:44-61)
//   * weight gradients of all of the above (reduction over pixels: MN-major operands, MODE 1)
//
// D[128 x BN] (fp32, TMEM) = sum over taps and 32-wide K blocks of A_tile[128 x 32] * B_tile[BN x 32]^T
//   MODE 0 (K-major)
//     A: activations. plain: row-major (M, K) matrix, 2-D TMA box {32, 128}; spatial: NHWC tensor, 4-D TMA
//        box {32 ch, TW, TH, TN} at (c, w0+dw, h0+dh, n0) - 3x3 / strided / transposed taps are shifted boxes,
//        image borders come from TMA zero fill.
//     B: weights (N, K_total) K-major, 2-D TMA box {32, BN}.
//   MODE 1 (MN-major, weight gradients): K runs over blocks of 32 output pixels; A = dY channels, B = the
//     taps of the (shifted) input stacked along N ("virtual im2col": column group g = (tap, 32-channel group),
//     one 5-D TMA box per group), so dY is read once per 256 output columns instead of once per tap.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2-5 = fp32 -> (tf32 hi, tf32 lo) splitters during the main loop, then TMEM -> register epilogue.
// Precision: PASSES=1 is plain TF32; PASSES=3 accumulates lo*hi + hi*lo + hi*hi ("3xTF32", ~fp32 accuracy).
//   TS=1 (MODE 0, PASSES 3, weights pre-split into hi/lo in global memory): the splitters write the A tile's
//   hi/lo halves straight into TENSOR MEMORY and the MMAs take A from TMEM (tcgen05.mma [d],[a_tmem],b_desc),
//   which removes the A re-reads and the splitter's shared-memory writes - the kernel was shared-memory
//   bandwidth bound with all four operand tiles in shared memory (profiles/ncu_full_r1_gemm_summary.txt).
#include <stdlib.h>

#include "gemm_shared.cuh"

namespace obman {

// OCC = CTAs per SM the configuration is sized for: with 2, one CTA's prologue / epilogue overlaps the other's
// main loop (the kernel is not persistent), at the price of a shallower per-CTA pipeline.
template <int BN, int PASSES, int TS, int OCC>
struct GemmCfg {
  static constexpr int B_TILE_BYTES = BN * BK * 4;
  // TS == 1: A_raw | B_hi | B_lo (tf32).  TS == 2: A_raw | B with bf16 hi|lo interleaved per 32-element block.
  // TS == 3 (experimental, OBMAN_GEMM_STACK64): as 2, but the B tile is landed as 2*BN rows x 64 bytes (all hi rows,
  // then all lo rows) so that ONE N = 2*BN MMA yields a_hi*b_hi and a_hi*b_lo side by side: N = 64 MMAs only reach
  // 54 % of the tensor rate (profiles/mma_chain_r1j.txt), N = 128 MMAs the full rate.  Two accumulators (2*BN columns).
  static constexpr bool BF = TS == 2 || TS == 3;
  static constexpr int ACC_COLS = TS == 3 ? 2 * BN : BN;
  static constexpr int STAGE_BYTES = BF ? (A_TILE_BYTES + B_TILE_BYTES)
                                     : TS ? (A_TILE_BYTES + 2 * B_TILE_BYTES)
                                          : (A_TILE_BYTES + B_TILE_BYTES) * (PASSES == 3 ? 2 : 1);
  static constexpr int A_COLS = BF ? 32 : 64;   // TMEM columns per stage: A hi | A lo
  static constexpr int STAGES_RAW = ((OCC == 2 ? 104 : 200) * 1024) / STAGE_BYTES;
  static constexpr int STAGES_SMEM = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  // TS: every stage also owns A_COLS TMEM columns; 512 columns per SM in total
  static constexpr int STAGES_TMEM = ((OCC == 2 ? 256 : 512) - ACC_COLS) / A_COLS;
  static constexpr int STAGES = TS ? (STAGES_SMEM < STAGES_TMEM ? STAGES_SMEM : STAGES_TMEM) : STAGES_SMEM;
  static constexpr int TMEM_NEED = TS ? ACC_COLS + STAGES * A_COLS : BN;
  static constexpr int TMEM_COLS = TMEM_NEED <= 64 ? 64 : (TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512));
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// CL > 1 (TS path only): thread-block cluster of CL CTAs along the M-tile axis.  They share the weight tile, so
// each CTA fetches 1/CL of it and TMA-multicasts it to all of them: the kernels were L2->SM bandwidth bound
// (~9 TB/s measured against ~42 B/clk/SM), the weight tile being 2/3 of the bytes of every K block.
//
// TAIL = 1 (TS == 2, plain matrices): N = 256 j + (1..4) columns, as in the point decoder (515 -> 257 -> 128: the "+3" / "+1"
// of the concatenated grid coordinates survive the halvings).  A column tile of their own would repeat the whole
// conversion of the A tile into tensor memory - which paces this kernel - for 1-3 useful columns; instead the
// splitter thread of row r, which holds the 32 fp32 values of its row in registers anyway, accumulates those columns
// with plain FMAs.  Their weights are unpacked per K block by the otherwise idle lanes of the producer warp into a
// small shared-memory ring that travels with the pipeline stages.
template <int BN, int PASSES, int MODE, int TS, int OCC, int CL, int TAIL>
__global__ void __launch_bounds__(GEMM_THREADS, OCC)
gemm_tc_kernel(const __grid_constant__ GemmMaps maps, const GemmProgram prog, const GemmEpilogue epi) {
  using Cfg = GemmCfg<BN, PASSES, TS, OCC>;
  static_assert(Cfg::STAGES >= 2, "pipeline needs at least two stages");
  static_assert(CL == 1 || (TS >= 1 && MODE == 0), "clusters are only used on the TS path");
  static_assert(TAIL == 0 || (TS == 2 && MODE == 0 && CL == 1), "tail columns exist on the plain 3xBF16 path only");
  constexpr uint16_t CL_MASK = (uint16_t)((1u << CL) - 1);
  const uint32_t cl_rank = CL > 1 ? cluster_ctarank() : 0;
  constexpr int S = Cfg::STAGES;
  constexpr int B_TILE_BYTES = Cfg::B_TILE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cfg::STAGE_BYTES);
  uint64_t* full = bars;            // [S] TMA bytes landed
  uint64_t* conv = bars + S;        // [S] hi/lo split done (PASSES==3)
  uint64_t* empty = bars + 2 * S;   // [S] MMAs reading the stage retired
  uint64_t* accum = bars + 3 * S;   // accumulator complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * S + 1);
  float* tailw = reinterpret_cast<float*>(smem + S * Cfg::STAGE_BYTES + 256);   // TAIL: [S][4 columns][32 k] fp32

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  int n_iters = prog.num_taps * prog.kblocks;
  if (threadIdx.x == 0) trace_stamp(epi, 0);

  // tile coordinates.  raster_n: column tiles vary fastest (blockIdx.x), so the CTAs that share an A tile run back to
  // back and the tile is fetched from HBM once (the decoder GEMM 41088 x 257 x 515 read its A operand 2.6x, round 1)
  const int bx = prog.raster_n ? blockIdx.y : blockIdx.x;
  const int by = prog.raster_n ? blockIdx.x : blockIdx.y;
  int m0 = 0, n_img0 = 0, h0 = 0, w0 = 0;
  int n0 = by * BN;
  int pb_begin = 0;
  if (MODE >= 1) {
    m0 = (blockIdx.x / prog.n_tiles) * BM;
    n0 = (blockIdx.x % prog.n_tiles) * BN;
    const int total = prog.kblocks_n * prog.kblocks_h * prog.kblocks_w;
    const int per = (total + gridDim.z - 1) / gridDim.z;
    pb_begin = blockIdx.z * per;
    n_iters = min(total, pb_begin + per) - pb_begin;
    if (n_iters <= 0) return;  // uniform for the whole CTA, before any barrier / TMEM allocation
  } else if (prog.spatial) {
    int t = bx;
    const int tw_i = t % prog.tiles_w; t /= prog.tiles_w;
    const int th_i = t % prog.tiles_h; t /= prog.tiles_h;
    n_img0 = t * prog.TN;
    h0 = th_i * prog.TH;
    w0 = tw_i * prog.TW;
  } else {
    m0 = bx * BM;
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&conv[s], 128);
      mbar_init(&empty[s], CL);   // every CTA of the cluster commits to every CTA's empty barriers
    }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // peers' barriers must exist before any multicast / remote arrive
  tc_fence_after();
  // read through a shuffle: tells the compiler the value is warp-uniform, so the tcgen05 operands derived
  // from it live in uniform registers (otherwise every MMA is wrapped in an elect / broadcast loop)
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  if (threadIdx.x == 0) trace_stamp(epi, 1);

  // stage layout.  SS: A | B | A_lo | B_lo.   TS: A_raw | B_hi | B_lo.
  auto stage_a = [&](int s) { return smem + s * Cfg::STAGE_BYTES; };
  auto stage_b = [&](int s) { return smem + s * Cfg::STAGE_BYTES + A_TILE_BYTES; };
  auto stage_alo = [&](int s) { return smem + s * Cfg::STAGE_BYTES + A_TILE_BYTES + B_TILE_BYTES; };
  auto stage_blo = [&](int s) {
    return smem + s * Cfg::STAGE_BYTES + (TS ? A_TILE_BYTES + B_TILE_BYTES : 2 * A_TILE_BYTES + B_TILE_BYTES);
  };
  // TS: TMEM columns of stage s: [BN + A_COLS s, + A_COLS/2) = A hi, next A_COLS/2 = A lo
  auto tmem_a = [&](int s) { return tmem_base + (uint32_t)(Cfg::ACC_COLS + Cfg::A_COLS * s); };

  // MODE 1: number of 32-column groups of this tile that exist (B = stacked input taps).
  // MODE 2 (roles swapped: A = stacked input taps, B = dY channels): number of 32-row groups that exist.
  int valid_groups = BN / 32;
  if (MODE == 1) valid_groups = min(BN / 32, prog.total_groups - n0 / 32);
  if (MODE == 2) valid_groups = min(BM / 32, prog.total_groups - m0 / 32);

  const bool tail_on = TAIL && n0 == 0;   // CTA-uniform
  if (warp == 0) {
    // ===== TMA producer =====
    if (TAIL || lane == 0) {
      if (lane == 0) {
        tma_prefetch_desc(&maps.b);
        tma_prefetch_desc(&maps.a[0]);
      }
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        float tw[4] = {0.f, 0.f, 0.f, 0.f};
        if (TAIL && tail_on) {
          // tail-column weights of this K block: lane = k, packed row = 32 bf16 hi then 32 bf16 lo per 128 bytes.
          // Issued before the wait for the stage so that the loads overlap it.
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (e < prog.n_tail) {
              const unsigned short* rowp =
                  reinterpret_cast<const unsigned short*>(prog.tail_w + e * prog.tail_ldw + (long long)it * 128);
              tw[e] = __uint_as_float((uint32_t)__ldg(rowp + lane) << 16) +
                      __uint_as_float((uint32_t)__ldg(rowp + 32 + lane) << 16);
            }
          }
        }
        if (lane == 0) mbar_wait(&empty[s], ph ^ 1);
        if (TAIL) {
          __syncwarp();
          if (tail_on) {
#pragma unroll
            for (int e = 0; e < 4; ++e) tailw[(s * 4 + e) * 32 + lane] = tw[e];
          }
          __syncwarp();   // lane 0's arrive on full[s] below publishes the whole warp's stores
          if (lane != 0) continue;
        }
        if (MODE == 2) {
          mbar_arrive_expect_tx(&full[s], 4096 * valid_groups + B_TILE_BYTES);
          int pb = pb_begin + it;
          const int bw = pb % prog.kblocks_w; pb /= prog.kblocks_w;
          const int bh = pb % prog.kblocks_h; pb /= prog.kblocks_h;
          const int pw = bw * prog.kTW, ph_ = bh * prog.kTH, pn = pb * prog.kTN;
          for (int j = 0; j < valid_groups; ++j) {
            const int g = m0 / 32 + j;
            const int tap = g / prog.cg_in, cg = g - tap * prog.cg_in;
            tma_load_5d(stage_a(s) + j * 4096, &maps.a[1 + prog.tap_map[tap]], &full[s], 0, pw + prog.tap_dw[tap],
                        ph_ + prog.tap_dh[tap], pn, cg);
          }
          tma_load_5d(stage_b(s), &maps.a[0], &full[s], 0, pw, ph_, pn, n0 / 32);
          continue;
        }
        if (MODE == 1) {
          mbar_arrive_expect_tx(&full[s], A_TILE_BYTES + 4096 * valid_groups);
          int pb = pb_begin + it;
          const int bw = pb % prog.kblocks_w; pb /= prog.kblocks_w;
          const int bh = pb % prog.kblocks_h; pb /= prog.kblocks_h;
          const int pw = bw * prog.kTW, ph_ = bh * prog.kTH, pn = pb * prog.kTN;
          tma_load_5d(stage_a(s), &maps.a[0], &full[s], 0, pw, ph_, pn, m0 / 32);
          for (int j = 0; j < valid_groups; ++j) {
            const int g = n0 / 32 + j;
            const int tap = g / prog.cg_in, cg = g - tap * prog.cg_in;
            tma_load_5d(stage_b(s) + j * 4096, &maps.a[1 + prog.tap_map[tap]], &full[s], 0, pw + prog.tap_dw[tap],
                        ph_ + prog.tap_dh[tap], pn, cg);
          }
          continue;
        }
        mbar_arrive_expect_tx(&full[s], A_TILE_BYTES + B_TILE_BYTES * (TS == 1 ? 2 : 1));
        const int tap = it / prog.kblocks;
        const int kb = it - tap * prog.kblocks;
        if (prog.spatial) {
          tma_load_4d(stage_a(s), &maps.a[prog.tap_map[tap]], &full[s], kb * BK, w0 + prog.tap_dw[tap],
                      h0 + prog.tap_dh[tap], n_img0);
        } else {
          tma_load_2d(stage_a(s), &maps.a[0], &full[s], kb * BK, m0);
        }
        if (CL > 1) {
          // this CTA's 1/CL slice of the weight tile, multicast to the whole cluster
          constexpr int ROWS = BN / CL;
          const int kc = prog.tap_bk[tap] + kb * BK;
          tma_load_2d_mc(stage_b(s) + cl_rank * ROWS * 128, &maps.b, &full[s], kc, n0 + cl_rank * ROWS, CL_MASK);
          if (TS == 1)
            tma_load_2d_mc(stage_blo(s) + cl_rank * ROWS * 128, &maps.b_lo, &full[s], kc, n0 + cl_rank * ROWS, CL_MASK);
        } else {
          if (TS == 3) {
            // packed row of this K block = 64 bytes of hi then 64 bytes of lo: two 64-byte-wide boxes (SW64 map) land
            // them as rows 0..BN-1 (hi) and BN..2BN-1 (lo) of one 2*BN x 64-byte tile
            tma_load_2d(stage_b(s), &maps.b, &full[s], prog.tap_bk[tap] + kb * BK, n0);
            tma_load_2d(stage_b(s) + BN * 64, &maps.b, &full[s], prog.tap_bk[tap] + kb * BK + 16, n0);
          } else {
          tma_load_2d(stage_b(s), &maps.b, &full[s], prog.tap_bk[tap] + kb * BK, n0);
          if (TS == 1) tma_load_2d(stage_blo(s), &maps.b_lo, &full[s], prog.tap_bk[tap] + kb * BK, n0);
          }
        }
        if (it == 0) trace_stamp(epi, 2);
        if (it == n_iters - 1) trace_stamp(epi, 3);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = Cfg::BF ? umma_idesc_bf16(BM, BN) : umma_idesc_tf32(BM, BN, MODE != 0, MODE != 0);
    const uint32_t idesc_wide = umma_idesc_bf16(BM, 2 * BN);   // TS == 3 only
    // K-major: 4 k-steps of 32 bytes inside the 128-byte row (SW128, 8-row groups 1024 B apart).
    // MN-major (TF32 => SW128 with 32-byte atoms): 4 k-steps of 8 pixel rows (1024 B), 4-row K atoms 512 B
    // apart (SBO), 32-channel groups 4096 B apart (LBO)
    constexpr uint32_t KSTEP = MODE != 0 ? 1024 : 32;
    const uint32_t LBO = MODE != 0 ? prog.mn_lbo : 16;
    const uint32_t SBO = MODE != 0 ? prog.mn_sbo : 1024;
    const uint32_t LT = MODE != 0 ? prog.mn_layout : 2;
    for (int it = 0; it < n_iters; ++it) {
      const int s = it % S;
      const uint32_t ph = (it / S) & 1;
      mbar_wait(&full[s], ph);
      if (PASSES == 3) mbar_wait(&conv[s], ph);
      tc_fence_after();
      if (lane == 0) {
        if (it == 0) trace_stamp(epi, 4);
        const uint32_t a_hi = smem_u32(stage_a(s)), b_hi = smem_u32(stage_b(s));
        const uint32_t a_lo = smem_u32(stage_alo(s)), b_lo = smem_u32(stage_blo(s));
        if (TS == 3) {
          // stacked B: rows 0..BN-1 = b_hi, rows BN..2BN-1 = b_lo, 64-byte rows (SW64: 8-row groups 512 B apart)
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t db = umma_desc(b_hi + k * 32, 16, 512, 4);
            const uint32_t ta_hi = tmem_a(s) + k * 8, ta_lo = tmem_a(s) + 16 + k * 8;
            umma_f16_ts(tmem_base, ta_hi, db, idesc_wide, (it > 0 || k > 0) ? 1u : 0u);   // [a_hi*b_hi | a_hi*b_lo]
            umma_f16_ts(tmem_base, ta_lo, db, idesc, 1u);                                  // columns 0..BN-1 += a_lo*b_hi
          }
        } else if (TS == 2) {
          // bf16 hi|lo: every 128-byte B row holds 32 hi then 32 lo values of this K block; K = 16 per MMA
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t db = umma_desc(b_hi + k * 32, 16, 1024, 2);
            const uint64_t dbl = umma_desc(b_hi + 64 + k * 32, 16, 1024, 2);
            const uint32_t ta_hi = tmem_a(s) + k * 8, ta_lo = tmem_a(s) + 16 + k * 8;
            umma_f16_ts(tmem_base, ta_lo, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
            umma_f16_ts(tmem_base, ta_hi, dbl, idesc, 1u);
            umma_f16_ts(tmem_base, ta_hi, db, idesc, 1u);
          }
        } else {
#pragma unroll
        for (int k = 0; k < BK / 8; ++k) {
          const uint64_t db = umma_desc(b_hi + k * KSTEP, LBO, SBO, LT);
          const uint32_t acc0 = (it > 0 || k > 0) ? 1u : 0u;
          if (TS) {
            const uint64_t dbl = umma_desc(b_lo + k * KSTEP, LBO, SBO, LT);
            const uint32_t ta_hi = tmem_a(s) + k * 8, ta_lo = tmem_a(s) + 32 + k * 8;
            umma_tf32_ts(tmem_base, ta_lo, db, idesc, acc0);
            umma_tf32_ts(tmem_base, ta_hi, dbl, idesc, 1u);
            umma_tf32_ts(tmem_base, ta_hi, db, idesc, 1u);
          } else if (PASSES == 3) {
            const uint64_t da = umma_desc(a_hi + k * KSTEP, LBO, SBO, LT);
            const uint64_t dal = umma_desc(a_lo + k * KSTEP, LBO, SBO, LT);
            const uint64_t dbl = umma_desc(b_lo + k * KSTEP, LBO, SBO, LT);
            umma_tf32(tmem_base, dal, db, idesc, acc0);
            umma_tf32(tmem_base, da, dbl, idesc, 1u);
            umma_tf32(tmem_base, da, db, idesc, 1u);
          } else {
            const uint64_t da = umma_desc(a_hi + k * KSTEP, LBO, SBO, LT);
            umma_tf32(tmem_base, da, db, idesc, acc0);
          }
        }
        }
        if (CL > 1) umma_commit_mc(&empty[s], CL_MASK);
        else umma_commit(&empty[s]);
        if (it == n_iters - 1) { umma_commit(accum); trace_stamp(epi, 5); }
      }
      __syncwarp();
    }
  } else {
    // ===== splitter (main loop) + epilogue: warps 2..5, TMEM lane quadrant = warp % 4 =====
    const int tid = threadIdx.x - 64;  // 0..127
    const int q = warp & 3;
    const int r = q * 32 + lane;       // row of the tile == TMEM lane
    if (TS) {
      // A row r (32 fp32 along K, 8 swizzled 16-byte chunks) -> hi / lo -> tensor memory
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
      float ext[4] = {0.f, 0.f, 0.f, 0.f};   // TAIL: row r of the tail columns
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(&full[s], ph);
        const uint32_t row = smem_u32(stage_a(s)) + r * 128;
        if (Cfg::BF) {
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
          const uint32_t dst = lane_base + (uint32_t)(Cfg::ACC_COLS + Cfg::A_COLS * s);
          tmem_st_32x16(dst, hi);
          tmem_st_32x16(dst + 16, lo);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&conv[s]);
          if (TAIL && tail_on && it == n_iters - 1 && m0 + r < prog.M) {
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
        } else {
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = lds_v4(row + ((j ^ (r & 7)) << 4));
          const float h0f = to_tf32_rna(v.x), h1f = to_tf32_rna(v.y), h2f = to_tf32_rna(v.z), h3f = to_tf32_rna(v.w);
          hi[4 * j + 0] = __float_as_uint(h0f); lo[4 * j + 0] = __float_as_uint(v.x - h0f);
          hi[4 * j + 1] = __float_as_uint(h1f); lo[4 * j + 1] = __float_as_uint(v.y - h1f);
          hi[4 * j + 2] = __float_as_uint(h2f); lo[4 * j + 2] = __float_as_uint(v.z - h2f);
          hi[4 * j + 3] = __float_as_uint(h3f); lo[4 * j + 3] = __float_as_uint(v.w - h3f);
        }
        const uint32_t dst = lane_base + (uint32_t)(BN + 64 * s);
        tmem_st_32x32(dst, hi);
        tmem_st_32x32(dst + 32, lo);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&conv[s]);
        }
      }
    } else if (PASSES == 3) {
      const int a_v4 = (MODE == 2 ? valid_groups * 4096 : A_TILE_BYTES) / 16;
      const int b_v4 = (MODE == 1 ? valid_groups * 4096 : B_TILE_BYTES) / 16;
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(&full[s], ph);
        float4* a = reinterpret_cast<float4*>(stage_a(s));
        float4* alo = reinterpret_cast<float4*>(stage_alo(s));
        float4* b = reinterpret_cast<float4*>(stage_b(s));
        float4* blo = reinterpret_cast<float4*>(stage_blo(s));
#pragma unroll 4
        for (int i = tid; i < a_v4 + b_v4; i += 128) {
          float4* src = i < a_v4 ? a + i : b + (i - a_v4);
          float4* dst = i < a_v4 ? alo + i : blo + (i - a_v4);
          // The tensor core TRUNCATES fp32 operands to tf32 (measured: scripts/probe_dense.py rounding_mode),
          // so the raw tile already acts as `hi`; only lo = v - trunc(v) has to be written.
          const float4 v = *src;
          float4 l;
          l.x = v.x - tf32_trunc(v.x); l.y = v.y - tf32_trunc(v.y);
          l.z = v.z - tf32_trunc(v.z); l.w = v.w - tf32_trunc(v.w);
          *dst = l;
        }
        fence_proxy_async_smem();
        mbar_arrive(&conv[s]);
      }
    }
    if (TS == 3) gemm_epilogue_stacked<BN>(smem, tmem_base, accum, prog, epi, m0, n0, n_img0, h0, w0, q, lane, r);
    else gemm_epilogue<BN, MODE>(smem, tmem_base, accum, prog, epi, m0, n0, n_img0, h0, w0, q, lane, r);
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();   // no CTA may exit while peers can still multicast into it / arrive on its barriers
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  if (threadIdx.x == 0) trace_stamp(epi, 7);
}

template <int BN, int PASSES, int MODE, int TS, int OCC, int CL = 1, int TAIL = 0>
static int launch_gemm(const GemmMaps& maps, const GemmProgram& prog, const GemmEpilogue& epi, dim3 grid,
                       cudaStream_t st) {
  using Cfg = GemmCfg<BN, PASSES, TS, OCC>;
  constexpr int SMEM = Cfg::SMEM_BYTES + (TAIL ? Cfg::STAGES * 512 : 0);
  static_assert(SMEM <= 232448, "shared memory budget");
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, PASSES, MODE, TS, OCC, CL, TAIL>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      set_error("gemm_tc: cudaFuncSetAttribute(%d bytes) failed: %s", SMEM, cudaGetErrorString(e));
      return OBMAN_ERR_CUDA;
    }
    attr = true;
  }
  if (CL > 1) {
    grid.x = (grid.x + CL - 1) / CL * CL;  // padded CTAs work on out-of-range tiles: TMA zero fill, rows masked
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attrs[1];
    attrs[0].id = cudaLaunchAttributeClusterDimension;
    attrs[0].val.clusterDim.x = CL;
    attrs[0].val.clusterDim.y = 1;
    attrs[0].val.clusterDim.z = 1;
    cfg.attrs = attrs;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, PASSES, MODE, TS, OCC, CL, TAIL>, maps, prog, epi);
    if (e != cudaSuccess) {
      set_error("gemm_tc: cluster launch failed: %s", cudaGetErrorString(e));
      (void)cudaGetLastError();
      return OBMAN_ERR_CUDA;
    }
    return OBMAN_OK;
  }
  gemm_tc_kernel<BN, PASSES, MODE, TS, OCC, CL, TAIL><<<grid, GEMM_THREADS, SMEM, st>>>(maps, prog, epi);
  return check_launch("gemm_tc_kernel");
}

// ---- weight gradient on the 3xBF16 path -----------------------------------------------------------------------
// Both operands of a weight gradient are activations (dY and the shifted taps of x), channels contiguous and the
// reduction (pixels) along rows, so neither can be pre-split.  The raw fp32 tiles land UNswizzled in a small
// ring (TMA), the four splitter warps read them column-wise (one thread per channel: 32 pixels, conflict-free),
// split every value into bf16 hi + lo and write
//   * the A operand (128 channels x 32 pixels) into tensor memory, and
//   * the B operand as K-major rows [32 hi | 32 lo] (the layout of the packed weights of the fprop kernels) into
//     a second ring, hand-swizzled for the SW128 descriptor,
// so the MMA stream is the one of the fprop kernels.  The raw ring is released as soon as it has been read, the
// converted ring when its MMAs retire.  SWAP = 1: rows = stacked taps, columns = dY channels (c_out <= 64).
constexpr int TRACE_IT = 12;   // main-loop iteration whose inner phases obman_debug_trace records (slots 8..15)

// Epilogue of the swapped + stacked weight-gradient variant (SWAP == 2): D[m, n] = acc[m, n] + acc[m, BN + n] with
// m = stacked (tap, c_in) index, n = output channel; dw is (c_out, taps*c_in) row-major, element (m, n) at n*ld + m, so
// the 32 lanes of a warp (consecutive m) make every column one coalesced access (same as gemm_epilogue's MODE 2).
template <int BN>
__device__ __forceinline__ void wgrad_epilogue_swapped_stacked(uint32_t tmem_base, uint64_t* accum,
                                                               const GemmProgram& prog, const GemmEpilogue& epi,
                                                               int m0, int n0, int q, int r) {
  mbar_wait(accum, 0);
  tc_fence_after();
  if (r == 0) trace_stamp(epi, 6);
  const bool row_ok = (m0 + r) < prog.M;
  const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 16) {
    if (n0 + c0 >= prog.N) break;  // warp-uniform
    uint32_t v[16], v2[16];
    tmem_ld_32x16(lane_addr + (uint32_t)c0, v);
    tmem_ld_32x16(lane_addr + (uint32_t)(BN + c0), v2);
    tmem_ld_wait();
    if (!row_ok) continue;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int col = n0 + c0 + j;
      if (col >= prog.N) break;
      float* dst = epi.out + (long long)col * epi.ld + (m0 + r);
      const float y = epi.alpha * (__uint_as_float(v[j]) + __uint_as_float(v2[j]));
      if (epi.accumulate) atomicAdd(dst, y);
      else *dst = y;
    }
  }
}

template <int BN, int SWAP, int OCC>
struct WgradCfg {
  static constexpr int A_RAW = 128 * 128;              // 4 groups x 32 pixels x 32 channels fp32
  static constexpr int B_RAW = BN * 128;
  static constexpr int RAW_BYTES = A_RAW + B_RAW;
  static constexpr int B_CONV = BN * 128;
  static constexpr int BUDGET = (OCC == 2 ? 104 : 200) * 1024;
  static constexpr int R = 2;                          // raw stages
  static constexpr int C_SMEM = (BUDGET - R * RAW_BYTES) / B_CONV;
  // SWAP == 2 (experimental, OBMAN_WGRAD_STACK64): roles swapped AND the converted B operand stacked as 2*BN rows of
  // 64 bytes (hi rows, then lo rows), two accumulators - the weight-gradient counterpart of GemmCfg's TS == 3
  static constexpr int ACC_COLS = SWAP == 2 ? 2 * BN : BN;
  static constexpr int C_TMEM = ((OCC == 2 ? 256 : 512) - ACC_COLS) / 32;
  static constexpr int C0 = C_SMEM < C_TMEM ? C_SMEM : C_TMEM;
  static constexpr int C = C0 > 6 ? 6 : C0;            // converted stages
  static constexpr int TMEM_NEED = ACC_COLS + C * 32;
  static constexpr int TMEM_COLS = TMEM_NEED <= 64 ? 64 : (TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512));
  static constexpr int SMEM_BYTES = R * RAW_BYTES + C * B_CONV + 1024 + 256;
  // one splitter thread per operand row: 4 warps for the 128 A rows (also the epilogue warps), BN / 32 for B
  static constexpr int SPLIT_WARPS = 4 + BN / 32;
  static constexpr int SPLIT_THREADS = SPLIT_WARPS * 32;
  static constexpr int THREADS = 64 + SPLIT_THREADS;
};

// 32 pixels of one channel (column `lane` of a 32 x 32 fp32 group at shared address `grp`) -> 16 packed bf16x2 hi / lo
__device__ __forceinline__ void split_column(uint32_t grp, int lane, uint32_t* hi, uint32_t* lo) {
  const uint32_t a = grp + (uint32_t)lane * 4u;
#pragma unroll
  for (int p = 0; p < 16; ++p) split_bf16x2(lds_f32(a + (2 * p) * 128), lds_f32(a + (2 * p + 1) * 128), hi[p], lo[p]);
}
// same, also adding the 32 raw values to `sum` (column sum of dY = the BatchNorm-beta / bias gradient, for free)
__device__ __forceinline__ void split_column_sum(uint32_t grp, int lane, uint32_t* hi, uint32_t* lo, float& sum) {
  const uint32_t a = grp + (uint32_t)lane * 4u;
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const float v0 = lds_f32(a + (2 * p) * 128), v1 = lds_f32(a + (2 * p + 1) * 128);
    sum += v0 + v1;
    split_bf16x2(v0, v1, hi[p], lo[p]);
  }
}

template <int BN, int SWAP, int OCC>
__global__ void __launch_bounds__(WgradCfg<BN, SWAP, OCC>::THREADS, OCC)
wgrad_bf16_kernel(const __grid_constant__ GemmMaps maps, const GemmProgram prog, const GemmEpilogue epi) {
  using Cfg = WgradCfg<BN, SWAP, OCC>;
  static_assert(Cfg::C >= 2, "needs at least two converted stages");
  constexpr int R = Cfg::R, C = Cfg::C;
  constexpr int TAP_GROUPS = SWAP ? BM / 32 : BN / 32;   // 32-channel groups of the stacked-tap operand per tile
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* conv_base = smem + R * Cfg::RAW_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(conv_base + C * Cfg::B_CONV);
  uint64_t* full = bars;                // [R] TMA bytes landed
  uint64_t* rawfree = bars + R;         // [R] splitter has read the raw stage
  uint64_t* conv = bars + 2 * R;        // [C] converted operands written
  uint64_t* empty = bars + 2 * R + C;   // [C] MMAs reading the converted stage retired
  uint64_t* accum = bars + 2 * R + 2 * C;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  const int m0 = (blockIdx.x / prog.n_tiles) * BM;
  const int n0 = (blockIdx.x % prog.n_tiles) * BN;
  const int total = prog.kblocks_n * prog.kblocks_h * prog.kblocks_w;
  const int per = (total + gridDim.z - 1) / gridDim.z;
  const int pb_begin = blockIdx.z * per;
  const int n_iters = min(total, pb_begin + per) - pb_begin;
  if (n_iters <= 0) return;  // uniform for the whole CTA, before any barrier / TMEM allocation
  if (threadIdx.x == 0) trace_stamp(epi, 0);

  if (threadIdx.x == 0) {
    for (int i = 0; i < R; ++i) { mbar_init(&full[i], 1); mbar_init(&rawfree[i], Cfg::SPLIT_THREADS); }
    for (int i = 0; i < C; ++i) { mbar_init(&conv[i], Cfg::SPLIT_THREADS); mbar_init(&empty[i], 1); }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  if (threadIdx.x == 0) trace_stamp(epi, 1);
  auto raw_a = [&](int i) { return smem + i * Cfg::RAW_BYTES; };
  auto raw_b = [&](int i) { return smem + i * Cfg::RAW_BYTES + Cfg::A_RAW; };
  auto conv_b = [&](int i) { return conv_base + i * Cfg::B_CONV; };
  // groups of 32 channels of the stacked-tap operand that exist in this tile (B normally, A when swapped)
  const int valid_groups = min(TAP_GROUPS, prog.total_groups - (SWAP ? m0 : n0) / 32);

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&maps.a[0]);
      // one TMA per `gl` consecutive channel groups of a tap (gl = prog.grp_per_load divides cg_in and the tile)
      const int gl = prog.grp_per_load;
      const int nl = valid_groups / gl;
      const CUtensorMap* lmap[TAP_GROUPS];
      int ldw[TAP_GROUPS], ldh[TAP_GROUPS], lcg[TAP_GROUPS];
#pragma unroll
      for (int j = 0; j < TAP_GROUPS; ++j) {
        const int g = (SWAP ? m0 : n0) / 32 + j * gl;
        const int tap = (j < nl) ? g / prog.cg_in : 0;
        lmap[j] = &maps.a[1 + prog.tap_map[tap]];
        ldw[j] = prog.tap_dw[tap];
        ldh[j] = prog.tap_dh[tap];
        lcg[j] = g - tap * prog.cg_in;
      }
      int pb = pb_begin;
      int bw = pb % prog.kblocks_w; pb /= prog.kblocks_w;
      int bh = pb % prog.kblocks_h;
      int bn = pb / prog.kblocks_h;
      for (int it = 0; it < n_iters; ++it) {
        const int rs = it % R;
        mbar_wait(&rawfree[rs], ((it / R) & 1) ^ 1);
        if (it == TRACE_IT) trace_stamp(epi, 8);
        mbar_arrive_expect_tx(&full[rs], 4096 * valid_groups + (SWAP ? Cfg::B_RAW : Cfg::A_RAW));
        const int pw = bw * prog.kTW, ph_ = bh * prog.kTH, pn = bn * prog.kTN;
        uint8_t* taps_dst = SWAP ? raw_a(rs) : raw_b(rs);
#pragma unroll
        for (int j = 0; j < TAP_GROUPS; ++j) {
          if (j < nl) tma_load_5d(taps_dst + j * gl * 4096, lmap[j], &full[rs], 0, pw + ldw[j], ph_ + ldh[j], pn, lcg[j]);
        }
        tma_load_5d(SWAP ? raw_b(rs) : raw_a(rs), &maps.a[0], &full[rs], 0, pw, ph_, pn, (SWAP ? n0 : m0) / 32);
        if (it == 0) trace_stamp(epi, 2);
        if (it == TRACE_IT) trace_stamp(epi, 9);
        if (it == n_iters - 1) trace_stamp(epi, 3);
        if (++bw == prog.kblocks_w) {
          bw = 0;
          if (++bh == prog.kblocks_h) { bh = 0; ++bn; }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(BM, BN);
    for (int it = 0; it < n_iters; ++it) {
      const int cs = it % C;
      mbar_wait(&conv[cs], (it / C) & 1);
      tc_fence_after();
      if (lane == 0) {
        if (it == 0) trace_stamp(epi, 4);
        if (it == TRACE_IT) trace_stamp(epi, 13);
        const uint32_t b = smem_u32(conv_b(cs));
        const uint32_t ta = tmem_base + (uint32_t)(Cfg::ACC_COLS + 32 * cs);
        if (SWAP == 2) {
          const uint32_t idesc_wide = umma_idesc_bf16(BM, 2 * BN);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t db = umma_desc(b + k * 32, 16, 512, 4);   // 2*BN rows x 64 B, SW64
            umma_f16_ts(tmem_base, ta + k * 8, db, idesc_wide, (it > 0 || k > 0) ? 1u : 0u);   // [a_hi*b_hi | a_hi*b_lo]
            umma_f16_ts(tmem_base, ta + 16 + k * 8, db, idesc, 1u);                             // cols 0..BN-1 += a_lo*b_hi
          }
        } else {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const uint64_t db = umma_desc(b + k * 32, 16, 1024, 2);
          const uint64_t dbl = umma_desc(b + 64 + k * 32, 16, 1024, 2);
          umma_f16_ts(tmem_base, ta + 16 + k * 8, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
          umma_f16_ts(tmem_base, ta + k * 8, dbl, idesc, 1u);
          umma_f16_ts(tmem_base, ta + k * 8, db, idesc, 1u);
        }
        }
        umma_commit(&empty[cs]);
        if (it == TRACE_IT) trace_stamp(epi, 14);
        if (it == n_iters - 1) { umma_commit(accum); trace_stamp(epi, 5); }
      }
      __syncwarp();
    }
  } else if (warp < 6) {
    // ---- A rows: warps 2..5, TMEM lane quadrant = warp % 4; afterwards the epilogue ----
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const int a_groups = SWAP ? valid_groups : BM / 32;
    // dY is the A operand here: the CTAs of column tile 0 also accumulate its column sums (d beta / d bias)
    const bool csum_on = !SWAP && epi.colsum != nullptr && (blockIdx.x % prog.n_tiles) == 0;
    float csum = 0.f;
    for (int it = 0; it < n_iters; ++it) {
      const int rs = it % R, cs = it % C;
      mbar_wait(&full[rs], (it / R) & 1);
      if (r == 0 && it == TRACE_IT) trace_stamp(epi, 10);
      if (r == 0 && it == TRACE_IT + 1) trace_stamp(epi, 15);
      mbar_wait(&empty[cs], ((it / C) & 1) ^ 1);
      tc_fence_after();
      if (r == 0 && it == TRACE_IT) trace_stamp(epi, 11);
      uint32_t hi[16], lo[16];
      if (q < a_groups) {
        if (!SWAP && csum_on) split_column_sum(smem_u32(raw_a(rs)) + q * 4096, lane, hi, lo, csum);
        else split_column(smem_u32(raw_a(rs)) + q * 4096, lane, hi, lo);
      } else {
#pragma unroll
        for (int p = 0; p < 16; ++p) hi[p] = lo[p] = 0u;
      }
      mbar_arrive(&rawfree[rs]);
      const uint32_t dst = lane_base + (uint32_t)(Cfg::ACC_COLS + 32 * cs);
      tmem_st_32x16(dst, hi);
      tmem_st_32x16(dst + 16, lo);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&conv[cs]);
      if (r == 0 && it == TRACE_IT) trace_stamp(epi, 12);
    }
    if (csum_on && m0 + r < prog.M) atomicAdd(epi.colsum + m0 + r, csum);
    if (SWAP == 2) wgrad_epilogue_swapped_stacked<BN>(tmem_base, accum, prog, epi, m0, n0, q, r);
    else gemm_epilogue<BN, SWAP ? 2 : 1>(smem, tmem_base, accum, prog, epi, m0, n0, 0, 0, 0, q, lane, r);
  } else {
    // ---- B rows: one warp per 32-channel group; K-major rows [32 hi | 32 lo], hand-swizzled for SW128 ----
    const int g = warp - 6;
    const int n = g * 32 + lane;
    const int b_groups = SWAP ? BN / 32 : valid_groups;
    // roles swapped: dY is the B operand; the CTAs of row tile 0 accumulate its column sums
    const bool csum_on = SWAP && epi.colsum != nullptr && m0 == 0;
    float csum = 0.f;
    for (int it = 0; it < n_iters; ++it) {
      const int rs = it % R, cs = it % C;
      mbar_wait(&full[rs], (it / R) & 1);
      mbar_wait(&empty[cs], ((it / C) & 1) ^ 1);
      uint32_t hi[16], lo[16];
      if (g < b_groups) {
        if (csum_on) split_column_sum(smem_u32(raw_b(rs)) + g * 4096, lane, hi, lo, csum);
        else split_column(smem_u32(raw_b(rs)) + g * 4096, lane, hi, lo);
      } else {
#pragma unroll
        for (int p = 0; p < 16; ++p) hi[p] = lo[p] = 0u;
      }
      mbar_arrive(&rawfree[rs]);
      if (SWAP == 2) {
        // stacked: row n (64 B of hi) and row BN + n (64 B of lo), hand-swizzled for SW64 (16-byte chunk ^ row bits 1..2)
        const uint32_t row_hi = smem_u32(conv_b(cs)) + n * 64, row_lo = row_hi + BN * 64;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          sts_v4(row_hi + ((j ^ ((n >> 1) & 3)) << 4), hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
          sts_v4(row_lo + ((j ^ ((n >> 1) & 3)) << 4), lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
        }
      } else {
      const uint32_t row = smem_u32(conv_b(cs)) + n * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sts_v4(row + ((j ^ (n & 7)) << 4), hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
        sts_v4(row + (((j + 4) ^ (n & 7)) << 4), lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
      }
      }
      fence_proxy_async_smem();
      mbar_arrive(&conv[cs]);
    }
    if (csum_on && n0 + n < prog.N) atomicAdd(epi.colsum + n0 + n, csum);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  if (threadIdx.x == 0) trace_stamp(epi, 7);
}

template <int BN, int SWAP, int OCC>
static int launch_wgrad_bf16(const GemmMaps& maps, const GemmProgram& prog, const GemmEpilogue& epi, dim3 grid,
                             cudaStream_t st) {
  using Cfg = WgradCfg<BN, SWAP, OCC>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_bf16_kernel<BN, SWAP, OCC>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("wgrad_bf16: cudaFuncSetAttribute(%d bytes) failed: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return OBMAN_ERR_CUDA;
    }
    attr = true;
  }
  wgrad_bf16_kernel<BN, SWAP, OCC><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(maps, prog, epi);
  return check_launch("wgrad_bf16_kernel");
}

// ---- 3x3-neighbourhood convolution with the input tile shared by all taps ("halo" kernel) -------------------------
// The shifted-box kernel above fetches the A tile once per tap: 9 x 16 KB per 32 input channels, which (with the
// weight tile) runs into the L2->SM fabric limit (~60 B/clk/SM measured, profiles/cta_phases_r1_bf16.txt) long
// before the tensor pipe is busy.  Here ONE (TN, TH+2, TW+2, 32ch) box with a one-pixel halo is fetched per 32
// input channels and every tap is a different row offset into it: A traffic drops 6.4x, total L2->SM bytes 1.7x
// (C_out = 128) to 2.2x (C_out = 64).  Rings: halo boxes (RA slots, released by the splitter warps), weight tiles
// + A operands in tensor memory (C slots, released by tcgen05.commit).  Stride-1 taps with |dh|, |dw| <= 1 only
// (3x3 forward and its data gradient); everything else takes the shifted-box kernel.
template <int BN, int OCC>
struct HaloCfg {
  static constexpr int HALO_BYTES = 25600;   // up to 200 halo pixels x 32 channels fp32 (multiple of 1024: swizzle phase)
  static constexpr int RA = 2;
  static constexpr int B_TILE = BN * 128;
  static constexpr int BUDGET = (OCC == 2 ? 104 : 200) * 1024;
  static constexpr int C_SMEM = (BUDGET - RA * HALO_BYTES) / B_TILE;
  static constexpr int C_TMEM = ((OCC == 2 ? 256 : 512) - BN) / 32;
  static constexpr int C0 = C_SMEM < C_TMEM ? C_SMEM : C_TMEM;
  static constexpr int C = C0 > 8 ? 8 : C0;
  static constexpr int TMEM_NEED = BN + C * 32;
  static constexpr int TMEM_COLS = TMEM_NEED <= 64 ? 64 : (TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512));
  static constexpr int SMEM_BYTES = RA * HALO_BYTES + C * B_TILE + 1024 + 256;
};

template <int BN, int OCC>
__global__ void __launch_bounds__(GEMM_THREADS, OCC)
conv_halo_kernel(const __grid_constant__ GemmMaps maps, const GemmProgram prog, const GemmEpilogue epi) {
  using Cfg = HaloCfg<BN, OCC>;
  static_assert(Cfg::C >= 2, "needs at least two weight / operand slots");
  constexpr int RA = Cfg::RA, C = Cfg::C;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bring = smem + RA * Cfg::HALO_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bring + C * Cfg::B_TILE);
  uint64_t* afull = bars;                  // [RA] halo box landed
  uint64_t* afree = bars + RA;             // [RA] splitter warps are done with the halo box
  uint64_t* bfull = bars + 2 * RA;         // [C] weight tile landed
  uint64_t* conv = bars + 2 * RA + C;      // [C] A operand (bf16 hi | lo) written to tensor memory
  uint64_t* empty = bars + 2 * RA + 2 * C; // [C] MMAs reading slot c retired
  uint64_t* accum = bars + 2 * RA + 3 * C;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  const int T = prog.num_taps, KB = prog.kblocks;
  const int n_iters = T * KB;
  if (threadIdx.x == 0) trace_stamp(epi, 0);
  int t_ = blockIdx.x;
  const int tw_i = t_ % prog.tiles_w; t_ /= prog.tiles_w;
  const int th_i = t_ % prog.tiles_h; t_ /= prog.tiles_h;
  const int n_img0 = t_ * prog.TN, h0 = th_i * prog.TH, w0 = tw_i * prog.TW;
  const int n0 = blockIdx.y * BN;

  if (threadIdx.x == 0) {
    for (int i = 0; i < RA; ++i) { mbar_init(&afull[i], 1); mbar_init(&afree[i], 128); }
    for (int i = 0; i < C; ++i) { mbar_init(&bfull[i], 1); mbar_init(&conv[i], 128); mbar_init(&empty[i], 1); }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // read through a shuffle: tells the compiler the value is warp-uniform, so the tcgen05 operands derived
  // from it live in uniform registers (otherwise every MMA is wrapped in an elect / broadcast loop)
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  if (threadIdx.x == 0) trace_stamp(epi, 1);
  auto halo = [&](int i) { return smem + i * Cfg::HALO_BYTES; };
  auto bslot = [&](int i) { return bring + i * Cfg::B_TILE; };

  if (warp == 0) {
    if (lane == 0) {
      tma_prefetch_desc(&maps.a[0]);
      tma_prefetch_desc(&maps.b);
      // the halo box of K block kb + 1 is requested while the taps of kb are still being fed (at tap `pre`): by
      // then the splitter warps have long released the slot (they run C taps behind this thread)
      const int pre = (C < T - 1) ? C : T - 1;
      mbar_arrive_expect_tx(&afull[0], prog.halo_bytes);
      tma_load_4d(halo(0), &maps.a[0], &afull[0], 0, w0 - 1, h0 - 1, n_img0);
      int it = 0;
      for (int kb = 0; kb < KB; ++kb) {
        for (int tap = 0; tap < T; ++tap, ++it) {
          const int c = it % C;
          mbar_wait(&empty[c], ((it / C) & 1) ^ 1);
          mbar_arrive_expect_tx(&bfull[c], Cfg::B_TILE);
          tma_load_2d(bslot(c), &maps.b, &bfull[c], prog.tap_bk[tap] + kb * BK, n0);
          if (it == 0) trace_stamp(epi, 2);
          if (tap == pre && kb + 1 < KB) {
            const int a = (kb + 1) % RA;
            mbar_wait(&afree[a], (((kb + 1) / RA) & 1) ^ 1);
            mbar_arrive_expect_tx(&afull[a], prog.halo_bytes);
            tma_load_4d(halo(a), &maps.a[0], &afull[a], (kb + 1) * BK, w0 - 1, h0 - 1, n_img0);
          }
        }
      }
      trace_stamp(epi, 3);
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(BM, BN);
    for (int it = 0; it < n_iters; ++it) {
      const int c = it % C;
      const uint32_t ph = (it / C) & 1;
      mbar_wait(&bfull[c], ph);
      if (lane == 0 && it == TRACE_IT) trace_stamp(epi, 11);
      mbar_wait(&conv[c], ph);
      tc_fence_after();
      if (lane == 0) {
        if (it == 0) trace_stamp(epi, 4);
        if (it == TRACE_IT) trace_stamp(epi, 12);
        if (it == TRACE_IT + 1) trace_stamp(epi, 15);
        const uint32_t b = smem_u32(bslot(c));
        const uint32_t ta = tmem_base + (uint32_t)(BN + 32 * c);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const uint64_t db = umma_desc(b + k * 32, 16, 1024, 2);
          const uint64_t dbl = umma_desc(b + 64 + k * 32, 16, 1024, 2);
          umma_f16_ts(tmem_base, ta + 16 + k * 8, db, idesc, (it > 0 || k > 0) ? 1u : 0u);
          umma_f16_ts(tmem_base, ta + k * 8, dbl, idesc, 1u);
          umma_f16_ts(tmem_base, ta + k * 8, db, idesc, 1u);
        }
        umma_commit(&empty[c]);
        if (it == TRACE_IT) trace_stamp(epi, 13);
        if (it == n_iters - 1) { umma_commit(accum); trace_stamp(epi, 5); }
      }
      __syncwarp();
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    // pixel of this row inside the halo box (box = TN x (TH+2) x (TW+2) pixels, 128 bytes each, SW128)
    const int tw = r % prog.TW, th = (r / prog.TW) % prog.TH, tn = r / (prog.TW * prog.TH);
    const int hp0 = (tn * (prog.TH + 2) + th + 1) * (prog.TW + 2) + tw + 1;
    int it = 0;
    for (int kb = 0; kb < KB; ++kb) {
      const int a = kb % RA;
      mbar_wait(&afull[a], (kb / RA) & 1);
      const uint32_t box = smem_u32(halo(a));
      for (int tap = 0; tap < T; ++tap, ++it) {
        const int c = it % C;
        mbar_wait(&empty[c], ((it / C) & 1) ^ 1);
        tc_fence_after();
        if (r == 0 && it == TRACE_IT) trace_stamp(epi, 8);
        if (r == 0 && it == TRACE_IT + 1) trace_stamp(epi, 14);
        const int hp = hp0 + prog.tap_delta[tap];
        const uint32_t row = box + hp * 128;
        const int sw = hp & 7;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = lds_v4(row + ((j ^ sw) << 4));
          split_bf16x2(v.x, v.y, hi[2 * j], lo[2 * j]);
          split_bf16x2(v.z, v.w, hi[2 * j + 1], lo[2 * j + 1]);
        }
        const uint32_t dst = lane_base + (uint32_t)(BN + 32 * c);
        if (r == 0 && it == TRACE_IT) trace_stamp(epi, 9);
        tmem_st_32x16(dst, hi);
        tmem_st_32x16(dst + 16, lo);
        tmem_st_wait();
        if (r == 0 && it == TRACE_IT) trace_stamp(epi, 10);
        tc_fence_before();
        mbar_arrive(&conv[c]);
      }
      mbar_arrive(&afree[a]);
    }
    gemm_epilogue<BN, 0>(smem, tmem_base, accum, prog, epi, 0, n0, n_img0, h0, w0, q, lane, r);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  if (threadIdx.x == 0) trace_stamp(epi, 7);
}

template <int BN, int OCC>
static int launch_conv_halo(const GemmMaps& maps, const GemmProgram& prog, const GemmEpilogue& epi, dim3 grid,
                            cudaStream_t st) {
  using Cfg = HaloCfg<BN, OCC>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<BN, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("conv_halo: cudaFuncSetAttribute(%d bytes) failed: %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return OBMAN_ERR_CUDA;
    }
    attr = true;
  }
  conv_halo_kernel<BN, OCC><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(maps, prog, epi);
  return check_launch("conv_halo_kernel");
}

int try_conv64(const float* x, int n_img, int h_in, int w_in, int c_in, long long x_sN, long long x_sH, long long x_sW,
               const float* w, int c_out, int w_slots, int num_taps, const int* tap_dh, const int* tap_dw,
               const int* tap_wslot, float* out, int h_out, int w_out, long long o_sN, long long o_sH, long long o_sW,
               const float* bias, const float* addend, const float* mask_src, int relu, cudaStream_t st);   // conv64.cu

int try_gemm_persist(const GemmMaps& maps, const GemmProgram& prog, const GemmEpilogue& epi, long long m_tiles,
                     cudaStream_t st);   // gemm_persist.cu

long long* g_trace = nullptr;       // obman_debug_trace (also read by conv64.cu)
long long g_trace_cap = 0;

// Cluster size the TS path will run with for a grid of grid_x row tiles (the weight tensor maps are built with a
// box of BN / cluster rows, so the host code asks before encoding them).
static int cluster_for(long long grid_x) {
  static int cl = -1;
  if (cl < 0) {
    const char* c = getenv("OBMAN_GEMM_CLUSTER");   // 1 (default), 2 or 4 CTAs share the weight tile by TMA multicast
    cl = c ? atoi(c) : 1;   // measured: no gain on the 3xTF32 path (tensor-pipe bound), see DESIGN.md
    if (cl != 1 && cl != 2 && cl != 4) cl = 1;
  }
  if (cl == 4 && grid_x >= 4) return 4;
  if (cl >= 2 && grid_x >= 2) return 2;
  return 1;
}

// ts: use the A-in-TMEM path (MODE 0, passes 3, pre-split weights)
template <int MODE>
static int dispatch_gemm(int BN, int passes, int ts, const GemmMaps& maps, const GemmProgram& prog,
                         const GemmEpilogue& epi, dim3 grid, cudaStream_t st) {
  static int occ2 = -1;
  if (occ2 < 0) {
    const char* e = getenv("OBMAN_GEMM_OCC");
    occ2 = (e && e[0] == '1') ? 0 : 1;
  }
  const int cl = cluster_for(grid.x);
  GemmProgram prog_r = prog;
  if (MODE == 0 && cl == 1 && grid.y > 1 && grid.x <= 65535 && grid.z == 1) {
    prog_r.raster_n = 1;
    grid = dim3(grid.y, grid.x, 1);
  }
  static int deep_small = -1;
  if (deep_small < 0) {
    // off by default: measured (round 1j) -2 % on the isolated 8x8x512 kernels but +1 % on the step, because a
    // 198 KB CTA keeps the other streams' kernels off its SM
    const char* e = getenv("OBMAN_GEMM_DEEP_SMALL");
    deep_small = (e && e[0] == '1') ? 1 : 0;
  }
#define OBMAN_GEMM_CASE(bn)                                                                         \
  if (BN == bn) {                                                                                   \
    if (MODE == 0 && ts == 3) {                                                                     \
      if (bn == 64) return launch_gemm<64, 3, 0, 3, 2>(maps, prog_r, epi, grid, st);                  \
      set_error("gemm_tc: the stacked-N variant exists for 64-wide tiles only");                    \
      return OBMAN_ERR_UNSUPPORTED;                                                                 \
    }                                                                                               \
    if (MODE == 0 && ts == 2) {                                                                     \
      constexpr int occ = (bn <= 128 ? 2 : 1);                                                      \
      if (cl == 4) return launch_gemm<bn, 3, 0, 2, occ, 4>(maps, prog_r, epi, grid, st);              \
      if (cl == 2) return launch_gemm<bn, 3, 0, 2, occ, 2>(maps, prog_r, epi, grid, st);              \
      /* a grid that cannot put two CTAs on an SM anyway gets the one-CTA configuration: twice the */ \
      /* pipeline stages (6 instead of 3-4) for the same tile                                      */ \
      if (deep_small && (long long)grid.x * grid.y <= num_sms())                                    \
        return launch_gemm<bn, 3, 0, 2, 1>(maps, prog_r, epi, grid, st);                              \
      return launch_gemm<bn, 3, 0, 2, occ>(maps, prog_r, epi, grid, st);                              \
    }                                                                                               \
    if (MODE == 0 && ts && passes == 3) {                                                           \
      constexpr int occ = (bn <= 128 ? 2 : 1);                                                      \
      if (cl == 4) return launch_gemm<bn, 3, 0, 1, occ, 4>(maps, prog_r, epi, grid, st);              \
      if (cl == 2) return launch_gemm<bn, 3, 0, 1, occ, 2>(maps, prog_r, epi, grid, st);              \
      if (bn <= 128 && occ2) return launch_gemm<bn, 3, 0, 1, occ>(maps, prog_r, epi, grid, st);       \
      return launch_gemm<bn, 3, 0, 1, 1>(maps, prog_r, epi, grid, st);                                \
    }                                                                                               \
    if (MODE == 2 && bn == 64)                                                                      \
      return passes == 3 ? launch_gemm<bn, 3, MODE, 0, (bn == 64 ? 2 : 1)>(maps, prog_r, epi, grid, st) \
                         : launch_gemm<bn, 1, MODE, 0, (bn == 64 ? 2 : 1)>(maps, prog_r, epi, grid, st); \
    return passes == 3 ? launch_gemm<bn, 3, MODE, 0, 1>(maps, prog_r, epi, grid, st)                  \
                       : launch_gemm<bn, 1, MODE, 0, 1>(maps, prog_r, epi, grid, st);                 \
  }
  OBMAN_GEMM_CASE(64)
  OBMAN_GEMM_CASE(128)
  OBMAN_GEMM_CASE(256)
#undef OBMAN_GEMM_CASE
  set_error("gemm_tc: unsupported BN=%d", BN);
  return OBMAN_ERR_UNSUPPORTED;
}

static int pick_bn(int N, long long m_tiles) {
  // Column tile: minimise padded columns / relative tile efficiency (measured: 256-wide tiles are the most
  // efficient per column, 64-wide the least), but only take 256 when the grid still covers most of the SMs.
  if (N <= 64) return 64;
  {
    // experiment knob: column tile for grids that would not fill the GPU with 128-wide tiles
    static int small_grid_bn = -1;
    if (small_grid_bn < 0) {
      const char* e = getenv("OBMAN_GEMM_SMALLGRID_BN");
      small_grid_bn = e ? atoi(e) : 0;
    }
    if (small_grid_bn == 64 && m_tiles * ((N + 127) / 128) < num_sms()) return 64;
  }
  const double eff[3] = {0.6, 0.85, 1.0};
  const int bn[3] = {64, 128, 256};
  int best = 128;
  double best_cost = 1e30;
  for (int i = 1; i < 3; ++i) {
    const long long tiles_n = (N + bn[i] - 1) / bn[i];
    if (bn[i] == 256 && m_tiles * tiles_n < (num_sms() * 3) / 4) continue;
    const double cost = (double)(tiles_n * bn[i]) / eff[i];
    if (cost < best_cost) { best_cost = cost; best = bn[i]; }
  }
  return best;
}

// Experimental (default off, untested on hardware at the end of round 1): stacked-N MMAs for 64-wide column tiles on the
// 3xBF16 path, see GemmCfg / TS == 3.  OBMAN_GEMM_STACK64=1 switches it on.
static bool stack64_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OBMAN_GEMM_STACK64");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v != 0;
}

// column-tile width of the TAIL variant (0 = off): OBMAN_GEMM_TAIL = 0 | 128 (default) | 256
static int tail_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OBMAN_GEMM_TAIL");
    v = e ? atoi(e) : 128;
    if (v != 0 && v != 128 && v != 256) v = 128;
  }
  return v;
}

static bool ts_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OBMAN_GEMM_TS");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

}  // namespace obman

using namespace obman;

// out[M,N] = epilogue(alpha * A[M,K] * W[N,K]^T)   (row-major, leading dimensions in elements)
// epilogue: + bias[N] + addend[M,N] ; relu ; mask by (mask_src > 0) ; store or atomicAdd.
// W_lo (nullable): with passes == 3, W must then hold tf32-rounded values and W_lo the residual (W_true =
// W + W_lo, see obman_split_tf32 / obman_fold_conv); enables the A-in-TMEM path.
extern "C" int obman_gemm(const float* A, long long lda, const float* W, const float* W_lo, long long ldw, int M,
                          int N, int K, float* out, long long ldo, const float* bias, const float* addend,
                          const float* mask_src, float alpha, int relu, int accumulate, int passes,
                          void* stream) {
  OBMAN_REQUIRE(A && W && out, "obman_gemm: null argument");
  OBMAN_REQUIRE(M > 0 && N > 0 && K > 0, "obman_gemm: bad sizes M=%d N=%d K=%d", M, N, K);
  OBMAN_REQUIRE(passes == 1 || passes == 3 || passes == OBMAN_PREC_3XBF16,
                "obman_gemm: passes must be 1 (tf32), 3 (3xtf32) or 2 (3xbf16, packed weights)");
  OBMAN_REQUIRE(lda % 4 == 0 && ldw % 4 == 0, "obman_gemm: lda/ldw must be multiples of 4 floats (TMA 16-byte stride)");
  OBMAN_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)W & 15) == 0 && ((uintptr_t)W_lo & 15) == 0,
                "obman_gemm: A/W must be 16-byte aligned");
  GemmMaps maps;
  memset(&maps, 0, sizeof(maps));
  const long long m_tiles = (M + BM - 1) / BM;
  const bool bf = passes == OBMAN_PREC_3XBF16;
  // N = 128 j + (1..4) on a grid that fills the GPU: 128-wide tiles + tail columns on the CUDA cores (TAIL variant).
  // OBMAN_GEMM_TAIL=256 takes 256-wide tiles where N = 256 j + (1..4): measured SLOWER on the decoder layers (K <= 544:
  // one CTA per SM, the prologue and the 256-column epilogue are not hidden behind a second CTA's main loop).
  const int tail_bn = tail_enabled();
  const int n_tail = (bf && tail_bn && N > tail_bn && N % tail_bn >= 1 && N % tail_bn <= 4 && m_tiles >= num_sms() &&
                      cluster_for(m_tiles) == 1) ? N % tail_bn : 0;
  const int N_main = N - n_tail;
  const int BN = n_tail ? tail_bn : pick_bn(N, m_tiles);
  int ts = bf ? 2 : ((W_lo != nullptr) && passes == 3 && ts_enabled());
  if (ts == 2 && BN == 64 && stack64_enabled() && cluster_for(m_tiles) == 1) ts = 3;
  OBMAN_REQUIRE(W_lo == nullptr || passes == 1 || ts == 1, "obman_gemm: pre-split weights need the TS path (OBMAN_GEMM_TS=0 set?)");
  OBMAN_REQUIRE(!bf || (ldw % 32 == 0 && ldw >= (K + 31) / 32 * 32),
                "obman_gemm: packed bf16 weights need ldw = K rounded up to 32 (see obman_pack_bf16)");
  if (bf) passes = 3;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)lda * 4};
    uint32_t box[2] = {BK, BM};
    int rc = make_tensor_map(&maps.a[0], A, 2, dims, strides, box);
    if (rc) return rc;
    uint64_t dimsb[2] = {(uint64_t)(bf ? (K + 31) / 32 * 32 : K), (uint64_t)N};
    uint64_t stridesb[1] = {(uint64_t)ldw * 4};
    uint32_t boxb[2] = {BK, (uint32_t)(ts ? BN / cluster_for(m_tiles) : BN)};
    if (ts == 3) boxb[0] = 16;   // 64-byte-wide boxes: hi half and lo half of a packed K block are fetched separately
    rc = make_tensor_map(&maps.b, W, 2, dimsb, stridesb, boxb,
                         ts == 3 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if (ts == 1) {
      rc = make_tensor_map(&maps.b_lo, W_lo, 2, dimsb, stridesb, boxb);
      if (rc) return rc;
    }
  }
  GemmProgram prog;
  memset(&prog, 0, sizeof(prog));
  prog.spatial = 0;
  prog.num_taps = 1;
  prog.kblocks = (K + BK - 1) / BK;
  prog.M = M;
  prog.N = N_main;
  prog.n_tail = n_tail;
  prog.tail_w = reinterpret_cast<const unsigned char*>(W) + (size_t)N_main * ldw * 4;
  prog.tail_ldw = ldw * 4;
  GemmEpilogue epi;
  memset(&epi, 0, sizeof(epi));
  epi.out = out; epi.bias = bias; epi.addend = addend; epi.mask_src = mask_src;
  epi.alpha = alpha; epi.relu = relu; epi.accumulate = accumulate; epi.ld = ldo;
  epi.trace = g_trace; epi.trace_cap = g_trace_cap;
  dim3 grid((unsigned)m_tiles, (unsigned)((N_main + BN - 1) / BN), 1);
  if (bf && BN == 128 && cluster_for(m_tiles) == 1) {
    // many short tiles (the point decoder): persistent CTAs, epilogue overlapped with the next tile (gemm_persist.cu)
    prog.n_tiles = (int)grid.y;
    const int rc = try_gemm_persist(maps, prog, epi, m_tiles, (cudaStream_t)stream);
    if (rc != 0) return rc < 0 ? rc : OBMAN_OK;
  }
  if (n_tail) {
    if (grid.y > 1 && grid.x <= 65535) {
      prog.raster_n = 1;
      grid = dim3(grid.y, grid.x, 1);
    }
    if (BN == 256) return launch_gemm<256, 3, 0, 2, 1, 1, 1>(maps, prog, epi, grid, (cudaStream_t)stream);
    return launch_gemm<128, 3, 0, 2, 2, 1, 1>(maps, prog, epi, grid, (cudaStream_t)stream);
  }
  return dispatch_gemm<0>(BN, passes, ts, maps, prog, epi, grid, (cudaStream_t)stream);
}

// NHWC convolution as implicit GEMM (forward and data-gradient share this entry point).
//   x    : (n_img, h_in, w_in, c_in) NHWC, c_in % 4 == 0
//   w    : (c_out, w_slots * c_in) K-major; tap t reads weight slot tap_wslot[t] (default t), i.e.
//          columns [slot*c_in, (slot+1)*c_in).  w_lo (nullable): residual of pre-split weights (see obman_gemm).
//   taps : for tap t the input pixel of output (h, w) is (h + dh[t], w + dw[t]) of phase view `tap_phase[t]`
//          (ph*2+pw) when in_step == 2, i.e. of x[:, ph::2, pw::2, :]; with in_step == 1 there is a single view.
//   out  : written at element offset n*o_sN + h*o_sH + w*o_sW + c   (lets dgrad write strided phases)
extern "C" int obman_conv_nhwc(const float* x, int n_img, int h_in, int w_in, int c_in, int in_step,
                               long long x_sN, long long x_sH, long long x_sW, const float* w, const float* w_lo, int c_out, int w_slots, int num_taps,
                               const int* tap_dh, const int* tap_dw, const int* tap_phase, const int* tap_wslot,
                               float* out, int h_out, int w_out, long long o_sN, long long o_sH, long long o_sW,
                               const float* bias, const float* addend, const float* mask_src, int relu,
                               int passes, void* stream) {
  OBMAN_REQUIRE(x && w && out && tap_dh && tap_dw, "obman_conv_nhwc: null argument");
  OBMAN_REQUIRE(n_img > 0 && h_in > 0 && w_in > 0 && c_in > 0 && c_out > 0 && h_out > 0 && w_out > 0,
                "obman_conv_nhwc: bad sizes");
  OBMAN_REQUIRE(c_in % 4 == 0, "obman_conv_nhwc: c_in=%d must be a multiple of 4", c_in);
  OBMAN_REQUIRE(num_taps >= 1 && num_taps <= MAX_TAPS, "obman_conv_nhwc: num_taps=%d out of [1,%d]", num_taps, MAX_TAPS);
  OBMAN_REQUIRE(in_step == 1 || in_step == 2, "obman_conv_nhwc: in_step must be 1 or 2");
  OBMAN_REQUIRE(w_slots >= 1, "obman_conv_nhwc: w_slots must be >= 1");
  OBMAN_REQUIRE(in_step == 1 || (h_in % 2 == 0 && w_in % 2 == 0), "obman_conv_nhwc: phase views need even h_in/w_in");
  OBMAN_REQUIRE(passes == 1 || passes == 3 || passes == OBMAN_PREC_3XBF16, "obman_conv_nhwc: passes must be 1, 2 or 3");
  OBMAN_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0 && ((uintptr_t)w_lo & 15) == 0,
                "obman_conv_nhwc: x/w must be 16-byte aligned");
  // x strides in elements (0 = dense NHWC).  The pixel stride may be SMALLER than c_in: consecutive pixels then
  // overlap (a sliding window over a narrower tensor; the stem reads 4 neighbouring 16-channel pixels as 64 channels).
  const bool x_dense = x_sN == 0 && x_sH == 0 && x_sW == 0;
  if (x_dense) { x_sW = c_in; x_sH = (long long)w_in * c_in; x_sN = (long long)h_in * w_in * c_in; }
  OBMAN_REQUIRE(x_sW > 0 && x_sH > 0 && x_sN > 0 && x_sW % 4 == 0 && x_sH % 4 == 0 && x_sN % 4 == 0,
                "obman_conv_nhwc: x strides must be positive multiples of 4 elements (or all 0)");
  if (passes == OBMAN_PREC_3XBF16 && in_step == 1 && c_out <= 64) {
    // 64-output-channel layers (stem, layer1, the c_out = 64 data-gradient phases): persistent kernel with resident
    // weights, halo tiles and stacked-N MMAs (conv64.cu); shapes it does not cover fall through
    const int rc = try_conv64(x, n_img, h_in, w_in, c_in, x_sN, x_sH, x_sW, w, c_out, w_slots, num_taps, tap_dh, tap_dw,
                              tap_wslot, out, h_out, w_out, o_sN, o_sH, o_sW, bias, addend, mask_src, relu,
                              (cudaStream_t)stream);
    if (rc != 0) return rc < 0 ? rc : OBMAN_OK;
  }
  {
    // halo kernel: stride-1 taps inside the 3x3 neighbourhood, packed bf16 weights
    static int halo_on = -1;
    if (halo_on < 0) {
      const char* e = getenv("OBMAN_CONV_HALO");
      halo_on = (e && e[0] == '1') ? 1 : 0;   // off by default: measured no faster (the loop is MMA-issue bound)
    }
    bool ok = halo_on && x_dense && passes == OBMAN_PREC_3XBF16 && in_step == 1 && num_taps >= 2 && c_in % 32 == 0;
    for (int t = 0; ok && t < num_taps; ++t) ok = tap_dh[t] >= -1 && tap_dh[t] <= 1 && tap_dw[t] >= -1 && tap_dw[t] <= 1;
    int hTW = 1;
    while (hTW * 2 <= w_out && hTW * 2 <= 16) hTW *= 2;
    int hTH = 1;
    while (hTH * 2 <= h_out && hTW * hTH * 2 <= 128) hTH *= 2;
    const int hTN = 128 / (hTW * hTH);
    const int halo_pix = (hTW + 2) * (hTH + 2) * hTN;
    ok = ok && hTW >= 8 && halo_pix * 128 <= HaloCfg<64, 2>::HALO_BYTES;
    if (ok) {
      GemmProgram prog;
      memset(&prog, 0, sizeof(prog));
      prog.spatial = 1;
      prog.num_taps = num_taps;
      prog.kblocks = c_in / BK;
      prog.N = c_out;
      prog.TN = hTN; prog.TH = hTH; prog.TW = hTW;
      prog.tiles_h = (h_out + hTH - 1) / hTH;
      prog.tiles_w = (w_out + hTW - 1) / hTW;
      prog.n_img = n_img; prog.h_out = h_out; prog.w_out = w_out;
      prog.halo_bytes = halo_pix * 128;
      for (int t = 0; t < num_taps; ++t) {
        prog.tap_bk[t] = (tap_wslot ? tap_wslot[t] : t) * c_in;
        prog.tap_delta[t] = tap_dh[t] * (hTW + 2) + tap_dw[t];
      }
      const long long m_tiles = (long long)((n_img + hTN - 1) / hTN) * prog.tiles_h * prog.tiles_w;
      const int BN = c_out <= 64 ? 64 : 128;
      GemmMaps maps;
      memset(&maps, 0, sizeof(maps));
      uint64_t dims[4] = {(uint64_t)c_in, (uint64_t)w_in, (uint64_t)h_in, (uint64_t)n_img};
      uint64_t strides[3] = {(uint64_t)c_in * 4, (uint64_t)w_in * c_in * 4, (uint64_t)h_in * w_in * c_in * 4};
      uint32_t box[4] = {BK, (uint32_t)(hTW + 2), (uint32_t)(hTH + 2), (uint32_t)hTN};
      int rc = make_tensor_map(&maps.a[0], x, 4, dims, strides, box);
      if (rc) return rc;
      uint64_t dimsb[2] = {(uint64_t)w_slots * (uint64_t)c_in, (uint64_t)c_out};
      uint64_t stridesb[1] = {dimsb[0] * 4};
      uint32_t boxb[2] = {BK, (uint32_t)BN};
      rc = make_tensor_map(&maps.b, w, 2, dimsb, stridesb, boxb);
      if (rc) return rc;
      GemmEpilogue epi;
      memset(&epi, 0, sizeof(epi));
      epi.out = out; epi.bias = bias; epi.addend = addend; epi.mask_src = mask_src;
      epi.alpha = 1.f; epi.relu = relu; epi.accumulate = 0;
      epi.sN = o_sN; epi.sH = o_sH; epi.sW = o_sW;
      epi.trace = g_trace; epi.trace_cap = g_trace_cap;
      dim3 grid((unsigned)m_tiles, (unsigned)((c_out + BN - 1) / BN), 1);
      if (BN == 64) return launch_conv_halo<64, 2>(maps, prog, epi, grid, (cudaStream_t)stream);
      return launch_conv_halo<128, 2>(maps, prog, epi, grid, (cudaStream_t)stream);
    }
  }
  // output tile shape: TW = largest power of two <= min(w_out, 128) ... keep TN*TH*TW == 128
  int TW = 1;
  while (TW * 2 <= w_out && TW * 2 <= 128) TW *= 2;
  int TH = 1;
  while (TH * 2 <= h_out && TW * TH * 2 <= 128) TH *= 2;
  int TN = 128 / (TW * TH);
  GemmProgram prog;
  memset(&prog, 0, sizeof(prog));
  prog.spatial = 1;
  prog.num_taps = num_taps;
  prog.kblocks = (c_in + BK - 1) / BK;
  prog.N = c_out;
  prog.TN = TN; prog.TH = TH; prog.TW = TW;
  prog.tiles_h = (h_out + TH - 1) / TH;
  prog.tiles_w = (w_out + TW - 1) / TW;
  prog.n_img = n_img; prog.h_out = h_out; prog.w_out = w_out;
  const long long m_tiles = (long long)((n_img + TN - 1) / TN) * prog.tiles_h * prog.tiles_w;
  const int BN = pick_bn(c_out, m_tiles);
  const bool bf = passes == OBMAN_PREC_3XBF16;
  int ts = bf ? 2 : ((w_lo != nullptr) && passes == 3 && ts_enabled());
  if (ts == 2 && BN == 64 && stack64_enabled() && cluster_for(m_tiles) == 1) ts = 3;
  OBMAN_REQUIRE(w_lo == nullptr || passes == 1 || ts == 1, "obman_conv_nhwc: pre-split weights need the TS path");
  OBMAN_REQUIRE(!bf || c_in % 32 == 0, "obman_conv_nhwc: packed bf16 weights need c_in %% 32 == 0 (c_in=%d)", c_in);
  if (bf) passes = 3;
  GemmMaps maps;
  memset(&maps, 0, sizeof(maps));
  bool used[4] = {false, false, false, false};
  for (int t = 0; t < num_taps; ++t) {
    const int ph = (in_step == 2 && tap_phase) ? tap_phase[t] : 0;
    OBMAN_REQUIRE(ph >= 0 && ph < 4, "obman_conv_nhwc: bad tap phase");
    prog.tap_dh[t] = tap_dh[t];
    prog.tap_dw[t] = tap_dw[t];
    prog.tap_map[t] = ph;
    prog.tap_bk[t] = (tap_wslot ? tap_wslot[t] : t) * c_in;
    used[ph] = true;
  }
  for (int ph = 0; ph < 4; ++ph) {
    if (!used[ph]) continue;
    const int py = ph >> 1, px = ph & 1;
    uint64_t dims[4] = {(uint64_t)c_in, (uint64_t)(w_in / in_step), (uint64_t)(h_in / in_step), (uint64_t)n_img};
    uint64_t strides[3] = {(uint64_t)x_sW * 4 * in_step, (uint64_t)x_sH * 4 * in_step, (uint64_t)x_sN * 4};
    uint32_t box[4] = {BK, (uint32_t)TW, (uint32_t)TH, (uint32_t)TN};
    const float* base = x + (long long)py * x_sH + (long long)px * x_sW;
    int rc = make_tensor_map(&maps.a[ph], base, 4, dims, strides, box);
    if (rc) return rc;
  }
  {
    uint64_t dimsb[2] = {(uint64_t)w_slots * (uint64_t)c_in, (uint64_t)c_out};
    uint64_t stridesb[1] = {dimsb[0] * 4};
    uint32_t boxb[2] = {BK, (uint32_t)(ts ? BN / cluster_for(m_tiles) : BN)};
    if (ts == 3) boxb[0] = 16;   // see obman_gemm
    int rc = make_tensor_map(&maps.b, w, 2, dimsb, stridesb, boxb,
                             ts == 3 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if (ts == 1) {
      rc = make_tensor_map(&maps.b_lo, w_lo, 2, dimsb, stridesb, boxb);
      if (rc) return rc;
    }
  }
  GemmEpilogue epi;
  memset(&epi, 0, sizeof(epi));
  epi.out = out; epi.bias = bias; epi.addend = addend; epi.mask_src = mask_src;
  epi.alpha = 1.f; epi.relu = relu; epi.accumulate = 0;
  epi.sN = o_sN; epi.sH = o_sH; epi.sW = o_sW;
  epi.trace = g_trace; epi.trace_cap = g_trace_cap;
  dim3 grid((unsigned)m_tiles, (unsigned)((c_out + BN - 1) / BN), 1);
  return dispatch_gemm<0>(BN, passes, ts, maps, prog, epi, grid, (cudaStream_t)stream);
}

// Weight gradient of the NHWC convolution above (and, with h = 1, of any row-major matrix product):
//   dw[co, t*c_in + ci] = sum_{n,h,w} dy[n,h,w,co] * xview_t[n, h + dh[t], w + dw[t], ci]     t = 0..num_taps-1
// One GEMM with N = num_taps * c_in: the taps are stacked along N and fetched as shifted boxes, both operands are
// read MN-major (channels contiguous, pixels along K) straight from NHWC; the pixel reduction is split over
// gridDim.z (a whole number of waves) and combined with fp32 atomics.  c_out, c_in multiples of 32.
extern "C" int obman_wgrad_nhwc(const float* dy, int n_img, int h_out, int w_out, int c_out,
                                const float* x, int h_in, int w_in, int c_in, int in_step, long long x_sN,
                                long long x_sH, long long x_sW, int num_taps,
                                const int* tap_dh, const int* tap_dw, const int* tap_phase, float* dw,
                                float* dy_colsum, int passes, void* stream) {
  OBMAN_REQUIRE(dy && x && dw && tap_dh && tap_dw, "obman_wgrad_nhwc: null argument");
  OBMAN_REQUIRE(dy_colsum == nullptr || passes == OBMAN_PREC_3XBF16,
                "obman_wgrad_nhwc: the fused column sum of dy exists on the 3xBF16 path only");
  OBMAN_REQUIRE(n_img > 0 && h_out > 0 && w_out > 0 && h_in > 0 && w_in > 0, "obman_wgrad_nhwc: bad sizes");
  OBMAN_REQUIRE(c_out > 0 && c_in > 0 && c_out % 32 == 0 && c_in % 32 == 0,
                "obman_wgrad_nhwc: c_out=%d and c_in=%d must be multiples of 32", c_out, c_in);
  OBMAN_REQUIRE(num_taps >= 1 && num_taps <= MAX_TAPS, "obman_wgrad_nhwc: bad tap count");
  OBMAN_REQUIRE(in_step == 1 || in_step == 2, "obman_wgrad_nhwc: in_step must be 1 or 2");
  OBMAN_REQUIRE(in_step == 1 || (h_in % 2 == 0 && w_in % 2 == 0), "obman_wgrad_nhwc: phase views need even h_in/w_in");
  OBMAN_REQUIRE(passes == 1 || passes == 3 || passes == OBMAN_PREC_3XBF16, "obman_wgrad_nhwc: passes must be 1, 2 or 3");
  const bool bf = passes == OBMAN_PREC_3XBF16;
  OBMAN_REQUIRE(((uintptr_t)dy & 15) == 0 && ((uintptr_t)x & 15) == 0, "obman_wgrad_nhwc: dy/x must be 16-byte aligned");
  if (x_sN == 0 && x_sH == 0 && x_sW == 0) { x_sW = c_in; x_sH = (long long)w_in * c_in; x_sN = (long long)h_in * w_in * c_in; }
  OBMAN_REQUIRE(x_sW > 0 && x_sH > 0 && x_sN > 0 && x_sW % 4 == 0 && x_sH % 4 == 0 && x_sN % 4 == 0,
                "obman_wgrad_nhwc: x strides must be positive multiples of 4 elements (or all 0)");
  cudaStream_t st = (cudaStream_t)stream;
  int kTW = 1;
  while (kTW * 2 <= w_out && kTW * 2 <= 32) kTW *= 2;
  int kTH = 1;
  while (kTH * 2 <= h_out && kTW * kTH * 2 <= 32) kTH *= 2;
  const int kTN = 32 / (kTW * kTH);
  GemmProgram prog;
  memset(&prog, 0, sizeof(prog));
  prog.num_taps = num_taps;
  prog.kblocks = 1;
  prog.M = c_out;
  prog.cg_in = c_in / 32;
  prog.total_groups = num_taps * prog.cg_in;
  prog.N = prog.total_groups * 32;
  prog.kTN = kTN; prog.kTH = kTH; prog.kTW = kTW;
  prog.mn_lbo = 4096; prog.mn_sbo = 512; prog.mn_layout = 1;
  // 3xBF16: the tiles are read by the splitter warps, not by the tensor core: no swizzle (column reads are
  // conflict-free); 3xTF32: MN-major TF32 operands need the 128-byte swizzle with 32-byte atoms
  const CUtensorMapSwizzle mn_swizzle = bf ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  prog.kblocks_n = (n_img + kTN - 1) / kTN;
  prog.kblocks_h = (h_out + kTH - 1) / kTH;
  prog.kblocks_w = (w_out + kTW - 1) / kTW;
  // c_out <= 64 would leave half of the 128 accumulator rows empty: swap the roles (rows = stacked taps x c_in,
  // columns = output channels) so that the tile is full; dw is then written transposed by the epilogue.
  const bool swapped = c_out <= 64;
  const int n_total = prog.N;
  int m_tiles = (c_out + BM - 1) / BM;
  int BN = prog.N > 128 ? 256 : (prog.N > 64 ? 128 : 64);
  prog.n_tiles = (prog.N + BN - 1) / BN;
  if (swapped) {
    prog.M = n_total;
    prog.N = c_out;
    m_tiles = (n_total + BM - 1) / BM;
    BN = 64;
    prog.n_tiles = 1;
  }
  // 3xBF16: one TMA fetches `grp_per_load` consecutive 32-channel groups of a tap (largest power of two that
  // divides c_in / 32 and the tile's group count); the 3xTF32 kernels load group by group
  prog.grp_per_load = 1;
  if (bf) {
    const int cap = swapped ? BM / 32 : 8;
    while (prog.grp_per_load * 2 <= cap && prog.cg_in % (prog.grp_per_load * 2) == 0) prog.grp_per_load *= 2;
  }
  GemmMaps maps;
  memset(&maps, 0, sizeof(maps));
  {
    uint64_t dims[5] = {32, (uint64_t)w_out, (uint64_t)h_out, (uint64_t)n_img, (uint64_t)(c_out / 32)};
    uint64_t strides[4] = {(uint64_t)c_out * 4, (uint64_t)w_out * c_out * 4, (uint64_t)h_out * w_out * c_out * 4, 128};
    uint32_t box[5] = {32, (uint32_t)kTW, (uint32_t)kTH, (uint32_t)kTN, swapped ? (uint32_t)(BN / 32) : 4u};
    int rc = make_tensor_map(&maps.a[0], dy, 5, dims, strides, box, mn_swizzle);
    if (rc) return rc;
  }
  bool used[4] = {false, false, false, false};
  for (int t = 0; t < num_taps; ++t) {
    const int ph = (in_step == 2 && tap_phase) ? tap_phase[t] : 0;
    OBMAN_REQUIRE(ph >= 0 && ph < 4, "obman_wgrad_nhwc: bad tap phase");
    prog.tap_dh[t] = tap_dh[t];
    prog.tap_dw[t] = tap_dw[t];
    prog.tap_map[t] = ph;
    used[ph] = true;
  }
  for (int ph = 0; ph < 4; ++ph) {
    if (!used[ph]) continue;
    const int py = ph >> 1, px = ph & 1;
    uint64_t dims[5] = {32, (uint64_t)(w_in / in_step), (uint64_t)(h_in / in_step), (uint64_t)n_img, (uint64_t)(c_in / 32)};
    uint64_t strides[4] = {(uint64_t)x_sW * 4 * in_step, (uint64_t)x_sH * 4 * in_step, (uint64_t)x_sN * 4, 128};
    uint32_t box[5] = {32, (uint32_t)kTW, (uint32_t)kTH, (uint32_t)kTN, (uint32_t)prog.grp_per_load};
    const float* base = x + (long long)py * x_sH + (long long)px * x_sW;
    int rc = make_tensor_map(&maps.a[1 + ph], base, 5, dims, strides, box, mn_swizzle);
    if (rc) return rc;
  }
  const long long total_blocks = (long long)prog.kblocks_n * prog.kblocks_h * prog.kblocks_w;
  const long long tiles = (long long)m_tiles * prog.n_tiles;
  // split the pixel reduction so that tiles * splits is a whole number of waves (one CTA per SM)
  const long long slots = (long long)num_sms() * (swapped ? 2 : 1);  // the swapped configuration runs 2 CTAs / SM
  const long long waves = (tiles + slots - 1) / slots;
  long long splits = (waves * slots) / tiles;
  if (splits > total_blocks / 4) splits = total_blocks / 4;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  const long long ld = (long long)n_total;
  if (splits > 1) cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)c_out * ld, st);
  GemmEpilogue epi;
  memset(&epi, 0, sizeof(epi));
  epi.out = dw;
  epi.colsum = dy_colsum;
  if (dy_colsum) cudaMemsetAsync(dy_colsum, 0, sizeof(float) * (size_t)c_out, st);
  epi.alpha = 1.f;
  epi.accumulate = splits > 1;
  epi.ld = ld;
  epi.trace = g_trace; epi.trace_cap = g_trace_cap;
  dim3 grid((unsigned)(m_tiles * prog.n_tiles), 1, (unsigned)splits);
  if (bf) {
    if (swapped) {
      static int wstack = -1;
      if (wstack < 0) {
        const char* e = getenv("OBMAN_WGRAD_STACK64");   // experimental, default off (not yet run on hardware)
        wstack = (e && e[0] == '1') ? 1 : 0;
      }
      if (wstack) return launch_wgrad_bf16<64, 2, 2>(maps, prog, epi, grid, st);
      return launch_wgrad_bf16<64, 1, 2>(maps, prog, epi, grid, st);
    }
    if (BN == 256) return launch_wgrad_bf16<256, 0, 1>(maps, prog, epi, grid, st);
    if (BN == 128) return launch_wgrad_bf16<128, 0, 1>(maps, prog, epi, grid, st);
    return launch_wgrad_bf16<64, 0, 1>(maps, prog, epi, grid, st);
  }
  if (swapped) return dispatch_gemm<2>(BN, passes, 0, maps, prog, epi, grid, st);
  return dispatch_gemm<1>(BN, passes, 0, maps, prog, epi, grid, st);
}

// hi = tf32-rounded copy of w, lo = w - hi (elementwise, n floats): weights pre-split for the TS path.
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ w, long long n,
                                                         float* __restrict__ hi, float* __restrict__ lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = w[i];
  const float h = obman::sm100::to_tf32_rna(v);
  hi[i] = h;
  lo[i] = v - h;
}

// out (rows, ld_out) <- w (rows, K; row stride ldw) in the packed layout of the 3xBF16 path: per 32-element block
// of a row, 32 bf16 hi values then 32 bf16 lo = bf16(w - hi) values (128 bytes, the bytes of 32 floats); columns
// K .. ld_out-1 are zero.  ld_out = K rounded up to 32.
__global__ void __launch_bounds__(256) pack_bf16_kernel(const float* __restrict__ w, long long ldw, int rows, int K,
                                                        uint32_t* __restrict__ out, long long ld_out) {
  // one thread per PAIR of consecutive elements
  const long long pairs_per_row = ld_out / 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pairs_per_row * rows) return;
  const long long r = i / pairs_per_row;
  const int k = (int)(i - r * pairs_per_row) * 2;
  const float v0 = k < K ? w[r * ldw + k] : 0.f;
  const float v1 = k + 1 < K ? w[r * ldw + k + 1] : 0.f;
  uint32_t hi, lo;
  obman::sm100::split_bf16x2(v0, v1, hi, lo);
  const int blk = k >> 5, in = (k & 31) >> 1;
  out[r * ld_out + blk * 32 + in] = hi;
  out[r * ld_out + blk * 32 + 16 + in] = lo;
}

extern "C" int obman_pack_bf16(const float* w, long long ldw, int rows, int K, float* out, long long ld_out,
                               void* stream) {
  OBMAN_REQUIRE(w && out && rows > 0 && K > 0, "obman_pack_bf16: bad arguments");
  OBMAN_REQUIRE(ld_out % 32 == 0 && ld_out >= K && ldw >= K, "obman_pack_bf16: ld_out must be a multiple of 32 >= K");
  const long long n = ld_out / 2 * rows;
  pack_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      w, ldw, rows, K, reinterpret_cast<uint32_t*>(out), ld_out);
  return check_launch("pack_bf16_kernel");
}

// Diagnostics: while buf != NULL every tensor-core kernel launch writes 8 clock64 stamps per CTA into
// buf[cta * 16 + k] (k = 0 entry, 1 setup done, 2 first / 3 last TMA issued, 4 first operands ready, 5 last MMA
// issued, 6 accumulator complete, 7 exit | smid << 48).  cap = capacity in 8-byte entries.  NULL switches it off.
extern "C" int obman_debug_trace(long long* buf, long long cap) {
  g_trace = buf;
  g_trace_cap = buf ? cap : 0;
  return OBMAN_OK;
}

extern "C" int obman_split_tf32(const float* w, long long n, float* hi, float* lo, void* stream) {
  OBMAN_REQUIRE(w && hi && lo && n > 0, "obman_split_tf32: bad arguments");
  split_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, n, hi, lo);
  return check_launch("split_tf32_kernel");
}
