// Declarations shared by the tensor-core kernels (gemm_tc.cu, conv64.cu): tile constants, the kernel parameter
// structs, the per-CTA trace hook, the epilogue operand prefetch and the epilogue of the stacked-N variants.
#pragma once
#include "common.cuh"
#include "sm100.cuh"
#include "tensormap.cuh"

namespace obman {
using namespace sm100;

constexpr int BM = 128;
constexpr int BK = 32;                  // fp32 elements per 128-byte swizzle row
constexpr int A_TILE_BYTES = BM * BK * 4;
constexpr int GEMM_THREADS = 192;
constexpr int MAX_TAPS = 16;

struct alignas(64) GemmMaps {
  CUtensorMap a[5];   // MODE 0: up to 4 phase views of the activations.  MODE 1: a[0] = dY, a[1..4] = x views
  CUtensorMap b;      // MODE 0: weights (hi part when TS)
  CUtensorMap b_lo;   // MODE 0, TS: weights lo part
};

struct GemmProgram {
  int spatial;          // 0: A is a (M,K) matrix; 1: A is NHWC with shifted-box taps
  int num_taps;
  int kblocks;          // 32-wide K blocks per tap
  int tap_dh[MAX_TAPS], tap_dw[MAX_TAPS], tap_map[MAX_TAPS], tap_bk[MAX_TAPS];
  int M, N;             // rows / cols of D that exist
  // spatial output tiling: tile = TN images x TH rows x TW cols (TN*TH*TW == 128)
  int TN, TH, TW, tiles_h, tiles_w;
  int n_img, h_out, w_out;
  // wgrad (MODE 1): K runs over blocks of 32 output pixels (kTN x kTH x kTW), split over gridDim.z
  int kTN, kTH, kTW, kblocks_n, kblocks_h, kblocks_w;
  int n_tiles;          // column tiles (blockIdx.x = m_tile * n_tiles + n_tile)
  int cg_in;            // 32-channel groups per tap of the input (c_in / 32)
  int total_groups;     // num_taps * cg_in
  unsigned mn_lbo, mn_sbo, mn_layout;  // MN-major smem descriptor fields (bytes, bytes, layout type)
  int grp_per_load;     // bf16 wgrad: consecutive channel groups of a tap fetched by one TMA (divides cg_in)
  // halo kernel: bytes of one halo box and, per tap, its pixel offset inside the box (dh * (TW + 2) + dw)
  int halo_bytes;
  int tap_delta[MAX_TAPS];
  // persistent 64-wide kernel (conv64.cu): halo box = halo_w x halo_h pixels per image of the tile, its origin is the
  // tile origin + (halo_dw0, halo_dh0); halo_pix = pixels of the whole box (all TN images)
  int halo_w, halo_h, halo_dw0, halo_dh0, halo_pix;
  int halo_stride, halo_ring;   // bytes between halo buffers (box rounded up to 1 KB), number of buffers
  int run_cols;                 // conv64 second generation: accumulator columns = flattened halo run of a tile
  int raster_n;         // MODE 0: blockIdx.x = column tile, blockIdx.y = row tile
  // gemm_tc_kernel TAIL variant: the last n_tail (1..4) output columns N .. N+n_tail-1 are not given a column tile of their
  // own; the splitter threads of column tile 0 evaluate them on the CUDA cores from the rows they convert anyway.
  // tail_w = packed bf16 hi|lo weight rows of those columns, tail_ldw = their row stride in bytes
  int n_tail;
  const unsigned char* tail_w;
  long long tail_ldw;
  int debug_skip;       // conv64.cu diagnostics (OBMAN_CONV64_DEBUG), 0 in normal operation
};

struct GemmEpilogue {
  float* out;
  const float* bias;      // per output column, nullable
  const float* addend;    // same indexing as out, nullable
  const float* mask_src;  // same indexing as out, nullable: out = mask_src > 0 ? v : 0
  float alpha;            // v = alpha * acc + bias + addend
  int relu;
  int accumulate;         // atomicAdd into out instead of store
  // row -> element offset: plain: row * ld ; spatial: n*sN + h*sH + w*sW  (+ column)
  long long ld, sN, sH, sW;
  float* colsum;          // weight-gradient kernels: per-channel sum of dY over all pixels (atomicAdd), nullable
  // diagnostics (obman_debug_trace): 16 clock64 stamps per CTA, NULL in normal operation
  long long* trace;
  long long trace_cap;
};

// stamp slot `k` of this CTA's trace record (slot 7 also gets the SM id in its upper bits)
__device__ __forceinline__ void trace_stamp(const GemmEpilogue& epi, int k) {
  if (epi.trace == nullptr) return;
  const long long cta = blockIdx.x + (long long)gridDim.x * (blockIdx.y + (long long)gridDim.y * blockIdx.z);
  if ((cta + 1) * 16 > epi.trace_cap) return;
  long long t = clock64();
  if (k == 7) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    t = (t & 0x0000ffffffffffffLL) | ((long long)smid << 48);
  }
  epi.trace[cta * 16 + k] = t;
}

// wait-time accounting (obman_debug_trace): a timed mbarrier wait adds the cycles it blocked to `acc`
__device__ __forceinline__ void mbar_wait_timed(uint64_t* bar, uint32_t parity, bool timed, long long& acc) {
  if (!timed) { mbar_wait(bar, parity); return; }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t0;
}
__device__ __forceinline__ void trace_put(const GemmEpilogue& epi, int k, long long v) {
  if (epi.trace == nullptr || ((long long)blockIdx.x + 1) * 16 > epi.trace_cap) return;
  epi.trace[(long long)blockIdx.x * 16 + k] = v;
}

// Epilogue operand prefetch for one row (16 bytes per lane): residual addend and ReLU-mask source of columns
// col .. col+3 of the row at element offset row_off; neutral values when the row / column / alignment rules it out.
__device__ __forceinline__ void epilogue_prefetch(const GemmEpilogue& epi, const GemmProgram& prog, long long row_off,
                                                  int col, bool ptr_ok, float4& add, float4& msk) {
  add = make_float4(0.f, 0.f, 0.f, 0.f);
  msk = make_float4(1.f, 1.f, 1.f, 1.f);
  if (ptr_ok && row_off >= 0 && (row_off & 3) == 0 && col + 3 < prog.N) {
    if (epi.addend) add = __ldg(reinterpret_cast<const float4*>(epi.addend + row_off + col));
    if (epi.mask_src) msk = __ldg(reinterpret_cast<const float4*>(epi.mask_src + row_off + col));
  }
}

// Epilogue shared by the GEMM kernels: wait for the accumulator, then TMEM -> registers -> global.
// Called by the four epilogue warps (q = TMEM lane quadrant, r = q * 32 + lane = tile row); `smem` is the tile
// buffer base (its first 16 KB are reused for staging, every operand read has retired by then; the persistent kernel
// passes a dedicated 16 KB staging area, its accumulator slot as tmem_base and the slot's barrier phase).
template <int BN, int MODE>
__device__ __forceinline__ void gemm_epilogue(uint8_t* smem, uint32_t tmem_base, uint64_t* accum,
                                              const GemmProgram& prog, const GemmEpilogue& epi, int m0, int n0,
                                              int n_img0, int h0, int w0, int q, int lane, int r,
                                              uint32_t acc_phase = 0) {
    // ---- epilogue ----
    // TMEM -> registers (lane = tile row) -> XOR-swizzled shared-memory transpose -> each store / addend / mask
    // instruction touches 4 rows x 128 contiguous bytes (the row-per-lane layout would touch 32 rows x 16 bytes).
    // The global reads (residual addend, ReLU-mask source) of the first 32-column chunk are issued BEFORE the wait for
    // the accumulator, so their latency hides behind the last MMAs.  (Fetching chunk c+1 row group by row group while
    // chunk c is consumed was measured slower: the late rows' loads are exposed again and interleave with the stores.)
    const int cc = lane & 7;     // 16-byte column chunk handled by this lane after the transpose
    const int rsub = lane >> 3;  // row within each group of 4
    long long ro[8];             // element offsets of the 8 rows this lane stores (row 4 i + rsub of the warp's 32), -1 = none
    float4 add4[8], msk4[8];
    bool ptr_ok = false;
    if (MODE != 2) {
      bool row_ok;
      long long row_off;
      if (MODE == 0 && prog.spatial) {
        const int tw = r % prog.TW;
        const int th = (r / prog.TW) % prog.TH;
        const int tn = r / (prog.TW * prog.TH);
        const int n = n_img0 + tn, h = h0 + th, w = w0 + tw;
        row_ok = n < prog.n_img && h < prog.h_out && w < prog.w_out;
        row_off = n * epi.sN + h * epi.sH + w * epi.sW;
      } else {
        row_ok = (m0 + r) < prog.M;
        row_off = (long long)(m0 + r) * epi.ld;
      }
      const long long mine = row_ok ? row_off : -1;
#pragma unroll
      for (int i = 0; i < 8; ++i) ro[i] = __shfl_sync(0xffffffffu, mine, 4 * i + rsub);
      ptr_ok = ((reinterpret_cast<uintptr_t>(epi.out) | reinterpret_cast<uintptr_t>(epi.addend) |
                 reinterpret_cast<uintptr_t>(epi.mask_src)) & 15) == 0 && !epi.accumulate;
#pragma unroll
      for (int i = 0; i < 8; ++i) epilogue_prefetch(epi, prog, ro[i], n0 + 4 * cc, ptr_ok, add4[i], msk4[i]);
    }
    mbar_wait(accum, acc_phase);
    tc_fence_after();
    if (r == 0) trace_stamp(epi, 6);
    if (MODE == 2) {
      // D[m, n] with m = stacked (tap, c_in) index and n = output channel; dw is (c_out, taps*c_in) row-major, so
      // element (m, n) lives at n*ld + m: the 32 lanes of a warp (consecutive m) make every column one coalesced access.
      const bool row_ok = (m0 + r) < prog.M;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (n0 + c0 >= prog.N) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
        if (!row_ok) continue;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = n0 + c0 + j;
          if (col >= prog.N) break;
          float* dst = epi.out + (long long)col * epi.ld + (m0 + r);
          const float y = epi.alpha * __uint_as_float(v[j]);
          if (epi.accumulate) atomicAdd(dst, y);
          else *dst = y;
        }
      }
    } else {
    // all MMAs (and therefore all TMA loads and operand reads) have retired: stage 0 is free for staging
    uint8_t* wbase = smem + q * 4096;                                   // 32 rows x 128 B per warp
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n0 + c0 >= prog.N) break;  // warp-uniform
      const int col = n0 + c0 + 4 * cc;
      const bool colvec = (col + 3 < prog.N) && ptr_ok;
      // issue every global read of this chunk (residual, ReLU-mask source, bias) before touching the accumulator:
      // they are independent, so their latencies overlap instead of forming 8 serial round trips (chunk 0 was
      // fetched before the accumulator wait)
      if (c0 > 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) epilogue_prefetch(epi, prog, ro[i], col, ptr_ok, add4[i], msk4[i]);
      }
      float bias4[4] = {0.f, 0.f, 0.f, 0.f};
      if (epi.bias) {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (col + e < prog.N) bias4[e] = __ldg(epi.bias + col + e);
      }
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      // Keep the mask words opaque until here.  Without this the compiler folds every mask load into predicate bits
      // right behind the load (FSETP on the freshly loaded registers), which turns the 8 independent loads into 8
      // serial memory round trips per chunk: +42 % on the 64-channel data-gradient kernels (scripts/ab_conv.py).
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("" : "+f"(msk4[i].x), "+f"(msk4[i].y), "+f"(msk4[i].z), "+f"(msk4[i].w));
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(wbase + lane * 128 + ((c ^ (lane & 7)) << 4)) =
            make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = 4 * i + rsub;
        const long long row_off = ro[i];
        if (row_off >= 0 && col < prog.N) {
          const float4 a = *reinterpret_cast<const float4*>(wbase + rr * 128 + ((cc ^ (rr & 7)) << 4));
          float x[4] = {epi.alpha * a.x + bias4[0], epi.alpha * a.y + bias4[1], epi.alpha * a.z + bias4[2],
                        epi.alpha * a.w + bias4[3]};
          if (colvec && ((row_off & 3) == 0)) {
            x[0] += add4[i].x; x[1] += add4[i].y; x[2] += add4[i].z; x[3] += add4[i].w;
            if (epi.relu) {
#pragma unroll
              for (int e = 0; e < 4; ++e) x[e] = fmaxf(x[e], 0.f);
            }
            x[0] = msk4[i].x > 0.f ? x[0] : 0.f; x[1] = msk4[i].y > 0.f ? x[1] : 0.f;
            x[2] = msk4[i].z > 0.f ? x[2] : 0.f; x[3] = msk4[i].w > 0.f ? x[3] : 0.f;
            *reinterpret_cast<float4*>(epi.out + row_off + col) = make_float4(x[0], x[1], x[2], x[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (col + e >= prog.N) break;
              float y = x[e];
              if (epi.addend) y += epi.addend[row_off + col + e];
              if (epi.relu) y = fmaxf(y, 0.f);
              if (epi.mask_src) y = epi.mask_src[row_off + col + e] > 0.f ? y : 0.f;
              if (epi.accumulate) atomicAdd(epi.out + row_off + col + e, y);
              else epi.out[row_off + col + e] = y;
            }
          }
        }
      }
      __syncwarp();  // staging is overwritten by the next chunk
    }
    }  // MODE != 2
}

// Epilogue of the stacked-N variant (TS == 3, MODE 0): the result is the SUM of accumulator columns [c] and [BN + c]
// (a_hi*b_hi + a_lo*b_hi and a_hi*b_lo).  16-column chunks (two 16-register TMEM reads instead of one 32-register read),
// staging 32 rows x 64 bytes per warp, every store / addend / mask instruction covers 8 rows x 64 contiguous bytes.
template <int BN>
__device__ __forceinline__ void gemm_epilogue_stacked(uint8_t* smem, uint32_t tmem_base, uint64_t* accum,
                                                      const GemmProgram& prog, const GemmEpilogue& epi, int m0, int n0,
                                                      int n_img0, int h0, int w0, int q, int lane, int r,
                                                      uint32_t acc_phase = 0, int c_begin = 0, int c_end = BN) {
  const int cc = lane & 3;     // 16-byte column chunk of the 64-byte staged row
  const int rsub = lane >> 2;  // row within each group of 8
  long long ro[4];
  float4 add4[4], msk4[4];
  bool row_ok;
  long long row_off;
  if (prog.spatial) {
    const int tw = r % prog.TW;
    const int th = (r / prog.TW) % prog.TH;
    const int tn = r / (prog.TW * prog.TH);
    const int n = n_img0 + tn, h = h0 + th, w = w0 + tw;
    row_ok = n < prog.n_img && h < prog.h_out && w < prog.w_out;
    row_off = n * epi.sN + h * epi.sH + w * epi.sW;
  } else {
    row_ok = (m0 + r) < prog.M;
    row_off = (long long)(m0 + r) * epi.ld;
  }
  const long long mine = row_ok ? row_off : -1;
#pragma unroll
  for (int i = 0; i < 4; ++i) ro[i] = __shfl_sync(0xffffffffu, mine, 8 * i + rsub);
  const bool ptr_ok = ((reinterpret_cast<uintptr_t>(epi.out) | reinterpret_cast<uintptr_t>(epi.addend) |
                        reinterpret_cast<uintptr_t>(epi.mask_src)) & 15) == 0 && !epi.accumulate;
#pragma unroll
  for (int i = 0; i < 4; ++i) epilogue_prefetch(epi, prog, ro[i], n0 + c_begin + 4 * cc, ptr_ok, add4[i], msk4[i]);
  mbar_wait(accum, acc_phase);
  tc_fence_after();
  if (r == 0) trace_stamp(epi, 6);
  uint8_t* wbase = smem + q * 2048;   // 32 rows x 64 B per warp (stage 0 is free: every operand read has retired)
#pragma unroll 1
  for (int c0 = c_begin; c0 < c_end; c0 += 16) {
    if (n0 + c0 >= prog.N) break;  // warp-uniform
    const int col = n0 + c0 + 4 * cc;
    const bool colvec = (col + 3 < prog.N) && ptr_ok;
    if (c0 > c_begin) {
#pragma unroll
      for (int i = 0; i < 4; ++i) epilogue_prefetch(epi, prog, ro[i], col, ptr_ok, add4[i], msk4[i]);
    }
    float bias4[4] = {0.f, 0.f, 0.f, 0.f};
    if (epi.bias) {
#pragma unroll
      for (int e = 0; e < 4; ++e) if (col + e < prog.N) bias4[e] = __ldg(epi.bias + col + e);
    }
    uint32_t v[16], v2[16];
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    tmem_ld_32x16(lane_addr + (uint32_t)c0, v);
    tmem_ld_32x16(lane_addr + (uint32_t)(BN + c0), v2);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("" : "+f"(msk4[i].x), "+f"(msk4[i].y), "+f"(msk4[i].z), "+f"(msk4[i].w));
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float4 t;
      t.x = __uint_as_float(v[4 * c]) + __uint_as_float(v2[4 * c]);
      t.y = __uint_as_float(v[4 * c + 1]) + __uint_as_float(v2[4 * c + 1]);
      t.z = __uint_as_float(v[4 * c + 2]) + __uint_as_float(v2[4 * c + 2]);
      t.w = __uint_as_float(v[4 * c + 3]) + __uint_as_float(v2[4 * c + 3]);
      *reinterpret_cast<float4*>(wbase + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) = t;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = 8 * i + rsub;
      const long long roff = ro[i];
      if (roff < 0 || col >= prog.N) continue;
      const float4 a = *reinterpret_cast<const float4*>(wbase + rr * 64 + ((cc ^ ((rr >> 1) & 3)) << 4));
      float x[4] = {epi.alpha * a.x + bias4[0], epi.alpha * a.y + bias4[1], epi.alpha * a.z + bias4[2],
                    epi.alpha * a.w + bias4[3]};
      if (colvec && ((roff & 3) == 0)) {
        x[0] += add4[i].x; x[1] += add4[i].y; x[2] += add4[i].z; x[3] += add4[i].w;
        if (epi.relu) {
#pragma unroll
          for (int e = 0; e < 4; ++e) x[e] = fmaxf(x[e], 0.f);
        }
        x[0] = msk4[i].x > 0.f ? x[0] : 0.f; x[1] = msk4[i].y > 0.f ? x[1] : 0.f;
        x[2] = msk4[i].z > 0.f ? x[2] : 0.f; x[3] = msk4[i].w > 0.f ? x[3] : 0.f;
        *reinterpret_cast<float4*>(epi.out + roff + col) = make_float4(x[0], x[1], x[2], x[3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (col + e >= prog.N) break;
          float y = x[e];
          if (epi.addend) y += epi.addend[roff + col + e];
          if (epi.relu) y = fmaxf(y, 0.f);
          if (epi.mask_src) y = epi.mask_src[roff + col + e] > 0.f ? y : 0.f;
          if (epi.accumulate) atomicAdd(epi.out + roff + col + e, y);
          else epi.out[roff + col + e] = y;
        }
      }
    }
    __syncwarp();  // staging is overwritten by the next chunk
  }
}

}  // namespace obman
