// Tiled pairwise-distance / arg-min kernel shared by the Chamfer loss and the contact loss.
//
// Replaces the reference's materialised (B,M,N) distance matrix
//   ChamferLoss.forward / batch_pairwise_dist   mano_train/networks/branches/atlasutils.py:11-39
//   batch_pairwise_dist + 2x torch.min          mano_train/networks/branches/contactloss.py:60-79,164-166
// by a register-tiled search that never leaves the SM: every thread owns Q query points in
// registers, the CTA streams the other cloud through shared memory in 1024-point chunks
// (AoS xyz -> float4, one broadcast LDS.128 per candidate), distances use the direct
// (x-y)^2 form (more accurate than the reference's |x|^2+|y|^2-2x.y expansion), running
// (min, argmin) live in registers, ties keep the lowest candidate index.
//
// Algorithmic traffic per launch: 12*B*(N+M) bytes read, 8*B*(N+M) bytes written (min + idx);
// work: 2*B*N*M pair evaluations when both directions are requested.
#include <stdlib.h>

#include "common.cuh"

namespace obman {

constexpr int NN_THREADS = 128;
constexpr int NN_CHUNK = 1024;

template <int Q>
__global__ void __launch_bounds__(NN_THREADS)
nn_kernel(const float* __restrict__ x, const float* __restrict__ y, int N, int M,
          float* __restrict__ minx, int* __restrict__ idxx,
          float* __restrict__ miny, int* __restrict__ idxy, int dirs) {
  const int dir = (dirs == 3) ? (int)blockIdx.z : (dirs == 2 ? 1 : 0);
  const int nq = dir == 0 ? N : M;
  const int nc = dir == 0 ? M : N;
  const int q0 = blockIdx.x * (NN_THREADS * Q);
  if (q0 >= nq) return;
  const int b = blockIdx.y;
  const float* __restrict__ qp = (dir == 0 ? x : y) + (size_t)b * nq * 3;
  const float* __restrict__ cp = (dir == 0 ? y : x) + (size_t)b * nc * 3;
  float* __restrict__ omin = (dir == 0 ? minx : miny) + (size_t)b * nq;
  int* __restrict__ oidx = (dir == 0 ? idxx : idxy) + (size_t)b * nq;

  __shared__ float4 sc[NN_CHUNK];
  // fminf drops NaNs, so a NaN coordinate would come out as a huge finite (or an arbitrary neighbour's) distance; the
  // reference's torch.min over the distance matrix gives NaN for the NaN query itself and - a NaN candidate poisons
  // every row - for all queries of the sample.  Detected at load time (outside the inner loop) and applied at the end.
  __shared__ int s_nan;
  if (threadIdx.x == 0) s_nan = 0;

  float qx[Q], qy[Q], qz[Q], best[Q];
  int bi[Q];
  bool nan_seen[Q];
#pragma unroll
  for (int k = 0; k < Q; ++k) {
    int j = q0 + k * NN_THREADS + threadIdx.x;
    bool ok = j < nq;
    qx[k] = ok ? __ldg(qp + 3 * j + 0) : 0.f;
    qy[k] = ok ? __ldg(qp + 3 * j + 1) : 0.f;
    qz[k] = ok ? __ldg(qp + 3 * j + 2) : 0.f;
    best[k] = 3.0e38f;
    bi[k] = 0;
    nan_seen[k] = (qx[k] != qx[k]) | (qy[k] != qy[k]) | (qz[k] != qz[k]);
  }

  for (int c0 = 0; c0 < nc; c0 += NN_CHUNK) {
    const int n = min(NN_CHUNK, nc - c0);
    const int n4 = (n + 3) & ~3;
    __syncthreads();
    const float* __restrict__ src = cp + (size_t)c0 * 3;
    for (int e = threadIdx.x; e < 3 * n; e += NN_THREADS) {
      float v = __ldg(src + e);
      int p = e / 3;
      reinterpret_cast<float*>(sc)[p * 4 + (e - 3 * p)] = v;
      if (v != v) s_nan = 1;
    }
    for (int p = n + threadIdx.x; p < n4; p += NN_THREADS)
      sc[p] = make_float4(1.0e18f, 1.0e18f, 1.0e18f, 0.f);  // never the minimum
    __syncthreads();

#pragma unroll 2
    for (int i = 0; i < n4; i += 4) {
      const float4 p0 = sc[i], p1 = sc[i + 1], p2 = sc[i + 2], p3 = sc[i + 3];
#pragma unroll
      for (int k = 0; k < Q; ++k) {
        float ax, ay, az;
        ax = qx[k] - p0.x; ay = qy[k] - p0.y; az = qz[k] - p0.z;
        const float d0 = fmaf(az, az, fmaf(ay, ay, ax * ax));
        ax = qx[k] - p1.x; ay = qy[k] - p1.y; az = qz[k] - p1.z;
        const float d1 = fmaf(az, az, fmaf(ay, ay, ax * ax));
        ax = qx[k] - p2.x; ay = qy[k] - p2.y; az = qz[k] - p2.z;
        const float d2 = fmaf(az, az, fmaf(ay, ay, ax * ax));
        ax = qx[k] - p3.x; ay = qy[k] - p3.y; az = qz[k] - p3.z;
        const float d3 = fmaf(az, az, fmaf(ay, ay, ax * ax));
        const float m = fminf(fminf(d0, d1), fminf(d2, d3));
        if (m < best[k]) {
          best[k] = m;
          bi[k] = c0 + i + (d0 == m ? 0 : (d1 == m ? 1 : (d2 == m ? 2 : 3)));
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < Q; ++k) {
    int j = q0 + k * NN_THREADS + threadIdx.x;
    if (j < nq) {
      omin[j] = (nan_seen[k] || s_nan) ? __int_as_float(0x7fc00000) : best[k];
      oidx[j] = bi[k];
    }
  }
}

// ---- packed-math variant ----------------------------------------------------------------------------------------
// Same search, same arithmetic per pair ((q - p) per coordinate, fma chain z, y, x: bit-identical distances), but two
// candidates per instruction with the sm_100 packed fp32 forms (sub / mul / fma .f32x2 -> FADD2 / FMUL2 / FFMA2): the
// scalar kernel is bound by instruction ISSUE (6 FMA-pipe + ~1.3 other instructions per pair at 57-61 % of the issue
// peak), the packed forms halve the FMA-pipe instruction count.  Candidates sit in shared memory as structure of arrays
// in groups of four ({x0..x3}, {y0..y3}, {z0..z3}: three broadcast LDS.128 per four candidates instead of four).
__device__ __forceinline__ unsigned long long pk2(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(unsigned long long v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

template <int Q>
__global__ void __launch_bounds__(NN_THREADS)
nn_packed_kernel(const float* __restrict__ x, const float* __restrict__ y, int N, int M,
                 float* __restrict__ minx, int* __restrict__ idxx,
                 float* __restrict__ miny, int* __restrict__ idxy, int dirs) {
  const int dir = (dirs == 3) ? (int)blockIdx.z : (dirs == 2 ? 1 : 0);
  const int nq = dir == 0 ? N : M;
  const int nc = dir == 0 ? M : N;
  const int q0 = blockIdx.x * (NN_THREADS * Q);
  if (q0 >= nq) return;
  const int b = blockIdx.y;
  const float* __restrict__ qp = (dir == 0 ? x : y) + (size_t)b * nq * 3;
  const float* __restrict__ cp = (dir == 0 ? y : x) + (size_t)b * nc * 3;
  float* __restrict__ omin = (dir == 0 ? minx : miny) + (size_t)b * nq;
  int* __restrict__ oidx = (dir == 0 ? idxx : idxy) + (size_t)b * nq;

  __shared__ float4 sc[3][NN_CHUNK / 4];   // [coordinate][group of four candidates]
  __shared__ int s_nan;                    // see nn_kernel
  if (threadIdx.x == 0) s_nan = 0;

  unsigned long long qx[Q], qy[Q], qz[Q];  // the query coordinate in both halves
  float best[Q];
  int bi[Q];
  bool nan_seen[Q];
#pragma unroll
  for (int k = 0; k < Q; ++k) {
    int j = q0 + k * NN_THREADS + threadIdx.x;
    bool ok = j < nq;
    const float fx = ok ? __ldg(qp + 3 * j + 0) : 0.f;
    const float fy = ok ? __ldg(qp + 3 * j + 1) : 0.f;
    const float fz = ok ? __ldg(qp + 3 * j + 2) : 0.f;
    qx[k] = pk2(fx, fx); qy[k] = pk2(fy, fy); qz[k] = pk2(fz, fz);
    best[k] = 3.0e38f;
    bi[k] = 0;
    nan_seen[k] = (fx != fx) | (fy != fy) | (fz != fz);
  }

  for (int c0 = 0; c0 < nc; c0 += NN_CHUNK) {
    const int n = min(NN_CHUNK, nc - c0);
    const int n4 = (n + 3) & ~3;
    __syncthreads();
    const float* __restrict__ src = cp + (size_t)c0 * 3;
    for (int e = threadIdx.x; e < 3 * n; e += NN_THREADS) {
      float v = __ldg(src + e);
      int p = e / 3;
      reinterpret_cast<float*>(sc[e - 3 * p])[p] = v;
      if (v != v) s_nan = 1;
    }
    for (int p = n + threadIdx.x; p < n4; p += NN_THREADS) {
#pragma unroll
      for (int c = 0; c < 3; ++c) reinterpret_cast<float*>(sc[c])[p] = 1.0e18f;   // never the minimum
    }
    __syncthreads();

#pragma unroll 2
    for (int i = 0; i < n4 / 4; ++i) {
      const float4 X = sc[0][i], Y = sc[1][i], Z = sc[2][i];
      const unsigned long long x01 = pk2(X.x, X.y), x23 = pk2(X.z, X.w);
      const unsigned long long y01 = pk2(Y.x, Y.y), y23 = pk2(Y.z, Y.w);
      const unsigned long long z01 = pk2(Z.x, Z.y), z23 = pk2(Z.z, Z.w);
#pragma unroll
      for (int k = 0; k < Q; ++k) {
        const unsigned long long ax01 = sub2(qx[k], x01), ay01 = sub2(qy[k], y01), az01 = sub2(qz[k], z01);
        const unsigned long long ax23 = sub2(qx[k], x23), ay23 = sub2(qy[k], y23), az23 = sub2(qz[k], z23);
        const unsigned long long d01 = fma2(az01, az01, fma2(ay01, ay01, mul2(ax01, ax01)));
        const unsigned long long d23 = fma2(az23, az23, fma2(ay23, ay23, mul2(ax23, ax23)));
        float d0, d1, d2, d3;
        upk2(d01, d0, d1);
        upk2(d23, d2, d3);
        const float m = fminf(fminf(d0, d1), fminf(d2, d3));
        // only the GROUP of the running minimum is tracked here (one compare + two selects); which of its four
        // candidates it was is resolved once, after the search
        const bool better = m < best[k];
        best[k] = better ? m : best[k];
        bi[k] = better ? c0 + 4 * i : bi[k];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < Q; ++k) {
    int j = q0 + k * NN_THREADS + threadIdx.x;
    if (j < nq) {
      // first candidate of the winning group whose distance (same operations, hence the same bits) is the minimum:
      // ties keep the lowest candidate index, as a strict "<" over all candidates in order would
      float fx, fy, fz, unused;
      upk2(qx[k], fx, unused); upk2(qy[k], fy, unused); upk2(qz[k], fz, unused);
      int idx = bi[k];
#pragma unroll
      for (int e = 3; e >= 0; --e) {
        const int c = bi[k] + e;
        if (c < nc) {
          const float ax = fx - __ldg(cp + 3 * c), ay = fy - __ldg(cp + 3 * c + 1), az = fz - __ldg(cp + 3 * c + 2);
          if (fmaf(az, az, fmaf(ay, ay, ax * ax)) == best[k]) idx = c;
        }
      }
      omin[j] = (nan_seen[k] || s_nan) ? __int_as_float(0x7fc00000) : best[k];
      oidx[j] = idx;
    }
  }
}

// loss[b] = mean(mins[b, :]) ; grid (B, 2), deterministic block reduction.
__global__ void __launch_bounds__(256)
chamfer_mean_kernel(const float* __restrict__ minx, const float* __restrict__ miny, int N, int M,
                    float* __restrict__ loss1, float* __restrict__ loss2) {
  __shared__ float scratch[32];
  const int b = blockIdx.x;
  const int dir = blockIdx.y;
  const int n = dir == 0 ? N : M;
  const float* __restrict__ v = (dir == 0 ? minx : miny) + (size_t)b * n;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) (dir == 0 ? loss1 : loss2)[b] = s / (float)n;
}

// d loss / d points for loss = sum_b gl1[b]*mean_j minx[b,j] + gl2[b]*mean_i miny[b,i], for ONE of the two clouds
// ("a", A points; the other cloud "o" has O points):
//   ga[t] = 2 s_a (a_t - o_{idx_a[t]})                                   direct term: own nearest neighbour
//         + sum over { i : idx_o[i] == t } of 2 s_o (a_t - o_i)          scatter term: points of o whose nearest is a_t
// with s_a = g_a[b] / A, s_o = g_o[b] / O.  The scatter term is turned into a gather: one CTA per sample builds the
// inverse of idx_o in shared memory (counting sort with integer atomics: counts -> exclusive scan -> bucket lists), every
// thread then owns output points and sums its bucket in ascending index order.  No float atomics, no zero-fill of the
// output, every output element written exactly once, bit-reproducible.  HBM bytes: 12 A + 12 O + 4 (A + O) read, 12 A
// written.  Shared memory: (2 A + 2 + O) ints.
constexpr int CH_BWD_THREADS = 512;

// index arrays in shared memory: 32-bit, or 16-bit (clouds below 65 536 points: half the shared memory, up to three
// CTAs per SM at 10^4 points - the kernel is latency-bound on its random gathers, so residency is what counts)
template <typename IT> struct IdxOps;
template <> struct IdxOps<int> {
  static __device__ __forceinline__ int add(int* arr, int j) { return atomicAdd(&arr[j], 1); }
};
template <> struct IdxOps<unsigned short> {
  // 16-bit atomic increment on the containing 32-bit word (counts stay below 65 536: no carry into the neighbour)
  static __device__ __forceinline__ int add(unsigned short* arr, int j) {
    unsigned* word = reinterpret_cast<unsigned*>(arr) + (j >> 1);
    const unsigned old = atomicAdd(word, (j & 1) ? 0x10000u : 1u);
    return (int)((j & 1) ? (old >> 16) : (old & 0xffffu));
  }
};

template <typename IT>
__global__ void __launch_bounds__(CH_BWD_THREADS)
chamfer_bwd_gather_kernel(const float* __restrict__ a, const float* __restrict__ o,
                          const int* __restrict__ idx_a, const int* __restrict__ idx_o,
                          const float* __restrict__ g_a, const float* __restrict__ g_o, int g_stride,
                          int A, int O, float* __restrict__ ga) {
  extern __shared__ int sh_raw[];
  const int A2 = (A + 2) & ~1;                       // even lengths keep every array 4-byte aligned
  IT* offs = reinterpret_cast<IT*>(sh_raw);          // [A + 1] bucket offsets (exclusive scan of the counts)
  IT* cursor = offs + A2;                            // [A]     counts, then fill cursors
  IT* list = cursor + A2;                            // [O]     indices of the other cloud, grouped by the point of `a` they map to
  __shared__ int warp_tot[CH_BWD_THREADS / 32];
  __shared__ int carry;
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const float* __restrict__ ab = a + (size_t)b * A * 3;
  const float* __restrict__ ob = o + (size_t)b * O * 3;
  const int* __restrict__ ia = idx_a + (size_t)b * A;
  const int* __restrict__ io = idx_o + (size_t)b * O;
  for (int t = tid; t < A2; t += CH_BWD_THREADS) cursor[t] = 0;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int i = tid; i < O; i += CH_BWD_THREADS) IdxOps<IT>::add(cursor, io[i]);
  __syncthreads();
  // exclusive scan of the counts, CH_BWD_THREADS entries per round
  for (int base = 0; base < A; base += CH_BWD_THREADS) {
    const int t = base + tid;
    const int c = t < A ? (int)cursor[t] : 0;
    int incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if ((tid & 31) >= d) incl += v;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
    __syncthreads();
    int before = carry;
    for (int w = 0; w < (tid >> 5); ++w) before += warp_tot[w];
    if (t < A) offs[t] = (IT)(before + incl - c);
    __syncthreads();
    if (tid == CH_BWD_THREADS - 1) carry = before + incl;
    __syncthreads();
  }
  if (tid == 0) offs[A] = (IT)carry;
  for (int t = tid; t < A2; t += CH_BWD_THREADS) cursor[t] = 0;
  __syncthreads();
  for (int i = tid; i < O; i += CH_BWD_THREADS) {
    const int j = io[i];
    list[(int)offs[j] + IdxOps<IT>::add(cursor, j)] = (IT)i;
  }
  __syncthreads();
  const float sa = 2.f * g_a[(size_t)b * g_stride] / (float)A;
  const float so = 2.f * g_o[(size_t)b * g_stride] / (float)O;
  float* __restrict__ gab = ga + (size_t)b * A * 3;
  // four output points per thread and iteration: their (dependent) index -> coordinate gathers are issued together, the
  // kernel is bound by the latency of those random reads, not by bandwidth
  constexpr int U = 4;
  for (int t0 = tid; t0 < A; t0 += U * CH_BWD_THREADS) {
    int n1[U], lo[U], hi[U];
    float px[U], py[U], pz[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t0 + u * CH_BWD_THREADS;
      const bool in = t < A;
      const int tc = in ? t : 0;
      n1[u] = ia[tc];
      px[u] = ab[3 * tc]; py[u] = ab[3 * tc + 1]; pz[u] = ab[3 * tc + 2];
      lo[u] = in ? (int)offs[tc] : 0;
      hi[u] = in ? (int)offs[tc + 1] : 0;
    }
    float gx[U], gy[U], gz[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      gx[u] = sa * (px[u] - ob[3 * n1[u]]);
      gy[u] = sa * (py[u] - ob[3 * n1[u] + 1]);
      gz[u] = sa * (pz[u] - ob[3 * n1[u] + 2]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      // ascending order of the bucket's indices whatever order the fill left them in (buckets hold ~O/A entries)
      int prev = -1;
      for (int e = lo[u]; e < hi[u]; ++e) {
        int best = 0x7fffffff;
        for (int f = lo[u]; f < hi[u]; ++f) {
          const int v = (int)list[f];
          if (v > prev && v < best) best = v;
        }
        prev = best;
        gx[u] = fmaf(so, px[u] - ob[3 * best], gx[u]);
        gy[u] = fmaf(so, py[u] - ob[3 * best + 1], gy[u]);
        gz[u] = fmaf(so, pz[u] - ob[3 * best + 2], gz[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t0 + u * CH_BWD_THREADS;
      if (t < A) { gab[3 * t] = gx[u]; gab[3 * t + 1] = gy[u]; gab[3 * t + 2] = gz[u]; }
    }
  }
}

// Fallback for clouds whose inverse index does not fit in shared memory (> ~18 k points): float atomics.
// gx / gy must be zero-initialised; gy may be null (targets without grad).
__global__ void __launch_bounds__(256)
chamfer_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                   const int* __restrict__ idxx, const int* __restrict__ idxy,
                   const float* __restrict__ gl1, const float* __restrict__ gl2, int g_stride, int N, int M,
                   float* __restrict__ gx, float* __restrict__ gy) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const float* xb = x + (size_t)b * N * 3;
  const float* yb = y + (size_t)b * M * 3;
  float* gxb = gx + (size_t)b * N * 3;
  float* gyb = gy ? gy + (size_t)b * M * 3 : nullptr;
  if (t < N) {
    const float s = 2.f * gl1[(size_t)b * g_stride] / (float)N;
    const int i = idxx[(size_t)b * N + t];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float g = s * (xb[3 * t + c] - yb[3 * i + c]);
      atomicAdd(gxb + 3 * t + c, g);
      if (gyb) atomicAdd(gyb + 3 * i + c, -g);
    }
  }
  if (t < M) {
    const float s = 2.f * gl2[(size_t)b * g_stride] / (float)M;
    const int j = idxy[(size_t)b * M + t];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float g = s * (yb[3 * t + c] - xb[3 * j + c]);
      atomicAdd(gxb + 3 * j + c, -g);
      if (gyb) atomicAdd(gyb + 3 * t + c, g);
    }
  }
}

int launch_nn(const float* x, const float* y, int B, int N, int M, float* minx, int* idxx,
              float* miny, int* idxy, int dirs, cudaStream_t st) {
  const int nq_max = dirs == 3 ? max(N, M) : (dirs == 1 ? N : M);
  const int ndir = dirs == 3 ? 2 : 1;
  // queries per thread: more register tiling for big clouds, finer tiles for small ones
  int Q = 4;
  if (nq_max <= 1024 || (long long)B * nq_max < 128LL * 4 * 2 * num_sms()) Q = 2;
  if (nq_max <= 320) Q = 1;
  dim3 grid((nq_max + NN_THREADS * Q - 1) / (NN_THREADS * Q), B, ndir);
  static int packed = -1;
  if (packed < 0) {
    const char* e = getenv("OBMAN_NN_PACKED");   // 0: the scalar kernel
    packed = (e && e[0] == '0') ? 0 : 1;
  }
  if (packed) {
    if (Q == 4)
      nn_packed_kernel<4><<<grid, NN_THREADS, 0, st>>>(x, y, N, M, minx, idxx, miny, idxy, dirs);
    else if (Q == 2)
      nn_packed_kernel<2><<<grid, NN_THREADS, 0, st>>>(x, y, N, M, minx, idxx, miny, idxy, dirs);
    else
      nn_packed_kernel<1><<<grid, NN_THREADS, 0, st>>>(x, y, N, M, minx, idxx, miny, idxy, dirs);
    return check_launch("nn_packed_kernel");
  }
  if (Q == 4)
    nn_kernel<4><<<grid, NN_THREADS, 0, st>>>(x, y, N, M, minx, idxx, miny, idxy, dirs);
  else if (Q == 2)
    nn_kernel<2><<<grid, NN_THREADS, 0, st>>>(x, y, N, M, minx, idxx, miny, idxy, dirs);
  else
    nn_kernel<1><<<grid, NN_THREADS, 0, st>>>(x, y, N, M, minx, idxx, miny, idxy, dirs);
  return check_launch("nn_kernel");
}

}  // namespace obman

using namespace obman;

extern "C" int obman_nn_fwd(const float* x, const float* y, int B, int N, int M, float* minx,
                            int* idxx, float* miny, int* idxy, int dirs, void* stream) {
  OBMAN_REQUIRE(B > 0 && N > 0 && M > 0, "obman_nn_fwd: empty input (B=%d N=%d M=%d)", B, N, M);
  OBMAN_REQUIRE(B <= 65535, "obman_nn_fwd: B=%d exceeds grid.y", B);
  OBMAN_REQUIRE(dirs >= 1 && dirs <= 3, "obman_nn_fwd: dirs must be 1, 2 or 3");
  OBMAN_REQUIRE(x && y, "obman_nn_fwd: null input");
  OBMAN_REQUIRE(!(dirs & 1) || (minx && idxx), "obman_nn_fwd: null x-side output");
  OBMAN_REQUIRE(!(dirs & 2) || (miny && idxy), "obman_nn_fwd: null y-side output");
  return launch_nn(x, y, B, N, M, minx, idxx, miny, idxy, dirs, (cudaStream_t)stream);
}

extern "C" int obman_chamfer_fwd(const float* preds, const float* gts, int B, int N, int M,
                                 float* loss1, float* loss2, float* min1, int* idx1, float* min2,
                                 int* idx2, void* stream) {
  OBMAN_REQUIRE(loss1 && loss2 && min1 && idx1 && min2 && idx2, "obman_chamfer_fwd: null output");
  int rc = obman_nn_fwd(preds, gts, B, N, M, min1, idx1, min2, idx2, 3, stream);
  if (rc) return rc;
  chamfer_mean_kernel<<<dim3(B, 2), 256, 0, (cudaStream_t)stream>>>(min1, min2, N, M, loss1, loss2);
  return check_launch("chamfer_mean_kernel");
}

static size_t chamfer_bwd_smem(int A, int O, bool narrow) {
  const size_t A2 = (size_t)((A + 2) & ~1);
  return (narrow ? sizeof(unsigned short) : sizeof(int)) * (2 * A2 + (size_t)((O + 1) & ~1)) + 16;
}

static int chamfer_bwd_side(const float* a, const float* o, const int* idx_a, const int* idx_o, const float* g_a,
                            const float* g_o, int g_stride, int B, int A, int O, float* ga, cudaStream_t st) {
  const bool narrow = A < 65535 && O < 65535;
  const size_t smem = chamfer_bwd_smem(A, O, narrow);
  static size_t configured[2] = {0, 0};
  if (smem > configured[narrow]) {
    const int want = (int)(smem > 48 * 1024 ? smem : 48 * 1024);
    cudaError_t e = narrow ? cudaFuncSetAttribute(chamfer_bwd_gather_kernel<unsigned short>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, want)
                           : cudaFuncSetAttribute(chamfer_bwd_gather_kernel<int>,
                                                  cudaFuncAttributeMaxDynamicSharedMemorySize, want);
    if (e != cudaSuccess) {
      set_error("chamfer_bwd: cudaFuncSetAttribute(%zu bytes) failed: %s", smem, cudaGetErrorString(e));
      return OBMAN_ERR_CUDA;
    }
    configured[narrow] = (size_t)want;
  }
  if (narrow)
    chamfer_bwd_gather_kernel<unsigned short><<<B, CH_BWD_THREADS, smem, st>>>(a, o, idx_a, idx_o, g_a, g_o, g_stride, A, O, ga);
  else
    chamfer_bwd_gather_kernel<int><<<B, CH_BWD_THREADS, smem, st>>>(a, o, idx_a, idx_o, g_a, g_o, g_stride, A, O, ga);
  return check_launch("chamfer_bwd_gather_kernel");
}

extern "C" int obman_chamfer_bwd(const float* preds, const float* gts, const int* idx1,
                                 const int* idx2, const float* gloss1, const float* gloss2, int g_stride, int B,
                                 int N, int M, float* gpreds, float* ggts, void* stream) {
  OBMAN_REQUIRE(B > 0 && N > 0 && M > 0 && B <= 65535, "obman_chamfer_bwd: bad sizes");
  OBMAN_REQUIRE(preds && gts && idx1 && idx2 && gloss1 && gloss2 && gpreds,
                "obman_chamfer_bwd: null argument");
  OBMAN_REQUIRE(g_stride == 0 || g_stride == 1, "obman_chamfer_bwd: g_stride must be 0 (one scalar) or 1 (per sample)");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem_max = 200 * 1024;
  const bool narrow = N < 65535 && M < 65535;
  if (chamfer_bwd_smem(N, M, narrow) <= smem_max && chamfer_bwd_smem(M, N, narrow) <= smem_max) {
    int rc = chamfer_bwd_side(preds, gts, idx1, idx2, gloss1, gloss2, g_stride, B, N, M, gpreds, st);
    if (rc || !ggts) return rc;
    return chamfer_bwd_side(gts, preds, idx2, idx1, gloss2, gloss1, g_stride, B, M, N, ggts, st);
  }
  cudaMemsetAsync(gpreds, 0, sizeof(float) * (size_t)B * N * 3, st);
  if (ggts) cudaMemsetAsync(ggts, 0, sizeof(float) * (size_t)B * M * 3, st);
  int n = max(N, M);
  chamfer_bwd_kernel<<<dim3((n + 255) / 256, B), 256, 0, st>>>(preds, gts, idx1, idx2, gloss1,
                                                               gloss2, g_stride, N, M, gpreds, ggts);
  return check_launch("chamfer_bwd_kernel");
}
