// Contact (attraction / repulsion) loss kernels.
//
// Replaces, with three fused kernels and no (B, 778*F, 3) temporaries:
//   batch_mesh_contains_points   mano_train/networks/branches/contactutils.py:62-159   (ray parity)
//   compute_contact_loss         mano_train/networks/branches/contactloss.py:149-308   (values, masks)
//   masked_mean_loss             mano_train/networks/branches/contactloss.py:50-57     (batch-global means)
// The hand->object nearest-vertex search (contactloss.py:164-166) is the shared kernel in nn_pairs.cu.
#include <stdlib.h>

#include "common.cuh"

namespace obman {

// ---- ray / triangle parity --------------------------------------------------------------------
constexpr int RC_THREADS = 256;  // upper bound; the launch picks the block size that wastes the fewest lanes (raycast_block)
constexpr int RC_CHUNK = 512;  // triangles per smem stage: 4 float4 each = 32 KB

// Fixed ray direction and tolerances of the reference (contactutils.py:65,78,104).
#define RC_DX 0.4395064455f
#define RC_DY 0.617598629942f
#define RC_DZ 0.652231566745f
#define RC_TOL 0.0000001f
#define RC_DET_EPS 0.00000001f

__global__ void __launch_bounds__(RC_THREADS)
raycast_kernel(const float* __restrict__ pts, const float* __restrict__ obj,
               const int* __restrict__ faces, int P, int N, int F, int f_per_split,
               int* __restrict__ hits) {
  __shared__ float4 sA[RC_CHUNK];  // v0.xyz, invdet
  __shared__ float4 sB[RC_CHUNK];  // e1.xyz, parallel flag
  __shared__ float4 sC[RC_CHUNK];  // e2.xyz
  __shared__ float4 sD[RC_CHUNK];  // pvec.xyz
  const int b = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int f_begin = blockIdx.z * f_per_split;
  const int f_end = min(F, f_begin + f_per_split);
  const float* __restrict__ ob = obj + (size_t)b * N * 3;
  float px = 0.f, py = 0.f, pz = 0.f;
  if (p < P) {
    const float* q = pts + ((size_t)b * P + p) * 3;
    px = q[0]; py = q[1]; pz = q[2];
  }
  int count = 0;
  for (int f0 = f_begin; f0 < f_end; f0 += RC_CHUNK) {
    const int n = min(RC_CHUNK, f_end - f0);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int* fc = faces + (size_t)(f0 + i) * 3;
      const float* a = ob + 3 * fc[0];
      const float* bb = ob + 3 * fc[1];
      const float* c = ob + 3 * fc[2];
      const float v0x = a[0], v0y = a[1], v0z = a[2];
      const float e1x = bb[0] - v0x, e1y = bb[1] - v0y, e1z = bb[2] - v0z;
      const float e2x = c[0] - v0x, e2y = c[1] - v0y, e2z = c[2] - v0z;
      // pvec = dir x e2
      const float pvx = RC_DY * e2z - RC_DZ * e2y;
      const float pvy = RC_DZ * e2x - RC_DX * e2z;
      const float pvz = RC_DX * e2y - RC_DY * e2x;
      const float det = e1x * pvx + e1y * pvy + e1z * pvz;
      const float invdet = 1.0f / (det + RC_DET_EPS);
      sA[i] = make_float4(v0x, v0y, v0z, invdet);
      sB[i] = make_float4(e1x, e1y, e1z, fabsf(det) < RC_TOL ? 1.f : 0.f);
      sC[i] = make_float4(e2x, e2y, e2z, 0.f);
      sD[i] = make_float4(pvx, pvy, pvz, 0.f);
    }
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < n; ++i) {
      const float4 A = sA[i], Bv = sB[i], C = sC[i], D = sD[i];
      const float tx = px - A.x, ty = py - A.y, tz = pz - A.z;
      const float u = (tx * D.x + ty * D.y + tz * D.z) * A.w;
      const float qx = ty * Bv.z - tz * Bv.y;
      const float qy = tz * Bv.x - tx * Bv.z;
      const float qz = tx * Bv.y - ty * Bv.x;
      const float v = (RC_DX * qx + RC_DY * qy + RC_DZ * qz) * A.w;
      const float t = (C.x * qx + C.y * qy + C.z * qz) * A.w;
      const bool hit = (u > 0.f) & (u < 1.f) & (v > 0.f) & (u + v < 1.f) & (t >= RC_TOL) &
                       (Bv.w == 0.f);
      count += hit ? 1 : 0;
    }
  }
  if (p < P && count) atomicAdd(hits + (size_t)b * P + p, count);
}

// Packed-math variant: two triangles per instruction with the sm_100 packed fp32 forms (FADD2 / FMUL2 / FFMA2), as in
// nn_pairs.cu's nn_packed_kernel - the scalar kernel is bound by instruction issue (~30 instructions per ray / triangle
// test).  Per pair of triangles the staged record is 8 float4 = {a, b} interleaved per component:
//   v0.xyz, invdet | pvec.xyz | e1.xyz | -e1.xyz | e2.xyz
// A triangle parallel to the ray is staged with invdet = 0 (u = 0 fails "u > 0"), an odd triangle out likewise, so the
// inner loop has no flag test; "u < 1" is implied by v > 0 and u + v < 1 (rounding is monotonic) and "u + v < 1" is
// evaluated as 1 - (u + v) > 0 (exactly equivalent), which turns four of the comparisons into one three-way minimum.
__device__ __forceinline__ unsigned long long rc_pk(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void rc_upk(unsigned long long v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long rc_add(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long rc_sub(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long rc_mul(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long rc_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// PT = points per thread (p, p + blockDim.x, ...): the eight broadcast LDS.128 of a triangle pair - the shared-memory
// pipe, not instruction issue, bounds this kernel once the arithmetic is packed - are amortised over PT points.
constexpr int RC_PACKED_THREADS = 448;
template <int PT>
__global__ void __launch_bounds__(RC_PACKED_THREADS)
raycast_packed_kernel(const float* __restrict__ pts, const float* __restrict__ obj,
                      const int* __restrict__ faces, int P, int N, int F, int f_per_split,
                      int* __restrict__ hits) {
  __shared__ float4 sp[RC_CHUNK / 2][8];   // 32 KB
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * (blockDim.x * PT) + threadIdx.x;
  const int f_begin = blockIdx.z * f_per_split;
  const int f_end = min(F, f_begin + f_per_split);
  const float* __restrict__ ob = obj + (size_t)b * N * 3;
  unsigned long long px[PT], py[PT], pz[PT];
  int count[PT];
#pragma unroll
  for (int k = 0; k < PT; ++k) {
    const int p = p0 + k * blockDim.x;
    float fx = 0.f, fy = 0.f, fz = 0.f;
    if (p < P) {
      const float* q = pts + ((size_t)b * P + p) * 3;
      fx = q[0]; fy = q[1]; fz = q[2];
    }
    px[k] = rc_pk(fx, fx); py[k] = rc_pk(fy, fy); pz[k] = rc_pk(fz, fz);
    count[k] = 0;
  }
  const unsigned long long dirx = rc_pk(RC_DX, RC_DX), diry = rc_pk(RC_DY, RC_DY), dirz = rc_pk(RC_DZ, RC_DZ);
  const unsigned long long one = rc_pk(1.f, 1.f);
  for (int f0 = f_begin; f0 < f_end; f0 += RC_CHUNK) {
    const int n = min(RC_CHUNK, f_end - f0);
    const int n2 = (n + 1) & ~1;
    __syncthreads();
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
      float rec[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) rec[k] = 0.f;   // the odd one out: invdet = 0 never hits
      if (i < n) {
        const int* fc = faces + (size_t)(f0 + i) * 3;
        const float* a = ob + 3 * fc[0];
        const float* bb = ob + 3 * fc[1];
        const float* c = ob + 3 * fc[2];
        const float v0x = a[0], v0y = a[1], v0z = a[2];
        const float e1x = bb[0] - v0x, e1y = bb[1] - v0y, e1z = bb[2] - v0z;
        const float e2x = c[0] - v0x, e2y = c[1] - v0y, e2z = c[2] - v0z;
        // pvec = dir x e2
        const float pvx = RC_DY * e2z - RC_DZ * e2y;
        const float pvy = RC_DZ * e2x - RC_DX * e2z;
        const float pvz = RC_DX * e2y - RC_DY * e2x;
        const float det = e1x * pvx + e1y * pvy + e1z * pvz;
        const float invdet = fabsf(det) < RC_TOL ? 0.f : 1.0f / (det + RC_DET_EPS);
        rec[0] = v0x; rec[1] = v0y; rec[2] = v0z; rec[3] = invdet;
        rec[4] = pvx; rec[5] = pvy; rec[6] = pvz;
        rec[7] = e1x; rec[8] = e1y; rec[9] = e1z;
        rec[10] = -e1x; rec[11] = -e1y; rec[12] = -e1z;
        rec[13] = e2x; rec[14] = e2y; rec[15] = e2z;
      }
      float* dst = reinterpret_cast<float*>(sp[i >> 1]) + (i & 1);
#pragma unroll
      for (int k = 0; k < 16; ++k) dst[2 * k] = rec[k];
    }
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < n2 / 2; ++j) {
      const float4 L0 = sp[j][0], L1 = sp[j][1], L2 = sp[j][2], L3 = sp[j][3];
      const float4 L4 = sp[j][4], L5 = sp[j][5], L6 = sp[j][6], L7 = sp[j][7];
      const unsigned long long v0x = rc_pk(L0.x, L0.y), v0y = rc_pk(L0.z, L0.w), v0z = rc_pk(L1.x, L1.y);
      const unsigned long long inv = rc_pk(L1.z, L1.w);
      const unsigned long long Dx = rc_pk(L2.x, L2.y), Dy = rc_pk(L2.z, L2.w), Dz = rc_pk(L3.x, L3.y);
      const unsigned long long e1x = rc_pk(L3.z, L3.w), e1y = rc_pk(L4.x, L4.y), e1z = rc_pk(L4.z, L4.w);
      const unsigned long long m1x = rc_pk(L5.x, L5.y), m1y = rc_pk(L5.z, L5.w), m1z = rc_pk(L6.x, L6.y);
      const unsigned long long e2x = rc_pk(L6.z, L6.w), e2y = rc_pk(L7.x, L7.y), e2z = rc_pk(L7.z, L7.w);
#pragma unroll
      for (int k = 0; k < PT; ++k) {
        const unsigned long long tx = rc_sub(px[k], v0x), ty = rc_sub(py[k], v0y), tz = rc_sub(pz[k], v0z);
        const unsigned long long u = rc_mul(rc_fma(tz, Dz, rc_fma(ty, Dy, rc_mul(tx, Dx))), inv);
        // q = t x e1
        const unsigned long long qx = rc_fma(tz, m1y, rc_mul(ty, e1z));
        const unsigned long long qy = rc_fma(tx, m1z, rc_mul(tz, e1x));
        const unsigned long long qz = rc_fma(ty, m1x, rc_mul(tx, e1y));
        const unsigned long long v = rc_mul(rc_fma(dirz, qz, rc_fma(diry, qy, rc_mul(dirx, qx))), inv);
        const unsigned long long t = rc_mul(rc_fma(e2z, qz, rc_fma(e2y, qy, rc_mul(e2x, qx))), inv);
        const unsigned long long w = rc_sub(one, rc_add(u, v));
        float ua, ub, va, vb, wa, wb, ta, tb;
        rc_upk(u, ua, ub); rc_upk(v, va, vb); rc_upk(w, wa, wb); rc_upk(t, ta, tb);
        count[k] += (fminf(fminf(ua, va), wa) > 0.f) & (ta >= RC_TOL);
        count[k] += (fminf(fminf(ub, vb), wb) > 0.f) & (tb >= RC_TOL);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < PT; ++k) {
    const int p = p0 + k * blockDim.x;
    if (p < P && count[k]) atomicAdd(hits + (size_t)b * P + p, count[k]);
  }
}

// Streamed variant: the triangle-pair records are built ONCE per sample by raycast_prep_kernel into a scratch buffer
// ((B, ceil(F/2), 8) float4, 84 MB at B = 256 / 5120 faces: L2 resident), the search kernel reads them with uniform
// (broadcast) global loads - no shared memory, no barriers, any number of warps per CTA.  Measured on the staged kernel
// above (ncu, profiles/ncu_raycast_r2.txt): a quarter of the warp samples sat at the barriers around the per-chunk
// record set-up (13 warps per CTA on 4 schedulers, 2 CTAs per SM), issue slots 50 % busy.
__global__ void __launch_bounds__(256)
raycast_prep_kernel(const float* __restrict__ obj, const int* __restrict__ faces, int N, int F, int F2,
                    float* __restrict__ rec) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * F2) return;
  const float* __restrict__ ob = obj + (size_t)b * N * 3;
  float r[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) r[k] = 0.f;   // the odd one out: invdet = 0 never hits
  if (i < F) {
    const int* fc = faces + (size_t)i * 3;
    const float* a = ob + 3 * fc[0];
    const float* bb = ob + 3 * fc[1];
    const float* c = ob + 3 * fc[2];
    const float v0x = a[0], v0y = a[1], v0z = a[2];
    const float e1x = bb[0] - v0x, e1y = bb[1] - v0y, e1z = bb[2] - v0z;
    const float e2x = c[0] - v0x, e2y = c[1] - v0y, e2z = c[2] - v0z;
    const float pvx = RC_DY * e2z - RC_DZ * e2y;
    const float pvy = RC_DZ * e2x - RC_DX * e2z;
    const float pvz = RC_DX * e2y - RC_DY * e2x;
    const float det = e1x * pvx + e1y * pvy + e1z * pvz;
    const float invdet = fabsf(det) < RC_TOL ? 0.f : 1.0f / (det + RC_DET_EPS);
    r[0] = v0x; r[1] = v0y; r[2] = v0z; r[3] = invdet;
    r[4] = pvx; r[5] = pvy; r[6] = pvz;
    r[7] = e1x; r[8] = e1y; r[9] = e1z;
    r[10] = -e1x; r[11] = -e1y; r[12] = -e1z;
    r[13] = e2x; r[14] = e2y; r[15] = e2z;
  }
  float* dst = rec + ((size_t)b * F2 + (i >> 1)) * 32 + (i & 1);
#pragma unroll
  for (int k = 0; k < 16; ++k) dst[2 * k] = r[k];
}

template <int PT>
__global__ void __launch_bounds__(256)
raycast_stream_kernel(const float* __restrict__ pts, const float4* __restrict__ rec, int P, int F2, int pairs_per_split,
                      int* __restrict__ hits) {
  const int b = blockIdx.y;
  const int p0 = blockIdx.x * (blockDim.x * PT) + threadIdx.x;
  const int j_begin = blockIdx.z * pairs_per_split;
  const int j_end = min(F2, j_begin + pairs_per_split);
  unsigned long long px[PT], py[PT], pz[PT];
  int count[PT];
#pragma unroll
  for (int k = 0; k < PT; ++k) {
    const int p = p0 + k * blockDim.x;
    float fx = 0.f, fy = 0.f, fz = 0.f;
    if (p < P) {
      const float* q = pts + ((size_t)b * P + p) * 3;
      fx = q[0]; fy = q[1]; fz = q[2];
    }
    px[k] = rc_pk(fx, fx); py[k] = rc_pk(fy, fy); pz[k] = rc_pk(fz, fz);
    count[k] = 0;
  }
  const unsigned long long dirx = rc_pk(RC_DX, RC_DX), diry = rc_pk(RC_DY, RC_DY), dirz = rc_pk(RC_DZ, RC_DZ);
  const unsigned long long one = rc_pk(1.f, 1.f);
  const float4* __restrict__ r = rec + ((size_t)b * F2 + j_begin) * 8;
#pragma unroll 2
  for (int j = j_begin; j < j_end; ++j, r += 8) {
    const float4 L0 = __ldg(r), L1 = __ldg(r + 1), L2 = __ldg(r + 2), L3 = __ldg(r + 3);
    const float4 L4 = __ldg(r + 4), L5 = __ldg(r + 5), L6 = __ldg(r + 6), L7 = __ldg(r + 7);
    const unsigned long long v0x = rc_pk(L0.x, L0.y), v0y = rc_pk(L0.z, L0.w), v0z = rc_pk(L1.x, L1.y);
    const unsigned long long inv = rc_pk(L1.z, L1.w);
    const unsigned long long Dx = rc_pk(L2.x, L2.y), Dy = rc_pk(L2.z, L2.w), Dz = rc_pk(L3.x, L3.y);
    const unsigned long long e1x = rc_pk(L3.z, L3.w), e1y = rc_pk(L4.x, L4.y), e1z = rc_pk(L4.z, L4.w);
    const unsigned long long m1x = rc_pk(L5.x, L5.y), m1y = rc_pk(L5.z, L5.w), m1z = rc_pk(L6.x, L6.y);
    const unsigned long long e2x = rc_pk(L6.z, L6.w), e2y = rc_pk(L7.x, L7.y), e2z = rc_pk(L7.z, L7.w);
#pragma unroll
    for (int k = 0; k < PT; ++k) {
      const unsigned long long tx = rc_sub(px[k], v0x), ty = rc_sub(py[k], v0y), tz = rc_sub(pz[k], v0z);
      const unsigned long long u = rc_mul(rc_fma(tz, Dz, rc_fma(ty, Dy, rc_mul(tx, Dx))), inv);
      const unsigned long long qx = rc_fma(tz, m1y, rc_mul(ty, e1z));
      const unsigned long long qy = rc_fma(tx, m1z, rc_mul(tz, e1x));
      const unsigned long long qz = rc_fma(ty, m1x, rc_mul(tx, e1y));
      const unsigned long long v = rc_mul(rc_fma(dirz, qz, rc_fma(diry, qy, rc_mul(dirx, qx))), inv);
      const unsigned long long t = rc_mul(rc_fma(e2z, qz, rc_fma(e2y, qy, rc_mul(e2x, qx))), inv);
      const unsigned long long w = rc_sub(one, rc_add(u, v));
      float ua, ub, va, vb, wa, wb, ta, tb;
      rc_upk(u, ua, ub); rc_upk(v, va, vb); rc_upk(w, wa, wb); rc_upk(t, ta, tb);
      count[k] += (fminf(fminf(ua, va), wa) > 0.f) & (ta >= RC_TOL);
      count[k] += (fminf(fminf(ub, vb), wb) > 0.f) & (tb >= RC_TOL);
    }
  }
#pragma unroll
  for (int k = 0; k < PT; ++k) {
    const int p = p0 + k * blockDim.x;
    if (p < P && count[k]) atomicAdd(hits + (size_t)b * P + p, count[k]);
  }
}

// ---- per-vertex values, masks, per-sample partial sums ------------------------------------------
enum { MODE_DIST_SQ = 0, MODE_DIST = 1, MODE_DIST_TANH = 2 };
enum { ZONES_ALL = 0, ZONES_TIPS = 1, ZONES_ZONES = 2 };
enum { TARGET_ALL = 0, TARGET_OBJ = 1, TARGET_HAND = 2 };

__device__ __forceinline__ float contact_value(int mode, float thresh, float sq, float anchor) {
  if (mode == MODE_DIST_SQ) return sq;
  if (mode == MODE_DIST) return anchor;
  return thresh * tanhf(anchor / thresh);
}
// d value / d diff = w * diff
__device__ __forceinline__ float contact_dweight(int mode, float thresh, float anchor) {
  if (mode == MODE_DIST_SQ) return 2.f;
  if (anchor == 0.f) return 0.f;  // torch.norm backward is masked to 0 at the origin
  if (mode == MODE_DIST) return 1.f / anchor;
  const float th = tanhf(anchor / thresh);
  return (1.f - th * th) / anchor;
}

constexpr int CV_THREADS = 256;

__global__ void __launch_bounds__(CV_THREADS)
contact_values_kernel(const float* __restrict__ hand, const float* __restrict__ obj,
                      const float* __restrict__ mins21, const int* __restrict__ idx21,
                      const int* __restrict__ hits, const int* __restrict__ zone_ids,
                      const int* __restrict__ zone_ptr, int n_zones, int zones_mode, int P, int N,
                      float contact_thresh, int contact_mode, float collision_thresh,
                      int collision_mode, unsigned char* __restrict__ attr_mask,
                      unsigned char* __restrict__ rep_mask, float* __restrict__ close,
                      float* __restrict__ anchor_out, float* __restrict__ partial) {
  extern __shared__ unsigned char s_match[];  // P flags
  __shared__ float scratch[32];
  const int b = blockIdx.x;
  const float* hb = hand + (size_t)b * P * 3;
  const float* ob = obj + (size_t)b * N * 3;
  const float* mb = mins21 + (size_t)b * P;
  const int* ib = idx21 + (size_t)b * P;
  const int* hc = hits + (size_t)b * P;

  for (int v = threadIdx.x; v < P; v += CV_THREADS) s_match[v] = zones_mode == ZONES_ALL ? 1 : 0;
  __syncthreads();
  if (zones_mode == ZONES_TIPS) {
    for (int i = threadIdx.x; i < zone_ptr[n_zones]; i += CV_THREADS) s_match[zone_ids[i]] = 1;
  } else if (zones_mode == ZONES_ZONES) {
    // per zone: the zone vertex with the smallest mins21 (first in list order on ties)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int z = warp; z < n_zones; z += CV_THREADS / 32) {
      const int beg = zone_ptr[z], end = zone_ptr[z + 1];
      float bv = 3.0e38f;
      int bp = 0x7fffffff;
      for (int i = beg + lane; i < end; i += 32) {
        float val = mb[zone_ids[i]];
        if (val < bv) { bv = val; bp = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int op = __shfl_xor_sync(0xffffffffu, bp, o);
        if (ov < bv || (ov == bv && op < bp)) { bv = ov; bp = op; }
      }
      if (lane == 0 && bp != 0x7fffffff) s_match[zone_ids[bp]] = 1;
    }
  }
  __syncthreads();

  float s_av = 0.f, s_ac = 0.f, s_rv = 0.f, s_rc = 0.f, s_pen = 0.f, m_pen = 0.f;
  for (int v = threadIdx.x; v < P; v += CV_THREADS) {
    const int j = ib[v];
    const float cx = ob[3 * j], cy = ob[3 * j + 1], cz = ob[3 * j + 2];
    const float dx = cx - hb[3 * v], dy = cy - hb[3 * v + 1], dz = cz - hb[3 * v + 2];
    const float sq = dx * dx + dy * dy + dz * dz;
    const float anchor = sqrtf(sq);
    const bool exterior = (hc[v] & 1) == 0;
    bool below = true;
    if (contact_mode == MODE_DIST_SQ) below = mb[v] < contact_thresh * contact_thresh;
    else if (contact_mode == MODE_DIST) below = mb[v] < contact_thresh;  // reference quirk kept
    const bool attr = below && exterior && s_match[v];
    const bool rep = !exterior;
    const size_t o = (size_t)b * P + v;
    attr_mask[o] = attr;
    rep_mask[o] = rep;
    close[3 * o] = cx; close[3 * o + 1] = cy; close[3 * o + 2] = cz;
    anchor_out[o] = anchor;
    if (attr) { s_av += contact_value(contact_mode, contact_thresh, sq, anchor); s_ac += 1.f; }
    if (rep) {
      s_rv += contact_value(collision_mode, collision_thresh, sq, anchor);
      s_rc += 1.f;
      s_pen += anchor;
      m_pen = fmaxf(m_pen, anchor);
    }
  }
  float r;
  r = block_sum(s_av, scratch); if (threadIdx.x == 0) partial[b * 6 + 0] = r;
  r = block_sum(s_ac, scratch); if (threadIdx.x == 0) partial[b * 6 + 1] = r;
  r = block_sum(s_rv, scratch); if (threadIdx.x == 0) partial[b * 6 + 2] = r;
  r = block_sum(s_rc, scratch); if (threadIdx.x == 0) partial[b * 6 + 3] = r;
  r = block_sum(s_pen, scratch); if (threadIdx.x == 0) partial[b * 6 + 4] = r / (float)P;
  // block max
  m_pen = warp_max(m_pen);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = m_pen;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.f;
    for (int w = 0; w < CV_THREADS / 32; ++w) m = fmaxf(m, scratch[w]);
    partial[b * 6 + 5] = m;
  }
}

// out[0]=missed_loss out[1]=penetr_loss out[2]=max_penetr out[3]=mean_penetr
// out[4]=#attraction verts out[5]=#repulsion verts    (batch-global, fixed summation order)
__global__ void __launch_bounds__(256)
contact_finalize_kernel(const float* __restrict__ partial, int B, float* __restrict__ out) {
  __shared__ float scratch[32];
  float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int b = threadIdx.x; b < B; b += blockDim.x)
#pragma unroll
    for (int k = 0; k < 6; ++k) acc[k] += partial[b * 6 + k];
  float tot[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) tot[k] = block_sum(acc[k], scratch);
  if (threadIdx.x == 0) {
    out[0] = tot[1] > 0.f ? tot[0] / tot[1] : 0.f;
    out[1] = tot[3] > 0.f ? tot[2] / tot[3] : 0.f;
    out[2] = tot[5] / (float)B;
    out[3] = tot[4] / (float)B;
    out[4] = tot[1];
    out[5] = tot[3];
  }
}

__global__ void __launch_bounds__(256)
contact_bwd_kernel(const float* __restrict__ hand, const float* __restrict__ close,
                   const float* __restrict__ anchor, const int* __restrict__ idx21,
                   const unsigned char* __restrict__ attr_mask,
                   const unsigned char* __restrict__ rep_mask, const float* __restrict__ fwd_out,
                   const float* __restrict__ g_missed, const float* __restrict__ g_penetr, int B,
                   int P, int N, float contact_thresh, int contact_mode, float collision_thresh,
                   int collision_mode, int target, float* __restrict__ ghand,
                   float* __restrict__ gobj) {
  const size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= (size_t)B * P) return;
  const int b = (int)(o / P);
  const float n_attr = fwd_out[4], n_rep = fwd_out[5];
  float w = 0.f;
  const float a = anchor[o];
  if (attr_mask[o] && n_attr > 0.f)
    w += g_missed[0] / n_attr * contact_dweight(contact_mode, contact_thresh, a);
  if (rep_mask[o] && n_rep > 0.f)
    w += g_penetr[0] / n_rep * contact_dweight(collision_mode, collision_thresh, a);
  float gd[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) gd[c] = w * (close[3 * o + c] - hand[3 * o + c]);
  if (ghand) {
#pragma unroll
    for (int c = 0; c < 3; ++c) ghand[3 * o + c] = target == TARGET_OBJ ? 0.f : -gd[c];
  }
  if (gobj && target != TARGET_HAND && w != 0.f) {
    float* g = gobj + ((size_t)b * N + idx21[o]) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) atomicAdd(g + c, gd[c]);
  }
}

// Contact-IoU metric (meshiou / thresh_ious, contactloss.py:20-47): for every sample and threshold the IoU of the two
// thresholded contact maps (dists <= thresh; 0 where the union is empty).  One CTA per sample, all thresholds at once.
constexpr int IOU_MAX_T = 16;
struct IouThresholds { float t[IOU_MAX_T]; int n; };

__global__ void __launch_bounds__(256)
contact_iou_kernel(const float* __restrict__ gt, const float* __restrict__ pred, int P, const IouThresholds th,
                   float* __restrict__ iou /* (B, T) */) {
  __shared__ int s_inter[IOU_MAX_T], s_union[IOU_MAX_T];
  const int b = blockIdx.x;
  if (threadIdx.x < IOU_MAX_T) { s_inter[threadIdx.x] = 0; s_union[threadIdx.x] = 0; }
  __syncthreads();
  int inter[IOU_MAX_T], uni[IOU_MAX_T];
#pragma unroll
  for (int k = 0; k < IOU_MAX_T; ++k) { inter[k] = 0; uni[k] = 0; }
  for (int i = threadIdx.x; i < P; i += blockDim.x) {
    const float g = gt[(size_t)b * P + i], p = pred[(size_t)b * P + i];
#pragma unroll
    for (int k = 0; k < IOU_MAX_T; ++k) {
      if (k < th.n) {
        const bool cg = g <= th.t[k], cp = p <= th.t[k];
        inter[k] += (cg && cp) ? 1 : 0;
        uni[k] += (cg || cp) ? 1 : 0;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < IOU_MAX_T; ++k) {
    if (k < th.n) {
      int a = inter[k], u = uni[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        u += __shfl_xor_sync(0xffffffffu, u, o);
      }
      if ((threadIdx.x & 31) == 0) { atomicAdd(&s_inter[k], a); atomicAdd(&s_union[k], u); }   // integer: order-free
    }
  }
  __syncthreads();
  if (threadIdx.x < th.n)
    iou[(size_t)b * th.n + threadIdx.x] =
        s_union[threadIdx.x] != 0 ? (float)s_inter[threadIdx.x] / (float)s_union[threadIdx.x] : 0.f;
}

// batch_ious[t] = mean_b iou[b, t]; auc = trapezoid of batch_ious over the thresholds (= the reference's mean over the
// batch of the per-sample trapezoids).  One warp per threshold, fixed summation order.
__global__ void __launch_bounds__(32 * IOU_MAX_T)
contact_iou_finalize_kernel(const float* __restrict__ iou, int B, const IouThresholds th,
                            float* __restrict__ batch_ious, float* __restrict__ auc) {
  __shared__ float means[IOU_MAX_T];
  const int t = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (t < th.n) {
    float s = 0.f;
    for (int b = lane; b < B; b += 32) s += iou[(size_t)b * th.n + t];
    s = warp_sum(s) / (float)B;
    if (lane == 0) { means[t] = s; batch_ious[t] = s; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int k = 0; k + 1 < th.n; ++k) a += 0.5f * (means[k] + means[k + 1]) * (th.t[k + 1] - th.t[k]);
    *auc = a;
  }
}

}  // namespace obman

using namespace obman;

extern "C" int obman_raycast_hits(const float* points, const float* obj_verts, const int* faces,
                                  int B, int P, int N, int F, int* hits, float* tri_scratch, void* stream) {
  OBMAN_REQUIRE(B > 0 && P > 0 && N > 0 && F > 0 && B <= 65535, "obman_raycast_hits: bad sizes");
  OBMAN_REQUIRE(points && obj_verts && faces && hits, "obman_raycast_hits: null argument");
  OBMAN_REQUIRE(((uintptr_t)tri_scratch & 15) == 0, "obman_raycast_hits: tri_scratch must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(hits, 0, sizeof(int) * (size_t)B * P, st);
  static int stream_on = -1;
  if (stream_on < 0) {
    const char* e = getenv("OBMAN_RAYCAST_STREAM");   // 0: stage the triangle records per CTA even when scratch is given
    stream_on = (e && e[0] == '0') ? 0 : 1;
  }
  if (tri_scratch && stream_on) {
    const int F2 = (F + 1) / 2;
    raycast_prep_kernel<<<dim3((2 * F2 + 255) / 256, B), 256, 0, st>>>(obj_verts, faces, N, F, F2, tri_scratch);
    int rc = check_launch("raycast_prep_kernel");
    if (rc) return rc;
    // two points per lane; warps per CTA (<= 8) chosen for the fewest idle lanes, pairs split until the grid is
    // several waves deep (short CTAs: no tail)
    const int w2 = (P + 63) / 64;
    int t2 = (w2 + 7) / 8, thr2 = 256, best = 1 << 30;
    for (int t = t2; t <= w2; ++t) {
      const int w = (w2 + t - 1) / t;
      if (w * t - w2 < best || (w * t - w2 == best && w > thr2 / 32)) { best = w * t - w2; t2 = t; thr2 = w * 32; }
      if (best == 0) break;
    }
    int sp = 1;
    while (sp < 16 && (long long)t2 * B * sp * (thr2 / 32) < 8LL * 32 * num_sms() && F2 / (sp + 1) >= 128) ++sp;
    const int per2 = (F2 + sp - 1) / sp;
    sp = (F2 + per2 - 1) / per2;
    raycast_stream_kernel<2><<<dim3(t2, B, sp), thr2, 0, st>>>(points, reinterpret_cast<const float4*>(tri_scratch), P,
                                                              F2, per2, hits);
    return check_launch("raycast_stream_kernel");
  }
  // point tiles: the block size (whole warps, <= RC_THREADS) that leaves the fewest idle lanes.  778 hand vertices =
  // 25 warps: 4 tiles of 256 threads would idle 24 % of the lanes, 5 tiles of 160 threads idle 3 %.
  const int warps = (P + 31) / 32;
  int ptiles = (warps + RC_THREADS / 32 - 1) / (RC_THREADS / 32), threads = RC_THREADS;
  {
    int best_waste = 1 << 30;
    for (int t = ptiles; t <= ptiles + 3 && t <= warps; ++t) {
      const int w = (warps + t - 1) / t;          // warps per tile
      if (w * t - warps < best_waste) { best_waste = w * t - warps; ptiles = t; threads = w * 32; }
    }
  }
  // split the triangle list across CTAs until the grid covers ~2 waves
  int chunks = (F + RC_CHUNK - 1) / RC_CHUNK;
  int splits = 1;
  while (splits < chunks && (long long)ptiles * B * splits < 2LL * num_sms()) ++splits;
  int per = ((chunks + splits - 1) / splits) * RC_CHUNK;
  splits = (F + per - 1) / per;
  static int packed = -1;
  if (packed < 0) {
    const char* e = getenv("OBMAN_RAYCAST_PACKED");   // 0: the scalar kernel
    packed = (e && e[0] == '0') ? 0 : 1;
  }
  if (packed) {
    // PT points per thread; tiles of whole warps (<= RC_PACKED_THREADS) that leave the fewest idle lanes
    static int pt = -1;
    if (pt < 0) {
      const char* e = getenv("OBMAN_RAYCAST_PT");
      pt = (e && atoi(e) == 4) ? 4 : 2;
    }
    const int w2 = (P + 32 * pt - 1) / (32 * pt);        // warps when every lane owns PT points
    const int maxw = RC_PACKED_THREADS / 32;
    int t2 = (w2 + maxw - 1) / maxw, thr2 = RC_PACKED_THREADS, best = 1 << 30;
    for (int t = t2; t <= t2 + 3 && t <= w2; ++t) {
      const int w = (w2 + t - 1) / t;
      if (w * t - w2 < best) { best = w * t - w2; t2 = t; thr2 = w * 32; }
    }
    int sp2 = 1;
    while (sp2 < chunks && (long long)t2 * B * sp2 < 2LL * num_sms()) ++sp2;
    const int per2 = ((chunks + sp2 - 1) / sp2) * RC_CHUNK;
    sp2 = (F + per2 - 1) / per2;
    if (pt == 4) raycast_packed_kernel<4><<<dim3(t2, B, sp2), thr2, 0, st>>>(points, obj_verts, faces, P, N, F, per2, hits);
    else raycast_packed_kernel<2><<<dim3(t2, B, sp2), thr2, 0, st>>>(points, obj_verts, faces, P, N, F, per2, hits);
    return check_launch("raycast_packed_kernel");
  }
  raycast_kernel<<<dim3(ptiles, B, splits), threads, 0, st>>>(points, obj_verts, faces, P, N, F,
                                                                per, hits);
  return check_launch("raycast_kernel");
}

extern "C" int obman_contact_fwd(const float* hand, const float* obj, const float* mins21,
                                 const int* idx21, const int* hits, const int* zone_ids,
                                 const int* zone_ptr, int n_zones, int zones_mode, int B, int P,
                                 int N, float contact_thresh, int contact_mode,
                                 float collision_thresh, int collision_mode,
                                 unsigned char* attr_mask, unsigned char* rep_mask, float* close,
                                 float* anchor, float* partial, float* out, void* stream) {
  OBMAN_REQUIRE(B > 0 && P > 0 && N > 0 && P <= 32768, "obman_contact_fwd: bad sizes");
  OBMAN_REQUIRE(hand && obj && mins21 && idx21 && hits && attr_mask && rep_mask && close &&
                    anchor && partial && out, "obman_contact_fwd: null argument");
  OBMAN_REQUIRE(contact_mode >= 0 && contact_mode <= 2 && collision_mode >= 0 &&
                    collision_mode <= 2, "obman_contact_fwd: mode not in [dist_sq|dist|dist_tanh]");
  OBMAN_REQUIRE(zones_mode >= 0 && zones_mode <= 2, "obman_contact_fwd: zones not in [all|tips|zones]");
  OBMAN_REQUIRE(zones_mode == ZONES_ALL || (zone_ids && zone_ptr && n_zones > 0),
                "obman_contact_fwd: zone table required for tips/zones");
  cudaStream_t st = (cudaStream_t)stream;
  contact_values_kernel<<<B, CV_THREADS, P, st>>>(hand, obj, mins21, idx21, hits, zone_ids,
                                                  zone_ptr, n_zones, zones_mode, P, N,
                                                  contact_thresh, contact_mode, collision_thresh,
                                                  collision_mode, attr_mask, rep_mask, close,
                                                  anchor, partial);
  int rc = check_launch("contact_values_kernel");
  if (rc) return rc;
  contact_finalize_kernel<<<1, 256, 0, st>>>(partial, B, out);
  return check_launch("contact_finalize_kernel");
}

extern "C" int obman_contact_bwd(const float* hand, const float* close, const float* anchor,
                                 const int* idx21, const unsigned char* attr_mask,
                                 const unsigned char* rep_mask, const float* fwd_out,
                                 const float* g_missed, const float* g_penetr, int B, int P, int N,
                                 float contact_thresh, int contact_mode, float collision_thresh,
                                 int collision_mode, int target, float* ghand, float* gobj,
                                 void* stream) {
  OBMAN_REQUIRE(B > 0 && P > 0 && N > 0, "obman_contact_bwd: bad sizes");
  OBMAN_REQUIRE(hand && close && anchor && idx21 && attr_mask && rep_mask && fwd_out && g_missed &&
                    g_penetr, "obman_contact_bwd: null argument");
  OBMAN_REQUIRE(target >= 0 && target <= 2, "obman_contact_bwd: contact_target not in [all|obj|hand]");
  cudaStream_t st = (cudaStream_t)stream;
  if (gobj) cudaMemsetAsync(gobj, 0, sizeof(float) * (size_t)B * N * 3, st);
  size_t total = (size_t)B * P;
  contact_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
      hand, close, anchor, idx21, attr_mask, rep_mask, fwd_out, g_missed, g_penetr, B, P, N,
      contact_thresh, contact_mode, collision_thresh, collision_mode, target, ghand, gobj);
  return check_launch("contact_bwd_kernel");
}

extern "C" int obman_contact_iou(const float* gt_dists, const float* pred_dists, int B, int P, const float* threshs,
                                 int n_thresh, float* iou_ws, float* batch_ious, float* auc, void* stream) {
  OBMAN_REQUIRE(gt_dists && pred_dists && threshs && iou_ws && batch_ious && auc, "obman_contact_iou: null argument");
  OBMAN_REQUIRE(B > 0 && P > 0 && n_thresh >= 1 && n_thresh <= IOU_MAX_T,
                "obman_contact_iou: bad sizes (B=%d P=%d thresholds=%d, at most %d)", B, P, n_thresh, IOU_MAX_T);
  IouThresholds th;
  th.n = n_thresh;
  for (int k = 0; k < IOU_MAX_T; ++k) th.t[k] = k < n_thresh ? threshs[k] : 0.f;
  cudaStream_t st = (cudaStream_t)stream;
  contact_iou_kernel<<<B, 256, 0, st>>>(gt_dists, pred_dists, P, th, iou_ws);
  int rc = check_launch("contact_iou_kernel");
  if (rc) return rc;
  contact_iou_finalize_kernel<<<1, 32 * IOU_MAX_T, 0, st>>>(iou_ws, B, th, batch_ious, auc);
  return check_launch("contact_iou_finalize_kernel");
}
