// MANO layer: batched Rodrigues + blend shapes + kinematic chain + linear blend skinning, fwd and bwd.
//
// Replaces the ~60 ATen launches of manopth.manolayer.ManoLayer.forward (external dependency; call
// site mano_train/networks/branches/manobranch.py:170-182, algorithm restated in SURVEY.md §8a-M and
// oracle/mano.py) by three kernels per direction:
//   pose   : one warp per sample  - PCA -> axis-angle -> 16 Rodrigues -> joints -> chain -> G'
//   lbs    : vertex-tiled         - 64-vertex slices of posedirs/shapedirs/weights live in shared
//                                   memory (104 KB), samples stream through the CTA
//   finish : one CTA per sample   - tips / palm root / 21-joint reorder / centring / metres->mm
// Joints are regressed through the precomputed J_template = Jreg*v_template and
// J_shapedirs = Jreg*shapedirs so no kernel ever needs all 778 vertices of a sample at once.
#include "common.cuh"

namespace obman {

constexpr int NJ = 16;
constexpr int NPM = 135;   // pose-map entries (15 joints x 9)
constexpr int NB = 10;     // betas
constexpr int VT = 64;     // vertices per CTA tile
constexpr int SL = 4;      // sample lanes per CTA
constexpr int GP = NJ * 12;

__device__ __forceinline__ int mano_parent(int j) { return j == 0 ? -1 : (((j - 1) % 3) == 0 ? 0 : j - 1); }
__device__ __forceinline__ int mano_depth(int j) { return j == 0 ? 0 : ((j - 1) % 3) + 1; }

struct ManoTables {
  const float* v_template;   // (V,3)
  const float* shapedirs;    // (V,3,10)
  const float* posedirs;     // (V,3,135)
  const float* weights;      // (V,16)
  const float* j_template;   // (16,3)
  const float* j_shapedirs;  // (16,3,10)
  const float* hands_mean;   // (45)
  const float* comps;        // (C,45)
  const float* default_betas;  // (10)
  int V;
  int ncomps;
};

// ---- per-warp pose stage (shared by fwd and bwd) -----------------------------------------------------
struct PoseSmem {
  float pose[48];
  float R[NJ][9];    // local rotations
  float Rw[NJ][9];   // world rotations
  float J[NJ][3];    // rest joints
  float tw[NJ][3];   // world joint positions
  float beta[NB];
};

__device__ __forceinline__ void rodrigues_fwd(const float* a, float* R) {
  const float e = 1e-8f;
  const float ax = a[0] + e, ay = a[1] + e, az = a[2] + e;
  const float th = sqrtf(ax * ax + ay * ay + az * az);
  const float nx = a[0] / th, ny = a[1] / th, nz = a[2] / th;
  const float h = 0.5f * th;
  float sn, cs;
  sincosf(h, &sn, &cs);
  float w = cs, x = sn * nx, y = sn * ny, z = sn * nz;
  const float inv = 1.0f / sqrtf(w * w + x * x + y * y + z * z);
  w *= inv; x *= inv; y *= inv; z *= inv;
  R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * x * y - 2 * w * z; R[2] = 2 * w * y + 2 * x * z;
  R[3] = 2 * w * z + 2 * x * y; R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * y * z - 2 * w * x;
  R[6] = 2 * x * z - 2 * w * y; R[7] = 2 * w * x + 2 * y * z; R[8] = w * w - x * x - y * y + z * z;
}

// g (dL/dR, 9) -> ga (dL/d axis-angle, 3)
__device__ __forceinline__ void rodrigues_bwd(const float* a, const float* g, float* ga) {
  const float e = 1e-8f;
  const float ax = a[0] + e, ay = a[1] + e, az = a[2] + e;
  const float th = sqrtf(ax * ax + ay * ay + az * az);
  const float nx = a[0] / th, ny = a[1] / th, nz = a[2] / th;
  const float h = 0.5f * th;
  float sn, cs;
  sincosf(h, &sn, &cs);
  const float qw = cs, qx = sn * nx, qy = sn * ny, qz = sn * nz;
  const float qn = sqrtf(qw * qw + qx * qx + qy * qy + qz * qz);
  const float w = qw / qn, x = qx / qn, y = qy / qn, z = qz / qn;
  // dL/d(normalised quaternion)
  const float gw = 2 * w * (g[0] + g[4] + g[8]) - 2 * z * g[1] + 2 * y * g[2] + 2 * z * g[3] - 2 * x * g[5] - 2 * y * g[6] + 2 * x * g[7];
  const float gx = 2 * x * (g[0] - g[4] - g[8]) + 2 * y * g[1] + 2 * z * g[2] + 2 * y * g[3] - 2 * w * g[5] + 2 * z * g[6] + 2 * w * g[7];
  const float gy = 2 * y * (-g[0] + g[4] - g[8]) + 2 * x * g[1] + 2 * w * g[2] + 2 * x * g[3] + 2 * z * g[5] - 2 * w * g[6] + 2 * z * g[7];
  const float gz = 2 * z * (-g[0] - g[4] + g[8]) - 2 * w * g[1] + 2 * x * g[2] + 2 * w * g[3] + 2 * y * g[5] + 2 * x * g[6] + 2 * y * g[7];
  // through the normalisation
  const float dot = w * gw + x * gx + y * gy + z * gz;
  const float uw = (gw - w * dot) / qn, ux = (gx - x * dot) / qn, uy = (gy - y * dot) / qn, uz = (gz - z * dot) / qn;
  // q = (cos h, sin h * n)
  const float gh = -sn * uw + cs * (nx * ux + ny * uy + nz * uz);
  const float gnx = sn * ux, gny = sn * uy, gnz = sn * uz;
  float gth = 0.5f * gh;
  // n = a / th
  gth += -(a[0] * gnx + a[1] * gny + a[2] * gnz) / (th * th);
  ga[0] = gnx / th + gth * ax / th;
  ga[1] = gny / th + gth * ay / th;
  ga[2] = gnz / th + gth * az / th;
}

__device__ __forceinline__ void mat3_mul(const float* A, const float* B, float* C) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c)
      C[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
}

// Forward pose stage for one sample, executed by one warp. Fills S.
__device__ void pose_stage(const ManoTables& T, const float* __restrict__ pose_in,
                           const float* __restrict__ betas_in, PoseSmem& S) {
  const int lane = threadIdx.x & 31;
  if (lane < NB) S.beta[lane] = betas_in ? betas_in[lane] : T.default_betas[lane];
  // full pose = [global rot (3), hands_mean + coeffs @ comps]
  for (int i = lane; i < 48; i += 32) {
    float v;
    if (i < 3) v = pose_in[i];
    else {
      v = T.hands_mean[i - 3];
      for (int c = 0; c < T.ncomps; ++c) v = fmaf(pose_in[3 + c], T.comps[c * 45 + (i - 3)], v);
    }
    S.pose[i] = v;
  }
  __syncwarp();
  if (lane < NJ) {
    rodrigues_fwd(&S.pose[3 * lane], S.R[lane]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = T.j_template[lane * 3 + c];
#pragma unroll
      for (int k = 0; k < NB; ++k) v = fmaf(T.j_shapedirs[(lane * 3 + c) * NB + k], S.beta[k], v);
      S.J[lane][c] = v;
    }
  }
  __syncwarp();
  for (int d = 0; d <= 3; ++d) {
    if (lane < NJ && mano_depth(lane) == d) {
      const int p = mano_parent(lane);
      if (p < 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) S.Rw[0][i] = S.R[0][i];
#pragma unroll
        for (int c = 0; c < 3; ++c) S.tw[0][c] = S.J[0][c];
      } else {
        mat3_mul(S.Rw[p], S.R[lane], S.Rw[lane]);
        const float dx = S.J[lane][0] - S.J[p][0], dy = S.J[lane][1] - S.J[p][1], dz = S.J[lane][2] - S.J[p][2];
#pragma unroll
        for (int r = 0; r < 3; ++r)
          S.tw[lane][r] = S.tw[p][r] + S.Rw[p][3 * r] * dx + S.Rw[p][3 * r + 1] * dy + S.Rw[p][3 * r + 2] * dz;
      }
    }
    __syncwarp();
  }
}

constexpr int POSE_WARPS = 4;

// Outputs: pose_map (B,135), gp (B,16,12) = [Rw | tw - Rw*J], tw (B,16,3), betas_used (B,10)
__global__ void __launch_bounds__(POSE_WARPS * 32)
mano_pose_fwd_kernel(ManoTables T, const float* __restrict__ pose, int pose_stride,
                     const float* __restrict__ betas, int B, float* __restrict__ pose_map,
                     float* __restrict__ gp, float* __restrict__ tw_out,
                     float* __restrict__ betas_used) {
  __shared__ PoseSmem smem[POSE_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * POSE_WARPS + warp;
  if (b >= B) return;
  PoseSmem& S = smem[warp];
  pose_stage(T, pose + (size_t)b * pose_stride, betas ? betas + (size_t)b * NB : nullptr, S);
  for (int i = lane; i < NPM; i += 32) {
    const int j = 1 + i / 9, e = i % 9;
    pose_map[(size_t)b * NPM + i] = S.R[j][e] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
  }
  if (lane < NJ) {
    float* g = gp + ((size_t)b * NJ + lane) * 12;
#pragma unroll
    for (int i = 0; i < 9; ++i) g[i] = S.Rw[lane][i];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      g[9 + r] = S.tw[lane][r] - (S.Rw[lane][3 * r] * S.J[lane][0] + S.Rw[lane][3 * r + 1] * S.J[lane][1] +
                                  S.Rw[lane][3 * r + 2] * S.J[lane][2]);
      tw_out[((size_t)b * NJ + lane) * 3 + r] = S.tw[lane][r];
    }
  }
  if (lane < NB) betas_used[(size_t)b * NB + lane] = S.beta[lane];
}

// ---- vertex-tiled LBS ----------------------------------------------------------------------------------
struct LbsSmem {
  float P[VT * 3 * NPM];   // posedirs slice  [v][c][k]
  float Sd[VT * 3 * NB];   // shapedirs slice [v][c][k]
  float W[VT * NJ];        // weights slice   [v][j]
  float vt[VT * 3];
  float pm[SL][NPM];       // per-sample pose map
  float be[SL][NB];
  float g[SL][GP];         // per-sample G'
  float gv[SL][VT][16];    // bwd: per-vertex (g_vposed[3], gT[12]) (+1 pad)
};

__device__ __forceinline__ void lbs_load_tables(const ManoTables& T, int v0, int nv, LbsSmem& S) {
  const int tid = threadIdx.y * VT + threadIdx.x, nt = VT * SL;
  for (int i = tid; i < VT * 3 * NPM; i += nt) S.P[i] = i < nv * 3 * NPM ? T.posedirs[(size_t)v0 * 3 * NPM + i] : 0.f;
  for (int i = tid; i < VT * 3 * NB; i += nt) S.Sd[i] = i < nv * 3 * NB ? T.shapedirs[(size_t)v0 * 3 * NB + i] : 0.f;
  for (int i = tid; i < VT * NJ; i += nt) S.W[i] = i < nv * NJ ? T.weights[(size_t)v0 * NJ + i] : 0.f;
  for (int i = tid; i < VT * 3; i += nt) S.vt[i] = i < nv * 3 ? T.v_template[(size_t)v0 * 3 + i] : 0.f;
}

__device__ __forceinline__ void lbs_load_samples(const float* pose_map, const float* betas_used,
                                                 const float* gp, int b0, int B, LbsSmem& S) {
  const int tid = threadIdx.y * VT + threadIdx.x, nt = VT * SL;
  for (int i = tid; i < SL * NPM; i += nt) {
    int s = i / NPM, k = i % NPM;
    S.pm[s][k] = (b0 + s) < B ? pose_map[(size_t)(b0 + s) * NPM + k] : 0.f;
  }
  for (int i = tid; i < SL * NB; i += nt) {
    int s = i / NB, k = i % NB;
    S.be[s][k] = (b0 + s) < B ? betas_used[(size_t)(b0 + s) * NB + k] : 0.f;
  }
  for (int i = tid; i < SL * GP; i += nt) {
    int s = i / GP, k = i % GP;
    S.g[s][k] = (b0 + s) < B ? gp[(size_t)(b0 + s) * GP + k] : 0.f;
  }
}

// v_posed and blended transform for (vertex threadIdx.x, sample lane threadIdx.y)
__device__ __forceinline__ void lbs_point(const LbsSmem& S, int v, int s, float* vp, float* Tm) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float acc = S.vt[v * 3 + c];
    const float* sd = &S.Sd[(v * 3 + c) * NB];
#pragma unroll
    for (int k = 0; k < NB; ++k) acc = fmaf(sd[k], S.be[s][k], acc);
    const float* pd = &S.P[(v * 3 + c) * NPM];
#pragma unroll 9
    for (int k = 0; k < NPM; ++k) acc = fmaf(pd[k], S.pm[s][k], acc);
    vp[c] = acc;
  }
#pragma unroll
  for (int e = 0; e < 12; ++e) Tm[e] = 0.f;
#pragma unroll 4
  for (int j = 0; j < NJ; ++j) {
    const float w = S.W[v * NJ + j];
#pragma unroll
    for (int e = 0; e < 12; ++e) Tm[e] = fmaf(w, S.g[s][j * 12 + e], Tm[e]);
  }
}

__global__ void __launch_bounds__(VT * SL)
mano_lbs_fwd_kernel(ManoTables T, const float* __restrict__ pose_map,
                    const float* __restrict__ betas_used, const float* __restrict__ gp, int B,
                    int samples_per_cta, float* __restrict__ verts_raw) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LbsSmem& S = *reinterpret_cast<LbsSmem*>(smem_raw);
  const int v0 = blockIdx.x * VT;
  const int nv = min(VT, T.V - v0);
  lbs_load_tables(T, v0, nv, S);
  const int b_begin = blockIdx.y * samples_per_cta;
  const int b_end = min(B, b_begin + samples_per_cta);
  for (int b0 = b_begin; b0 < b_end; b0 += SL) {
    __syncthreads();
    lbs_load_samples(pose_map, betas_used, gp, b0, b_end, S);
    __syncthreads();
    const int v = threadIdx.x, s = threadIdx.y, b = b0 + s;
    if (v < nv && b < b_end) {
      float vp[3], Tm[12];
      lbs_point(S, v, s, vp, Tm);
      float* o = verts_raw + ((size_t)b * T.V + v0 + v) * 3;
#pragma unroll
      for (int r = 0; r < 3; ++r)
        o[r] = Tm[3 * r] * vp[0] + Tm[3 * r + 1] * vp[1] + Tm[3 * r + 2] * vp[2] + Tm[9 + r];
    }
  }
}

// ---- finish: joints, centring, units -------------------------------------------------------------------
struct FinishParams {
  int tips[5];
  int palm_a, palm_b;
  int root_palm;
  int center_idx;   // -1: none
  int use_trans;    // 1: add trans (B,3) instead of centring
};
__constant__ int kJointReorder[21] = {0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5, 6, 18, 10, 11, 12, 19, 7, 8, 9, 20};

__global__ void __launch_bounds__(256)
mano_finish_fwd_kernel(FinishParams fp, const float* __restrict__ tw, const float* __restrict__ trans,
                       int V, float* __restrict__ verts, float* __restrict__ joints) {
  __shared__ float pre[21][3];
  __shared__ float centre[3];
  const int b = blockIdx.x;
  float* vb = verts + (size_t)b * V * 3;
  if (threadIdx.x < 63) {
    const int j = threadIdx.x / 3, c = threadIdx.x % 3;
    float val;
    if (j < NJ) {
      val = tw[((size_t)b * NJ + j) * 3 + c];
      if (j == 0 && fp.root_palm) val = 0.5f * (vb[fp.palm_a * 3 + c] + vb[fp.palm_b * 3 + c]);
    } else {
      val = vb[fp.tips[j - NJ] * 3 + c];
    }
    pre[j][c] = val;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    float c = 0.f;
    if (fp.use_trans) c = -trans[(size_t)b * 3 + threadIdx.x];
    else if (fp.center_idx >= 0) c = pre[kJointReorder[fp.center_idx]][threadIdx.x];
    centre[threadIdx.x] = c;
  }
  __syncthreads();
  if (threadIdx.x < 63) {
    const int k = threadIdx.x / 3, c = threadIdx.x % 3;
    joints[((size_t)b * 21 + k) * 3 + c] = (pre[kJointReorder[k]][c] - centre[c]) * 1000.f;
  }
  for (int i = threadIdx.x; i < V * 3; i += blockDim.x) vb[i] = (vb[i] - centre[i % 3]) * 1000.f;
}

// gverts (B,V,3) / gjoints (B,21,3) (either may be null) -> gv_raw (B,V,3), gtw (B,16,3)
__global__ void __launch_bounds__(256)
mano_finish_bwd_kernel(FinishParams fp, const float* __restrict__ gverts,
                       const float* __restrict__ gjoints, int V, float* __restrict__ gv_raw,
                       float* __restrict__ gtw) {
  __shared__ float scratch[32];
  __shared__ float gsum[3];
  __shared__ float gpre[21][3];
  const int b = blockIdx.x;
  float acc[3] = {0.f, 0.f, 0.f};
  float* go = gv_raw + (size_t)b * V * 3;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float g = gverts ? gverts[((size_t)b * V + v) * 3 + c] * 1000.f : 0.f;
      go[v * 3 + c] = g;
      acc[c] += g;
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float r = block_sum(acc[c], scratch);
    if (threadIdx.x == 0) gsum[c] = r;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int c = threadIdx.x;
    float gj[21];
    float tot = gsum[c];
    for (int k = 0; k < 21; ++k) {
      gj[k] = gjoints ? gjoints[((size_t)b * 21 + k) * 3 + c] * 1000.f : 0.f;
      tot += gj[k];
    }
    if (!fp.use_trans && fp.center_idx >= 0) gj[fp.center_idx] -= tot;
    for (int k = 0; k < 21; ++k) gpre[kJointReorder[k]][c] = gj[k];
  }
  __syncthreads();  // also orders the go[] writes above before the tip/palm updates below
  if (threadIdx.x < 3) {
    const int c = threadIdx.x;
    for (int t = 0; t < 5; ++t) go[fp.tips[t] * 3 + c] += gpre[NJ + t][c];
    if (fp.root_palm) {
      go[fp.palm_a * 3 + c] += 0.5f * gpre[0][c];
      go[fp.palm_b * 3 + c] += 0.5f * gpre[0][c];
    }
  }
  if (threadIdx.x < 48) {
    const int j = threadIdx.x / 3, c = threadIdx.x % 3;
    gtw[((size_t)b * NJ + j) * 3 + c] = (j == 0 && fp.root_palm) ? 0.f : gpre[j][c];
  }
}

// ---- LBS backward: per-sample accumulators gacc (B, 135 + 10 + 192), zero-initialised ---------------------
constexpr int NACC = NPM + NB + GP;  // 337

__global__ void __launch_bounds__(VT * SL)
mano_lbs_bwd_kernel(ManoTables T, const float* __restrict__ pose_map,
                    const float* __restrict__ betas_used, const float* __restrict__ gp,
                    const float* __restrict__ gv_raw, int B, int samples_per_cta,
                    float* __restrict__ gacc) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  LbsSmem& S = *reinterpret_cast<LbsSmem*>(smem_raw);
  const int v0 = blockIdx.x * VT;
  const int nv = min(VT, T.V - v0);
  lbs_load_tables(T, v0, nv, S);
  const int b_begin = blockIdx.y * samples_per_cta;
  const int b_end = min(B, b_begin + samples_per_cta);
  const int tid = threadIdx.y * VT + threadIdx.x;
  for (int b0 = b_begin; b0 < b_end; b0 += SL) {
    __syncthreads();
    lbs_load_samples(pose_map, betas_used, gp, b0, b_end, S);
    __syncthreads();
    {
      const int v = threadIdx.x, s = threadIdx.y, b = b0 + s;
      float out[15];
#pragma unroll
      for (int i = 0; i < 15; ++i) out[i] = 0.f;
      if (v < nv && b < b_end) {
        float vp[3], Tm[12];
        lbs_point(S, v, s, vp, Tm);
        const float* g = gv_raw + ((size_t)b * T.V + v0 + v) * 3;
        const float g0 = g[0], g1 = g[1], g2 = g[2];
#pragma unroll
        for (int c = 0; c < 3; ++c) out[c] = Tm[c] * g0 + Tm[3 + c] * g1 + Tm[6 + c] * g2;
        const float gg[3] = {g0, g1, g2};
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int c = 0; c < 3; ++c) out[3 + 3 * r + c] = gg[r] * vp[c];
          out[12 + r] = gg[r];
        }
      }
#pragma unroll
      for (int i = 0; i < 15; ++i) S.gv[s][v][i] = out[i];
    }
    __syncthreads();
    for (int o = tid; o < SL * NACC; o += VT * SL) {
      const int s = o / NACC, k = o % NACC;
      if (b0 + s >= b_end) continue;
      float acc = 0.f;
      if (k < NPM) {
        for (int v = 0; v < nv; ++v)
#pragma unroll
          for (int c = 0; c < 3; ++c) acc = fmaf(S.P[(v * 3 + c) * NPM + k], S.gv[s][v][c], acc);
      } else if (k < NPM + NB) {
        const int kk = k - NPM;
        for (int v = 0; v < nv; ++v)
#pragma unroll
          for (int c = 0; c < 3; ++c) acc = fmaf(S.Sd[(v * 3 + c) * NB + kk], S.gv[s][v][c], acc);
      } else {
        const int kk = k - NPM - NB, j = kk / 12, e = kk % 12;
        for (int v = 0; v < nv; ++v) acc = fmaf(S.W[v * NJ + j], S.gv[s][v][3 + e], acc);
      }
      atomicAdd(&gacc[(size_t)(b0 + s) * NACC + k], acc);
    }
  }
}

// ---- pose backward: one warp per sample ------------------------------------------------------------------
struct PoseBwdSmem {
  float gRw[NJ][9];
  float gtw[NJ][3];
  float gJ[NJ][3];
  float gR[NJ][9];
  float gfull[48];
};

__global__ void __launch_bounds__(POSE_WARPS * 32)
mano_pose_bwd_kernel(ManoTables T, const float* __restrict__ pose, int pose_stride,
                     const float* __restrict__ betas, int B, const float* __restrict__ gacc,
                     const float* __restrict__ gtw_in, float* __restrict__ gpose,
                     float* __restrict__ gbetas) {
  __shared__ PoseSmem fsm[POSE_WARPS];
  __shared__ PoseBwdSmem bsm[POSE_WARPS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * POSE_WARPS + warp;
  if (b >= B) return;
  PoseSmem& S = fsm[warp];
  PoseBwdSmem& G = bsm[warp];
  pose_stage(T, pose + (size_t)b * pose_stride, betas ? betas + (size_t)b * NB : nullptr, S);
  const float* acc = gacc + (size_t)b * NACC;
  const float* gGp = acc + NPM + NB;
  if (lane < NJ) {
    const int j = lane;
    // G' = [Rw | tw - Rw*J]
    float gt[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) gt[r] = gGp[j * 12 + 9 + r];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int c = 0; c < 3; ++c) G.gRw[j][3 * r + c] = gGp[j * 12 + 3 * r + c] - gt[r] * S.J[j][c];
      G.gtw[j][r] = gt[r] + gtw_in[((size_t)b * NJ + j) * 3 + r];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
      G.gJ[j][c] = -(S.Rw[j][c] * gt[0] + S.Rw[j][3 + c] * gt[1] + S.Rw[j][6 + c] * gt[2]);
  }
  __syncwarp();
  for (int d = 3; d >= 1; --d) {
    if (lane < NJ && mano_depth(lane) == d) {
      const int j = lane, p = mano_parent(j);
      const float* Rp = S.Rw[p];
      // Rw_j = Rw_p * R_j
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          // gR_j = Rw_p^T * gRw_j
          G.gR[j][3 * r + c] = Rp[r] * G.gRw[j][c] + Rp[3 + r] * G.gRw[j][3 + c] + Rp[6 + r] * G.gRw[j][6 + c];
          // gRw_p += gRw_j * R_j^T
          float v = G.gRw[j][3 * r] * S.R[j][3 * c] + G.gRw[j][3 * r + 1] * S.R[j][3 * c + 1] +
                    G.gRw[j][3 * r + 2] * S.R[j][3 * c + 2];
          // tw_j = tw_p + Rw_p (J_j - J_p)
          v += G.gtw[j][r] * (S.J[j][c] - S.J[p][c]);
          atomicAdd(&G.gRw[p][3 * r + c], v);
        }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        atomicAdd(&G.gtw[p][c], G.gtw[j][c]);
        const float v = Rp[c] * G.gtw[j][0] + Rp[3 + c] * G.gtw[j][1] + Rp[6 + c] * G.gtw[j][2];
        atomicAdd(&G.gJ[j][c], v);
        atomicAdd(&G.gJ[p][c], -v);
      }
    }
    __syncwarp();
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 9; ++i) G.gR[0][i] = G.gRw[0][i];
#pragma unroll
    for (int c = 0; c < 3; ++c) G.gJ[0][c] += G.gtw[0][c];
  }
  __syncwarp();
  if (lane < NJ) {
    float g[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) g[i] = G.gR[lane][i] + (lane > 0 ? acc[(lane - 1) * 9 + i] : 0.f);
    rodrigues_bwd(&S.pose[3 * lane], g, &G.gfull[3 * lane]);
  }
  __syncwarp();
  const int nc = T.ncomps;
  for (int c = lane; c < 3 + nc; c += 32) {
    float v;
    if (c < 3) v = G.gfull[c];
    else {
      v = 0.f;
      for (int i = 0; i < 45; ++i) v = fmaf(T.comps[(c - 3) * 45 + i], G.gfull[3 + i], v);
    }
    gpose[(size_t)b * (3 + nc) + c] = v;
  }
  if (gbetas && lane < NB) {
    float v = acc[NPM + lane];
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c) v = fmaf(T.j_shapedirs[(j * 3 + c) * NB + lane], G.gJ[j][c], v);
    gbetas[(size_t)b * NB + lane] = v;
  }
}

static int lbs_samples_per_cta(int B, int vtiles) {
  int want_ctas = 2 * num_sms();
  int chunks = max(1, want_ctas / vtiles);
  int per = (B + chunks - 1) / chunks;
  per = ((per + SL - 1) / SL) * SL;
  return max(SL, per);
}

}  // namespace obman

using namespace obman;

namespace {
struct ManoArgs {
  ManoTables T;
  FinishParams fp;
};
int fill_args(ManoArgs& a, const float* v_template, const float* shapedirs, const float* posedirs,
              const float* weights, const float* j_template, const float* j_shapedirs,
              const float* hands_mean, const float* comps, const float* default_betas, int V,
              int ncomps, int side_left, int root_palm, int center_idx, int use_trans) {
  a.T = {v_template, shapedirs, posedirs, weights, j_template, j_shapedirs, hands_mean, comps,
         default_betas, V, ncomps};
  const int tips_r[5] = {745, 317, 444, 556, 673};
  for (int i = 0; i < 5; ++i) a.fp.tips[i] = tips_r[i];
  if (side_left) a.fp.tips[2] = 445;
  a.fp.palm_a = 95;
  a.fp.palm_b = 22;
  a.fp.root_palm = root_palm;
  a.fp.center_idx = center_idx;
  a.fp.use_trans = use_trans;
  return 0;
}
}  // namespace

// Workspace sizes (floats): pose_map B*135, gp B*192, tw B*48, betas_used B*10.
extern "C" int obman_mano_fwd(const float* v_template, const float* shapedirs, const float* posedirs,
                              const float* weights, const float* j_template,
                              const float* j_shapedirs, const float* hands_mean, const float* comps,
                              const float* default_betas, int V, int ncomps, const float* pose,
                              const float* betas, const float* trans, int B, int side_left,
                              int root_palm, int center_idx, float* ws_pose_map, float* ws_gp,
                              float* ws_tw, float* ws_betas, float* verts, float* joints,
                              void* stream) {
  OBMAN_REQUIRE(B > 0 && V >= 746 && ncomps >= 0 && ncomps <= 45, "obman_mano_fwd: bad sizes (B=%d V=%d ncomps=%d)", B, V, ncomps);
  OBMAN_REQUIRE(center_idx >= -1 && center_idx < 21, "obman_mano_fwd: center_idx out of range");
  OBMAN_REQUIRE(v_template && shapedirs && posedirs && weights && j_template && j_shapedirs &&
                    hands_mean && comps && default_betas && pose && ws_pose_map && ws_gp && ws_tw &&
                    ws_betas && verts && joints, "obman_mano_fwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  ManoArgs a;
  fill_args(a, v_template, shapedirs, posedirs, weights, j_template, j_shapedirs, hands_mean, comps,
            default_betas, V, ncomps, side_left, root_palm, center_idx, trans != nullptr);
  mano_pose_fwd_kernel<<<(B + POSE_WARPS - 1) / POSE_WARPS, POSE_WARPS * 32, 0, st>>>(
      a.T, pose, 3 + ncomps, betas, B, ws_pose_map, ws_gp, ws_tw, ws_betas);
  int rc = check_launch("mano_pose_fwd_kernel");
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(mano_lbs_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LbsSmem));
    cudaFuncSetAttribute(mano_lbs_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LbsSmem));
    attr_set = true;
  }
  const int vtiles = (V + VT - 1) / VT;
  const int per = lbs_samples_per_cta(B, vtiles);
  mano_lbs_fwd_kernel<<<dim3(vtiles, (B + per - 1) / per), dim3(VT, SL), sizeof(LbsSmem), st>>>(
      a.T, ws_pose_map, ws_betas, ws_gp, B, per, verts);
  rc = check_launch("mano_lbs_fwd_kernel");
  if (rc) return rc;
  mano_finish_fwd_kernel<<<B, 256, 0, st>>>(a.fp, ws_tw, trans, V, verts, joints);
  return check_launch("mano_finish_fwd_kernel");
}

// Extra workspace (floats): ws_gv B*V*3, ws_gtw B*48, ws_gacc B*337.
extern "C" int obman_mano_bwd(const float* v_template, const float* shapedirs, const float* posedirs,
                              const float* weights, const float* j_template,
                              const float* j_shapedirs, const float* hands_mean, const float* comps,
                              const float* default_betas, int V, int ncomps, const float* pose,
                              const float* betas, int has_trans, int B, int side_left, int root_palm,
                              int center_idx, const float* ws_pose_map, const float* ws_gp,
                              const float* ws_betas, const float* gverts, const float* gjoints,
                              float* ws_gv, float* ws_gtw, float* ws_gacc, float* gpose,
                              float* gbetas, void* stream) {
  OBMAN_REQUIRE(B > 0 && V >= 746 && ncomps >= 0 && ncomps <= 45, "obman_mano_bwd: bad sizes");
  OBMAN_REQUIRE(pose && ws_pose_map && ws_gp && ws_betas && ws_gv && ws_gtw && ws_gacc && gpose,
                "obman_mano_bwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  ManoArgs a;
  fill_args(a, v_template, shapedirs, posedirs, weights, j_template, j_shapedirs, hands_mean, comps,
            default_betas, V, ncomps, side_left, root_palm, center_idx, has_trans);
  mano_finish_bwd_kernel<<<B, 256, 0, st>>>(a.fp, gverts, gjoints, V, ws_gv, ws_gtw);
  int rc = check_launch("mano_finish_bwd_kernel");
  if (rc) return rc;
  cudaMemsetAsync(ws_gacc, 0, sizeof(float) * (size_t)B * NACC, st);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(mano_lbs_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LbsSmem));
    cudaFuncSetAttribute(mano_lbs_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LbsSmem));
    attr_set = true;
  }
  const int vtiles = (V + VT - 1) / VT;
  const int per = lbs_samples_per_cta(B, vtiles);
  mano_lbs_bwd_kernel<<<dim3(vtiles, (B + per - 1) / per), dim3(VT, SL), sizeof(LbsSmem), st>>>(
      a.T, ws_pose_map, ws_betas, ws_gp, ws_gv, B, per, ws_gacc);
  rc = check_launch("mano_lbs_bwd_kernel");
  if (rc) return rc;
  mano_pose_bwd_kernel<<<(B + POSE_WARPS - 1) / POSE_WARPS, POSE_WARPS * 32, 0, st>>>(
      a.T, pose, 3 + ncomps, betas, B, ws_gacc, ws_gtw, gpose, gbetas);
  return check_launch("mano_pose_bwd_kernel");
}
