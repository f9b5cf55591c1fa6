// Mesh regularisers of AtlasLoss on the predicted object mesh (fixed icosphere topology):
//   laplacian  cotangent-Laplacian loss  mean_{b,i} || (L V_b)_i ||_2      (laplacianloss.py:24-41,137-150)
//   edge       edge-length regulariser   mean |e - mean_b(e)|, e = squared edge lengths (atlasbranch.py:153-167)
// Both are HBM/L2-latency bound gathers over a CONSTANT topology: L is the cotangent Laplacian of the unit
// icosphere (<= 7 non-zeros per row), passed as an ELL table (neighbour ids + off-diagonal weights); the edge
// gradient uses a vertex -> incident-faces table so that no float atomics are needed (bit-reproducible).
// Algorithmic bytes per sample: laplacian fwd 12 N read + 12 N written (+ K*(4+4) N table bytes, L2 resident),
// edge fwd 12 N read + 12 F index bytes.
#include "common.cuh"

namespace obman {

// Lx[b,i] = sum_k w[i,k] * (V[b,nbr[i,k]] - V[b,i])  ( = (L V_b)_i because L_ii = -sum_j L_ij ),
// partial[block] = sum over the block's rows of ||Lx||.
__global__ void __launch_bounds__(256)
laplacian_fwd_kernel(const float* __restrict__ V, const int* __restrict__ nbr, const float* __restrict__ w, int B,
                     int N, int K, float* __restrict__ Lx, float* __restrict__ partial) {
  __shared__ float scratch[32];
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float norm = 0.f;
  if (t < (long long)B * N) {
    const int i = (int)(t % N);
    const float* Vb = V + (t - i) * 3;
    const float x = Vb[i * 3], y = Vb[i * 3 + 1], z = Vb[i * 3 + 2];
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int k = 0; k < K; ++k) {
      const int j = __ldg(nbr + (size_t)i * K + k);
      const float wk = __ldg(w + (size_t)i * K + k);
      ax = fmaf(wk, Vb[j * 3] - x, ax);
      ay = fmaf(wk, Vb[j * 3 + 1] - y, ay);
      az = fmaf(wk, Vb[j * 3 + 2] - z, az);
    }
    Lx[t * 3] = ax;
    Lx[t * 3 + 1] = ay;
    Lx[t * 3 + 2] = az;
    norm = sqrtf(ax * ax + ay * ay + az * az);
  }
  const float s = block_sum(norm, scratch);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// out[0] = scale * sum_i partial[i], summed in a fixed order by one block (deterministic).
__global__ void __launch_bounds__(256)
sum_scale_kernel(const float* __restrict__ partial, int n, float scale, float* __restrict__ out) {
  __shared__ float scratch[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
  const float s = block_sum(acc, scratch);
  if (threadIdx.x == 0) out[0] = s * scale;
}

// gV = L^T u = L u with u[b,j] = gloss * inv_count * Lx[b,j] / ||Lx[b,j]||  (0 where the norm is 0, like torch.norm).
__global__ void __launch_bounds__(256)
laplacian_bwd_kernel(const float* __restrict__ Lx, const float* __restrict__ gloss, float inv_count,
                     const int* __restrict__ nbr, const float* __restrict__ w, int B, int N, int K,
                     float* __restrict__ gV) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)B * N) return;
  const int i = (int)(t % N);
  const float* Lb = Lx + (t - i) * 3;
  const float g = gloss[0] * inv_count;
  float ux, uy, uz;
  {
    const float x = Lb[i * 3], y = Lb[i * 3 + 1], z = Lb[i * 3 + 2];
    const float n = sqrtf(x * x + y * y + z * z);
    const float s = n > 0.f ? g / n : 0.f;
    ux = x * s; uy = y * s; uz = z * s;
  }
  float ax = 0.f, ay = 0.f, az = 0.f;
  for (int k = 0; k < K; ++k) {
    const int j = __ldg(nbr + (size_t)i * K + k);
    const float wk = __ldg(w + (size_t)i * K + k);
    const float x = Lb[j * 3], y = Lb[j * 3 + 1], z = Lb[j * 3 + 2];
    const float n = sqrtf(x * x + y * y + z * z);
    const float s = n > 0.f ? g / n : 0.f;
    ax = fmaf(wk, x * s - ux, ax);
    ay = fmaf(wk, y * s - uy, ay);
    az = fmaf(wk, z * s - uz, az);
  }
  gV[t * 3] = ax;
  gV[t * 3 + 1] = ay;
  gV[t * 3 + 2] = az;
}

__device__ __forceinline__ void face_edges(const float* __restrict__ Vb, int a, int b, int c, float& eA, float& eB,
                                           float& eC) {
  const float ax = Vb[a * 3], ay = Vb[a * 3 + 1], az = Vb[a * 3 + 2];
  const float bx = Vb[b * 3], by = Vb[b * 3 + 1], bz = Vb[b * 3 + 2];
  const float cx = Vb[c * 3], cy = Vb[c * 3 + 1], cz = Vb[c * 3 + 2];
  eA = (bx - ax) * (bx - ax) + (by - ay) * (by - ay) + (bz - az) * (bz - az);
  eB = (cx - bx) * (cx - bx) + (cy - by) * (cy - by) + (cz - bz) * (cz - bz);
  eC = (ax - cx) * (ax - cx) + (ay - cy) * (ay - cy) + (az - cz) * (az - cz);
}

// One block per sample: pass 1 mean of the 3F squared edge lengths, pass 2 sum |e - mean| and sum sign(e - mean).
// stats[b] = {mean, sum_abs_dev, sum_sign}.
__global__ void __launch_bounds__(256)
edge_stats_kernel(const float* __restrict__ V, const int* __restrict__ faces, int N, int F,
                  float* __restrict__ stats) {
  __shared__ float scratch[32];
  __shared__ float mean_sh;
  const float* Vb = V + (size_t)blockIdx.x * N * 3;
  float acc = 0.f;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float eA, eB, eC;
    face_edges(Vb, __ldg(faces + f * 3), __ldg(faces + f * 3 + 1), __ldg(faces + f * 3 + 2), eA, eB, eC);
    acc += eA + eB + eC;
  }
  float s = block_sum(acc, scratch);
  if (threadIdx.x == 0) mean_sh = s / (3.f * F);
  __syncthreads();
  const float mean = mean_sh;
  float dev = 0.f, sgn = 0.f;
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    float e[3];
    face_edges(Vb, __ldg(faces + f * 3), __ldg(faces + f * 3 + 1), __ldg(faces + f * 3 + 2), e[0], e[1], e[2]);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const float d = e[q] - mean;
      dev += fabsf(d);
      sgn += (d > 0.f) ? 1.f : (d < 0.f ? -1.f : 0.f);
    }
  }
  dev = block_sum(dev, scratch);
  sgn = block_sum(sgn, scratch);
  if (threadIdx.x == 0) {
    stats[blockIdx.x * 3] = mean;
    stats[blockIdx.x * 3 + 1] = dev;
    stats[blockIdx.x * 3 + 2] = sgn;
  }
}

__global__ void __launch_bounds__(256)
edge_loss_finish_kernel(const float* __restrict__ stats, int B, float scale, float* __restrict__ out) {
  __shared__ float scratch[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) acc += stats[i * 3 + 1];
  const float s = block_sum(acc, scratch);
  if (threadIdx.x == 0) out[0] = s * scale;
}

// d loss / d e_k = gloss/(B*3F) * (sign_k - sum_sign_b/(3F)); d e/d p = 2 (p - q) for the edge (q -> p).
// Vertex-centric gather over the incident faces (vf (N,Kf), -1 padded): no atomics.
__global__ void __launch_bounds__(256)
edge_loss_bwd_kernel(const float* __restrict__ V, const int* __restrict__ faces, const int* __restrict__ vf,
                     const float* __restrict__ stats, const float* __restrict__ gloss, int B, int N, int F, int Kf,
                     float* __restrict__ gV) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)B * N) return;
  const int i = (int)(t % N);
  const int b = (int)(t / N);
  const float* Vb = V + (size_t)b * N * 3;
  const float mean = stats[b * 3];
  const float corr = stats[b * 3 + 2] / (3.f * F);
  const float g = gloss[0] / ((float)B * 3.f * F);
  const float px = Vb[i * 3], py = Vb[i * 3 + 1], pz = Vb[i * 3 + 2];
  float ax = 0.f, ay = 0.f, az = 0.f;
  for (int k = 0; k < Kf; ++k) {
    const int f = __ldg(vf + (size_t)i * Kf + k);
    if (f < 0) break;
    const int c0 = __ldg(faces + f * 3), c1 = __ldg(faces + f * 3 + 1), c2 = __ldg(faces + f * 3 + 2);
    // the two edges of face f that touch vertex i: towards its other two corners
    int o1, o2;
    if (c0 == i) { o1 = c1; o2 = c2; } else if (c1 == i) { o1 = c0; o2 = c2; } else { o1 = c0; o2 = c1; }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int o = q == 0 ? o1 : o2;
      const float dx = px - Vb[o * 3], dy = py - Vb[o * 3 + 1], dz = pz - Vb[o * 3 + 2];
      const float d = dx * dx + dy * dy + dz * dz - mean;
      const float sg = (d > 0.f) ? 1.f : (d < 0.f ? -1.f : 0.f);
      const float coef = 2.f * g * (sg - corr);
      ax = fmaf(coef, dx, ax);
      ay = fmaf(coef, dy, ay);
      az = fmaf(coef, dz, az);
    }
  }
  gV[t * 3] = ax;
  gV[t * 3 + 1] = ay;
  gV[t * 3 + 2] = az;
}

}  // namespace obman

using namespace obman;

extern "C" int obman_laplacian_fwd(const float* V, const int* nbr, const float* w, int B, int N, int K, float* Lx,
                                   float* partial, float* loss, void* stream) {
  OBMAN_REQUIRE(V && nbr && w && Lx && partial && loss, "obman_laplacian_fwd: null pointer");
  OBMAN_REQUIRE(B > 0 && N > 0 && K > 0 && K <= 64, "obman_laplacian_fwd: bad sizes B=%d N=%d K=%d", B, N, K);
  const long long rows = (long long)B * N;
  const unsigned blocks = (unsigned)((rows + 255) / 256);
  laplacian_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(V, nbr, w, B, N, K, Lx, partial);
  sum_scale_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(partial, (int)blocks, 1.f / (float)rows, loss);
  return check_launch("laplacian_fwd_kernel");
}

extern "C" int obman_laplacian_bwd(const float* Lx, const float* gloss, const int* nbr, const float* w, int B, int N,
                                   int K, float* gV, void* stream) {
  OBMAN_REQUIRE(Lx && gloss && nbr && w && gV, "obman_laplacian_bwd: null pointer");
  OBMAN_REQUIRE(B > 0 && N > 0 && K > 0 && K <= 64, "obman_laplacian_bwd: bad sizes B=%d N=%d K=%d", B, N, K);
  const long long rows = (long long)B * N;
  laplacian_bwd_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      Lx, gloss, 1.f / (float)rows, nbr, w, B, N, K, gV);
  return check_launch("laplacian_bwd_kernel");
}

extern "C" int obman_edge_loss_fwd(const float* V, const int* faces, int B, int N, int F, float* stats, float* loss,
                                   void* stream) {
  OBMAN_REQUIRE(V && faces && stats && loss, "obman_edge_loss_fwd: null pointer");
  OBMAN_REQUIRE(B > 0 && N > 0 && F > 0, "obman_edge_loss_fwd: bad sizes B=%d N=%d F=%d", B, N, F);
  edge_stats_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(V, faces, N, F, stats);
  edge_loss_finish_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(stats, B, 1.f / ((float)B * 3.f * F), loss);
  return check_launch("edge_stats_kernel");
}

extern "C" int obman_edge_loss_bwd(const float* V, const int* faces, const int* vf, const float* stats,
                                   const float* gloss, int B, int N, int F, int Kf, float* gV, void* stream) {
  OBMAN_REQUIRE(V && faces && vf && stats && gloss && gV, "obman_edge_loss_bwd: null pointer");
  OBMAN_REQUIRE(B > 0 && N > 0 && F > 0 && Kf > 0, "obman_edge_loss_bwd: bad sizes B=%d N=%d F=%d Kf=%d", B, N, F, Kf);
  const long long rows = (long long)B * N;
  edge_loss_bwd_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      V, faces, vf, stats, gloss, B, N, F, Kf, gV);
  return check_launch("edge_loss_bwd_kernel");
}
