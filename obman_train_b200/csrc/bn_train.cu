// BatchNorm with BATCH statistics (training without --freeze_batchnorm: model.train() in epochpass3d.py:48-52,
// torch.nn.BatchNorm2d / BatchNorm1d of mano_train/networks/bases/resnet.py:25-54,110-130 and
// branches/atlasutils.py:57-63) around the tensor-core convolutions.  The convolution writes its raw output z
// (rows = pixels or points, C channels, row stride ld); these HBM-bound kernels do the rest:
//
//   bn_stats           per-channel mean / biased variance of z (shifted sums: no cancellation), scale = gamma * rstd,
//                      shift = beta - mean * scale, running-statistics update (momentum, unbiased variance)
//   bn_apply_fwd       y = [relu](z * scale + shift [+ addend])
//   bn_bwd_reduce      g' = g [where y > 0];  sum_g[c] = sum g',  sum_gz[c] = sum g' * zhat   (= d beta, d gamma)
//   bn_bwd_apply       dz = scale * (g' - sum_g / M - zhat * sum_gz / M)   [and g' itself for the residual branch]
//
// Reductions are two-stage with a fixed order (per-chunk partials, then a finalising kernel): bit-reproducible.
#include <string.h>

#include "common.cuh"

namespace obman {

constexpr int BN_THREADS = 256;

// partial[(chunk * C + c) * 2 + {0,1}] = sum over the chunk's rows of (x - pivot[c]) and (x - pivot[c])^2, pivot = row 0.
// C % 4 == 0: a thread owns 4 channels (float4) and walks rows; 256 / (C/4) rows in flight per block iteration.
__global__ void __launch_bounds__(BN_THREADS)
bn_stats_partial_kernel(const float* __restrict__ x, long long rows, int C, long long ld, long long rows_per_chunk,
                        float* __restrict__ partial) {
  extern __shared__ float sh[];   // [rpb][C][2]
  const int groups = C / 4;
  const int rpb = BN_THREADS / groups;
  const int g = threadIdx.x % groups, ry = threadIdx.x / groups;
  const long long r0 = (long long)blockIdx.x * rows_per_chunk;
  const long long r1 = min(rows, r0 + rows_per_chunk);
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  if (ry < rpb) {
    const float4 pv = *reinterpret_cast<const float4*>(x + 4 * g);
    for (long long r = r0 + ry; r < r1; r += rpb) {
      const float4 v = *reinterpret_cast<const float4*>(x + r * ld + 4 * g);
      const float d0 = v.x - pv.x, d1 = v.y - pv.y, d2 = v.z - pv.z, d3 = v.w - pv.w;
      s[0] += d0; s[1] += d1; s[2] += d2; s[3] += d3;
      q[0] = fmaf(d0, d0, q[0]); q[1] = fmaf(d1, d1, q[1]); q[2] = fmaf(d2, d2, q[2]); q[3] = fmaf(d3, d3, q[3]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      sh[((size_t)ry * C + 4 * g + e) * 2] = s[e];
      sh[((size_t)ry * C + 4 * g + e) * 2 + 1] = q[e];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += BN_THREADS) {
    float a = 0.f, b = 0.f;
    for (int y = 0; y < rpb; ++y) { a += sh[((size_t)y * C + c) * 2]; b += sh[((size_t)y * C + c) * 2 + 1]; }
    partial[((size_t)blockIdx.x * C + c) * 2] = a;
    partial[((size_t)blockIdx.x * C + c) * 2 + 1] = b;
  }
}

__global__ void __launch_bounds__(BN_THREADS)
bn_stats_finalize_kernel(const float* __restrict__ x, const float* __restrict__ partial, int chunks, long long rows,
                         int C, const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                         float momentum, float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ scale,
                         float* __restrict__ shift, float* __restrict__ running_mean, float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int k = 0; k < chunks; ++k) { s += partial[((size_t)k * C + c) * 2]; q += partial[((size_t)k * C + c) * 2 + 1]; }
  const double m = (double)rows;
  const double dm = s / m;                        // mean - pivot
  const double var = fmax(q / m - dm * dm, 0.0);  // biased variance (normalisation)
  const float mu = (float)((double)x[c] + dm);
  const float rs = rsqrtf((float)var + eps);
  mean[c] = mu;
  rstd[c] = rs;
  const float sc = (gamma ? gamma[c] : 1.f) * rs;
  scale[c] = sc;
  shift[c] = (beta ? beta[c] : 0.f) - mu * sc;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
  if (running_var) {
    const float unbiased = (float)(rows > 1 ? var * m / (m - 1.0) : var);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
}

// y = [relu](z * scale[c] + shift[c] [+ addend]); one float4 per thread
__global__ void __launch_bounds__(BN_THREADS)
bn_apply_fwd_kernel(const float* __restrict__ z, long long rows, int C, long long ld, const float* __restrict__ scale,
                    const float* __restrict__ shift, const float* __restrict__ addend, int relu,
                    float* __restrict__ y) {
  const int groups = C / 4;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * groups) return;
  const long long r = t / groups;
  const int c = (int)(t - r * groups) * 4;
  const float4 v = *reinterpret_cast<const float4*>(z + r * ld + c);
  const float4 sc = *reinterpret_cast<const float4*>(scale + c);
  const float4 sf = *reinterpret_cast<const float4*>(shift + c);
  float o[4] = {fmaf(v.x, sc.x, sf.x), fmaf(v.y, sc.y, sf.y), fmaf(v.z, sc.z, sf.z), fmaf(v.w, sc.w, sf.w)};
  if (addend) {
    const float4 a = *reinterpret_cast<const float4*>(addend + r * ld + c);
    o[0] += a.x; o[1] += a.y; o[2] += a.z; o[3] += a.w;
  }
  if (relu) {
#pragma unroll
    for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
  }
  *reinterpret_cast<float4*>(y + r * ld + c) = make_float4(o[0], o[1], o[2], o[3]);
}

// partial sums of g' and g' * zhat per chunk and channel (g' = g where mask_src > 0; mask_src nullable)
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_partial_kernel(const float* __restrict__ g, const float* __restrict__ mask_src, const float* __restrict__ z,
                      long long rows, int C, long long ld, const float* __restrict__ mean,
                      const float* __restrict__ rstd, long long rows_per_chunk, float* __restrict__ partial) {
  extern __shared__ float sh[];
  const int groups = C / 4;
  const int rpb = BN_THREADS / groups;
  const int gi = threadIdx.x % groups, ry = threadIdx.x / groups;
  const long long r0 = (long long)blockIdx.x * rows_per_chunk;
  const long long r1 = min(rows, r0 + rows_per_chunk);
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
  if (ry < rpb) {
    const float4 mu = *reinterpret_cast<const float4*>(mean + 4 * gi);
    const float4 rs = *reinterpret_cast<const float4*>(rstd + 4 * gi);
    for (long long r = r0 + ry; r < r1; r += rpb) {
      float4 gv = *reinterpret_cast<const float4*>(g + r * ld + 4 * gi);
      if (mask_src) {
        const float4 m = *reinterpret_cast<const float4*>(mask_src + r * ld + 4 * gi);
        gv.x = m.x > 0.f ? gv.x : 0.f; gv.y = m.y > 0.f ? gv.y : 0.f;
        gv.z = m.z > 0.f ? gv.z : 0.f; gv.w = m.w > 0.f ? gv.w : 0.f;
      }
      const float4 zv = *reinterpret_cast<const float4*>(z + r * ld + 4 * gi);
      s[0] += gv.x; s[1] += gv.y; s[2] += gv.z; s[3] += gv.w;
      q[0] = fmaf(gv.x, (zv.x - mu.x) * rs.x, q[0]); q[1] = fmaf(gv.y, (zv.y - mu.y) * rs.y, q[1]);
      q[2] = fmaf(gv.z, (zv.z - mu.z) * rs.z, q[2]); q[3] = fmaf(gv.w, (zv.w - mu.w) * rs.w, q[3]);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      sh[((size_t)ry * C + 4 * gi + e) * 2] = s[e];
      sh[((size_t)ry * C + 4 * gi + e) * 2 + 1] = q[e];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += BN_THREADS) {
    float a = 0.f, b = 0.f;
    for (int y = 0; y < rpb; ++y) { a += sh[((size_t)y * C + c) * 2]; b += sh[((size_t)y * C + c) * 2 + 1]; }
    partial[((size_t)blockIdx.x * C + c) * 2] = a;
    partial[((size_t)blockIdx.x * C + c) * 2 + 1] = b;
  }
}

__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_finalize_kernel(const float* __restrict__ partial, int chunks, int C, float* __restrict__ sum_g,
                       float* __restrict__ sum_gz) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int k = 0; k < chunks; ++k) { s += partial[((size_t)k * C + c) * 2]; q += partial[((size_t)k * C + c) * 2 + 1]; }
  sum_g[c] = (float)s;
  sum_gz[c] = (float)q;
}

// dz = scale * (g' - sum_g / M - zhat * sum_gz / M); gmasked (nullable) receives g' (the residual branch's gradient)
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_apply_kernel(const float* __restrict__ g, const float* __restrict__ mask_src, const float* __restrict__ z,
                    long long rows, int C, long long ld, const float* __restrict__ mean,
                    const float* __restrict__ rstd, const float* __restrict__ scale, const float* __restrict__ sum_g,
                    const float* __restrict__ sum_gz, float* __restrict__ dz, float* __restrict__ gmasked) {
  const int groups = C / 4;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * groups) return;
  const long long r = t / groups;
  const int c = (int)(t - r * groups) * 4;
  const float inv_m = 1.f / (float)rows;
  float4 gv = *reinterpret_cast<const float4*>(g + r * ld + c);
  if (mask_src) {
    const float4 m = *reinterpret_cast<const float4*>(mask_src + r * ld + c);
    gv.x = m.x > 0.f ? gv.x : 0.f; gv.y = m.y > 0.f ? gv.y : 0.f;
    gv.z = m.z > 0.f ? gv.z : 0.f; gv.w = m.w > 0.f ? gv.w : 0.f;
  }
  if (gmasked) *reinterpret_cast<float4*>(gmasked + r * ld + c) = gv;
  const float4 zv = *reinterpret_cast<const float4*>(z + r * ld + c);
  const float4 mu = *reinterpret_cast<const float4*>(mean + c);
  const float4 rs = *reinterpret_cast<const float4*>(rstd + c);
  const float4 sc = *reinterpret_cast<const float4*>(scale + c);
  const float4 sg = *reinterpret_cast<const float4*>(sum_g + c);
  const float4 sq = *reinterpret_cast<const float4*>(sum_gz + c);
  float4 o;
  o.x = sc.x * (gv.x - sg.x * inv_m - (zv.x - mu.x) * rs.x * sq.x * inv_m);
  o.y = sc.y * (gv.y - sg.y * inv_m - (zv.y - mu.y) * rs.y * sq.y * inv_m);
  o.z = sc.z * (gv.z - sg.z * inv_m - (zv.z - mu.z) * rs.z * sq.z * inv_m);
  o.w = sc.w * (gv.w - sg.w * inv_m - (zv.w - mu.w) * rs.w * sq.w * inv_m);
  *reinterpret_cast<float4*>(dz + r * ld + c) = o;
}

static int chunks_for(long long rows, int C, long long* per) {
  const int rpb = BN_THREADS / (C / 4);
  long long chunks = 4LL * num_sms();
  const long long min_rows = (long long)rpb * 8;
  if (chunks > (rows + min_rows - 1) / min_rows) chunks = (rows + min_rows - 1) / min_rows;
  if (chunks < 1) chunks = 1;
  *per = (rows + chunks - 1) / chunks;
  return (int)((rows + *per - 1) / *per);
}

}  // namespace obman

using namespace obman;

extern "C" int obman_bn_chunks(long long rows, int C) {
  if (rows <= 0 || C <= 0 || C % 4 != 0 || C > 1024) return 0;
  long long per;
  return chunks_for(rows, C, &per);
}

extern "C" int obman_bn_stats(const float* z, long long rows, int C, long long ld, const float* gamma,
                              const float* beta, float eps, float momentum, float* partial, float* mean, float* rstd,
                              float* scale, float* shift, float* running_mean, float* running_var, void* stream) {
  OBMAN_REQUIRE(z && partial && mean && rstd && scale && shift && rows > 0, "obman_bn_stats: bad arguments");
  OBMAN_REQUIRE(C > 0 && C % 4 == 0 && C <= 1024 && ld % 4 == 0 && ld >= C && (((uintptr_t)z) & 15) == 0,
                "obman_bn_stats: C=%d must be a multiple of 4 (<= 1024) and z 16-byte aligned", C);
  long long per;
  const int chunks = chunks_for(rows, C, &per);
  const int rpb = BN_THREADS / (C / 4);
  cudaStream_t st = (cudaStream_t)stream;
  bn_stats_partial_kernel<<<chunks, BN_THREADS, sizeof(float) * 2 * rpb * C, st>>>(z, rows, C, ld, per, partial);
  int rc = check_launch("bn_stats_partial_kernel");
  if (rc) return rc;
  bn_stats_finalize_kernel<<<(C + BN_THREADS - 1) / BN_THREADS, BN_THREADS, 0, st>>>(
      z, partial, chunks, rows, C, gamma, beta, eps, momentum, mean, rstd, scale, shift, running_mean, running_var);
  return check_launch("bn_stats_finalize_kernel");
}

extern "C" int obman_bn_apply_fwd(const float* z, long long rows, int C, long long ld, const float* scale,
                                  const float* shift, const float* addend, int relu, float* y, void* stream) {
  OBMAN_REQUIRE(z && scale && shift && y && rows > 0 && C > 0 && C % 4 == 0 && ld % 4 == 0 && ld >= C,
                "obman_bn_apply_fwd: bad arguments");
  const long long total = rows * (C / 4);
  bn_apply_fwd_kernel<<<(unsigned)((total + BN_THREADS - 1) / BN_THREADS), BN_THREADS, 0, (cudaStream_t)stream>>>(
      z, rows, C, ld, scale, shift, addend, relu, y);
  return check_launch("bn_apply_fwd_kernel");
}

extern "C" int obman_bn_bwd(const float* g, const float* mask_src, const float* z, long long rows, int C, long long ld,
                            const float* mean, const float* rstd, const float* scale, float* partial, float* sum_g,
                            float* sum_gz, float* dz, float* gmasked, void* stream) {
  OBMAN_REQUIRE(g && z && mean && rstd && scale && partial && sum_g && sum_gz && dz && rows > 0,
                "obman_bn_bwd: bad arguments");
  OBMAN_REQUIRE(C > 0 && C % 4 == 0 && C <= 1024 && ld % 4 == 0 && ld >= C, "obman_bn_bwd: C=%d must be a multiple of 4", C);
  long long per;
  const int chunks = chunks_for(rows, C, &per);
  const int rpb = BN_THREADS / (C / 4);
  cudaStream_t st = (cudaStream_t)stream;
  bn_bwd_partial_kernel<<<chunks, BN_THREADS, sizeof(float) * 2 * rpb * C, st>>>(g, mask_src, z, rows, C, ld, mean, rstd,
                                                                                per, partial);
  int rc = check_launch("bn_bwd_partial_kernel");
  if (rc) return rc;
  bn_bwd_finalize_kernel<<<(C + BN_THREADS - 1) / BN_THREADS, BN_THREADS, 0, st>>>(partial, chunks, C, sum_g, sum_gz);
  rc = check_launch("bn_bwd_finalize_kernel");
  if (rc) return rc;
  const long long total = rows * (C / 4);
  bn_bwd_apply_kernel<<<(unsigned)((total + BN_THREADS - 1) / BN_THREADS), BN_THREADS, 0, st>>>(
      g, mask_src, z, rows, C, ld, mean, rstd, scale, sum_g, sum_gz, dz, gmasked);
  return check_launch("bn_bwd_apply_kernel");
}
