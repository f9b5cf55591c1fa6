// Bandwidth-bound helpers around the tensor-core convolutions of the ResNet-18 encoder
// (mano_train/networks/bases/resnet.py:154-188) and the parameter update:
//   stem_pack        NCHW image -> space-to-depth NHWC (16 channels, row padding) whose overlapping 64-channel view
//                    turns the 7x7/2 stem into a 4-tap shifted-box convolution (K = 256) on the same
//                    TMA path as every other conv
//   fold_conv        BatchNorm(eval) folding + OIHW -> (O, KH*KW*I) / (I, KH*KW*O) re-layout, once per step
//   maxpool 3x3/2    forward (with arg-max) and backward
//   meanpool         spatial mean forward; backward fused with the ReLU mask of the last block
//   colsum           per-channel sums of a (rows, C) tensor (BatchNorm beta / bias gradients)
//   bn_wgrad_finish  raw weight gradient -> gradients of conv weight (OIHW), gamma, beta
//   adam             fused Adam step on a flat parameter buffer (torch.optim.Adam semantics, traineval.py:113-116)
#include <cuda_bf16.h>
#include <string.h>

#include "common.cuh"

namespace obman {

__device__ __forceinline__ float to_tf32_rna_dev(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// x (B,3,H,W) NCHW -> out (B, H/2, W/2 + 4, 16): 2x2 space-to-depth with two zero pixels on either side of a row,
//   out[b, i, j + 2, (ph*2+pw)*3 + c] = x[b, c, 2i+ph, 2j+pw],  channels 12..15 = 0.
// The 7x7/2 stem reads it through an overlapping view (pixel stride 16 floats, 64 channels = pixels j-2 .. j+1 of
// the unpadded row): 4 vertical taps x 64 channels, K = 256, without materialising the 4x replicated tensor.
// One thread per output pixel: 12 image reads (consecutive threads = consecutive j: 8-byte stride per (c, ph) row
// segment, both pw of a segment are consumed by the same thread) and four float4 stores.
__global__ void __launch_bounds__(256)
stem_pack_kernel(const float* __restrict__ x, int B, int H, int W, float* __restrict__ out) {
  const int Ho = H / 2, Wo = W / 2, Wp = Wo + 4;
  const int jp = blockIdx.x * blockDim.x + threadIdx.x;   // padded column
  const int i = blockIdx.y, b = blockIdx.z;
  if (jp >= Wp) return;
  float4 v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) v[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int j = jp - 2;
  if (j >= 0 && j < Wo) {
    float ch[12];
#pragma unroll
    for (int ph = 0; ph < 2; ++ph)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float2 p2 = __ldg(reinterpret_cast<const float2*>(x + (((size_t)b * 3 + c) * H + (2 * i + ph)) * W) + j);
        ch[(ph * 2 + 0) * 3 + c] = p2.x;
        ch[(ph * 2 + 1) * 3 + c] = p2.y;
      }
    v[0] = make_float4(ch[0], ch[1], ch[2], ch[3]);
    v[1] = make_float4(ch[4], ch[5], ch[6], ch[7]);
    v[2] = make_float4(ch[8], ch[9], ch[10], ch[11]);
  }
  float4* dst = reinterpret_cast<float4*>(out + (((size_t)b * Ho + i) * Wp + jp) * 16);
#pragma unroll
  for (int e = 0; e < 4; ++e) dst[e] = v[e];
}

// Folded weights for one convolution.  w (O,I,KH,KW); scale s[o] = gamma*rsqrt(var+eps) (or 1 without BN).
// When wf_lo / wft_lo are given, wf / wft hold the tf32-rounded value and the *_lo arrays the residual
// (pre-split operands of the 3xTF32 A-in-TMEM GEMM path).
//   wf [o][(kh*KW+kw)*Ip + i] = s[o]*w[o,i,kh,kw]        (fprop B operand; Ip = padded input channels)
//   wft[i][(kh*KW+kw)*O  + o] = s[o]*w[o,i,kh,kw]        (dgrad B operand), i < I only
//   shift[o] = beta + (conv_bias - mean)*s ; scale[o] = s ; rstd[o]
// stem == 1: w is the (64,3,7,7) stem filter, written in the 4-tap x 64-channel layout of stem_pack_kernel.
// packed == 1 (3xBF16 path): wf / wft hold, per 32-element block of a row, 32 bf16 `hi` values followed by the
// 32 bf16 `lo` = bf16(v - hi) values (same bytes per row as fp32); wft rows are taps * Op long, Op = O rounded up
// to 32 (the caller zero-fills the padding).
__device__ __forceinline__ void store_packed_bf16(float* row, size_t k, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  __nv_bfloat16* r16 = reinterpret_cast<__nv_bfloat16*>(row);
  const size_t blk = k >> 5, in = k & 31;
  r16[blk * 64 + in] = h;
  r16[blk * 64 + 32 + in] = l;
}

__device__ __forceinline__ void
fold_conv_row(int o, const float* __restrict__ w, const float* __restrict__ cbias,
              const float* __restrict__ gamma,
              const float* __restrict__ beta, const float* __restrict__ mean,
              const float* __restrict__ var, float eps, int O, int I, int KH, int KW, int Ip,
              int stem, int packed, float* __restrict__ wf, float* __restrict__ wf_lo, float* __restrict__ wft,
              float* __restrict__ wft_lo, float* __restrict__ shift, float* __restrict__ scale,
              float* __restrict__ rstd_out) {
  float s = 1.f, rs = 1.f;
  if (gamma) {
    rs = rsqrtf(var[o] + eps);
    s = gamma[o] * rs;
  }
  if (threadIdx.x == 0) {
    scale[o] = s;
    rstd_out[o] = rs;
    const float cb = cbias ? cbias[o] : 0.f;
    shift[o] = gamma ? beta[o] + (cb - mean[o]) * s : cb;
  }
  if (stem) {
    // vertical tap a in [-2,1] -> slot a+2; channel q*16 + (ph*2+pw)*3 + c <-> kh = 2a+ph+3, kw = 2(q-2)+pw+3
    for (int k = threadIdx.x; k < 4 * 64; k += blockDim.x) {
      const int slot = k >> 6, chn = k & 63;
      float v = 0.f;
      if ((chn & 15) < 12) {
        const int a = slot - 2, q = chn >> 4, ch = chn & 15;
        const int ph = ch / 6, pw = (ch / 3) & 1, c = ch % 3;
        const int kh = 2 * a + ph + 3, kw = 2 * (q - 2) + pw + 3;
        if (kh >= 0 && kh < 7 && kw >= 0 && kw < 7) v = s * w[((o * 3 + c) * 7 + kh) * 7 + kw];
      }
      if (packed) { store_packed_bf16(wf + (size_t)o * 256, k, v); continue; }
      const float h = to_tf32_rna_dev(v);
      wf[(size_t)o * 256 + k] = wf_lo ? h : v;
      if (wf_lo) wf_lo[(size_t)o * 256 + k] = v - h;
    }
    return;
  }
  const int taps = KH * KW;
  const int Op = (O + 31) / 32 * 32;
  for (int k = threadIdx.x; k < taps * Ip; k += blockDim.x) {
    const int t = k / Ip, i = k - t * Ip;
    float v = 0.f;
    if (i < I) v = s * w[((size_t)o * I + i) * taps + t];
    if (packed) {
      store_packed_bf16(wf + (size_t)o * taps * Ip, k, v);
      if (i < I && wft) store_packed_bf16(wft + (size_t)i * taps * Op, (size_t)t * Op + o, v);
      continue;
    }
    const float h = to_tf32_rna_dev(v);
    if (i < I && wft) {
      const size_t ot = (size_t)i * taps * O + (size_t)t * O + o;
      wft[ot] = wft_lo ? h : v;
      if (wft_lo) wft_lo[ot] = v - h;
    }
    wf[(size_t)o * taps * Ip + k] = wf_lo ? h : v;
    if (wf_lo) wf_lo[(size_t)o * taps * Ip + k] = v - h;
  }
}

__global__ void __launch_bounds__(256)
fold_conv_kernel(const float* __restrict__ w, const float* __restrict__ cbias, const float* __restrict__ gamma,
                 const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ var, float eps,
                 int O, int I, int KH, int KW, int Ip, int stem, int packed, float* __restrict__ wf,
                 float* __restrict__ wf_lo, float* __restrict__ wft, float* __restrict__ wft_lo,
                 float* __restrict__ shift, float* __restrict__ scale, float* __restrict__ rstd_out) {
  fold_conv_row(blockIdx.x, w, cbias, gamma, beta, mean, var, eps, O, I, KH, KW, Ip, stem, packed, wf, wf_lo, wft, wft_lo,
                shift, scale, rstd_out);
}

// All conv+BN units of an encoder in ONE launch (was one launch per unit, 20 per encoder and step): block b works on
// output channel b - row0[u] of unit u, found through the prefix table.  Packed bf16 layout only.
constexpr int FOLD_MAX_UNITS = 24;
struct FoldBatch {
  const float* w[FOLD_MAX_UNITS];
  const float* gamma[FOLD_MAX_UNITS];
  const float* beta[FOLD_MAX_UNITS];
  const float* mean[FOLD_MAX_UNITS];
  const float* var[FOLD_MAX_UNITS];
  float* wf[FOLD_MAX_UNITS];
  float* wft[FOLD_MAX_UNITS];     // nullable
  float* shift[FOLD_MAX_UNITS];
  float* scale[FOLD_MAX_UNITS];
  float* rstd[FOLD_MAX_UNITS];
  int row0[FOLD_MAX_UNITS + 1];   // first block of every unit (prefix sums of O)
  int O[FOLD_MAX_UNITS], I[FOLD_MAX_UNITS], K[FOLD_MAX_UNITS], Ip[FOLD_MAX_UNITS], stem[FOLD_MAX_UNITS];
  int n;
};

__global__ void __launch_bounds__(256) fold_conv_batch_kernel(const __grid_constant__ FoldBatch t, float eps) {
  int u = 0;
  while (u + 1 < t.n && (int)blockIdx.x >= t.row0[u + 1]) ++u;
  fold_conv_row((int)blockIdx.x - t.row0[u], t.w[u], nullptr, t.gamma[u], t.beta[u], t.mean[u], t.var[u], eps, t.O[u],
                t.I[u], t.K[u], t.K[u], t.Ip[u], t.stem[u], 1, t.wf[u], nullptr, t.wft[u], nullptr, t.shift[u],
                t.scale[u], t.rstd[u]);
}

// 3x3 stride-2 pad-1 max pooling, NHWC; idx = winning window position 0..8
__global__ void __launch_bounds__(256)
maxpool_fwd_kernel(const float* __restrict__ x, int B, int H, int W, int C, float* __restrict__ out,
                   unsigned char* __restrict__ idx) {
  // grid (ceil(Wo*C/4 / 256), Ho, B): 32-bit index arithmetic only (the flat 64-bit div/mod version was
  // instruction bound at ~1/3 of the HBM rate)
  const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= Wo * C4) return;
  const int c4 = u % C4, j = u / C4;
  const int i = blockIdx.y, b = blockIdx.z;
  const size_t t = ((size_t)b * Ho + i) * (size_t)(Wo * C4) + u;
  float best[4] = {-3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f};
  int bi[4] = {0, 0, 0, 0};
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int h = 2 * i + dy - 1;
    if (h < 0 || h >= H) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int w = 2 * j + dx - 1;
      if (w < 0 || w >= W) continue;
      const float4 v = *reinterpret_cast<const float4*>(x + (((size_t)b * H + h) * W + w) * C + c4 * 4);
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (vv[e] > best[e]) { best[e] = vv[e]; bi[e] = dy * 3 + dx; }
    }
  }
  reinterpret_cast<float4*>(out)[t] = make_float4(best[0], best[1], best[2], best[3]);
  reinterpret_cast<uchar4*>(idx)[t] = make_uchar4(bi[0], bi[1], bi[2], bi[3]);
}

// gx[b,h,w,c] = sum over the (<= 4) windows that contain (h,w) and whose arg-max is (h,w)
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const float* __restrict__ gout, const unsigned char* __restrict__ idx, int B, int H,
                   int W, int C, float* __restrict__ gx) {
  // grid (ceil(W*C/4 / 256), H, B), see maxpool_fwd_kernel
  const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= W * C4) return;
  const int c4 = u % C4, w = u / C4;
  const int h = blockIdx.y, b = blockIdx.z;
  const size_t t = ((size_t)b * H + h) * (size_t)(W * C4) + u;
  // (two restructurings that put the loads of all candidate windows in flight together - unconditional with clamped
  // addresses, and predicated without a loop - were measured 13 % and 30 % SLOWER than this loop, gpurun r2w / r2aa)
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = (h + 1) / 2 - 1; i <= (h + 1) / 2; ++i) {
    if (i < 0 || i >= Ho) continue;
    const int dy = h - (2 * i - 1);
    if (dy < 0 || dy > 2) continue;
    for (int j = (w + 1) / 2 - 1; j <= (w + 1) / 2; ++j) {
      if (j < 0 || j >= Wo) continue;
      const int dx = w - (2 * j - 1);
      if (dx < 0 || dx > 2) continue;
      const size_t o = (((size_t)b * Ho + i) * Wo + j) * C4 + c4;
      const uchar4 k = reinterpret_cast<const uchar4*>(idx)[o];
      const float4 g = reinterpret_cast<const float4*>(gout)[o];
      const int me = dy * 3 + dx;
      if (k.x == me) acc[0] += g.x;
      if (k.y == me) acc[1] += g.y;
      if (k.z == me) acc[2] += g.z;
      if (k.w == me) acc[3] += g.w;
    }
  }
  reinterpret_cast<float4*>(gx)[t] = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// out[b,c] = mean over P pixels of x[b,p,c]
__global__ void __launch_bounds__(256)
meanpool_fwd_kernel(const float* __restrict__ x, int B, int P, int C, float* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B * C) return;
  const int b = t / C, c = t % C;
  float s = 0.f;
  for (int p = 0; p < P; ++p) s += x[((size_t)b * P + p) * C + c];
  out[t] = s / (float)P;
}

// gx[b,p,c] = x[b,p,c] > 0 ? gout[b,c] / P : 0     (mean backward fused with the ReLU mask of x)
__global__ void __launch_bounds__(256)
meanpool_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ x, int B, int P, int C,
                    float* __restrict__ gx) {
  const size_t total = (size_t)B * P * C;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int c = (int)(t % C);
  const int b = (int)(t / ((size_t)P * C));
  gx[t] = x[t] > 0.f ? gout[(size_t)b * C + c] / (float)P : 0.f;
}

// Vector variant (C % 4 == 0, ld % 4 == 0, 16-byte aligned x): every thread owns one float4 column group and walks
// rows with stride (256 / groups); four independent loads in flight per thread.  grid (1, row chunks).
template <int UNROLL>
__global__ void __launch_bounds__(256)
colsum_v4_kernel(const float* __restrict__ x, long long rows, int C, long long ld, long long rows_per_block,
                 float* __restrict__ out) {
  extern __shared__ float4 red4[];            // [rows_in_flight][groups]
  const int groups = C >> 2;                  // float4 column groups, <= 256
  const int rpb = 256 / groups;               // rows handled per block iteration
  const int g = threadIdx.x % groups, rr = threadIdx.x / groups;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rr < rpb) {
    long long r = r0 + rr;
    for (; r + (long long)(UNROLL - 1) * rpb < r1; r += (long long)UNROLL * rpb) {
      float4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(x + (r + (long long)u * rpb) * ld) + g);
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    for (; r < r1; r += rpb) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ld) + g);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    red4[rr * groups + g] = acc;
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < rpb; ++k) {
      const float4 v = red4[k * groups + threadIdx.x];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    float* o = out + 4 * threadIdx.x;
    atomicAdd(o, t.x); atomicAdd(o + 1, t.y); atomicAdd(o + 2, t.z); atomicAdd(o + 3, t.w);
  }
}

// out[c] (+)= sum_r x[r, c] ; x has row stride ld. grid (col tiles of 32, row chunks); out zero-initialised.
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, long long rows, int C, long long ld, long long rows_per_block,
              float* __restrict__ out) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = min(rows, r0 + rows_per_block);
  float s = 0.f;
  if (c < C)
    for (long long r = r0 + ry; r < r1; r += 8) s += x[r * ld + c];
  red[ry][threadIdx.x & 31] = s;
  __syncthreads();
  if (ry == 0 && c < C) {
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += red[k][threadIdx.x & 31];
    atomicAdd(out + c, tot);
  }
}

// dwraw [O][(kh*KW+kw)*Ip + i] -> gw (O,I,KH,KW) = s[o]*dwraw ; ggamma[o] = rstd*(sum_k w*dwraw - mean*gbeta) ; gbeta = colsum
// stem == 1: dwraw is in the 4-tap x 64-channel layout of stem_pack_kernel.
__global__ void __launch_bounds__(256)
bn_wgrad_finish_kernel(const float* __restrict__ dwraw, long long dw_ld, const float* __restrict__ w,
                       const float* __restrict__ cbias,
                       const float* __restrict__ scale, const float* __restrict__ rstd,
                       const float* __restrict__ mean, const float* __restrict__ gbeta_sum, int O, int I,
                       int KH, int KW, int Ip, int stem, float* __restrict__ gw,
                       float* __restrict__ ggamma, float* __restrict__ gbeta, float* __restrict__ gcbias) {
  __shared__ float scratch[32];
  const int o = blockIdx.x;
  const float s = scale[o];
  const int taps = KH * KW;
  float dot = 0.f;
  for (int k = threadIdx.x; k < I * taps; k += blockDim.x) {
    const int i = k / taps, t = k - i * taps;  // OIHW order of this output channel
    float raw;
    if (stem) {
      const int kh = t / 7, kw = t % 7;
      const int a = (kh - 3) >> 1, ph = (kh - 3) & 1, q = ((kw - 3) >> 1) + 2, pw = (kw - 3) & 1;
      raw = dwraw[(size_t)o * dw_ld + (a + 2) * 64 + q * 16 + (ph * 2 + pw) * 3 + i];
    } else {
      raw = dwraw[(size_t)o * dw_ld + (size_t)t * Ip + i];
    }
    const float wv = w[(size_t)o * I * taps + k];
    dot = fmaf(wv, raw, dot);
    gw[(size_t)o * I * taps + k] = s * raw;
  }
  dot = block_sum(dot, scratch);
  if (threadIdx.x == 0) {
    const float gb = gbeta_sum ? gbeta_sum[o] : 0.f;
    if (ggamma) {
      const float cb = cbias ? cbias[o] : 0.f;
      ggamma[o] = rstd[o] * (dot + (cb - mean[o]) * gb);
      gbeta[o] = gb;
    }
    if (gcbias) gcbias[o] = s * gb;
  }
}

// torch.optim.Adam (no amsgrad, no weight decay unless wd != 0), bias-corrected, on flat buffers.
// hyper_dev = {1-based step number, learning-rate multiplier}: both live in device memory so that a captured CUDA
// graph applies the right bias correction and follows an lr schedule (StepLR, traineval.py:179-182) without re-capture.
// One element of torch.optim.Adam's update, operation for operation as torch's kernels evaluate it in fp32:
//   grad = g * gscale (+ wd * p);  m = lerp(m, grad, 1 - beta1);  v = beta2 * v + (1 - beta2) * grad^2;
//   p -= step_size * m / (sqrt(v) / sqrt(bias_correction2) + eps)
// with the scalars (1 - beta, step_size = lr / bias_correction1, sqrt(bias_correction2)) formed in DOUBLE and rounded to
// fp32 once, as Python does for torch (1.f - 0.999f would already be off by 1.3e-5 relative).
__device__ __forceinline__ void adam_one(float& pv, float gv, float& mv, float& vv, float step_size, float omb1,
                                         float beta2, float omb2, float eps, float wd, float bc2_sqrt, float gscale) {
  float grad = gv * gscale;
  if (wd != 0.f) grad = fmaf(wd, pv, grad);
  mv = fmaf(omb1, grad - mv, mv);
  vv = fmaf(omb2 * grad, grad, beta2 * vv);
  const float denom = sqrtf(vv) / bc2_sqrt + eps;
  pv = pv - step_size * (mv / denom);
}

// One float4 per thread per array (28 B of HBM traffic per parameter: p, g, m, v read; p, m, v written).
// hyper_dev = {1-based step number, learning-rate multiplier}: read at execution time (CUDA-graph replay).
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, long long n, double lr, double beta1, double beta2, float eps, float wd,
            const float* __restrict__ hyper_dev, float gscale) {
  __shared__ float s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    const double step = (double)hyper_dev[0];
    s_step_size = (float)(lr * (double)hyper_dev[1] / (1.0 - pow(beta1, step)));
    s_bc2_sqrt = (float)sqrt(1.0 - pow(beta2, step));
  }
  __syncthreads();
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  const float omb1 = (float)(1.0 - beta1), omb2 = (float)(1.0 - beta2), b2 = (float)beta2;
  if (i4 + 4 <= n) {
    float4 pv = *reinterpret_cast<float4*>(p + i4);
    const float4 gv = *reinterpret_cast<const float4*>(g + i4);
    float4 mv = *reinterpret_cast<float4*>(m + i4);
    float4 vv = *reinterpret_cast<float4*>(v + i4);
    adam_one(pv.x, gv.x, mv.x, vv.x, step_size, omb1, b2, omb2, eps, wd, bc2_sqrt, gscale);
    adam_one(pv.y, gv.y, mv.y, vv.y, step_size, omb1, b2, omb2, eps, wd, bc2_sqrt, gscale);
    adam_one(pv.z, gv.z, mv.z, vv.z, step_size, omb1, b2, omb2, eps, wd, bc2_sqrt, gscale);
    adam_one(pv.w, gv.w, mv.w, vv.w, step_size, omb1, b2, omb2, eps, wd, bc2_sqrt, gscale);
    *reinterpret_cast<float4*>(p + i4) = pv;
    *reinterpret_cast<float4*>(m + i4) = mv;
    *reinterpret_cast<float4*>(v + i4) = vv;
  } else {
    for (long long i = i4; i < n; ++i)
      adam_one(p[i], g[i], m[i], v[i], step_size, omb1, b2, omb2, eps, wd, bc2_sqrt, gscale);
  }
}

// AtlasNet decoder, first layer after the algebraic split of conv1 (atlasutils.py:65-67 on the input of
// atlasbranch.py:117-131): h1[b,n,c] = relu(sum_k grid[n,k] * W1[c,k] + F[b,c]) for c < C, 0 for C <= c < ld.
// Wg = the three grid columns of the folded conv1 weights, compacted to (C, 4) rows {w0, w1, w2, 0} (one coalesced float4
// per channel; reading them out of the (C, 516) weight rows made the kernel 4x slower), the grid is batch-independent
// or per-sample (grid_bstride != 0), F = feat * (s*W[:, 3:])^T + shift.
constexpr int L1F_ROWS = 4;   // points per thread: the weights / feature part of a channel quad are loaded once for all of them

__global__ void __launch_bounds__(256)
pointmlp_l1_fwd_kernel(const float* __restrict__ grid, long long grid_bstride, const float4* __restrict__ Wg,
                       const float* __restrict__ F, int B, int N, int C, int ld, int relu,
                       float* __restrict__ out) {
  // grid (ceil(ceil(N/4) * ld/4 / 256), B): thread = (group of 4 consecutive points, 4 consecutive channels); 32-bit
  // index arithmetic per sample
  const int ld4 = ld / 4;
  const int groups = (N + L1F_ROWS - 1) / L1F_ROWS;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= groups * ld4) return;
  const int b = blockIdx.y;
  const int ng = t / ld4;
  const int c4 = t - ng * ld4;
  const int c = c4 * 4;
  const float* __restrict__ Fb = F + (size_t)b * C;
  float4 w[4];
  float f[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const bool ok = c + e < C;
    w[e] = ok ? __ldg(Wg + c + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    f[e] = ok ? __ldg(Fb + c + e) : 0.f;
  }
  const float* __restrict__ gp = grid + (size_t)b * grid_bstride;
  float4* __restrict__ ob = reinterpret_cast<float4*>(out) + (size_t)b * N * ld4 + c4;
#pragma unroll
  for (int j = 0; j < L1F_ROWS; ++j) {
    const int n = ng * L1F_ROWS + j;
    if (n >= N) break;
    const float gx = __ldg(gp + 3 * n), gy = __ldg(gp + 3 * n + 1), gz = __ldg(gp + 3 * n + 2);
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float r = fmaf(gz, w[e].z, fmaf(gy, w[e].y, fmaf(gx, w[e].x, f[e])));   // padding channels: all-zero operands
      if (relu) r = fmaxf(r, 0.f);
      v[e] = r;
    }
    ob[(size_t)n * ld4] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// Backward of the split first decoder layer, one pass over g = d loss / d (pre-activation of layer 1), row stride ld:
//   gF[b,c]   = sum_n g[b,n,c]                   gradient of the per-sample feature part F (and, summed over b, of the shift)
//   gW[b,k,c] = sum_n g[b,n,c] * grid[b?,n,k]    per-sample partial of the grid columns of the conv1 weight gradient
// grid (ceil(C/32), B); fixed summation order.  (Was two kernels that each read g: a per-sample column sum and a sum over
// the batch per point followed by a weighted column sum.)
__global__ void __launch_bounds__(256)
pointmlp_l1_bwd_kernel(const float* __restrict__ g, const float* __restrict__ grid, long long grid_bstride, int N, int C,
                       int ld, float* __restrict__ gF, float* __restrict__ gW) {
  // grid (ceil(ld / 128), B): a warp covers 128 consecutive columns (one float4 per lane) of one row, the 8 warps of the
  // CTA take every 8th row; the three grid coordinates of a row are one broadcast load per 4 columns.  Columns >= C of
  // the padded rows may hold anything: they are summed but never written.
  __shared__ float4 red[8][4][32];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * 128 + lane * 4;
  const int ry = threadIdx.x >> 5;
  const float* __restrict__ gp = grid + (size_t)b * grid_bstride;
  float4 acc[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < ld) {
    const float* __restrict__ gb = g + (size_t)b * N * ld + c;
#pragma unroll 4
    for (int n = ry; n < N; n += 8) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(gb + (size_t)n * ld));
      const float w0 = __ldg(gp + 3 * n + 0), w1 = __ldg(gp + 3 * n + 1), w2 = __ldg(gp + 3 * n + 2);
      acc[0].x += v.x; acc[0].y += v.y; acc[0].z += v.z; acc[0].w += v.w;
      acc[1].x = fmaf(v.x, w0, acc[1].x); acc[1].y = fmaf(v.y, w0, acc[1].y);
      acc[1].z = fmaf(v.z, w0, acc[1].z); acc[1].w = fmaf(v.w, w0, acc[1].w);
      acc[2].x = fmaf(v.x, w1, acc[2].x); acc[2].y = fmaf(v.y, w1, acc[2].y);
      acc[2].z = fmaf(v.z, w1, acc[2].z); acc[2].w = fmaf(v.w, w1, acc[2].w);
      acc[3].x = fmaf(v.x, w2, acc[3].x); acc[3].y = fmaf(v.y, w2, acc[3].y);
      acc[3].z = fmaf(v.z, w2, acc[3].z); acc[3].w = fmaf(v.w, w2, acc[3].w);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) red[ry][k][lane] = acc[k];
  __syncthreads();
  if (ry < 4) {
    float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int y = 0; y < 8; ++y) {
      const float4 t = red[y][ry][lane];
      tot.x += t.x; tot.y += t.y; tot.z += t.z; tot.w += t.w;
    }
    float* dst = ry == 0 ? gF + (size_t)b * C : gW + ((size_t)b * 3 + (ry - 1)) * C;
    const float tv[4] = {tot.x, tot.y, tot.z, tot.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) if (c + e < C) dst[c + e] = tv[e];
  }
}

// dw[c * ld_dw + k] = sum_b gW[b,k,c]
__global__ void __launch_bounds__(256)
pointmlp_l1_bwd_finish_kernel(const float* __restrict__ gW, int B, int C, float* __restrict__ dw, long long ld_dw) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * C) return;
  const int k = t / C, c = t - k * C;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += gW[((size_t)b * 3 + k) * C + c];
  dw[(size_t)c * ld_dw + k] = s;
}

// dst (rows, ld_dst) <- alpha * src[:, :C] (row stride ld_src), entries whose mask value is <= 0 zeroed (mask nullable, row
// stride ld_mask), columns C .. ld_dst-1 zero: pad / scale / ReLU-mask glue of the Linear and decoder backward passes in
// one launch (was fill + compare + multiply + strided copy).
__global__ void __launch_bounds__(256)
pad_scale_mask_kernel(const float* __restrict__ src, long long ld_src, const float* __restrict__ mask,
                      long long ld_mask, long long rows, int C, float alpha, float* __restrict__ dst, int ld_dst) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * ld_dst) return;
  const long long r = t / ld_dst;
  const int c = (int)(t - r * ld_dst);
  float v = 0.f;
  if (c < C) {
    v = alpha * src[r * ld_src + c];
    if (mask && !(mask[r * ld_mask + c] > 0.f)) v = 0.f;
  }
  dst[t] = v;
}

// out (K rows, ld_out) <- transpose of w (N, K; row stride ldw) in the packed bf16 hi|lo layout (see pack_bf16_kernel):
// the B operand of a Linear layer's data gradient.  One thread per pair of consecutive output columns (n, n+1).
__global__ void __launch_bounds__(256)
pack_bf16_t_kernel(const float* __restrict__ w, long long ldw, int N, int K, uint32_t* __restrict__ out,
                   long long ld_out) {
  const long long pairs_per_row = ld_out / 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pairs_per_row * K) return;
  const long long k = i / pairs_per_row;
  const int n = (int)(i - k * pairs_per_row) * 2;
  const float v0 = n < N ? w[(long long)n * ldw + k] : 0.f;
  const float v1 = n + 1 < N ? w[(long long)(n + 1) * ldw + k] : 0.f;
  const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
  const __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
  const uint32_t hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
  const uint32_t lo = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  const int blk = n >> 5, in = (n & 31) >> 1;
  out[k * ld_out + blk * 32 + in] = hi;
  out[k * ld_out + blk * 32 + 16 + in] = lo;
}

}  // namespace obman

using namespace obman;

extern "C" int obman_pointmlp_l1_fwd(const float* grid, long long grid_bstride, const float* Wg4,
                                     const float* F, int B, int N, int C, int ld, int relu, float* out, void* stream) {
  OBMAN_REQUIRE(grid && Wg4 && F && out && B > 0 && N > 0 && C > 0 && ld >= C && ld % 4 == 0 && (((uintptr_t)Wg4) & 15) == 0,
                "obman_pointmlp_l1_fwd: bad arguments");
  OBMAN_REQUIRE(B <= 65535 && (long long)N * (ld / 4) < 0x7fffffffLL, "obman_pointmlp_l1_fwd: batch / cloud too large for the grid");
  const unsigned per_sample = (unsigned)(((long long)((N + L1F_ROWS - 1) / L1F_ROWS) * (ld / 4) + 255) / 256);
  pointmlp_l1_fwd_kernel<<<dim3(per_sample, (unsigned)B), 256, 0, (cudaStream_t)stream>>>(
      grid, grid_bstride, reinterpret_cast<const float4*>(Wg4), F, B, N, C, ld, relu, out);
  return check_launch("pointmlp_l1_fwd_kernel");
}

extern "C" int obman_pad_scale_mask(const float* src, long long ld_src, const float* mask, long long ld_mask,
                                    long long rows, int C, float alpha, float* dst, int ld_dst, void* stream) {
  OBMAN_REQUIRE(src && dst && rows > 0 && C > 0 && ld_dst >= C && ld_src >= C && (!mask || ld_mask >= C),
                "obman_pad_scale_mask: bad arguments");
  const long long total = rows * ld_dst;
  pad_scale_mask_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, ld_src, mask, ld_mask,
                                                                                        rows, C, alpha, dst, ld_dst);
  return check_launch("pad_scale_mask_kernel");
}

extern "C" int obman_pack_bf16_t(const float* w, long long ldw, int N, int K, float* out, long long ld_out,
                                 void* stream) {
  OBMAN_REQUIRE(w && out && N > 0 && K > 0 && ldw >= K, "obman_pack_bf16_t: bad arguments");
  OBMAN_REQUIRE(ld_out % 32 == 0 && ld_out >= N, "obman_pack_bf16_t: ld_out must be a multiple of 32 >= N");
  const long long n = ld_out / 2 * K;
  pack_bf16_t_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      w, ldw, N, K, reinterpret_cast<uint32_t*>(out), ld_out);
  return check_launch("pack_bf16_t_kernel");
}

extern "C" int obman_pointmlp_l1_bwd(const float* g, const float* grid, long long grid_bstride, int B, int N, int C,
                                     int ld, float* gF, float* gW, float* dw, long long ld_dw, void* stream) {
  OBMAN_REQUIRE(g && grid && gF && gW && dw && B > 0 && N > 0 && C > 0 && ld >= C && ld_dw >= 3 && B <= 65535,
                "obman_pointmlp_l1_bwd: bad arguments");
  OBMAN_REQUIRE(ld % 4 == 0 && ((uintptr_t)g & 15) == 0, "obman_pointmlp_l1_bwd: g must be 16-byte aligned, ld a multiple of 4");
  cudaStream_t st = (cudaStream_t)stream;
  pointmlp_l1_bwd_kernel<<<dim3((ld + 127) / 128, B), 256, 0, st>>>(g, grid, grid_bstride, N, C, ld, gF, gW);
  int rc = check_launch("pointmlp_l1_bwd_kernel");
  if (rc) return rc;
  pointmlp_l1_bwd_finish_kernel<<<(3 * C + 255) / 256, 256, 0, st>>>(gW, B, C, dw, ld_dw);
  return check_launch("pointmlp_l1_bwd_finish_kernel");
}

extern "C" int obman_stem_pack(const float* x, int B, int H, int W, float* out, void* stream) {
  OBMAN_REQUIRE(x && out && B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "obman_stem_pack: bad arguments");
  OBMAN_REQUIRE(((uintptr_t)x & 7) == 0 && ((uintptr_t)out & 15) == 0, "obman_stem_pack: misaligned pointers");
  OBMAN_REQUIRE(B <= 65535 && H / 2 <= 65535, "obman_stem_pack: batch / height too large for the grid");
  const int Wp = W / 2 + 4;
  dim3 grid((unsigned)((Wp + 127) / 128), (unsigned)(H / 2), (unsigned)B);
  stem_pack_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, B, H, W, out);
  return check_launch("stem_pack_kernel");
}

extern "C" int obman_fold_conv(const float* w, const float* cbias, const float* gamma, const float* beta, const float* mean,
                               const float* var, float eps, int O, int I, int KH, int KW, int Ip, int stem,
                               int packed, float* wf, float* wf_lo, float* wft, float* wft_lo, float* shift,
                               float* scale, float* rstd, void* stream) {
  OBMAN_REQUIRE(w && wf && shift && scale && rstd && O > 0 && I > 0 && KH > 0 && KW > 0 && Ip >= I,
                "obman_fold_conv: bad arguments");
  OBMAN_REQUIRE(!stem || (I == 3 && KH == 7 && KW == 7), "obman_fold_conv: stem layout needs a (O,3,7,7) filter");
  OBMAN_REQUIRE(!wft_lo || wft, "obman_fold_conv: wft_lo without wft");
  OBMAN_REQUIRE(!packed || (Ip % 32 == 0 && !wf_lo && !wft_lo),
                "obman_fold_conv: packed bf16 layout needs Ip %% 32 == 0 and no *_lo outputs");
  fold_conv_kernel<<<O, 256, 0, (cudaStream_t)stream>>>(w, cbias, gamma, beta, mean, var, eps, O, I, KH, KW, Ip, stem,
                                                        packed, wf, wf_lo, wft, wft_lo, shift, scale, rstd);
  return check_launch("fold_conv_kernel");
}

extern "C" int obman_fold_conv_batch(int n_units, const float* const* w, const float* const* gamma,
                                     const float* const* beta, const float* const* mean, const float* const* var,
                                     float eps, const int* O, const int* I, const int* K, const int* Ip, const int* stem,
                                     float* const* wf, float* const* wft, float* const* shift, float* const* scale,
                                     float* const* rstd, void* stream) {
  OBMAN_REQUIRE(n_units >= 1 && n_units <= FOLD_MAX_UNITS, "obman_fold_conv_batch: n_units=%d out of [1,%d]", n_units,
                FOLD_MAX_UNITS);
  OBMAN_REQUIRE(w && gamma && beta && mean && var && O && I && K && Ip && stem && wf && wft && shift && scale && rstd,
                "obman_fold_conv_batch: null table");
  FoldBatch t;
  memset(&t, 0, sizeof(t));
  t.n = n_units;
  int rows = 0;
  for (int u = 0; u < n_units; ++u) {
    OBMAN_REQUIRE(w[u] && gamma[u] && beta[u] && mean[u] && var[u] && wf[u] && shift[u] && scale[u] && rstd[u],
                  "obman_fold_conv_batch: unit %d has a null pointer", u);
    OBMAN_REQUIRE(O[u] > 0 && I[u] > 0 && K[u] > 0 && Ip[u] >= I[u] && Ip[u] % 32 == 0,
                  "obman_fold_conv_batch: unit %d has bad sizes (packed layout needs Ip %% 32 == 0)", u);
    OBMAN_REQUIRE(!stem[u] || (I[u] == 3 && K[u] == 7), "obman_fold_conv_batch: stem layout needs a (O,3,7,7) filter");
    t.w[u] = w[u]; t.gamma[u] = gamma[u]; t.beta[u] = beta[u]; t.mean[u] = mean[u]; t.var[u] = var[u];
    t.wf[u] = wf[u]; t.wft[u] = wft[u]; t.shift[u] = shift[u]; t.scale[u] = scale[u]; t.rstd[u] = rstd[u];
    t.O[u] = O[u]; t.I[u] = I[u]; t.K[u] = K[u]; t.Ip[u] = Ip[u]; t.stem[u] = stem[u];
    t.row0[u] = rows;
    rows += O[u];
  }
  t.row0[n_units] = rows;
  fold_conv_batch_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(t, eps);
  return check_launch("fold_conv_batch_kernel");
}

extern "C" int obman_maxpool_fwd(const float* x, int B, int H, int W, int C, float* out,
                                 unsigned char* idx, void* stream) {
  OBMAN_REQUIRE(x && out && idx && B > 0 && H % 2 == 0 && W % 2 == 0 && C % 4 == 0, "obman_maxpool_fwd: bad arguments");
  OBMAN_REQUIRE(B <= 65535 && H / 2 <= 65535, "obman_maxpool_fwd: batch / height too large for the grid");
  dim3 grid((unsigned)(((W / 2) * (C / 4) + 255) / 256), (unsigned)(H / 2), (unsigned)B);
  maxpool_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, B, H, W, C, out, idx);
  return check_launch("maxpool_fwd_kernel");
}

extern "C" int obman_maxpool_bwd(const float* gout, const unsigned char* idx, int B, int H, int W, int C,
                                 float* gx, void* stream) {
  OBMAN_REQUIRE(gout && idx && gx && B > 0 && H % 2 == 0 && W % 2 == 0 && C % 4 == 0, "obman_maxpool_bwd: bad arguments");
  OBMAN_REQUIRE(B <= 65535 && H <= 65535, "obman_maxpool_bwd: batch / height too large for the grid");
  dim3 grid((unsigned)((W * (C / 4) + 255) / 256), (unsigned)H, (unsigned)B);
  maxpool_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gout, idx, B, H, W, C, gx);
  return check_launch("maxpool_bwd_kernel");
}

extern "C" int obman_meanpool_fwd(const float* x, int B, int P, int C, float* out, void* stream) {
  OBMAN_REQUIRE(x && out && B > 0 && P > 0 && C > 0, "obman_meanpool_fwd: bad arguments");
  meanpool_fwd_kernel<<<(B * C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x, B, P, C, out);
  return check_launch("meanpool_fwd_kernel");
}

extern "C" int obman_meanpool_bwd(const float* gout, const float* x, int B, int P, int C, float* gx,
                                  void* stream) {
  OBMAN_REQUIRE(gout && x && gx && B > 0 && P > 0 && C > 0, "obman_meanpool_bwd: bad arguments");
  const size_t total = (size_t)B * P * C;
  meanpool_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gout, x, B, P, C, gx);
  return check_launch("meanpool_bwd_kernel");
}

extern "C" int obman_colsum(const float* x, long long rows, int C, long long ld, float* out, void* stream) {
  OBMAN_REQUIRE(x && out && rows > 0 && C > 0 && ld >= C, "obman_colsum: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(out, 0, sizeof(float) * C, st);
  if (C % 4 == 0 && ld % 4 == 0 && C <= 1024 && (((uintptr_t)x) & 15) == 0) {
    const int groups = C / 4, rpb = 256 / groups;
    long long chunks = 4LL * num_sms();
    const long long min_rows = (long long)rpb * 8;   // at least 8 block iterations per chunk
    if (chunks > (rows + min_rows - 1) / min_rows) chunks = (rows + min_rows - 1) / min_rows;
    if (chunks < 1) chunks = 1;
    const long long per = (rows + chunks - 1) / chunks;
    colsum_v4_kernel<4><<<dim3(1, (unsigned)((rows + per - 1) / per)), 256, sizeof(float4) * rpb * groups, st>>>(
        x, rows, C, ld, per, out);
    return check_launch("colsum_v4_kernel");
  }
  const int ctiles = (C + 31) / 32;
  long long chunks = (4LL * num_sms() + ctiles - 1) / ctiles;
  if (chunks > (rows + 63) / 64) chunks = (rows + 63) / 64;
  if (chunks < 1) chunks = 1;
  const long long per = (rows + chunks - 1) / chunks;
  colsum_kernel<<<dim3(ctiles, (unsigned)((rows + per - 1) / per)), 256, 0, st>>>(x, rows, C, ld, per, out);
  return check_launch("colsum_kernel");
}

extern "C" int obman_bn_wgrad_finish(const float* dwraw, long long dw_ld, const float* w, const float* cbias,
                                     const float* scale,
                                     const float* rstd, const float* mean, const float* gbeta_sum, int O,
                                     int I, int KH, int KW, int Ip, int stem, float* gw, float* ggamma,
                                     float* gbeta, float* gcbias, void* stream) {
  OBMAN_REQUIRE(dwraw && w && scale && gw && O > 0 && I > 0 && Ip >= I, "obman_bn_wgrad_finish: bad arguments");
  OBMAN_REQUIRE(!ggamma || (rstd && mean && gbeta_sum && gbeta), "obman_bn_wgrad_finish: BatchNorm outputs need rstd/mean/gbeta_sum");
  OBMAN_REQUIRE(!gcbias || gbeta_sum, "obman_bn_wgrad_finish: conv-bias gradient needs gbeta_sum");
  bn_wgrad_finish_kernel<<<O, 256, 0, (cudaStream_t)stream>>>(dwraw, dw_ld, w, cbias, scale, rstd, mean, gbeta_sum,
                                                              O, I, KH, KW, Ip, stem, gw, ggamma, gbeta, gcbias);
  return check_launch("bn_wgrad_finish_kernel");
}

extern "C" int obman_adam_step(float* p, const float* g, float* m, float* v, long long n, double lr,
                               double beta1, double beta2, double eps, double weight_decay,
                               const float* hyper_dev, double grad_scale, void* stream) {
  OBMAN_REQUIRE(p && g && m && v && n > 0 && hyper_dev, "obman_adam_step: bad arguments");
  OBMAN_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0,
                "obman_adam_step: buffers must be 16-byte aligned");
  OBMAN_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0, "obman_adam_step: betas must be in [0, 1)");
  const long long n4 = (n + 3) / 4;
  adam_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      p, g, m, v, n, lr, beta1, beta2, (float)eps, (float)weight_decay, hyper_dev, (float)grad_scale);
  return check_launch("adam_kernel");
}
