// Scalar-loss stage of the training step as device-side kernels (no host arithmetic, no .item()):
//
//   object_targets   centroid / scale / centred copy of the GT object points
//                    (AtlasLoss.compute_loss, mano_train/networks/branches/atlasbranch.py:211-227)
//   sq_terms         up to 8 mean-squared-error terms in one launch, plus their weighted sum
//                    (ManoLoss.compute_loss, manobranch.py:251-324: verts / joints / shape / pose regulariser;
//                     AtlasLoss: translation and scale terms, atlasbranch.py:211-227)
//   loss_combine     total = sum_k w[slot_k] * scale_k * sum(p_k) (+ sum(q_k)) over scalar or per-sample vector terms
//                    (the `final_loss = lambda * ... + ...` lines, atlasbranch.py:247-280, handnet.py:279-283,363-383)
//
// The loss weights live in a DEVICE vector read at execution time, so a captured CUDA graph follows
// HandNet.decay_regul (traineval.py:401-404) and any other change of a lambda without re-capture.
// All reductions use a fixed summation order (bit-reproducible run to run).
#include <string.h>

#include "common.cuh"

namespace obman {

constexpr int LOSS_MAX_TERMS = 8;
constexpr int SQ_CHUNKS = 32;        // partial sums per term
constexpr int COMBINE_MAX_TERMS = 12;
constexpr int COMBINE_GROUPS = 4;

// ---- GT object statistics --------------------------------------------------------------------------------------
// one CTA per sample: centroid = mean_i gt_i ; centred_i = gt_i - centroid ; scale = max_i |centred_i|
__global__ void __launch_bounds__(256)
object_targets_kernel(const float* __restrict__ gt, int M, float* __restrict__ centroid,
                      float* __restrict__ scale, float* __restrict__ centred) {
  __shared__ float scratch[32];
  __shared__ float cen[3];
  const int b = blockIdx.x;
  const float* __restrict__ g = gt + (size_t)b * M * 3;
  float s[3] = {0.f, 0.f, 0.f};
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    s[0] += g[3 * i]; s[1] += g[3 * i + 1]; s[2] += g[3 * i + 2];
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float t = block_sum(s[c], scratch);
    if (threadIdx.x == 0) cen[c] = t / (float)M;
  }
  __syncthreads();
  const float cx = cen[0], cy = cen[1], cz = cen[2];
  float m = 0.f;
  float* __restrict__ o = centred ? centred + (size_t)b * M * 3 : nullptr;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const float x = g[3 * i] - cx, y = g[3 * i + 1] - cy, z = g[3 * i + 2] - cz;
    if (o) { o[3 * i] = x; o[3 * i + 1] = y; o[3 * i + 2] = z; }
    m = fmaxf(m, x * x + y * y + z * z);
  }
  m = warp_max(m);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mm = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mm = fmaxf(mm, scratch[w]);
    if (centroid) { centroid[3 * b] = cx; centroid[3 * b + 1] = cy; centroid[3 * b + 2] = cz; }
    if (scale) scale[b] = sqrtf(mm);
  }
}

// ---- mean-squared-error terms ------------------------------------------------------------------------------------
struct SqTerms {
  const float* a[LOSS_MAX_TERMS];
  const float* b[LOSS_MAX_TERMS];   // nullable: compare with zero
  float* ga[LOSS_MAX_TERMS];        // backward only
  int rows[LOSS_MAX_TERMS], width[LOSS_MAX_TERMS], col0[LOSS_MAX_TERMS], col1[LOSS_MAX_TERMS];
  int slot[LOSS_MAX_TERMS];         // index into the device weight vector
  int n;
};

// grid (SQ_CHUNKS, n_terms).  term k = mean over rows x [col0, col1) of (a - b)^2; wsum = sum_k w[slot_k] * term_k.
// partial: n_terms * SQ_CHUNKS floats; ticket: one int, zero before the first launch, left zero by the kernel.
__global__ void __launch_bounds__(256)
sq_terms_fwd_kernel(const SqTerms t, const float* __restrict__ weights, float* __restrict__ partial,
                    int* __restrict__ ticket, float* __restrict__ terms, float* __restrict__ wsum) {
  __shared__ float scratch[32];
  __shared__ int last;
  const int k = blockIdx.y;
  const int cols = t.col1[k] - t.col0[k];
  const long long total = (long long)t.rows[k] * cols;
  const float* __restrict__ a = t.a[k];
  const float* __restrict__ b = t.b[k];
  float s = 0.f;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)SQ_CHUNKS * blockDim.x) {
    const long long r = e / cols;
    const long long off = r * t.width[k] + t.col0[k] + (e - r * cols);
    const float d = a[off] - (b ? b[off] : 0.f);
    s = fmaf(d, d, s);
  }
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) {
    partial[k * SQ_CHUNKS + blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(ticket, 1) == (int)(gridDim.x * gridDim.y) - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // the last CTA to finish folds the partial sums in a fixed order
  if (threadIdx.x < 32) {
    float ws = 0.f;
    for (int kk = 0; kk < t.n; ++kk) {
      float v = __ldcg(partial + kk * SQ_CHUNKS + threadIdx.x);   // SQ_CHUNKS == 32: one lane per partial
      v = warp_sum(v);
      const float mean = v / ((float)t.rows[kk] * (float)(t.col1[kk] - t.col0[kk]));
      if (threadIdx.x == 0) terms[kk] = mean;
      ws = fmaf(weights[t.slot[kk]], mean, ws);
    }
    if (threadIdx.x == 0) {
      *wsum = ws;
      *ticket = 0;
    }
  }
}

// ga_k[r, c] = g * w[slot_k] * 2 (a - b) / count_k inside [col0, col1), 0 elsewhere in the row.  grid (SQ_CHUNKS * 4, n).
__global__ void __launch_bounds__(256)
sq_terms_bwd_kernel(const SqTerms t, const float* __restrict__ weights, const float* __restrict__ gwsum) {
  const int k = blockIdx.y;
  float* __restrict__ ga = t.ga[k];
  if (ga == nullptr) return;
  const int width = t.width[k];
  const long long total = (long long)t.rows[k] * width;
  const float* __restrict__ a = t.a[k];
  const float* __restrict__ b = t.b[k];
  const float coef = 2.f * gwsum[0] * weights[t.slot[k]] / ((float)t.rows[k] * (float)(t.col1[k] - t.col0[k]));
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % width);
    float v = 0.f;
    if (c >= t.col0[k] && c < t.col1[k]) v = coef * (a[e] - (b ? b[e] : 0.f));
    ga[e] = v;
  }
}

// ---- weighted total ------------------------------------------------------------------------------------------------
struct CombineTerms {
  const float* p[COMBINE_MAX_TERMS];
  const float* q[COMBINE_MAX_TERMS];   // nullable second vector of the same length
  int len[COMBINE_MAX_TERMS];
  float scale[COMBINE_MAX_TERMS];
  int slot[COMBINE_MAX_TERMS];         // device weight index
  int group[COMBINE_MAX_TERMS];        // 0 .. COMBINE_GROUPS-1
  int n;
};

// one CTA, one warp per term (fixed lane-strided order, then a warp tree): vals[k] = scale_k * (sum p_k + sum q_k);
// total = sum_k w[slot_k] * vals[k]; groups[g] = the same sum restricted to group g.
__global__ void __launch_bounds__(32 * COMBINE_MAX_TERMS)
loss_combine_fwd_kernel(const CombineTerms t, const float* __restrict__ weights, float* __restrict__ total,
                        float* __restrict__ aux /* [COMBINE_GROUPS + n] */) {
  __shared__ float vals[COMBINE_MAX_TERMS];
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (k < t.n) {
    float s = 0.f;
    for (int i = lane; i < t.len[k]; i += 32) s += t.p[k][i] + (t.q[k] ? t.q[k][i] : 0.f);
    s = warp_sum(s) * t.scale[k];
    if (lane == 0) vals[k] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f, grp[COMBINE_GROUPS];
#pragma unroll
    for (int g = 0; g < COMBINE_GROUPS; ++g) grp[g] = 0.f;
    for (int kk = 0; kk < t.n; ++kk) {
      const float w = weights[t.slot[kk]] * vals[kk];
      tot += w;
#pragma unroll
      for (int g = 0; g < COMBINE_GROUPS; ++g) if (t.group[kk] == g) grp[g] += w;
      aux[COMBINE_GROUPS + kk] = vals[kk];
    }
    *total = tot;
#pragma unroll
    for (int g = 0; g < COMBINE_GROUPS; ++g) aux[g] = grp[g];
  }
}

// gterm[k] = gtotal * w[slot_k] * scale_k : the gradient of every element of term k's vector(s)
__global__ void loss_combine_bwd_kernel(const CombineTerms t, const float* __restrict__ weights,
                                        const float* __restrict__ gtotal, float* __restrict__ gterm) {
  const int k = threadIdx.x;
  if (k < t.n) gterm[k] = gtotal[0] * weights[t.slot[k]] * t.scale[k];
}

// ---- per-sample similarity transform of a point cloud ------------------------------------------------------------
// out[b,n,:] = s[b] * v[b,n,:] + t[b,:]  (AtlasBranch.forward_inference, atlasbranch.py:133-138; s and / or t nullable)
__global__ void __launch_bounds__(256)
affine_points_fwd_kernel(const float* __restrict__ v, const float* __restrict__ s, const float* __restrict__ t, int N,
                         float* __restrict__ out) {
  const int b = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 3 * N) return;
  const float sc = s ? s[b] : 1.f;
  const float tr = t ? t[3 * b + e % 3] : 0.f;
  out[(size_t)b * 3 * N + e] = fmaf(sc, v[(size_t)b * 3 * N + e], tr);
}

// gv = s[b] * g ; gs[b] = sum_{n,c} v * g ; gt[b,c] = sum_n g[b,n,c] : one CTA per sample, fixed summation order
__global__ void __launch_bounds__(256)
affine_points_bwd_kernel(const float* __restrict__ g, const float* __restrict__ v, const float* __restrict__ s, int N,
                         float* __restrict__ gv, float* __restrict__ gs, float* __restrict__ gt) {
  __shared__ float scratch[32];
  const int b = blockIdx.x;
  const float sc = s ? s[b] : 1.f;
  const float* __restrict__ gb = g + (size_t)b * 3 * N;
  const float* __restrict__ vb = v + (size_t)b * 3 * N;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};   // x, y, z sums of g; sum of v * g
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float gg = gb[3 * n + c];
      acc[c] += gg;
      acc[3] = fmaf(vb[3 * n + c], gg, acc[3]);
      if (gv) gv[(size_t)b * 3 * N + 3 * n + c] = sc * gg;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float tot = block_sum(acc[k], scratch);
    if (threadIdx.x == 0) {
      if (k < 3) { if (gt) gt[3 * b + k] = tot; }
      else if (gs) gs[b] = tot;
    }
  }
}

static int fill_sq(SqTerms& t, const float* const* a, const float* const* b, float* const* ga, const int* rows,
                   const int* width, const int* col0, const int* col1, const int* slot, int n) {
  OBMAN_REQUIRE(n >= 1 && n <= LOSS_MAX_TERMS, "sq_terms: n_terms=%d out of [1,%d]", n, LOSS_MAX_TERMS);
  OBMAN_REQUIRE(a && rows && width && col0 && col1 && slot, "sq_terms: null table");
  t.n = n;
  for (int k = 0; k < n; ++k) {
    OBMAN_REQUIRE(a[k] != nullptr, "sq_terms: term %d has no input", k);
    OBMAN_REQUIRE(rows[k] > 0 && width[k] > 0 && col0[k] >= 0 && col1[k] > col0[k] && col1[k] <= width[k],
                  "sq_terms: term %d has a bad shape (rows=%d width=%d cols=[%d,%d))", k, rows[k], width[k], col0[k], col1[k]);
    OBMAN_REQUIRE(slot[k] >= 0, "sq_terms: term %d has a negative weight slot", k);
    t.a[k] = a[k];
    t.b[k] = b ? b[k] : nullptr;
    t.ga[k] = ga ? ga[k] : nullptr;
    t.rows[k] = rows[k]; t.width[k] = width[k]; t.col0[k] = col0[k]; t.col1[k] = col1[k]; t.slot[k] = slot[k];
  }
  return OBMAN_OK;
}

static int fill_combine(CombineTerms& t, const float* const* p, const float* const* q, const int* len,
                        const float* scale, const int* slot, const int* group, int n) {
  OBMAN_REQUIRE(n >= 1 && n <= COMBINE_MAX_TERMS, "loss_combine: n_terms=%d out of [1,%d]", n, COMBINE_MAX_TERMS);
  OBMAN_REQUIRE(p && len && scale && slot && group, "loss_combine: null table");
  t.n = n;
  for (int k = 0; k < n; ++k) {
    OBMAN_REQUIRE(p[k] != nullptr && len[k] > 0, "loss_combine: term %d is empty", k);
    OBMAN_REQUIRE(slot[k] >= 0 && group[k] >= 0 && group[k] < COMBINE_GROUPS, "loss_combine: term %d: bad slot / group", k);
    t.p[k] = p[k];
    t.q[k] = q ? q[k] : nullptr;
    t.len[k] = len[k]; t.scale[k] = scale[k]; t.slot[k] = slot[k]; t.group[k] = group[k];
  }
  return OBMAN_OK;
}

}  // namespace obman

using namespace obman;

extern "C" int obman_object_targets(const float* gt, int B, int M, float* centroid, float* scale, float* centred,
                                    void* stream) {
  OBMAN_REQUIRE(gt && B > 0 && M > 0, "obman_object_targets: bad arguments (B=%d M=%d)", B, M);
  object_targets_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(gt, M, centroid, scale, centred);
  return check_launch("object_targets_kernel");
}

extern "C" int obman_sq_terms_fwd(const float* const* a, const float* const* b, const int* rows, const int* width,
                                  const int* col0, const int* col1, const int* slot, int n_terms,
                                  const float* weights, float* partial, int* ticket, float* terms, float* wsum,
                                  void* stream) {
  SqTerms t;
  memset(&t, 0, sizeof(t));
  int rc = fill_sq(t, a, b, nullptr, rows, width, col0, col1, slot, n_terms);
  if (rc) return rc;
  OBMAN_REQUIRE(weights && partial && ticket && terms && wsum, "obman_sq_terms_fwd: null argument");
  sq_terms_fwd_kernel<<<dim3(SQ_CHUNKS, n_terms), 256, 0, (cudaStream_t)stream>>>(t, weights, partial, ticket, terms, wsum);
  return check_launch("sq_terms_fwd_kernel");
}

extern "C" int obman_sq_terms_bwd(const float* const* a, const float* const* b, float* const* ga, const int* rows,
                                  const int* width, const int* col0, const int* col1, const int* slot, int n_terms,
                                  const float* weights, const float* gwsum, void* stream) {
  SqTerms t;
  memset(&t, 0, sizeof(t));
  int rc = fill_sq(t, a, b, ga, rows, width, col0, col1, slot, n_terms);
  if (rc) return rc;
  OBMAN_REQUIRE(weights && gwsum && ga, "obman_sq_terms_bwd: null argument");
  sq_terms_bwd_kernel<<<dim3(SQ_CHUNKS * 4, n_terms), 256, 0, (cudaStream_t)stream>>>(t, weights, gwsum);
  return check_launch("sq_terms_bwd_kernel");
}

extern "C" int obman_loss_combine_fwd(const float* const* p, const float* const* q, const int* len, const float* scale,
                                      const int* slot, const int* group, int n_terms, const float* weights,
                                      float* total, float* aux, void* stream) {
  CombineTerms t;
  memset(&t, 0, sizeof(t));
  int rc = fill_combine(t, p, q, len, scale, slot, group, n_terms);
  if (rc) return rc;
  OBMAN_REQUIRE(weights && total && aux, "obman_loss_combine_fwd: null argument");
  loss_combine_fwd_kernel<<<1, 32 * COMBINE_MAX_TERMS, 0, (cudaStream_t)stream>>>(t, weights, total, aux);
  return check_launch("loss_combine_fwd_kernel");
}

extern "C" int obman_loss_combine_bwd(const int* slot, const float* scale, int n_terms, const float* weights,
                                      const float* gtotal, float* gterm, void* stream) {
  OBMAN_REQUIRE(slot && scale && n_terms >= 1 && n_terms <= COMBINE_MAX_TERMS && weights && gtotal && gterm,
                "obman_loss_combine_bwd: bad arguments");
  CombineTerms t;
  memset(&t, 0, sizeof(t));
  t.n = n_terms;
  for (int k = 0; k < n_terms; ++k) { t.slot[k] = slot[k]; t.scale[k] = scale[k]; }
  loss_combine_bwd_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(t, weights, gtotal, gterm);
  return check_launch("loss_combine_bwd_kernel");
}

extern "C" int obman_affine_points_fwd(const float* v, const float* s, const float* t, int B, int N, float* out,
                                       void* stream) {
  OBMAN_REQUIRE(v && out && B > 0 && N > 0 && B <= 65535, "obman_affine_points_fwd: bad arguments");
  affine_points_fwd_kernel<<<dim3((3 * N + 255) / 256, B), 256, 0, (cudaStream_t)stream>>>(v, s, t, N, out);
  return check_launch("affine_points_fwd_kernel");
}

extern "C" int obman_affine_points_bwd(const float* g, const float* v, const float* s, int B, int N, float* gv,
                                       float* gs, float* gt, void* stream) {
  OBMAN_REQUIRE(g && v && B > 0 && N > 0, "obman_affine_points_bwd: bad arguments");
  affine_points_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(g, v, s, N, gv, gs, gt);
  return check_launch("affine_points_bwd_kernel");
}
