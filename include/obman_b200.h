/* libobman_b200.so - C ABI of the B200-native obman_train hot path (sm_100a only).
 *
 * The reference (hassony2/obman_train) is pure Python: it has no FFI/plugin layer, every device op is
 * an ATen/cuDNN/cuBLAS call chain.  Each entry point below replaces one such chain and
 * is what a maintainer would bind from Python (ctypes stub in INTEGRATION.md).  Conventions:
 *   - plain C types only; every pointer is a DEVICE pointer owned by the caller (inputs, outputs and
 *     workspaces) EXCEPT the small descriptor tables that say so (convolution tap tables tap_dh / tap_dw /
 *     tap_phase / tap_wslot, the term tables of the scalar-loss stage): those are HOST arrays, read during the
 *     call and copied into the kernel's parameters; the library allocates nothing and keeps no reference after
 *     return;
 *   - `stream` is the caller's cudaStream_t (as void*); all work is enqueued asynchronously on it;
 *   - return 0 on success, a negative code on failure (OBMAN_ERR_*), message via obman_get_last_error();
 *   - tensors are dense row-major fp32 unless stated; point clouds are (B, P, 3) xyz-interleaved,
 *     geometry in millimetres, like the reference's public layouts.
 */
#ifndef OBMAN_B200_H_
#define OBMAN_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define OBMAN_OK 0
#define OBMAN_ERR_BAD_ARG -1
#define OBMAN_ERR_CUDA -2
#define OBMAN_ERR_UNSUPPORTED -3
#define OBMAN_ERR_DRIVER -4

#define OBMAN_PREC_TF32 1
#define OBMAN_PREC_3XBF16 2
#define OBMAN_PREC_3XTF32 3

int obman_version(void);
const char* obman_get_last_error(void);
/* 0 when the current device is a Blackwell sm_10x part. */
int obman_device_ok(void);

/* ---- Pairwise nearest neighbours --------------------------------------------------------------
 * Replaces batch_pairwise_dist + torch.min (mano_train/networks/branches/contactloss.py:60-79,164-166;
 * atlasutils.py:20-39) without materialising the (B,N,M) matrix.
 * x (B,N,3), y (B,M,3).  dirs bit0: minx/idxx (B,N) = nearest y for every x; bit1: miny/idxy (B,M). */
int obman_nn_fwd(const float* x, const float* y, int B, int N, int M, float* minx, int* idxx,
                 float* miny, int* idxy, int dirs, void* stream);

/* ChamferLoss.forward(preds, gts) -> (loss_1 (B), loss_2 (B))   atlasutils.py:11-18.
 * min1/idx1 (B,N): nearest gt of every pred; min2/idx2 (B,M): nearest pred of every gt (saved for bwd). */
int obman_chamfer_fwd(const float* preds, const float* gts, int B, int N, int M, float* loss1,
                      float* loss2, float* min1, int* idx1, float* min2, int* idx2, void* stream);
/* Autograd of the above: gloss1/gloss2 -> gpreds (B,N,3) and, if non-null, ggts (B,M,3).  g_stride 1: gloss1/gloss2
 * are (B) vectors; g_stride 0: each points to ONE float that applies to every sample (the gradient of
 * torch.mean(loss_1 + loss_2), atlasbranch.py:235,243, without materialising the expanded vector).  Scatter-free:
 * the inverse of the nearest-neighbour index is built per sample in shared memory, every output element is written
 * once in a fixed order (bit-reproducible); clouds beyond ~33 k points take a float-atomic fallback. */
int obman_chamfer_bwd(const float* preds, const float* gts, const int* idx1, const int* idx2,
                      const float* gloss1, const float* gloss2, int g_stride, int B, int N, int M, float* gpreds,
                      float* ggts, void* stream);

/* ---- Contact loss -------------------------------------------------------------------------------
 * batch_mesh_contains_points (mano_train/networks/branches/contactutils.py:62-159): hits (B,P) int32 =
 * number of triangles of the per-sample mesh (obj_verts (B,N,3), faces (F,3) int32 shared by the batch)
 * crossed by the fixed-direction ray from points (B,P,3); exterior <=> hits even.
 * tri_scratch (nullable): 32 * B * ceil(F/2) floats, 16-byte aligned; when given, the per-triangle records are built
 * once per sample there and streamed by the search kernel (faster); NULL = staged per CTA in shared memory. */
int obman_raycast_hits(const float* points, const float* obj_verts, const int* faces, int B, int P,
                       int N, int F, int* hits, float* tri_scratch, void* stream);
/* compute_contact_loss value/mask stage (contactloss.py:174-307).  modes: 0 dist_sq, 1 dist,
 * 2 dist_tanh; zones_mode: 0 all, 1 tips (zone_ids = tip ids, zone_ptr = {0,5}), 2 zones (CSR table).
 * Outputs: attr_mask/rep_mask (B,P) u8, close (B,P,3), anchor (B,P), partial (B,6) workspace,
 * out[6] = {missed_loss, penetr_loss, max_penetr, mean_penetr, n_attraction, n_repulsion}. */
int obman_contact_fwd(const float* hand, const float* obj, const float* mins21, const int* idx21,
                      const int* hits, const int* zone_ids, const int* zone_ptr, int n_zones,
                      int zones_mode, int B, int P, int N, float contact_thresh, int contact_mode,
                      float collision_thresh, int collision_mode, unsigned char* attr_mask,
                      unsigned char* rep_mask, float* close, float* anchor, float* partial,
                      float* out, void* stream);
/* Gradients of missed_loss / penetr_loss; target: 0 all, 1 obj, 2 hand (contact_target). */
int obman_contact_bwd(const float* hand, const float* close, const float* anchor, const int* idx21,
                      const unsigned char* attr_mask, const unsigned char* rep_mask,
                      const float* fwd_out, const float* g_missed, const float* g_penetr, int B,
                      int P, int N, float contact_thresh, int contact_mode, float collision_thresh,
                      int collision_mode, int target, float* ghand, float* gobj, void* stream);

/* ---- MANO layer -----------------------------------------------------------------------------------
 * manopth.manolayer.ManoLayer.forward (external; call site manobranch.py:170-182).  Tables:
 * v_template (V,3), shapedirs (V,3,10), posedirs (V,3,135), weights (V,16),
 * j_template (16,3) = J_regressor*v_template, j_shapedirs (16,3,10) = J_regressor*shapedirs,
 * hands_mean (45), comps (ncomps,45), default_betas (10).  pose (B,3+ncomps), betas (B,10) or NULL,
 * trans (B,3) or NULL (NULL => centre on joint center_idx, -1 for none).
 * Workspaces (floats): ws_pose_map B*135, ws_gp B*192, ws_tw B*48, ws_betas B*10.
 * Outputs: verts (B,V,3) mm, joints (B,21,3) mm. */
int obman_mano_fwd(const float* v_template, const float* shapedirs, const float* posedirs,
                   const float* weights, const float* j_template, const float* j_shapedirs,
                   const float* hands_mean, const float* comps, const float* default_betas, int V,
                   int ncomps, const float* pose, const float* betas, const float* trans, int B,
                   int side_left, int root_palm, int center_idx, float* ws_pose_map, float* ws_gp,
                   float* ws_tw, float* ws_betas, float* verts, float* joints, void* stream);
/* Backward: gverts (B,V,3) / gjoints (B,21,3) (either may be NULL) -> gpose (B,3+ncomps), gbetas (B,10)
 * (may be NULL).  Extra workspaces (floats): ws_gv B*V*3, ws_gtw B*48, ws_gacc B*337. */
int obman_mano_bwd(const float* v_template, const float* shapedirs, const float* posedirs,
                   const float* weights, const float* j_template, const float* j_shapedirs,
                   const float* hands_mean, const float* comps, const float* default_betas, int V,
                   int ncomps, const float* pose, const float* betas, int has_trans, int B,
                   int side_left, int root_palm, int center_idx, const float* ws_pose_map,
                   const float* ws_gp, const float* ws_betas, const float* gverts,
                   const float* gjoints, float* ws_gv, float* ws_gtw, float* ws_gacc, float* gpose,
                   float* gbetas, void* stream);

/* ---- Dense contractions on tcgen05 tensor cores (FP32 accumulation in TMEM, TMA feeds) ------------------
 * passes = 1: plain TF32; passes = 3: 3xTF32 split (hi*hi + lo*hi + hi*lo), fp32-equivalent accuracy;
 * passes = 2 (OBMAN_PREC_3XBF16): fp32 operands split into bf16 hi + bf16 lo (16 significant bits), the same
 * three products at the bf16 rate; W / w must then be in the packed layout of obman_pack_bf16 and W_lo NULL.
 *
 * obman_gemm: out[M,N] = epilogue(alpha * A[M,K] * W[N,K]^T); replaces nn.Linear / Conv1d(k=1) calls
 * (manobranch.py:124-147, atlasbranch.py:44-61, atlasutils.py:65-75).  Row-major, leading dimensions in
 * elements (lda, ldw multiples of 4; A, W 16-byte aligned).  epilogue: + bias[N] + addend[M,N] (ldo),
 * ReLU, zero where mask_src[M,N] (ldo) <= 0, then store or atomicAdd (accumulate).  W_lo (NULL or the
 * residual of pre-split weights, W then being the tf32-rounded part) selects the A-in-TMEM kernel. */
int obman_gemm(const float* A, long long lda, const float* W, const float* W_lo, long long ldw, int M,
               int N, int K, float* out, long long ldo, const float* bias, const float* addend,
               const float* mask_src, float alpha, int relu, int accumulate, int passes, void* stream);
/* hi = tf32-rounded w, lo = w - hi (n floats): weights pre-split for the 3xTF32 path whose A operand is
 * staged in tensor memory (pass them as W / W_lo, w / w_lo). */
int obman_split_tf32(const float* w, long long n, float* hi, float* lo, void* stream);
/* Diagnostics: while buf != NULL every tensor-core kernel launch writes 16 clock64 stamps per CTA into
 * buf[cta*16 + k] (0 entry, 1 setup done, 2 first / 3 last TMA issued, 4 first operands ready, 5 last MMA issued,
 * 6 accumulator complete, 7 exit | smid << 48; 8..15 inner phases of one main-loop iteration, see
 * scripts/trace_kernels.py); cap = capacity in 8-byte entries.  NULL switches it off. */
int obman_debug_trace(long long* buf, long long cap);
/* Packed weights of the 3xBF16 path: out (rows, ld_out), ld_out = K rounded up to 32; every 32-element block of
 * a row holds 32 bf16 hi values then 32 bf16 lo = bf16(w - hi) values (the bytes of 32 floats); zero padded. */
int obman_pack_bf16(const float* w, long long ldw, int rows, int K, float* out, long long ld_out,
                    void* stream);

/* obman_conv_nhwc: NHWC convolution as implicit GEMM; forward and data-gradient of nn.Conv2d
 * (mano_train/networks/bases/resnet.py:19-23,38-54,110-152) share it.  x (n_img,h_in,w_in,c_in),
 * c_in % 4 == 0; w (c_out, w_slots*c_in), tap t uses weight slot tap_wslot[t] (NULL: t) and reads input
 * pixel (h + tap_dh[t], w + tap_dw[t]) of view tap_phase[t] = ph*2+pw of x[:, ph::in_step, pw::in_step]
 * (zero outside the view).  x_sN / x_sH / x_sW: element strides of x (all 0 = dense NHWC; multiples of 4); the
 * pixel stride may be smaller than c_in, consecutive pixels then overlap (sliding window over a narrower tensor:
 * the stem reads 4 neighbouring 16-channel pixels of obman_stem_pack's output as one 64-channel pixel).
 * out is written at n*o_sN + h*o_sH + w*o_sW + c (elements) for h < h_out,
 * w < w_out; bias[c_out], addend / mask_src indexed like out (NULL to disable), relu flag. */
int obman_conv_nhwc(const float* x, int n_img, int h_in, int w_in, int c_in, int in_step,
                    long long x_sN, long long x_sH, long long x_sW, const float* w, const float* w_lo, int c_out, int w_slots, int num_taps,
                    const int* tap_dh, const int* tap_dw, const int* tap_phase, const int* tap_wslot,
                    float* out, int h_out, int w_out, long long o_sN, long long o_sH, long long o_sW,
                    const float* bias, const float* addend, const float* mask_src, int relu,
                    int passes, void* stream);

/* obman_wgrad_nhwc: dw[co, t*c_in + ci] = sum_{n,h,w} dy[n,h,w,co] * xview_t[n, h+dh[t], w+dw[t], ci];
 * weight gradient of the convolution above and (h = 1) of obman_gemm, as ONE GEMM with the taps stacked
 * along N.  c_out, c_in multiples of 32.  dw (c_out, num_taps*c_in) is overwritten.  x_sN / x_sH / x_sW as above.
 * dy_colsum (nullable, 3xBF16 path only): (c_out) receives sum_{n,h,w} dy[n,h,w,co] - the BatchNorm-beta / bias
 * gradient, accumulated by the threads that split dy anyway (no separate pass over dy). */
int obman_wgrad_nhwc(const float* dy, int n_img, int h_out, int w_out, int c_out, const float* x,
                     int h_in, int w_in, int c_in, int in_step, long long x_sN, long long x_sH,
                     long long x_sW, int num_taps, const int* tap_dh,
                     const int* tap_dw, const int* tap_phase, float* dw, float* dy_colsum, int passes,
                     void* stream);

/* ---- Encoder helpers (bandwidth-bound; mano_train/networks/bases/resnet.py:154-188) -----------------------
 * obman_stem_pack: x (B,3,H,W) NCHW -> out (B, H/2, W/2 + 4, 16) NHWC: 2x2 space-to-depth, channel
 * (ph*2+pw)*3 + c (12 real + 4 zero channels), two zero pixels of padding on either side of every row.  Read through
 * an overlapping view (pixel stride 16, 64 channels: obman_conv_nhwc's x_sW) the four horizontal taps of the 7x7/2
 * stem sit side by side, so the stem runs as a 4-tap (vertical), 64-channel convolution (K = 256) without ever
 * materialising the 4x replicated tensor. */
int obman_stem_pack(const float* x, int B, int H, int W, float* out, void* stream);
/* BatchNorm(eval) folding + weight re-layout, once per step.  w (O,I,KH,KW); gamma/beta/mean/var (O) or NULL
 * (no BN), cbias (O) conv bias or NULL: shift = beta + (cbias - mean)*scale (no BN: shift = cbias).
 * wf_lo / wft_lo (NULL or): pre-split mode, wf/wft = tf32-rounded value, *_lo = residual.
 * packed = 1: wf / wft in the obman_pack_bf16 layout instead (Ip % 32 == 0, *_lo NULL; wft rows are
 * KH*KW*Op long, Op = O rounded up to 32, padding zeroed by the caller).
 * wf (O, KH*KW*Ip) fprop operand, wft (I, KH*KW*O) dgrad operand
 * (NULL to skip), shift/scale/rstd (O).  stem=1: (O,3,7,7) filter -> (O, 4*64): slot = vertical tap, channel q*16 + (ph*2+pw)*3 + c. */
int obman_fold_conv(const float* w, const float* cbias, const float* gamma, const float* beta, const float* mean,
                    const float* var, float eps, int O, int I, int KH, int KW, int Ip, int stem,
                    int packed, float* wf, float* wf_lo, float* wft, float* wft_lo, float* shift,
                    float* scale, float* rstd, void* stream);
/* MaxPool2d(3, stride 2, pad 1), NHWC (resnet.py:107): idx (B,H/2,W/2,C) u8 = arg-max window slot. */
int obman_maxpool_fwd(const float* x, int B, int H, int W, int C, float* out, unsigned char* idx,
                      void* stream);
int obman_maxpool_bwd(const float* gout, const unsigned char* idx, int B, int H, int W, int C,
                      float* gx, void* stream);
/* x.mean(3).mean(2) (resnet.py:179) on (B,P,C); backward fused with the ReLU mask of x. */
int obman_meanpool_fwd(const float* x, int B, int P, int C, float* out, void* stream);
int obman_meanpool_bwd(const float* gout, const float* x, int B, int P, int C, float* gx,
                       void* stream);
/* obman_fold_conv (packed bf16 layout, BatchNorm present, square filters K x K) for up to 24 units in ONE launch: every
 * table is a HOST array of n_units entries (device pointers / sizes of unit u); wft[u] may be NULL. */
int obman_fold_conv_batch(int n_units, const float* const* w, const float* const* gamma, const float* const* beta,
                          const float* const* mean, const float* const* var, float eps, const int* O, const int* I,
                          const int* K, const int* Ip, const int* stem, float* const* wf, float* const* wft,
                          float* const* shift, float* const* scale, float* const* rstd, void* stream);
/* out[c] = sum_r x[r*ld + c]  (bias / BatchNorm-beta gradients). */
int obman_colsum(const float* x, long long rows, int C, long long ld, float* out, void* stream);
/* Raw weight gradient dwraw (O rows of stride dw_ld, KH*KW*Ip used) -> gw (O,I,KH,KW) = scale*dwraw and
 * (NULL to skip) ggamma = rstd*(sum_k w*dwraw + (cbias - mean)*gbeta_sum), gbeta = gbeta_sum,
 * gcbias = scale*gbeta_sum. */
int obman_bn_wgrad_finish(const float* dwraw, long long dw_ld, const float* w, const float* cbias,
                          const float* scale, const float* rstd, const float* mean,
                          const float* gbeta_sum, int O, int I, int KH, int KW, int Ip, int stem,
                          float* gw, float* ggamma, float* gbeta, float* gcbias, void* stream);
/* AtlasNet decoder layer 1 after the conv1 split (atlasbranch.py:117-131 + atlasutils.py:65-67):
 * out[b,n,c] = [relu](sum_{k<3} grid[b*grid_bstride + 3n + k] * Wg4[4c + k] + F[b*C + c]) (c < C), 0 (C <= c < ld);
 * Wg4 (C, 4) = the first three input channels (the grid point) of the BatchNorm-folded conv1 weights, zero-padded to 4
 * (16-byte aligned).  relu = 0 gives the pre-activation (BatchNorm with batch statistics normalises it afterwards). */
int obman_pointmlp_l1_fwd(const float* grid, long long grid_bstride, const float* Wg4, const float* F,
                          int B, int N, int C, int ld, int relu, float* out, void* stream);
/* dst (rows, ld_dst) = alpha * src[:, :C] where mask > 0 (mask NULL = everywhere), zero in columns C .. ld_dst-1:
 * pad / scale / ReLU-mask glue of the Linear and decoder backward passes in one launch. */
int obman_pad_scale_mask(const float* src, long long ld_src, const float* mask, long long ld_mask, long long rows,
                         int C, float alpha, float* dst, int ld_dst, void* stream);
/* out (K, ld_out) = transpose of w (N, K) in the obman_pack_bf16 layout (B operand of a Linear data gradient). */
int obman_pack_bf16_t(const float* w, long long ldw, int N, int K, float* out, long long ld_out, void* stream);
/* One pass over g (B,N, ld): gF[b,c] = sum_n g[b,n,c]; dw[c * ld_dw + k] = sum_{b,n} g[b,n,c] * grid[n,k], k < 3 (the grid
 * columns of the conv1 weight gradient).  grid (N,3), or (B,N,3) with grid_bstride = 3 N.  gW: scratch of 3 B C floats. */
int obman_pointmlp_l1_bwd(const float* g, const float* grid, long long grid_bstride, int B, int N, int C, int ld,
                          float* gF, float* gW, float* dw, long long ld_dw, void* stream);
/* Cotangent-Laplacian regulariser (laplacianloss.py:24-41: loss = mean_{b,i} ||(L V_b)_i||_2; L built once from the
 * unit icosphere, laplacianloss.py:100-131).  L is passed in ELL form: nbr (N,K) int32 neighbour ids, w (N,K) the
 * off-diagonal entries L_ij (rows padded with w = 0); the diagonal is -sum_j L_ij by construction.
 * V (B,N,3) -> Lx (B,N,3), loss (1); partial: workspace of ceil(B*N/256) floats.  Deterministic (no atomics). */
int obman_laplacian_fwd(const float* V, const int* nbr, const float* w, int B, int N, int K, float* Lx,
                        float* partial, float* loss, void* stream);
/* Backward (laplacianloss.py:137-150: L^T g = L g) fused with the gradient of the row norms: gloss (1) -> gV (B,N,3). */
int obman_laplacian_bwd(const float* Lx, const float* gloss, const int* nbr, const float* w, int B, int N,
                        int K, float* gV, void* stream);
/* edge_loss (atlasbranch.py:153-167): loss (1) = mean |e - mean_b(e)| over the 3F squared edge lengths of every
 * sample.  faces (F,3) int32; stats (B,3) = {mean, sum |dev|, sum sign(dev)} is kept for the backward. */
int obman_edge_loss_fwd(const float* V, const int* faces, int B, int N, int F, float* stats, float* loss,
                        void* stream);
/* vf (N,Kf) int32: ids of the faces incident to each vertex, -1 padded (vertex-centric gather, no atomics). */
int obman_edge_loss_bwd(const float* V, const int* faces, const int* vf, const float* stats,
                        const float* gloss, int B, int N, int F, int Kf, float* gV, void* stream);
/* Contact-IoU metric meshiou (contactloss.py:20-47): gt_dists / pred_dists (B,P) squared hand->object distances,
 * threshs = HOST array of n_thresh (<= 16) thresholds; iou_ws (B, n_thresh) workspace; batch_ious (n_thresh) = mean
 * over the batch of the per-sample IoU of the thresholded maps; auc[0] = their trapezoid over the thresholds. */
int obman_contact_iou(const float* gt_dists, const float* pred_dists, int B, int P, const float* threshs,
                      int n_thresh, float* iou_ws, float* batch_ious, float* auc, void* stream);

/* ---- BatchNorm with batch statistics (csrc/bn_train.cu) ---------------------------------------------------------
 * Training WITHOUT --freeze_batchnorm (model.train(), epochpass3d.py:48-52): torch.nn.BatchNorm2d / BatchNorm1d in
 * training mode around the convolutions of bases/resnet.py:25-54,154-171.  z = raw convolution output, (rows, C) with
 * row stride ld (NHWC flattened), C % 4 == 0 and <= 1024.  partial: workspace of 2 * C * obman_bn_chunks(rows, C) floats.
 *
 * obman_bn_stats: mean / rstd (= 1 / sqrt(biased var + eps)) of z per channel, scale = gamma * rstd, shift = beta -
 * mean * scale (gamma / beta nullable = 1 / 0), and when running_mean / running_var are given their update
 * r = (1 - momentum) * r + momentum * {mean, unbiased var}. */
int obman_bn_chunks(long long rows, int C);
int obman_bn_stats(const float* z, long long rows, int C, long long ld, const float* gamma, const float* beta,
                   float eps, float momentum, float* partial, float* mean, float* rstd, float* scale, float* shift,
                   float* running_mean, float* running_var, void* stream);
/* y = [relu](z * scale[c] + shift[c] [+ addend]) (addend: same layout as z, nullable). */
int obman_bn_apply_fwd(const float* z, long long rows, int C, long long ld, const float* scale, const float* shift,
                       const float* addend, int relu, float* y, void* stream);
/* Backward through y = bn(z) for the incoming gradient g' = g where mask_src > 0 (the ReLU that follows; mask_src
 * nullable): sum_g[c] = sum g' (d beta), sum_gz[c] = sum g' * zhat (d gamma), dz = scale * (g' - sum_g / rows -
 * zhat * sum_gz / rows); gmasked (nullable) receives g' (gradient of a residual branch added before the ReLU). */
int obman_bn_bwd(const float* g, const float* mask_src, const float* z, long long rows, int C, long long ld,
                 const float* mean, const float* rstd, const float* scale, float* partial, float* sum_g,
                 float* sum_gz, float* dz, float* gmasked, void* stream);

/* ---- Scalar-loss stage (csrc/loss_head.cu) -------------------------------------------------------------------
 * Tables (a, b, ga, p, q, rows, width, ... slot, group) are HOST arrays of n_terms entries whose pointer entries are
 * device pointers; they are copied into the kernel's parameters.  `weights` is a DEVICE float vector indexed by
 * `slot`: loss lambdas are read at execution time, so a captured CUDA graph follows HandNet.decay_regul
 * (traineval.py:401-404).  All reductions have a fixed order.
 *
 * GT object statistics of AtlasLoss.compute_loss (atlasbranch.py:211-227): gt (B,M,3) -> centroid (B,3) = mean_i gt,
 * scale (B) = max_i |gt_i - centroid|, centred (B,M,3) = gt - centroid (each output nullable). */
int obman_object_targets(const float* gt, int B, int M, float* centroid, float* scale, float* centred, void* stream);
/* Up to 8 torch mse_loss terms in one launch (ManoLoss.compute_loss manobranch.py:251-324; trans / scale terms
 * atlasbranch.py:211-227): terms[k] = mean over rows[k] x [col0[k], col1[k]) of (a_k - b_k)^2 on (rows, width) row-major
 * tensors (b_k NULL = compare with zero); wsum[0] = sum_k weights[slot[k]] * terms[k].  partial: 32 * n_terms floats;
 * ticket: one int that is zero before the first call (the kernel leaves it zero). */
int obman_sq_terms_fwd(const float* const* a, const float* const* b, const int* rows, const int* width,
                       const int* col0, const int* col1, const int* slot, int n_terms, const float* weights,
                       float* partial, int* ticket, float* terms, float* wsum, void* stream);
/* ga_k (rows, width) = gwsum[0] * weights[slot[k]] * d terms[k] / d a_k (zero outside the column range); ga_k NULL = skip. */
int obman_sq_terms_bwd(const float* const* a, const float* const* b, float* const* ga, const int* rows,
                       const int* width, const int* col0, const int* col1, const int* slot, int n_terms,
                       const float* weights, const float* gwsum, void* stream);
/* total[0] = sum_k weights[slot[k]] * vals[k], vals[k] = scale[k] * (sum_i p_k[i] + sum_i q_k[i]) over len[k] elements
 * (q_k nullable; len 1 = a scalar loss, len B = per-sample losses such as ChamferLoss's loss_1 / loss_2 with
 * scale = 1/B).  aux[0..3] = the weighted sum restricted to group 0..3 (e.g. contact_loss, handnet.py:363-367),
 * aux[4 + k] = vals[k].  At most 12 terms. */
int obman_loss_combine_fwd(const float* const* p, const float* const* q, const int* len, const float* scale,
                           const int* slot, const int* group, int n_terms, const float* weights, float* total,
                           float* aux, void* stream);
/* gterm[k] = gtotal[0] * weights[slot[k]] * scale[k]: the gradient of every element of term k. */
int obman_loss_combine_bwd(const int* slot, const float* scale, int n_terms, const float* weights,
                           const float* gtotal, float* gterm, void* stream);

/* Per-sample similarity transform of the predicted object (AtlasBranch.forward_inference, atlasbranch.py:133-138):
 * out[b,n,:] = s[b] * v[b,n,:] + t[b,:]; v, out (B,N,3), s (B) nullable (= 1), t (B,3) nullable (= 0). */
int obman_affine_points_fwd(const float* v, const float* s, const float* t, int B, int N, float* out, void* stream);
/* Its gradients (each output nullable): gv = s[b] * g, gs[b] = sum_{n,c} v * g, gt[b,c] = sum_n g[b,n,c]. */
int obman_affine_points_bwd(const float* g, const float* v, const float* s, int B, int N, float* gv, float* gs,
                            float* gt, void* stream);

/* Fused torch.optim.Adam step (traineval.py:113-116) on flat fp32 buffers (16-byte aligned); g is multiplied by
 * grad_scale.  hyper_dev: device float[2] = {1-based step number, learning-rate multiplier} - device memory so that a
 * captured CUDA graph keeps exact bias corrections and follows StepLR (traineval.py:179-182): lr_eff = lr * hyper_dev[1].
 * The scalar hyper-parameters are doubles: 1 - beta, lr / (1 - beta1^t) and sqrt(1 - beta2^t) are formed in double and
 * rounded to fp32 once, as torch.optim.Adam does (tests/test_gpu_adam.py: elementwise parity with torch). */
int obman_adam_step(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1,
                    double beta2, double eps, double weight_decay, const float* hyper_dev, double grad_scale,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OBMAN_B200_H_ */
