/* cofi_b200.h -- C ABI of libcofi_b200.so: hand-written sm_100a kernels for CoFiI2P's coarse-to-fine
 * correspondence hot path.
 *
 * The reference (WHU-USI3DV/CoFiI2P @ ed90edf) is pure PyTorch and has no FFI of its own; every entry
 * point below replaces the ATen call sequence of the cited reference lines (paths relative to the
 * reference repository).  Conventions:
 *   - every pointer is a DEVICE pointer unless marked host; the caller (PyTorch) allocates all outputs
 *     and workspaces; the library is stateless apart from a per-process cache of TMA descriptors;
 *   - every function returns 0 on success or a negative COFI_E* code and never throws;
 *     cofi_last_error() returns a thread-local message for the last failure;
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous on it and capturable in
 *     CUDA graphs (no allocation, no synchronisation inside);
 *   - matrices are row-major fp32 with an explicit leading dimension (elements); index tables are int64
 *     exactly as the reference's dataset produces them (model/kpconv/preprocess_data.py:82-99);
 *   - "frames": B independent frames stacked along the row axis with equal row counts per frame
 *     (the reference is batch-1; B>1 is the batched entry point of BASELINE config 2).  Index tables hold
 *     frame-local indices; index == rows_per_frame of the source means "shadow neighbour" (zero feature,
 *     point at infinity), as in model/kpconv/kpconv.py:91,103.
 */
#ifndef COFI_B200_H
#define COFI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COFI_OK 0
#define COFI_EINVAL (-1)  /* bad argument (shape, alignment, null pointer) */
#define COFI_ECUDA (-2)   /* CUDA runtime / driver error at launch */
#define COFI_EUNSUPPORTED (-3)

/* activation codes shared by the epilogues */
#define COFI_ACT_NONE 0
#define COFI_ACT_RELU 1
#define COFI_ACT_LRELU01 2 /* LeakyReLU(0.1), model/kpconv/modules.py:82,149,218 */
#define COFI_ACT_SIGMOID 3

/* GEMM engines */
#define COFI_GEMM_FP32 0   /* SIMT fp32 FMA: exact-order parity engine */
#define COFI_GEMM_TF32 1   /* tcgen05 kind::tf32, fp32 operands read by TMA, fp32 accumulate in TMEM */
#define COFI_GEMM_TF32X3 2 /* tcgen05 3xTF32 split (hi*hi + hi*lo + lo*hi): fp32-grade accuracy on tensor cores */
#define COFI_GEMM_TF32X3S 3 /* 3xTF32 with PRE-SPLIT weights: W points to the [2][N][K] output of cofi_split_tf32 (ldw == K);
                             * persistent kernel, A split in registers into tensor memory (csrc/gemm_x3.cu) */

int cofi_version(void);
const char* cofi_last_error(void);
/* number of kernel launches issued by this library since process start (bench.py's gpu_launches) */
int64_t cofi_launch_count(void);

/* ---------------------------------------------------------------------------------------------------
 * Point stream (model/kpconv)
 * ------------------------------------------------------------------------------------------------- */

/* packed[i] = (x, y, z, flag) with flag = 1.0 if sum_c feats[i,c] > 0 else 0.0.
 * Pre-pass of KPConv's neighbour count (model/kpconv/kpconv.py:113-114) folded into the coordinate table
 * so the neighbour gather is a single 16-byte load.  rows = total rows (all frames). */
int cofi_pack_points(const float* points, const float* feats, int64_t ldf, int C, int64_t rows,
                     float* packed /* [rows,4] */, void* stream);

/* Fused neighbour gather + kernel-point influence + aggregation of rigid KPConv
 * (model/kpconv/kpconv.py:91-105 and :113-115):
 *   w[m,h,k]   = max(0, 1 - |(s[nbr[m,h]] - q[m]) - kp[k]| / sigma)
 *   agg[m,k,c] = sum_h w[m,h,k] * feats[nbr[m,h], c]          (written as [M, K*C] row-major)
 *   cnt[m]     = max(1, #{h : sum_c feats[nbr[m,h],c] > 0})   (as float)
 * The weight application agg[M,K*C] x W[K*C,Cout] (:108-110), the division by cnt (:116) and the bias
 * (:119-120) are one cofi_gemm call with `rowdiv = cnt`.  No tensor cores here: influences are sparse,
 * zero weights are skipped exactly. */
int cofi_kpconv_aggregate(const float* feats, int64_t ldf, int C,
                          const float* s_packed /* [frames*Ns,4] from cofi_pack_points */,
                          const float* q_points /* [frames*Mq,3] */,
                          const int64_t* nbr /* [frames*Mq,H] */, int H,
                          int64_t Mq, int64_t Ns, int frames,
                          const float* kernel_points /* [K,3] */, int K, float sigma,
                          float kp_reach /* max_k |kernel_points[k]| (host-computed); neighbours beyond
                                            kp_reach + sigma are culled exactly; <= 0 disables */,
                          float* agg /* [frames*Mq, K*C] */, float* cnt /* [frames*Mq] */, void* stream);

/* Same, with the aggregate written as fp16 [frames*Mq, K*C] (K*C % 8 == 0).  Used by the tf32 engine: fp16 carries the
 * same 11 significand bits the tensor core keeps of a tf32 operand, halves the [M,15C] round trip through HBM and feeds
 * cofi_gemm_f16 at twice the MMA rate. */
int cofi_kpconv_aggregate_f16(const float* feats, int64_t ldf, int C, const float* s_packed, const float* q_points,
                              const int64_t* nbr, int H, int64_t Mq, int64_t Ns, int frames,
                              const float* kernel_points, int K, float sigma, float kp_reach,
                              void* agg_f16, float* cnt, void* stream);

/* out[m,c] = max_h x[nbr[m,h], c] with shadow rows = 0 (model/kpconv/functional.py:53-66). */
int cofi_maxpool_rows(const float* x, int64_t ldx, int C, const int64_t* nbr, int H,
                      int64_t Mq, int64_t Ns, int frames, float* out, int64_t ldo, void* stream);

/* Same on an fp16 copy of x (cofi_cast_f16): exact max of the fp16-rounded rows (rounding is monotonic), half the
 * gather bytes; fp32 output. tf32 engine only. C % 8 == 0. */
int cofi_maxpool_rows_f16(const void* x_f16, int64_t ldx, int C, const int64_t* nbr, int H, int64_t Mq, int64_t Ns,
                          int frames, float* out, int64_t ldo, void* stream);

/* out[i, 0:C] = x[idx[i*idx_stride], 0:C] (shadow -> 0), or x[i] when idx == NULL.
 * nearest_upsample reads only column 0 of its [N,128] table (model/kpconv/functional.py:18-20):
 * idx_stride = 128.  `out` may point into a wider concat buffer (ldo) -> torch.cat of
 * model/kpconv/kp_backbone.py:112,117,122 costs no extra pass. */
int cofi_gather_rows(const float* x, int64_t ldx, int C, const int64_t* idx, int64_t idx_stride,
                     int64_t Mq, int64_t Ns, int frames, float* out, int64_t ldo, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Pyramid index tables (the step immediately before the hot path)
 * ------------------------------------------------------------------------------------------------- */

/* distance used for ranking */
#define COFI_KNN_DIRECT 0   /* ((dx*dx + dy*dy) + dz*dz), rounded fp32 ops, no FMA: the true squared distance open3d's
                               KNNSearch ranks by (model/kpconv/preprocess_data.py:75-99) */
#define COFI_KNN_EXPANDED 1 /* ((-2 q.s + |q|^2) + |s|^2) clamped at 1e-12: `knn()` / `square_distance()` of
                               model/kpconv/preprocess_data.py:110-143 (precompute_point_cloud_cuda) */
#define COFI_KNN_NOCULL 0x100 /* OR-ed into `mode`: open every tile (brute force; same result, test/debug only) */
#define COFI_KNN_NOFAST 0x200 /* OR-ed into `mode`: skip the threshold-selection fast path of the query kernel and run the
                               * iterative sort/merge path for every query (same result, test/debug only) */

/* Random half-sampling of the pyramid on the device (model/kpconv/preprocess_data.py:52-68: level l+1 = level l at
 * n/2 indices drawn WITH replacement).  The draw is counter-based so that the host oracle restates it exactly:
 * u = Philox4x32-10(counter (j, frame, level, 0), key seed)[0], index = (u * n_{level-1}) >> 32 for output row j.
 * pts0 [frames*n0, 3]; out_levels / out_index: HOST arrays of `levels` device pointers (entry 0 unused; out_index or its
 * entries may be NULL): out_levels[l] [frames*(n0>>l), 3], out_index[l] [frames*(n0>>l)] = frame-local level-0 row copied.
 * One launch for all levels and frames. */
int cofi_half_sample_pyramid(const float* pts0, int64_t n0, int frames, int levels, uint64_t seed,
                             float* const* out_levels, int64_t* const* out_index, void* stream);

/* Exact k-nearest-neighbour tables of a whole point pyramid, all frames and all tables in four launches
 * (Morton sort in two steps, then two warp-per-query passes: the same-level and up-sampling tables first, then the sub-sampling
 * tables, whose queries copy the row of their twin in the finished same-level table when they coincide with a source point --
 * always the case in a pyramid built by sub-sampling -- and are searched otherwise).  Replaces the 13 KNNSearch / knn() calls of
 * model/kpconv/preprocess_data.py:75-99 (stack mode) and :172-190 (cuda mode).
 *   points[l]       device [frames*n[l], 3] fp32           (host array of `levels` device pointers)
 *   neighbors[l]    device [frames*n[l],   k] int64: level l   looks up level l     (levels entries)
 *   subsampling[l]  device [frames*n[l+1], k] int64: level l+1 looks up level l     (levels-1 entries)
 *   upsampling[l]   device [frames*n[l], k_up] int64: level l looks up level l+1   (levels-1 entries)
 * Rows are ascending in (distance, index): ties go to the lower index, the query itself comes first when it is in the
 * source set.  Indices are frame-local; when a source level has fewer than k points the tail holds n (the shadow
 * index of model/kpconv/kpconv.py:91).  Any of the three table arrays, or any entry, may be NULL (skipped).
 * k <= 128, levels <= 8, n[l] <= 2^20.  workspace: cofi_knn_pyramid_workspace() bytes, 256-byte aligned. */
int64_t cofi_knn_pyramid_workspace(const int64_t* n_per_level /* host */, int levels, int frames);
int cofi_knn_pyramid(const float* const* points /* host array */, const int64_t* n_per_level /* host */, int levels,
                     int frames, int k, int k_up /* columns of the upsampling tables: the model reads only column 0
                     (model/kpconv/functional.py:20), k_up = 1 builds just that; k_up = k gives the reference's tables */,
                     int mode, int64_t* const* neighbors, int64_t* const* subsampling, int64_t* const* upsampling,
                     void* workspace, void* stream);

/* One table: out[frames*nq, k] = the k nearest of src[frames*ns,3] for every row of qry[frames*nq,3]
 * (`knn(nodes, points, k)`, model/kpconv/preprocess_data.py:131-143; KNNSearch()(src, qry, k), :82). */
int64_t cofi_knn_table_workspace(int64_t ns, int64_t nq, int frames);
int cofi_knn_table(const float* src, int64_t ns, const float* qry, int64_t nq, int frames, int k, int mode,
                   int64_t* out, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Dense contractions
 * ------------------------------------------------------------------------------------------------- */

/* C[M,N] = act( (A[M,K] * W[N,K]^T) / rowdiv[m] + bias[n] + (accumulate ? C : 0) )
 * W is K-major, i.e. nn.Linear's weight as stored (model/kpconv/modules.py:78,108; transformer.py:26-36;
 * network.py:29); KPConv's [K*C,Cout] weights are passed pre-transposed by the host.
 * rowdiv, bias may be NULL.  engine: COFI_GEMM_*.  K, lda, ldw must be multiples of 4. */
int cofi_gemm(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc,
              int64_t M, int N, int K, const float* bias, const float* rowdiv, int accumulate, int act,
              int engine, void* stream);

/* out[2][N][K] (dense): plane 0 = W rounded to tf32, plane 1 = (W - plane 0) rounded to tf32 -- the weight operand of the
 * COFI_GEMM_TF32X3S engine (cofi_gemm, cofi_gemm_ln, cofi_gemm_colstats, cofi_conv2d_nhwc).  Weights are constants of
 * the forward pass (model/network.py:15-43), so the host splits them once per weights epoch. */
int cofi_split_tf32(const float* w, int64_t ldw, int N, int K, float* out, void* stream);

/* Perf triage of the COFI_GEMM_TF32X3S kernel (environment COFI_X3_PROFILE=1): 16 counters accumulated over all launches
 * since the last call -- clocks the TMA / MMA / epilogue / splitter roles spent waiting on each barrier, k-blocks, CTA
 * clocks (csrc/gemm_x3.cu).  Not part of the reference-facing surface. */
int cofi_debug_x3_profile(unsigned long long* out16);

/* fp16-operand variant of cofi_gemm on tcgen05 (kind::f16, fp32 accumulate and output): A [M,K] and W [N,K] are fp16,
 * K-major; K, lda, ldw multiples of 8; N >= 16. */
int cofi_gemm_f16(const void* A, int64_t lda, const void* W, int64_t ldw, float* C, int64_t ldc, int64_t M, int N, int K,
                  const float* bias, const float* rowdiv, int act, void* stream);

/* cofi_gemm / cofi_gemm_f16 whose epilogue ALSO writes, per 128-row tile, the column sums and sums of squares of the
 * stored output: stats[M/128, N, 2] (fp32).  The GroupNorm that follows every point-branch Linear / KPConv
 * (model/kpconv/modules.py:89-94,155-159,222-240) then needs no statistics pass: cofi_norm_rows_pre reduces the tiles
 * (fp64) and applies.  Tensor-core engines, M % 128 == 0, no activation. */
int cofi_gemm_colstats(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc, int64_t M, int N,
                       int K, const float* bias, const float* rowdiv, int engine, float* stats, void* stream);
/* cofi_gemm_colstats with C += (C is read and rewritten; the statistics describe the stored sum): the second half of a Linear
 * over a concatenated input.  The point-branch decoder (model/kpconv/kp_backbone.py:100-118) applies Linear(cat[up(x_c), x_f]);
 * a row gather commutes with a row-wise linear map, so W_c is applied at the COARSE resolution (half the rows), the result is
 * up-sampled into C and this call adds W_f x_f + bias: a third fewer flops and no concatenated buffer. */
int cofi_gemm_colstats_acc(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc, int64_t M, int N,
                           int K, const float* bias, int engine, float* stats, void* stream);
int cofi_gemm_f16_colstats(const void* A, int64_t lda, const void* W, int64_t ldw, float* C, int64_t ldc, int64_t M, int N,
                           int K, const float* bias, const float* rowdiv, float* stats, void* stream);

/* C[M,N] = act(LayerNorm_N(A W^T + bias) * gamma + beta) + residual : Linear -> LayerNorm(eps) -> act -> + residual as
 * ONE kernel when a row fits a tile (N <= 128, N % 32 == 0, tensor-core engines): each epilogue thread owns a full
 * output row in TMEM and normalises it in registers (model/transformer/transformer.py:57-58,61-64: merge+norm1 and
 * mlp[2]+norm2+residual).  Other shapes/engines run cofi_gemm followed by cofi_layer_norm_rows. */
int cofi_gemm_ln(const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc, int64_t M, int N, int K,
                 const float* bias, const float* gamma, const float* beta, float eps, int act, const float* residual,
                 int64_t ldr, int engine, void* stream);

/* NHWC convolution as implicit GEMM (model/imagenet.py:26-34,142-143,377-394):
 *   y[b,ho,wo,co] = sum_{kh,kw,ci} x[b, ho*stride+kh-pad, wo*stride+kw-pad, ci] * w[co, kh, kw, ci]
 * w is [Cout, KH*KW*Cin] (host repacks the reference's [Cout,Cin,KH,KW]).  Cin % 4 == 0.
 * Optional per-channel affine + residual + activation epilogue = eval-mode BatchNorm folded
 * (model/imagenet.py:398-411): y = act(conv*scale[co] + shift[co] + residual). */
int cofi_conv2d_nhwc(const float* x, int B, int H, int W, int Cin, const float* w, int Cout, int KH, int KW,
                     int stride, int pad, const float* scale, const float* shift, const float* residual,
                     int act, float* y, int engine, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Normalisations (all statistics in fp64)
 * ------------------------------------------------------------------------------------------------- */

/* Grouped normalisation over the rows of each frame: x is [frames*R, C]; statistics per (frame, group)
 * over R rows x (C/G) channels, biased variance, eps inside the sqrt.
 *   GroupNorm(32) over a cloud   (model/kpconv/modules.py:45-49)        G = 32, gamma/beta given
 *   affine-free InstanceNorm     (model/imagenet.py:123; network.py:42-43) G = C, gamma = beta = NULL
 *   train-mode BatchNorm         (model/imagenet.py:381-394)             frames = 1, G = C
 * y = act( (x-mean)*rstd*gamma + beta + residual ).  `partials` is a caller workspace of
 * cofi_norm_rows_workspace(frames, C) bytes.  If mean_out/var_out != NULL (size frames*G) the batch
 * statistics are also written (BatchNorm running-stat update is done by the host). */
int64_t cofi_norm_rows_workspace(int frames, int C);
/* Allocates the per-device election counters of the fused statistics/finalize kernel. Call once per device BEFORE a
 * CUDA-graph capture that contains cofi_norm_rows (the first cofi_norm_rows call does it implicitly otherwise). */
int cofi_norm_rows_init(void);
int cofi_norm_rows(const float* x, int64_t ldx, int64_t R, int C, int frames, int G, const float* gamma,
                   const float* beta, float eps, const float* residual, int64_t ldr, int act, float* y,
                   int64_t ldy, void* partials, float* mean_out, float* var_out, void* stream);

/* cofi_norm_rows with the statistics taken from cofi_gemm_colstats tiles (R % 128 == 0): finalize + apply only.
 * workspace: frames*G*8 bytes. */
int cofi_norm_rows_pre(const float* x, int64_t ldx, int64_t R, int C, int frames, int G, const float* gamma,
                       const float* beta, float eps, const float* residual, int64_t ldr, int act, float* y, int64_t ldy,
                       const float* tile_stats, void* workspace, void* stream);

/* y = act(x*scale[c] + shift[c] + residual): eval-mode BatchNorm as a per-channel affine. */
int cofi_affine_rows(const float* x, int64_t ldx, int64_t rows, int C, const float* scale, const float* shift,
                     const float* residual, int64_t ldr, int act, float* y, int64_t ldy, void* stream);

/* Row LayerNorm (eps 1e-5) + activation + optional residual added AFTER the norm:
 * y = act(LN(x)*gamma+beta) + residual  (model/network.py:29; model/transformer/transformer.py:58,62-64). */
int cofi_layer_norm_rows(const float* x, int64_t ldx, int64_t rows, int C, const float* gamma, const float* beta,
                         float eps, int act, const float* residual, int64_t ldr, float* y, int64_t ldy,
                         void* stream);

/* y = x / max(|x|_2, 1e-12) per row (+ add[row,:] if add != NULL): F.normalize(dim=channel)
 * (model/network.py:82,84,90,125,126,130) fused with the positional-encoding add (:113-114). */
int cofi_l2norm_rows(const float* x, int64_t ldx, int64_t rows, int C, const float* add, int64_t ldadd,
                     float* y, int64_t ldy, void* stream);

/* Same without `add`, also writing the fp16 copy of y that cofi_sim_argmin_exact streams through the tensor cores. */
int cofi_l2norm_rows_f16(const float* x, int64_t ldx, int64_t rows, int C, float* y, int64_t ldy, void* y_f16,
                         int64_t ldh, void* stream);

/* Column L2 normalisation over the L rows of each frame: F.normalize(q) with default dim=1 normalises Q
 * over the SEQUENCE axis (model/transformer/transformer.py:53).  In place allowed. `work` is a caller
 * workspace of cofi_colnorm_workspace(frames, C) bytes (8-byte aligned). */
int64_t cofi_colnorm_workspace(int frames, int C);
int cofi_colnorm_rows(const float* x, int64_t ldx, int64_t L, int C, int frames, void* work, float* y,
                      int64_t ldy, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Image stream helpers (model/imagenet.py)
 * ------------------------------------------------------------------------------------------------- */
int cofi_nchw_to_nhwc(const float* x, int B, int C, int H, int W, int Cpad, float* y, void* stream);
int cofi_nhwc_to_nchw(const float* x, int B, int H, int W, int C, float* y, void* stream);
/* MaxPool2d(3, stride 2, pad 1) (model/imagenet.py:145,204) */
int cofi_maxpool2d_3x3s2_nhwc(const float* x, int B, int H, int W, int C, float* y, void* stream);
/* y = cat(bilinear_x2(x1, align_corners=False), x2) along channels (model/imagenet.py:433,441-443) */
int cofi_upsample2x_cat_nhwc(const float* x1, int B, int H, int W, int C1, const float* x2, int C2, float* y,
                             void* stream);

/* ---------------------------------------------------------------------------------------------------
 * I2P transformer (model/transformer)
 * ------------------------------------------------------------------------------------------------- */

/* PositionEmbeddingCoordsSine (model/transformer/position_encoding.py:29-50): scale 2*pi, temperature 1e4,
 * interleaved sin/cos, zero pad to d_model. coords [rows, n_dim] fp32. Accurate sinf/cosf (arguments reach
 * hundreds of radians). */
int cofi_posenc_sine(const float* coords, int64_t rows, int n_dim, int d_model,
                     const float* dim_t /* [d_model/n_dim/2*2] = 1e4^(2*(i/2)/npf), built by the host with torch.pow */,
                     float* out, void* stream);

/* Multi-head softmax attention, out = softmax(scale * Q K^T) V per (frame, head)
 * (model/transformer/linear_attention.py:69-77).  q [frames*L, heads*D], k,v [frames*S, heads*D]. D == 32. */
int cofi_attention(const float* q, const float* k, const float* v, int64_t L, int64_t S, int frames, int heads,
                   int D, float scale, float* out, int engine, void* stream);
/* Same contraction on the tcgen05 engine (TMA + TMEM flash attention, tf32 operands, fp32 softmax).  The value
 * operand is passed K-major, vt = V^T [heads*D, frames*S]: the host obtains it for free from the v_proj GEMM with
 * swapped operands (cofi_gemm(A = W_v, W = source)), so no transpose pass exists.  (frames*S) % 4 == 0. */
int cofi_attention_vt(const float* q, const float* k, const float* vt, int64_t L, int64_t S, int frames, int heads,
                      int D, float scale, float* out, void* stream);
/* Training forward of the same kernel: also writes lse [frames*L, heads] = log sum_s exp(scale * q.k_s), the one
 * statistic cofi_attention_bwd needs to rebuild the probabilities. */
int cofi_attention_vt_lse(const float* q, const float* k, const float* vt, int64_t L, int64_t S, int frames, int heads,
                          int D, float scale, float* out, float* lse, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Matching (model/network.py:167-264, evaluation/eval_all.py:99-105)
 * ------------------------------------------------------------------------------------------------- */

/* Fused similarity + row arg-min; the [Npt,Npx] matrix is never materialised.
 * For every point row p: d[x] = 1 - sum_c pt[p,c]*px[x,c]; best_idx[p] = argmin_x d (lowest index on ties),
 * best_val[p] = min.  (model/network.py:174-179).  pt [frames*Npt, C], px [frames*Npx, C]. */
int cofi_sim_argmin(const float* pt, int64_t ldpt, const float* px, int64_t ldpx, int64_t Npt, int64_t Npx, int C,
                    int frames, int64_t* best_idx, float* best_val, int engine, void* stream);

/* fp16-operand throughput engine of the same fused similarity + arg-min (tcgen05 kind::f16, 256-point CTAs, TMEM fully
 * used, matrix never written).  pt/px are fp16 [rows, C] (C = 64 or 128), e.g. produced by cofi_cast_f16 from the
 * L2-normalised fp32 features. */
int cofi_sim_argmin_f16(const void* pt, int64_t ldpt, const void* px, int64_t ldpx, int64_t Npt, int64_t Npx, int C,
                        int frames, int64_t* best_idx, float* best_val,
                        int nsplit /* >1: the pixel range is split over nsplit CTAs per point tile (fills the machine when
                                      frames*Npt/256 < #SMs); partial results go to ws_* [nsplit, frames*Npt] and are merged */,
                        int64_t* ws_idx, float* ws_val, void* stream);
/* Exact fused similarity + arg-min on the tensor cores (model/network.py:174-179; the kernel on every product path).
 * Pass 1 (tcgen05 kind::f16 over the fp16 copies pt_h / px_h, matrix never written) is a candidate generator: for every
 * point row it keeps the pixels whose fp16 score lies within a margin (2 x the rigorous bound of |fp16 score - fp32
 * score|, 2.5e-3 for unit-norm rows) of the row's best fp16 score -- the only pixels that can win the exact comparison.
 * Pass 2 re-ranks those few candidates with the exact fp32 arithmetic of cofi_sim_argmin(engine = FP32) (rounded products,
 * ATen cascade-sum order, d = 1 - total, lowest index on ties) from the fp32 rows pt / px, so best_idx / best_val are
 * bit-identical to the fp32 engine's.  A row whose candidate list overflows (16 entries per pixel-range split) is scanned
 * exactly in full.  bound2: optional device float[2] = max squared row norm of pt, px (cofi_cast_f16_bound) scaling the
 * margin for un-normalised inputs; NULL = every row has norm <= 1 (cofi_l2norm_rows_f16 output).
 * work: cofi_sim_argmin_exact_workspace(Npt, Npx, frames) bytes, 16-byte aligned.  stats: optional device int32[2],
 * accumulated: candidates re-ranked, rows that needed the full scan.  C = 64 or 128. */
int64_t cofi_sim_argmin_exact_workspace(int64_t Npt, int64_t Npx, int frames);
int cofi_sim_argmin_exact(const float* pt, int64_t ldpt, const float* px, int64_t ldpx, const void* pt_h, int64_t ldpth,
                          const void* px_h, int64_t ldpxh, int64_t Npt, int64_t Npx, int C, int frames, const float* bound2,
                          int64_t* best_idx, float* best_val, void* work, int32_t* stats, void* stream);
/* y (fp16) = x (fp32) row by row, and *bound2_slot = max(*bound2_slot, |row|^2 * 1.0001) (caller zeroes the slot). */
int cofi_cast_f16_bound(const float* x, int64_t ldx, int64_t rows, int C, void* y, int64_t ldy, float* bound2_slot,
                        void* stream);
/* y[rows, C] (fp16) = x[rows, C] (fp32), round to nearest. C, ldx, ldy even. */
int cofi_cast_f16(const float* x, int64_t ldx, int64_t rows, int C, void* y, int64_t ldy, void* stream);

/* Test-mode selection loop of model/network.py:146-151 + :169,:184-186 in one kernel, per frame:
 * find the first threshold t in thresholds[0..nthr) with at least `min_count` points satisfying
 * score >= t AND the border mask on their matched pixel (2<=x<=xmax, 2<=y<=ymax);
 * emit those point indices in ascending order.  out_count[frame*2+0] = n, out_count[frame*2+1] = index of the
 * threshold used; out_index[frame*Npt ..]; out_xy[frame*2*Npt ..] laid out as [2][Npt] (x=col, y=row, fp32).
 * Rows n..Npt-1 of out_index / out_xy are padded with a valid dummy (index 0, centre (2,2)*xy_scale) so that fixed-shape
 * consumers inside a captured CUDA graph stay in bounds; at most 128 thresholds. */
int cofi_select_matches(const float* score, const int64_t* best_idx, int64_t Npt, int frames, int gridH, int gridW,
                        int xmax, int ymax /* border mask 2 <= x <= xmax, 2 <= y <= ymax; the reference hard-codes 62 / 18
                                              (network.py:184) for every grid */,
                        const float* thresholds, int nthr, int min_count, float xy_scale /* out_xy = pixel * xy_scale;
                        network.py:156 multiplies by 4 */, int32_t* out_count, int64_t* out_index, float* out_xy, void* stream);

/* point2node (model/network.py:250-264): idx[i] = argmin_j clamp(|p_i|^2 + |n_j|^2 - 2 p_i.n_j, 1e-12). */
int cofi_nn_argmin(const float* points, int64_t n, const float* nodes, int64_t M, int64_t* idx, void* stream);
/* Batched form: frame f matches points[f*n:(f+1)*n] against nodes[f*M:(f+1)*M]; idx is frame-local. */
int cofi_nn_argmin_batched(const float* points, int64_t n, const float* nodes, int64_t M, int frames, int64_t* idx,
                           void* stream);

/* extract_patch (model/network.py:206-226): out[i,c,dy,dx] = map[b, floor(cy-2)+dy, floor(cx-2)+dx, c],
 * map NHWC [B,H,W,C], centres [2,n] as fp32 (x row 0, y row 1), out [n,C,4,4].  Out-of-range windows
 * return COFI_EINVAL semantics via `err_flag` (device int set to 1), mirroring the reference's assert. */
int cofi_extract_patch(const float* map, int H, int W, int C, int b, const float* centers, int64_t n, float* out,
                       int32_t* err_flag, void* stream);
/* Batched form: centers [frames, 2, n], frame f reads image f of map [frames,H,W,C]; out [frames, n, C, 4, 4]. */
int cofi_extract_patch_batched(const float* map, int H, int W, int C, int frames, const float* centers, int64_t n,
                               float* out, int32_t* err_flag, void* stream);

/* Caller-side fine match (evaluation/eval_all.py:99-105): argmax over the 16 patch pixels of the cosine
 * similarity with the point feature; patch [n,C,16], pc [n,C] -> idx [n]. */
int cofi_fine_match(const float* patch, const float* pc, int64_t n, int C, int64_t* idx, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Training: backward of every forward op (the reference trains through ATen autograd, train.py:285)
 * All gradient outputs that are scattered into (dx of gathers, max-pools, bilinear up-sampling, patch extraction,
 * KPConv aggregate) must be ZEROED by the caller; `accumulate` flags add into an existing gradient.
 * ------------------------------------------------------------------------------------------------- */
int cofi_act_bwd(const float* dy, const float* y, int64_t n, int act, float* dx, void* stream);
int cofi_rowscale(const float* x, int64_t rows, int C, const float* rowdiv, float* y, void* stream);
int64_t cofi_colsum_workspace(int C);
int cofi_colsum(const float* x, int64_t ldx, int64_t rows, int C, float* out, int accumulate, void* work, void* stream);
/* grouped-norm backward (GroupNorm / InstanceNorm / train BatchNorm of cofi_norm_rows): x = forward input, y = forward
 * output (for the activation mask), mean_rstd = [frames*G] (mean, rstd) pairs from cofi_norm_rows_stats. */
int cofi_norm_rows_stats(const float* x, int64_t ldx, int64_t R, int C, int frames, int G, float eps, void* partials,
                         float* mean_rstd, void* stream);
int64_t cofi_norm_rows_bwd_workspace(int frames, int C);
int cofi_norm_rows_bwd(const float* x, const float* dy, const float* y, int64_t R, int C, int frames, int G,
                       const float* mean_rstd, const float* gamma, int act, float* dx, float* dres, float* dgamma,
                       float* dbeta, int accumulate, void* work, void* stream);
/* LayerNorm backward: dx plus t1 = dz*xhat, t2 = dz whose column sums (cofi_colsum) are dgamma / dbeta. */
int cofi_layer_norm_bwd(const float* x, const float* dy, int64_t rows, int C, const float* gamma, const float* beta,
                        float eps, int act, float* dx, float* t1, float* t2, void* stream);
int cofi_l2norm_bwd(const float* x, const float* dy, int64_t rows, int C, float* dx, void* stream);
int64_t cofi_colnorm_bwd_workspace(int frames, int C);
int cofi_colnorm_bwd(const float* x, const float* dy, int64_t L, int C, int frames, float* dx, void* work, void* stream);
int cofi_scatter_add_rows(const float* dy, int64_t ldy, int C, const int64_t* idx, int64_t idx_stride, int64_t Mq,
                          int64_t Ns, int frames, float* dx, void* stream);
int cofi_maxpool_rows_bwd(const float* x, int C, const int64_t* nbr, int H, int64_t Mq, int64_t Ns, int frames,
                          const float* dy, float* dx, void* stream);
int cofi_kpconv_aggregate_bwd(const float* dagg, int C, const float* s_packed, const float* q_points, const int64_t* nbr,
                              int H, int64_t Mq, int64_t Ns, int frames, const float* kernel_points, int K, float sigma,
                              float kp_reach, float* dfeats, void* stream);
int cofi_upsample2x_cat_bwd(const float* dy, int B, int H, int W, int C1, int C2, float* dx1, float* dx2, void* stream);
int cofi_maxpool2d_3x3s2_bwd(const float* x, const float* dy, int B, int H, int W, int C, float* dx, void* stream);
int cofi_dilate2_nhwc(const float* x, int B, int H, int W, int C, float* y, void* stream);
int cofi_extract_patch_bwd(const float* dpatch, int H, int W, int C, int b, const float* centers, int64_t n, float* dmap,
                           void* stream);
/* Batched form (cofi_extract_patch_batched): dpatch [frames, n, C, 4, 4], centers [frames, 2, n], dmap [frames,H,W,C] (+=). */
int cofi_extract_patch_batched_bwd(const float* dpatch, int H, int W, int C, int frames, const float* centers, int64_t n,
                                   float* dmap, void* stream);
/* C[Mo,No] (+)= A[R,Mo]^T B[R,No]: weight gradient of a Linear without transposes (A = dY, B = X).
 * engine COFI_GEMM_TF32: tcgen05 with both operands MN-major (the reduction index is the outer one in HBM), split over R;
 * any other engine: SIMT fp32.  Partial tiles are folded by a deterministic reduction. */
int64_t cofi_gemm_tn_workspace(int64_t R, int Mo, int No);
int cofi_gemm_tn(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t R, int Mo, int No,
                 int accumulate, int engine, void* work, void* stream);
/* weight gradient of cofi_conv2d_nhwc: dw[Cout, KH*KW*Cin] */
int64_t cofi_conv2d_wgrad_workspace(int B, int Ho, int Wo, int Cout, int KH, int KW, int Cin);
int cofi_conv2d_wgrad_nhwc(const float* x, int B, int H, int W, int Cin, const float* dy, int Cout, int KH, int KW,
                           int stride, int pad, float* dw, int accumulate, int engine, void* work, void* stream);
/* attention for training: forward that also returns the log-sum-exp [frames*L, heads], and its backward (D = 32). */
int cofi_attention_fwd_lse(const float* q, const float* k, const float* v, int64_t L, int64_t S, int frames, int heads,
                           int D, float scale, float* out, float* lse, void* stream);
int cofi_attention_bwd(const float* q, const float* k, const float* v, const float* out, const float* dout,
                       const float* lse, int64_t L, int64_t S, int frames, int heads, int D, float scale, float* dq,
                       float* dk, float* dv, float* dsum_work /* [frames*L*heads] */, void* stream);
/* The same backward on tcgen05 (tf32 operands, fp32 accumulate in TMEM; D == 32, L and S multiples of 4): two launches of
 * one kernel shaped like the forward -- dQ with the keys streamed, then dK/dV with the queries streamed; the [L,S]
 * matrices never reach HBM.  tr_work: cofi_attention_bwd_tc_workspace() bytes for the Q^T, dO^T, K^T copies. */
int64_t cofi_attention_bwd_tc_workspace(int64_t L, int64_t S, int frames, int heads, int D);
int cofi_attention_bwd_tc(const float* q, const float* k, const float* v, const float* out, const float* dout,
                          const float* lse, int64_t L, int64_t S, int frames, int heads, int D, float scale, float* dq,
                          float* dk, float* dv, float* dsum_work, void* tr_work, void* stream);
/* ---------------------------------------------------------------------------------------------------
 * Pose step after the hot path (evaluation/eval_all.py:107: cv2.solvePnPRansac, 10000 iterations per frame on the CPU)
 * ------------------------------------------------------------------------------------------------- */

/* Batched P3P-RANSAC: all `iterations` hypotheses of all frames in one launch (one thread per hypothesis), winner = most
 * inliers (squared reprojection error <= reproj_threshold^2, positive depth), lowest hypothesis index on ties.
 * image_points [frames, n_max, 2] pixels, object_points [frames, n_max, 3]; count (optional device int32, frame f reads
 * count[f * count_stride]) = real rows per frame, NULL = n_max; cam [frames, 4] = fx, fy, cx, cy.  Sampling is
 * counter-based (Philox4x32-10 keyed by seed) so that the host oracle restates it.  Outputs: out_count [frames] inliers of
 * the winner (0 = no valid hypothesis), out_hypothesis [frames] its index (-1), out_pose [frames, 12] float64 = R row-major
 * then t (X_cam = R X + t), out_inlier [frames, n_max] uint8.  work: 8 * frames bytes, 8-byte aligned.  The caller refines
 * on the inliers (cv2.solvePnP ITERATIVE -- what solvePnPRansac itself ends with). */
int cofi_pnp_ransac(const float* image_points, const float* object_points, const int32_t* count, int count_stride, int n_max,
                    int frames, const float* cam, int iterations, float reproj_threshold, uint64_t seed, int32_t* out_count,
                    int32_t* out_hypothesis, double* out_pose, uint8_t* out_inlier, void* work, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * Training losses (model/loss.py:9-93; called from train.py:254-283), fused forward + analytic backward,
 * batched over `frames` stacked frames: loss[frames] per frame, gradients w.r.t. the gathered rows in compact
 * buffers (NULL = forward only); cofi_scatter_scaled_rows applies the upstream gradient and scatters them.
 * ------------------------------------------------------------------------------------------------- */

/* desc_loss (model/loss.py:69-93).  img_tok [frames*img_rows, C] / pc_tok [frames*pc_rows, C]: token-layout descriptor maps;
 * pix / kpt [frames, n]: frame-local rows of the n supervised pixels / points; mask [frames, n, n] (1 = positive pair,
 * rows = pixels).  dists [frames, n, n] optional output (the matrix the reference returns); d_img / d_pc [frames, n, C]. */
int cofi_desc_loss(const float* img_tok, const int64_t* pix, int64_t img_rows, const float* pc_tok, const int64_t* kpt,
                   int64_t pc_rows, const float* mask, int n, int C, int frames, float pos_margin, float neg_margin,
                   float log_scale, float* loss, float* dists, float* d_img, float* d_pc, void* stream);

/* overlap_loss (model/loss.py:53-60): BCE of score[idx] against 1 for the first n_in indices, 0 for the next n_out.
 * score [frames*rows] token layout; idx [frames, n_in + n_out] frame-local; d_score [frames, n_in + n_out]. */
int cofi_overlap_loss(const float* score, const int64_t* idx, int64_t rows, int n_in, int n_out, int frames, float* loss,
                      float* d_score, void* stream);

/* fine_circle_loss (model/loss.py:9-51).  patch [frames*n, C, 16], fpc [frames*n, C], rel [frames*n] in 0..15 (outside:
 * *bad_flag = 1, the reference's label[...] indexing raises).  d_patch / d_fpc: same shapes as patch / fpc. */
int cofi_fine_circle_loss(const float* patch, const float* fpc, const int64_t* rel, int n, int C, int frames, float m,
                          float gamma, float* loss, float* d_patch, float* d_fpc, int32_t* bad_flag, void* stream);

/* dst[f*rows_dst + idx[f*R + r], :] += scale * src[f*R + r, :]  (dst zero-initialised by the caller; idx NULL:
 * dst[f*R + r, :] = scale * src[f*R + r, :]).  scale = scale_host * scale_dev[f] (scale_dev: device float[frames] or NULL):
 * the upstream gradient of every frame's loss is a device value inside a captured training step. */
int cofi_scatter_scaled_rows(const float* src, const int64_t* idx, int64_t R, int64_t rows_dst, int frames, int C,
                             const float* scale_dev, float scale_host, float* dst, void* stream);

/* fused Adam step (torch.optim.Adam semantics, reference train.py:154); grad_scale folds the 1/world_size of the
 * data-parallel gradient average. */
int cofi_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                   int step, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COFI_B200_H */
