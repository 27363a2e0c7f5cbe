/*
 * dic_b200.h - C ABI of libdic_b200.so: the B200 (sm_100a) hot path of
 * Deep_Interpolation_Clustering.
 *
 * The reference has no FFI: its hot path is Python nn.Modules and sklearn calls.
 * Each entry point below states the reference interface it stands behind
 * (file:line in the upstream repository).  A caller binds these with ctypes (see
 * INTEGRATION.md); the package's own nn.Module mirrors do exactly that.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - all matrices are dense, row-major, float32 unless stated;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*;
 *     NULL = legacy default stream) and keeps no global mutable state;
 *   - return value 0 = success, < 0 = error (dic_status); never throws; the
 *     message of the calling thread's last error is dic_last_error();
 *   - inputs are never modified; outputs are fully overwritten.
 *
 * Planar layouts used across the boundary (the nn.Module mirrors expose the
 * reference's permuted views of these buffers):
 *   x      (B, 4C, T)  planes [value | padding mask | time in hours | hold-out]
 *                       interpolation_layer.py:26-30; the hold-out plane is never read, so every
 *                       call that takes x also takes x_stride = floats between consecutive
 *                       encounters (0 = dense 4*C*T; 3*C*T for a buffer that holds only the
 *                       three live planes, see dic_upload_encounters)
 *   u      (B, 3C, R)  SCI output rows [y (C) | w (C) | y' (C)]; the reference returns
 *                       u.permute(0,2,1), interpolation_layer.py:84-85
 *   cci    (B, 3C, R)  rows [z | exp(w) | y' - z]; interpolation_layer.py:124-126
 *   v      (B, C, R)   compress_fc output, rbf.py:101-103
 *   rec    (B, C, T)   RBF output, rbf.py:107
 */
#ifndef DIC_B200_H
#define DIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DIC_B200_VERSION 100 /* major*10000 + minor*100 + patch */

typedef void* dic_stream_t;

enum dic_status {
  DIC_OK = 0,
  DIC_ERR_INVALID_ARGUMENT = -1, /* NULL pointer, non-positive size, misaligned buffer  */
  DIC_ERR_UNSUPPORTED = -2,      /* shape outside the kernels' limits (see each call)   */
  DIC_ERR_CUDA = -3,             /* a CUDA runtime call / launch failed                 */
  DIC_ERR_NO_DEVICE = -4         /* no sm_100 device visible                            */
};

/* ---- library ------------------------------------------------------------------- */
const char* dic_last_error(void);
int dic_version(void);
/* Fills SM count and compute capability of the current device. */
int dic_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- interpolation network -------------------------------------------------------
 * SingleChannelInterp.forward, interpolation_layer.py:31-86.
 *   kernel (C)  raw parameter; alpha = log(1+exp(kernel)) is applied inside (:51)
 *   ref_t  (R)  reference grid, torch.linspace(0, H, R) (:41)
 *   u      (B,3C,R) out
 *   stats  (B,3C,R) out, may be NULL: the three gradient-moment rows [U1 | U0 | U1'] from which
 *          dic_sci_bwd forms d kernel without a second sweep over the observations (saved-for-backward
 *          state, 12CR bytes per encounter; U0 < 0 marks an all-masked vital)
 * Limits: 12*C*round_up(T,4) + 64 bytes of shared memory <= 227 KB.
 */
int dic_sci_fwd(const float* x, const float* kernel, const float* ref_t, float* u, float* stats,
                int64_t B, int C, int T, int R, int64_t x_stride, dic_stream_t stream);

/* Bytes of scratch dic_sci_bwd / dic_rbf_bwd need (per-encounter partials + reduction). */
size_t dic_interp_bwd_workspace_bytes(int64_t B, int C);

/* Gradient of dic_sci_fwd wrt `kernel` (what autograd produces for
 * interpolation_layer.py:51-83).  grad_u (B,3C,R) is the upstream gradient in planar
 * layout.  No input gradient is produced.  Only kernel, stats and grad_u are
 * read: d alpha_c = - sum_{b,r} (gy U1 + gw U0 + gy' U1'); x, ref_t and u may be NULL.
 *   d_kernel (C) out (overwritten, deterministic two-stage reduction)
 */
int dic_sci_bwd(const float* x, const float* kernel, const float* ref_t, const float* u,
                const float* stats, const float* grad_u, float* d_kernel, void* workspace,
                int64_t B, int C, int T, int R, int64_t x_stride, dic_stream_t stream);

/* CrossChannelInterp.forward, interpolation_layer.py:99-127.
 *   u (B,3C,R) in, kernel (C,C) in, out (B,3C,R).  Limit: C <= 16.
 */
int dic_cci_fwd(const float* u, const float* kernel, float* out, int64_t B, int C, int R,
                dic_stream_t stream);

size_t dic_cci_bwd_workspace_bytes(int64_t B, int C);

/* Gradients of dic_cci_fwd: grad_u (B,3C,R) and d_kernel (C,C). */
/* CCI backward with the SCI backward folded in (the common chain cci(sci(x)), pretrain_interp.py:138-139): the
 * gradient with respect to the SCI output is contracted with the saved moment rows `stats` inside the kernel instead of
 * being written to HBM and read back by dic_sci_bwd.  Results equal dic_cci_bwd + dic_sci_bwd.  d_dim <= 8.
 * workspace: dic_cci_sci_bwd_workspace_bytes(B, C). */
size_t dic_cci_sci_bwd_workspace_bytes(int64_t B, int C);
int dic_cci_sci_bwd(const float* u, const float* cci_kernel, const float* sci_kernel, const float* stats,
                    const float* grad_out, float* d_cci_kernel, float* d_sci_kernel, void* workspace,
                    int64_t B, int C, int R, dic_stream_t stream);
int dic_cci_bwd(const float* u, const float* kernel, const float* grad_out, float* grad_u,
                float* d_kernel, void* workspace, int64_t B, int C, int R, dic_stream_t stream);

/* RBF.forward after compress_fc, rbf.py:69-80,95-97,104-107 with the gaussian basis
 * rbf.py:129-131.
 *   v (B,C,R) in, x (B,4C,T) in (mask and time planes are read), kernel (C), ref_t (R)
 *   rec (B,C,T) out
 *   inv_norm (B,C,T) out, may be NULL: 1/(sum_r phi + 1e-10), saved for the backward pass
 */
int dic_rbf_fwd(const float* v, const float* x, const float* kernel, const float* ref_t,
                float* rec, float* inv_norm, int64_t B, int C, int T, int R, int64_t x_stride,
                dic_stream_t stream);

/* Gradients of dic_rbf_fwd wrt v (flows into compress_fc) and kernel.
 *   grad_rec (B,C,T) upstream; grad_v (B,C,R) out; d_kernel (C) out
 */
int dic_rbf_bwd(const float* v, const float* x, const float* kernel, const float* ref_t,
                const float* rec, const float* inv_norm, const float* grad_rec, float* grad_v,
                float* d_kernel, void* workspace, int64_t B, int C, int T, int R, int64_t x_stride,
                dic_stream_t stream);

/* Host -> device transfer of a batch of encounters as ONE strided DMA that skips the hold-out
 * plane (a quarter of x; interpolation_layer.py:26-30 never reads it): row b of x_host
 * (B, host_planes, T) [pinned host memory for an asynchronous copy] lands at
 * x_dev + b * dev_planes * T; only the first 3*C planes are moved.  host_planes, dev_planes >= 3C.
 * The device buffer is then passed to the calls above with x_stride = dev_planes * T. */
int dic_upload_encounters(float* x_dev, const float* x_host, int64_t B, int C, int T,
                          int host_planes, int dev_planes, dic_stream_t stream);

/* ---- ragged (packed) encounters ---------------------------------------------------------
 * The pipeline's rows are left-packed (p0_data_process.py:44-67: observations first, zero padding after), so the
 * dense planes the trainer ships (pretrain_trainer.py:132-136) are on average half padding and the mask plane of a
 * row is one integer.  The packed form keeps, per (encounter b, vital c), only the valid prefix:
 *   n_obs   (B, C) int32    prefix length
 *   enc_off (B + 1) int64   float offset of encounter b in `packed` (multiples of 4: rows are 16-byte aligned)
 *   packed                  for b, for c: [ value[0..n4) | time[0..n4) ], n4 = round_up(n_obs[b,c], 4); pad slots
 *                           hold value 0 and time 3e18 (an observation whose Gaussian weight is exactly 0)
 * ~6.2 KB instead of 18.4 KB per encounter at c2 (C = 6, T = 256, mean 128.5 observations).
 *
 * dic_pack_encounters_host: HOST function (no device work).  Reads x_host (B, host_planes, T) float32, fills
 * n_obs_host and enc_off_host, and - when packed_host is not NULL - the packed rows (capacity in floats).  Returns
 * the number of floats the packed rows take (call once with packed_host = NULL to size the buffer), or a negative
 * dic_status: DIC_ERR_UNSUPPORTED when a mask is not a left-packed 0/1 prefix (general masks take the dense
 * upload).  *all_sorted (may be NULL) is set to 1 when every row's times are non-decreasing on its prefix.  C <= 16. */
int64_t dic_pack_encounters_host(const float* x_host, int64_t B, int C, int T, int host_planes,
                                 int32_t* n_obs_host, int64_t* enc_off_host, float* packed_host,
                                 int64_t capacity, int* all_sorted);

/* Host -> device copy of encounters [b0, b0 + B) of a packed set: pass n_obs_host + b0*C and enc_off_host + b0
 * (B + 1 entries are read; packed_host is the BASE of the packed rows, the slice is located through enc_off_host).
 * Three contiguous asynchronous copies (pin the host arrays); packed_dev needs enc_off_host[B] - enc_off_host[0]
 * floats and 16-byte alignment.  Offsets stay absolute: consumers subtract enc_off_dev[0]. */
int dic_upload_encounters_packed(const float* packed_host, const int32_t* n_obs_host, const int64_t* enc_off_host,
                                 float* packed_dev, int32_t* n_obs_dev, int64_t* enc_off_dev, int64_t B, int C,
                                 dic_stream_t stream);

/* Device-side expansion of packed encounters into the dense planes [value | mask | time] of x_dev
 * (B, dev_planes, T), dev_planes >= 3C (planes beyond 3C are left untouched): the buffer every dic_* call above
 * takes with x_stride = dev_planes * T.  Bit-identical to what dic_upload_encounters delivers for the same rows. */
int dic_expand_encounters(const float* packed_dev, const int32_t* n_obs_dev, const int64_t* enc_off_dev,
                          float* x_dev, int64_t B, int C, int T, int dev_planes, dic_stream_t stream);

/* ---- DEC soft assignment ----------------------------------------------------------
 * ClusterAssignment.forward, dec.py:49-63.  z (B,D), mu (K,D) -> q (B,K).
 *   labels (B) int32 out, may be NULL: argmax_j q (clustering_trainer.py:473-484)
 *   colsum (K) float64 out, may be NULL: f_j = sum_i q_ij (dec.py:73), so that
 *          target_distribution needs no second pass over q
 * Limits: K <= 32, D % 4 == 0, D <= 1024.  workspace: dic_dec_workspace_bytes(K, D).
 */
size_t dic_dec_workspace_bytes(int K, int D);
int dic_dec_q_fwd(const float* z, const float* mu, float* q, int32_t* labels, double* colsum,
                  void* workspace, int64_t B, int D, int K, float alpha, dic_stream_t stream);

/* target_distribution, dec.py:66-76, with the column sum f (K, float64) supplied by the
 * caller (local, or all-reduced across ranks for a sharded batch). */
int dic_dec_p(const float* q, const double* colsum, float* p, int64_t B, int K,
              dic_stream_t stream);

/* Backward of dic_dec_q_fwd for an arbitrary upstream grad_q (B,K):
 *   grad_z (B,D) out, grad_mu (K,D) out. */
int dic_dec_q_bwd(const float* z, const float* mu, const float* grad_q, float* grad_z,
                  float* grad_mu, void* workspace, int64_t B, int D, int K, float alpha,
                  dic_stream_t stream);

/* Fused DEC step: given z, mu and the (global) column sum f, computes
 *   p = target(q) (detached), kl_sum = sum_ij p (log p - log q)   [Net.kl_loss,
 *   clustering_interp.py:205-207, before the 1/B of 'batchmean'],
 *   grad_z = scale * dKL/dz, grad_mu = scale * dKL/dmu  (closed form).
 * Pass scale = weight / B_global.  p, grad_z may be NULL.  kl_sum is one float64. */
int dic_dec_kl_fwd_bwd(const float* z, const float* mu, const double* colsum, float* p,
                       double* kl_sum, float* grad_z, float* grad_mu, void* workspace,
                       int64_t B, int D, int K, float alpha, float scale, dic_stream_t stream);

/* ---- k-means / gap statistic (sklearn.cluster.KMeans as called at
 * p2_clustering_optK.py:260,284,372,377; clustering_trainer.py:75-82) ----------------
 * dtype: 0 = float32, 1 = float64 (the reference feeds float64 reference draws,
 * p2_clustering_optK.py:370).  X (N,D) and centers (K,D) share the dtype.
 *
 * One Lloyd pass (sklearn/cluster/_k_means_lloyd.pyx:23-218): E-step + accumulation of
 * everything the M-step and the stopping rule need.
 *   labels (N) int32 in/out: argmin_j ||c_j||^2 - 2 x.c_j, lowest index on ties
 *   sums (K,D) float64 out, counts (K) float64 out: per-cluster sums and sizes
 *   stats (4) float64 out: [ sum_i ||x_i - c_label||^2 (computed directly, like
 *          _k_means_common.pyx:96-124),  number of labels that changed,
 *          sum_i ||x_i - c_label|| (the elbow distortion numerator,
 *          p2_clustering_optK.py:261-264),  0 ]
 *   flags: DIC_KM_COUNT_CHANGES  compare with the previous content of `labels`
 *          DIC_KM_KEEP_LABELS    do not re-assign: accumulate for the given labels
 * sums/counts may be NULL (predict / inertia only).  Limits: K <= 64, D <= 512.
 */
#define DIC_KM_COUNT_CHANGES 1
#define DIC_KM_KEEP_LABELS 2
#define DIC_KM_NO_INERTIA 4 /* skip stats[0] and stats[2] (a Lloyd iteration does not need them) */
/* Kernel selector in bits 8..11 of `flags` (parity tests and benchmarks address every kernel through the ABI; the
 * library reads no environment variables): 0 = chosen by shape and measured cost, 1 = specialised tile kernel,
 * 2 = streaming half-warp-per-row, 3 = generic tile kernel, 4 = general kernel, 5 = tcgen05 pass (float32 rows of
 * 64 / 128 / 256 elements, K <= 16, DIC_KM_NO_INERTIA form), 6 = its float64 form (rows of 64 elements: tensor-core
 * screen + exact float64 evaluation of the rows inside the screen's error bound); a kernel that does not cover the
 * shape answers DIC_ERR_UNSUPPORTED. */
#define DIC_KM_KERNEL(k) ((k) << 8)
size_t dic_kmeans_workspace_bytes(int K, int D);
int dic_kmeans_assign(const void* X, const void* centers, int32_t* labels, double* sums,
                      double* counts, double* stats, void* workspace, int64_t N, int D, int K,
                      int dtype, int flags, dic_stream_t stream);

/* M-step (sklearn/cluster/_k_means_common.pyx:236-260, _kmeans.py:703-733): centers (K,D, dtype)
 * in/out <- sums / counts; if ANY cluster is empty the centres are left untouched (relocation is
 * the caller's rare path: it redoes the M-step); status (4) float64 out = [labels changed (stats[1]), sum of squared
 * centre shifts, number of empty clusters, inertia (stats[0])] - one small read per iteration. */
int dic_kmeans_update(const double* sums, const double* counts, const double* stats, void* centers,
                      double* status, int D, int K, int dtype, dic_stream_t stream);

/* One whole Lloyd iteration in one call: dic_kmeans_assign followed by dic_kmeans_update on the same
 * stream (single-process use: a sharded fit all-reduces sums / counts / stats between the two). */
int dic_kmeans_lloyd_step(const void* X, void* centers, int32_t* labels, double* sums, double* counts,
                          double* stats, double* status, void* workspace, int64_t N, int D, int K,
                          int dtype, int flags, dic_stream_t stream);

/* Up to n_steps Lloyd iterations enqueued back to back with a device-side stopping rule, so that the host
 * synchronises once per batch instead of once per iteration (at N <= 1e5 the per-iteration round trip is most of
 * the time).  status8 (8 doubles, zero status8[4..7] before the first batch of a run):
 *   [0..3] as dic_kmeans_update, [4] stop flag (0 = running, 1 = converged, 2 = a cluster came out empty: the
 *   centres were left untouched, the caller relocates and resumes), [5] iterations executed so far,
 *   [6] 1 if the run stopped because no label changed (strict convergence, _kmeans.py:700-712), [7] unused.
 * Once the flag is set every remaining kernel of the batch returns at once: labels, sums and centres are those
 * of the stopping iteration.  tol is the absolute tolerance on the squared centre shift (_kmeans.py:285-293). */
int dic_kmeans_lloyd_run(const void* X, void* centers, int32_t* labels, double* sums, double* counts,
                         double* stats, double* status8, void* workspace, int64_t N, int D, int K, int dtype,
                         int flags, int n_steps, double tol, dic_stream_t stream);

/* k-means++ potentials (sklearn/cluster/_kmeans.py:224-281): for each of L candidate centres
 *   pots[l] = sum_i min(min_d2[i], ||x_i - cand_l||^2)          (float64, L <= 16)
 * min_d2 (N, same dtype as X) may be NULL (treated as +inf: first centre).  If
 * min_d2_out is not NULL (L must be 1) it receives the element-wise minimum: the commit. */
int dic_kmeans_min_d2(const void* X, const void* cands, const void* min_d2, void* min_d2_out,
                      double* pots, void* workspace, int64_t N, int D, int L, int dtype,
                      dic_stream_t stream);

/* Sum of all pairwise Euclidean distances inside one cluster, the kernel behind
 * KM.compute_inertia_v1 / computer_intertia_v2, p2_clustering_optK.py:334-351:
 *   Xc (n,D) rows of one cluster (gathered by the caller), out (1) float64 =
 *   sum_{i,i'} ||x_i - x_i'||  over the FULL n x n matrix (zero diagonal included).
 * The n x n matrix is never materialised.  Limit: D <= 512.
 * dtype | DIC_PAIRWISE_EXACT: always the direct (x_i - x_j)^2 CUDA-core kernel in the data's own precision (float32
 * clusters of >= 512 rows otherwise take the tensor-core Gram form, which expects a centred cluster). */
#define DIC_PAIRWISE_EXACT 16
size_t dic_pairwise_workspace_bytes(int64_t n, int D);
int dic_pairwise_dist_sum(const void* Xc, double* out, void* workspace, int64_t n, int D,
                          int dtype, dic_stream_t stream);
/* Stripe `part` of `n_parts` of the same sum (pairwise inertia on several GPUs): every
 * rank holds all n rows of the cluster (all-gathered by the caller) and evaluates the tiles
 * part, part + n_parts, ... of the tile list; the n_parts results add up to dic_pairwise_dist_sum's. */
int dic_pairwise_dist_sum_part(const void* Xc, double* out, void* workspace, int64_t n, int D,
                               int dtype, int part, int n_parts, dic_stream_t stream);

/* Per-row, per-cluster distance sums - the O(N^2) part of the silhouette coefficient
 * (internal_eval.py:112-122 -> sklearn.metrics.silhouette_score, called inside the gap loop at
 * p2_clustering_optK.py:401-405) - on the tensor cores, without the n x n matrix:
 *   rowsum[i][k] = sum_{j : cluster(j) = k} ||x_perm[i] - x_perm[j]||          (n_pad, K) float64
 * The caller sorts the rows by cluster and pads every cluster to whole 128-row tiles:
 *   perm (n_pad) int32: source row of packed row i, -1 = padding (contributes exactly 0);
 *   tile_cluster (n_pad / 128) int32: cluster of each 128-row tile.
 * Limits: D <= 256, D % 4 == 0.  workspace: dic_pairwise_workspace_bytes(n_pad, D). */
int dic_cluster_rowsums(const float* X, const int32_t* perm, const int32_t* tile_cluster,
                        double* rowsum, void* workspace, int64_t n_pad, int D, int K,
                        dic_stream_t stream);

/* Per-cluster dispersion for Calinski-Harabasz / Davies-Bouldin (internal_eval.py:125-147 -> sklearn.metrics
 * calinski_harabasz_score / davies_bouldin_score, called inside the gap loop at p2_clustering_optK.py:401-405):
 *   out (K,2) float64:  out[k][0] = sum_{label(i)=k} ||x_i - c_k||,  out[k][1] = sum_{label(i)=k} ||x_i - c_k||^2
 * in one pass over X (N,D) with labels (N) int32 in [0,K) and centers (K,D) of the data's dtype (the per-cluster
 * means: dic_kmeans_assign with DIC_KM_KEEP_LABELS yields their sums / counts).  Deterministic.  K <= 64.
 * workspace: dic_cluster_scatter_workspace_bytes(K). */
size_t dic_cluster_scatter_workspace_bytes(int K);
int dic_cluster_scatter(const void* X, const int32_t* labels, const void* centers, double* out,
                        void* workspace, int64_t N, int D, int K, int dtype, dic_stream_t stream);

/* The O(N^2) part of the Dunn index (internal_eval.py:15-109, "nearest" inter-cluster distances over the
 * "farthest" diameter):
 *   out[a*K + b] (a < b) = min over x_i in cluster a, x_j in cluster b of ||x_i - x_j||   (+inf: no such pair;
 *                          entries with a >= b stay +inf)      - the reference's cluster_distances matrix (:37-53)
 *   out[K*K]             = max over pairs with equal labels of ||x_i - x_j||               (0 if none; :77-81)
 * from 64 x 64 distance tiles, upper triangle only; the n x n matrix the reference builds (euclidean_distances,
 * internal_eval.py:103) is never materialised.  labels (N) int32 in [0, K), K <= 32; dtype as above. */
int dic_dunn_minmax(const void* X, const int32_t* labels, double* out, int64_t N, int D, int K, int dtype,
                    dic_stream_t stream);

/* ---- bidirectional LSTM either side of the interpolation network ------------------------------
 * Replaces the recurrence of nn.LSTM(input, 128, num_layers=1, bidirectional=True) in EncoderRNN / DecoderRNN
 * (pretrain_interp.py:14-41, clustering_interp.py likewise); float32 semantics of torch.nn.LSTM, gate order i|f|g|o.
 *
 * dic_lstm_pack_whh: weight_hh_l0 and weight_hh_l0_reverse (512,128 each) -> the shared-memory image of the tensor-core
 *   operand (fp16 hi + lo halves of the power-of-two-scaled weights, per direction and per 32-unit slice) followed by the
 *   two inverse scales; `packed` holds dic_lstm_packed_bytes() bytes.  Call again whenever the weights change.
 * dic_lstm_fwd: all R steps of both directions in ONE persistent kernel (4-CTA clusters, W_hh resident in shared memory,
 *   h exchanged through distributed shared memory, tcgen05 MMAs with TMEM accumulators).
 *   pre  (R,B,1024): W_ih x_t + b_ih + b_hh of both directions (a library GEMM of the caller), columns ordered
 *        [direction][slice q = 0..3][half = 0,1][gate i,f,g,o][16 units]  <->  hidden unit 32 q + 16 half + u
 *   h0, c0 (2,B,128) or NULL (zeros); out (R,B,256) = [forward | reverse]; hn, cn (2,B,128);
 *   save (2,R,B,5,128) or NULL: i, f, g, o, c_t of every step for the backward pass.  hidden must be 128.
 * dic_lstm_bwd_step: the gate-gradient algebra of ONE backward step of one direction (closed form of autograd on the
 *   cell): da (B,512) = d a_t in gate order i|f|g|o, dc_rec (B,128) in/out; save_t (B,5,128) the step's saved row,
 *   c_prev (row stride c_prev_stride) the previous cell state or NULL (zeros), gh_out (row stride gh_stride) the upstream
 *   gradient of this direction's h_t or NULL, dh_rec (B,128) the recurrent gradient d a_(t+1) W_hh (a library GEMM). */
/* dic_lstm_pack_wih / dic_lstm_project: the input projection pre (M,1024) = A (M,K) Wp^T + bias of all R*B rows on the
 *   tensor cores with the recurrence's arithmetic (fp16 hi + lo operands, float32 accumulation in TMEM, K <= 1024):
 *   Wp (1024,K) row-major in the column order dic_lstm_fwd expects; packed holds dic_lstm_project_packed_bytes(K) bytes;
 *   A has row stride lda (elements); relu != 0 applies max(., 0) to A on the fly (DecoderRNN.forward, :38). */
size_t dic_lstm_project_packed_bytes(int K);
int dic_lstm_pack_wih(const float* wp, void* packed, int K, dic_stream_t stream);
int dic_lstm_project(const float* A, int64_t lda, const void* packed, const float* bias, float* out, int64_t M, int K,
                     int relu, dic_stream_t stream);
size_t dic_lstm_packed_bytes(void);
int dic_lstm_pack_whh(const float* w_hh, const float* w_hh_reverse, void* packed, dic_stream_t stream);
int dic_lstm_fwd(const float* pre, const void* packed, const float* h0, const float* c0, float* out, float* hn,
                 float* cn, float* save, int R, int64_t B, int hidden, dic_stream_t stream);
int dic_lstm_bwd_step(const float* save_t, const float* c_prev, int64_t c_prev_stride, const float* gh_out,
                      int64_t gh_stride, const float* dh_rec, float* dc_rec, float* da, int64_t B, int hidden,
                      dic_stream_t stream);

/* ---- utilities -------------------------------------------------------------------- */
/* Deterministic column sums of a (rows, cols) float32 matrix into float64. */
size_t dic_colsum_workspace_bytes(int cols);
int dic_colsum_f32(const float* a, double* out, void* workspace, int64_t rows, int cols,
                   dic_stream_t stream);

/* Roofline probes: run a MUFU.EX2-only / FFMA-only loop on every SM and report the
 * achieved rate (operations per second) measured with CUDA events. */
int dic_probe_mufu(double* ex2_per_second_host, dic_stream_t stream);
int dic_probe_ffma(double* ffma_per_second_host, dic_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DIC_B200_H */
