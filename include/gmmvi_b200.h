/* gmmvi_b200 -- C ABI of the B200-native GMMVI hot path.
 *
 * The reference (OlegArenz/gmmvi) has no FFI: its "plugin API" is a set of Python classes whose
 * numerical work is TensorFlow library calls.  Each entry point below replaces the TF/TFP call
 * sequence of one reference method (cited as file:line relative to /root/reference/src/gmmvi/) and is
 * what the Python mirror in gmmvi_b200/ binds through ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (fp32 / int32, contiguous row-major) unless marked "host";
 *   - kernels never allocate: scratch is passed in (`ws`, size from the matching *_workspace());
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant, and keeps no
 *     global mutable state;
 *   - return value: 0 ok, <0 error (GVI_ERR_*); the message is in gvi_last_error() (thread local);
 *     nothing throws across the boundary;
 *   - numerical failure is DATA, not an error: updaters report it per component in `success[K]`
 *     (non-positive pivot => reject, keep old parameters), mirroring the reference's NaN-Cholesky test
 *     (optimization/gmmvi_modules/ng_based_component_updater.py:120-138, 320-324, 488-511).
 */
#ifndef GMMVI_B200_H
#define GMMVI_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GVI_OK 0
#define GVI_ERR_INVALID (-1)
#define GVI_ERR_WORKSPACE (-2)
#define GVI_ERR_CUDA (-3)
#define GVI_ERR_UNSUPPORTED (-4)

int gvi_version(void);
const char* gvi_last_error(void);

/* ---- parameter preparation ---------------------------------------------------------------------
 * chol[K,D,D] (lower) -> linv = chol^-1 (lower), prec = linv^T linv (nullable), cst[k] = -sum_i log
 * chol[k,i,i] - D/2 log(2 pi).  Replaces tf.linalg.inv(chols) (optimization/sample_db.py:121,132;
 * ng_based_component_updater.py:106,178,457) and the const_parts of models/full_cov_gmm.py:60-61.
 * Computed in fp64 internally, stored fp32.  ok[k]=0 (nullable) flags a non-positive diagonal. */
size_t gvi_prepare_full_workspace(int K, int D);
int gvi_prepare_full_f32(const float* chol, int K, int D, float* linv, float* prec, float* cst, int32_t* ok,
                         void* ws, size_t ws_bytes, void* stream);

/* ---- log densities ----------------------------------------------------------------------------
 * lq[k,n] = cst[k] - 1/2 || linv_k (x_n - mu_k) ||^2          models/full_cov_gmm.py:56-62,
 * and with linv = SampleDB.inv_chols                           optimization/sample_db.py:154-162. */
int gvi_logdens_full_f32(const float* X, int N, int D, const float* means, const float* linv, const float* cst,
                         int K, float* lq, void* stream);
/* Same result on the tcgen05 tensor cores (3xTF32 split precision, fp32 accumulation in TMEM, TMA-staged
 * Linv).  linv_hi / linv_lo = gvi_split_tf32_f32(linv).  Supported when gvi_logdens_full_tc_supported(D). */
int gvi_split_tf32_f32(const float* in, long long n, float* hi, float* lo, void* stream);
int gvi_logdens_full_tc_supported(int D);
int gvi_logdens_full_tc_f32(const float* X, int N, int D, const float* means, const float* linv_hi,
                            const float* linv_lo, const float* cst, int K, float* lq, void* stream);
/* Same result with tcgen05 kind::f16 in "2 x fp16" split precision (22 significand bits like the TF32 split, at
 * twice the tensor rate) and the factor of the current component RESIDENT in shared memory, so that only the
 * sample tile streams from L2.  Operands: gvi_split_h16_f32(linv) -> zero-padded fp16 hi / lo copies
 * [K, Dp, Dp] (Dp = gvi_h16_padded_dim(D)), scaled per component by the power of two derived from
 * tmax[k] = max |linv_k|;  tileinf[t] = max |X[n,d]| over the 128 samples of tile t and minf[k] = max_d
 * |means[k,d]| (gvi_group_absmax_f32 with group 128 / 1) bound |x_n - mu_k| and give the power-of-two scale of
 * the A operand of work item (k, t).  D % 4 == 0, 4 <= D <= 256. */
int gvi_logdens_full_h16_supported(int D);
int gvi_h16_padded_dim(int D);
int gvi_split_h16_f32(const float* linv, int K, int D, void* hi, void* lo, float* tmax, void* stream);
int gvi_group_absmax_f32(const float* in, long long rows, int cols, int group, float* out, void* stream);
int gvi_logdens_full_h16_f32(const float* X, const float* tileinf, int N, int D, const float* means, const float* minf,
                             const void* linv_hi, const void* linv_lo, const float* tmax, const float* cst, int K,
                             float* lq, void* stream);
/* lq[k,n] = -D/2 log 2pi - sum log std_k - 1/2 sum_d ((mu_kd - x_nd)/std_kd)^2   models/diagonal_gmm.py:31-34,47-53 */
int gvi_logdens_diag_f32(const float* X, int N, int D, const float* means, const float* stds, int K, float* lq,
                         void* stream);
/* out[n] = logsumexp_k(lq[k,n] + logw[k])   models/gmm.py:199-201,214-216; sample_db.py:184-192 */
int gvi_mixture_lse_f32(const float* lq, const float* logw, int K, int N, float* out, void* stream);
/* grad[n,:] = -sum_k r_kn Sigma_k^-1 (x_n - mu_k), r = exp(lq + logw - logq): the analytic form of the
 * GradientTape in models/gmm.py:294-300.  prec = linv^T linv from gvi_prepare_full_f32. */
size_t gvi_mixture_grad_full_workspace(int N, int K);   /* bit mask of the components each 128-sample block touches */
int gvi_mixture_grad_full_f32(const float* X, int N, int D, const float* means, const float* prec, const float* lq,
                              const float* logw, const float* logq, int K, float* grad, void* ws, size_t ws_bytes,
                              void* stream);
int gvi_mixture_grad_diag_f32(const float* X, int N, int D, const float* means, const float* stds, const float* lq,
                              const float* logw, const float* logq, int K, float* grad, void* stream);

/* ---- importance weights -----------------------------------------------------------------------
 * Row-wise over lq[K,N] with background bg[N]:
 *   self_normalized=1: w = softmax_n(lq-bg) renormalised once more (quirk: ng_estimator.py:173-176,
 *                      weight_updater.py:60-63);   self_normalized=0: w = exp(lq-bg)/N (ng_estimator.py:147-152);
 *   rel_map != NULL ("only_use_own_samples", ng_estimator.py:113-118): w[k,n] = [rel_map[n]==k] / n_k.
 * Outputs (each nullable): W[K,N]; dot[k] = sum_n w[k,n] rho[n] (weight_updater.py:64,70);
 * ess[k] = 1 / sum_n softmax_n(lq-bg)^2 (sample_selector.py:154-158);
 * active[k, ceil(N/128)] = 1 when a 128-sample block carries weight above 1e-30 of the row maximum. */
int gvi_importance_weights_f32(const float* lq, const float* bg, const int32_t* rel_map, int K, int N,
                               int self_normalized, const float* rho, float* W, float* dot, float* ess,
                               uint8_t* active, void* stream);

/* The same computation split for sample-sharded (multi-GPU) runs: every rank holds a slice of the N samples,
 * the per-row maximum / sum of exponentials are all-reduced between the calls, and the weights are then formed
 * from the global normalisers:  w[k,n] = exp(lq[k,n] - bg[n] - lse[k]) * scale[k]  (scale nullable = 1). */
int gvi_row_max_f32(const float* lq, const float* bg, int K, int N, float* out, void* stream);
int gvi_row_sumexp_f32(const float* lq, const float* bg, int K, int N, const float* shift, float* out, void* stream);
int gvi_importance_weights_ext_f32(const float* lq, const float* bg, int K, int N, const float* lse,
                                   const float* scale, const float* rowmax, const float* rho, float* W, float* dot,
                                   uint8_t* active, void* stream);

/* ---- Stein natural-gradient statistics ---------------------------------------------------------
 * M[k] = sum_n W[k,n] (x_n-mu_k) G[n,:]^T  (D x D),  gneg[k] = -sum_n W[k,n] G[n,:]
 * then Hneg[k] = -sym(prec_k M[k]) (symmetrize=1, ng_estimator.py:183-187) or -(prec_k M[k])^T
 * (symmetrize=0, :164-168).  `active` (nullable) is the block mask from gvi_importance_weights_f32.
 * For 16 <= D <= 256 the [D x N] . [N x D] reduction runs on the tcgen05 tensor cores (2 x fp16 split precision,
 * weights and centring fused into the operand producer, accumulator drained with round-to-nearest adds). */
size_t gvi_stein_full_workspace(int N, int K, int D);
int gvi_stein_full_f32(const float* X, int N, int D, const float* means, const float* prec, const float* W,
                       const uint8_t* active, const float* G, int K, int symmetrize, float* Hneg, float* gneg,
                       void* ws, size_t ws_bytes, void* stream);
/* The two halves of gvi_stein_full_f32 for sample-sharded (multi-GPU) runs: M and gneg are sums over samples, so the ranks
 * reduce-scatter them by component ("reduce-scatter of (sum e g, sum e z g^T)") and every rank
 * finalises only the K / world components it updates.  M[K,D,D], gneg[K,D]; ws sizes from the *_workspace() calls. */
size_t gvi_stein_stats_full_workspace(int N, int K, int D);
int gvi_stein_stats_full_f32(const float* X, int N, int D, const float* means, const float* W, const uint8_t* active,
                             const float* G, int K, float* M, float* gneg, void* ws, size_t ws_bytes, void* stream);
size_t gvi_stein_finalize_full_workspace(int K, int D);
int gvi_stein_finalize_full_f32(const float* prec, const float* M, int K, int D, int symmetrize, float* Hneg, void* ws,
                                size_t ws_bytes, void* stream);
/* diagonal: Hneg[k,d] = -sum_n W (x-mu)_d / std_d^2 * G[n,d]   (ng_estimator.py:177-180).  With a workspace of
 * gvi_stein_diag_workspace(N, K, D) bytes the sums run as two matrix products W [X o G | G] on the GEMM engines (split over
 * the samples, partial products added in a fixed order); ws = NULL selects the serial per-(component, dimension) kernel. */
size_t gvi_stein_diag_workspace(int N, int K, int D);
int gvi_stein_diag_f32(const float* X, int N, int D, const float* means, const float* stds, const float* W,
                       const float* G, int K, float* Hneg, float* gneg, void* ws, size_t ws_bytes, void* stream);

/* ---- MORE natural-gradient estimator -----------------------------------------------------------
 * Weighted ridge regression on quadratic features of the whitened samples, per component (ng_estimator.py:296-376,
 * least_squares.py:34-191): quad[k] = reward_quad (= expected_hessian_neg), lin[k] = reward_lin.
 * W[K,N] importance weights, y[N] = target_lnpdf - log q, l2reg[K]; linv from gvi_prepare_full_f32.
 * Components are processed `chunk` at a time (workspace grows with chunk); ok[k] (caller-initialised to 1) is
 * cleared when the normal matrix of component k is not positive definite.
 * gvi_more_tensor_cores() != 0: the normal matrix and the trailing updates of its blocked Cholesky run on the tcgen05
 * tensor cores, features written pre-split and transposed: 2 = 2 x fp16 split (default), 1 = 3xTF32
 * (GMMVI_B200_MORE_TC=tf32); GMMVI_B200_MORE_TC=0 selects the SIMT fp32 engine.  The workspace size depends on that
 * switch. */
int gvi_more_tensor_cores(void);
size_t gvi_more_workspace(int chunk, int N, int D);
int gvi_more_fit_f32(const float* X, int N, int D, const float* means, const float* linv, const float* W,
                     const float* y, const float* l2reg, int K, int chunk, float* quad, float* lin, int32_t* ok,
                     void* ws, size_t ws_bytes, void* stream);

/* ---- component updates -------------------------------------------------------------------------
 * mode 0: KL-constrained (ng_based_component_updater.py:431-524; bisection :335-429, kl :244-333)
 * mode 1: direct NG step  (:97-141)        mode 2: iBLR (:160-223)
 * stepsizes[K] = eps (mode 0) or s (modes 1,2); last_etas[K] (mode 0, <0 => cold bracket);
 * num_updates[K] (mode 2: no mean step on a component's first update).
 * Outputs: new means/chols (old ones when success[k]==0), etas[k] / kls[k] (mode 0; -1 on failure). */
size_t gvi_update_full_workspace(int K, int D);
/* Householder tridiagonalisation T = P^T S P, hp = P^T h of S[i][j] = S[j][i] = B[min(i,j)][max(i,j)] (D <= 256): the
 * mode-0 update evaluates KL(eta) of its bisection (:244-333, :335-429) from (d, e, hp) in O(D) per eta.  d[K,D]
 * diagonal, e[K,D] sub-diagonal (e[k][D-1] = 0). */
int gvi_tridiag_f32(const float* B, const float* h, int K, int D, float* d, float* e, float* hp, void* stream);
int gvi_update_full_f32(int mode, const float* means, const float* chols, const float* Hneg, const float* gneg,
                        const float* stepsizes, const float* last_etas, const float* num_updates, int K, int D,
                        float temperature, float* out_means, float* out_chols, int32_t* success, float* etas,
                        float* kls, int32_t* evals /* nullable: KL evaluations spent per component */, void* ws,
                        size_t ws_bytes, void* stream);
int gvi_update_diag_f32(int mode, const float* means, const float* stds, const float* Hneg, const float* gneg,
                        const float* stepsizes, const float* last_etas, const float* num_updates, int K, int D,
                        float temperature, float* out_means, float* out_stds, int32_t* success, float* etas,
                        float* kls, void* stream);

/* ---- weight updates ----------------------------------------------------------------------------
 * direct (weight_updater.py:136-141) / trust region (:164-279).  `stepsize` is a DEVICE scalar
 * (the weight stepsize, i.e. the KL bound for the trust-region variant).  info[0]=kl, info[1]=eta. */
int gvi_weight_update_f32(int trust_region, const float* logw, const float* elr, int K, const float* stepsize,
                          float temperature, float* out_logw, float* info, void* stream);

/* ---- sampling -----------------------------------------------------------------------------------
 * Counter-based N(0,1) noise: element (row, d) depends only on (seed, subsequence, row_offset+row, d),
 * so a shard draws exactly the rows it owns regardless of the number of GPUs. */
int gvi_fill_normal_f32(float* out, long long rows, int D, unsigned long long seed,
                        unsigned long long subsequence, long long row_offset, void* stream);
/* Same generator with the draw counter in DEVICE memory: subsequence = *subsequence_dev + subsequence_add.  A captured CUDA
 * graph of the iteration replays with fresh noise by incrementing the counter on the device. */
int gvi_fill_normal_dev_f32(float* out, long long rows, int D, unsigned long long seed,
                            const unsigned long long* subsequence_dev, unsigned long long subsequence_add,
                            long long row_offset, void* stream);
/* x = mu_k + L_k eps for the rows [offsets[k], offsets[k+1]) of eps/X[N,D]; mapping[n]=k
 * (models/gmm.py:361-386, full_cov_gmm.py:36-39, diagonal_gmm.py:43-45).  offsets[K+1] is a device
 * prefix sum; max_rows_per_component (host) bounds the grid. */
int gvi_sample_f32(int diagonal, const float* eps, const int32_t* offsets, const float* means, const float* chols,
                   int K, int D, int max_rows_per_component, float* X, int32_t* mapping, void* stream);

/* ---- batched GEMM used by the estimators / updaters (also exported for tests) --------------------
 * C[b] = alpha * opA(A[b]) * opB(B[b])  with A[b] (M x Kd, transA=0) or (Kd x M, transA=1), likewise B. */
int gvi_bgemm_f32(int transA, int transB, int batch, int M, int N, int Kd, float alpha, const float* A, int lda,
                  long long strideA, const float* B, int ldb, long long strideB, float* C, int ldc,
                  long long strideC, void* stream);

/* The same product on the tcgen05 tensor cores in 3xTF32 split precision (fp32-grade result); `ws` holds the
 * TF32 hi / lo copies of both operands.  The estimator / updater entry points use it internally when the shape
 * allows (K % 4 == 0) and fall back to the SIMT engine otherwise. */
int gvi_tc_bgemm_supported(int M, int N, int Kd);
size_t gvi_tc_bgemm_workspace(int batch, int M, int N, int Kd);
int gvi_tc_bgemm_f32(int transA, int transB, int batch, int M, int N, int Kd, float alpha, const float* A, int lda,
                     long long strideA, const float* B, int ldb, long long strideB, float* C, int ldc,
                     long long strideC, void* ws, size_t ws_bytes, void* stream);
/* C[b] = alpha * opA(A[b]) opB(B[b]) + beta * C[b] (beta 0 or 1) with the reduction cut into segments of
 * `kseg_kblocks` blocks of 32 that are added to C with round-to-nearest adds (the tensor core's accumulator
 * truncates; 0 = one segment), and with `lower_only` != 0 only the 128 x 256 tiles that touch the lower triangle
 * written: the normal-equation build Phi^T W Phi and the trailing updates of the blocked Cholesky of the MORE
 * estimator (least_squares.py:60-75). */
int gvi_tc_bgemm_ex_f32(int transA, int transB, int batch, int M, int N, int Kd, float alpha, const float* A, int lda,
                        long long strideA, const float* B, int ldb, long long strideB, float* C, int ldc,
                        long long strideC, float beta, int kseg_kblocks, int lower_only, void* ws, size_t ws_bytes,
                        void* stream);

/* The same product in the "2 x fp16" split precision of the log-density kernel (kind::f16 at twice the TF32 rate, half
 * the operand bytes): each operand batch entry is scaled by a power of two derived from its largest magnitude and split
 * into fp16 hi / lo parts.  Dense A [b][M][Kd], B [b][N][Kd] (C = A B^T), C [b][M][N]; Kd % 8 == 0; kseg_kblocks counts
 * blocks of 64.  The MORE estimator uses this kernel with operands it writes pre-split. */
size_t gvi_tc_bgemm_h16_workspace(int batch, int M, int N, int Kd);
int gvi_tc_bgemm_h16_f32(int batch, int M, int N, int Kd, float alpha, const float* A, const float* B, float* C,
                         float beta, int kseg_kblocks, int lower_only, void* ws, size_t ws_bytes, void* stream);

/* ---- tensor-core mixture gradient (same contraction as gvi_mixture_grad_full_f32) ------------------------------
 * grad[n, :] = -sum_k r_kn P_k (x_n - mu_k) with tcgen05 MMAs in the 2 x fp16 split precision of the log-density kernel.
 * p_hi / p_lo / tmaxp: gvi_split_h16_full_f32(prec); tileinf[ceil(N/128)] / minf[K]: gvi_group_absmax_f32 of X (groups of
 * 128 rows) and of means; ws: gvi_mixture_grad_full_workspace(N, K) bytes.  D % 4 == 0, 64 < D <= 128 or 192 < D <= 256. */
int gvi_mixture_grad_full_h16_supported(int D);
int gvi_split_h16_full_f32(const float* mat, int K, int D, void* hi, void* lo, float* tmax, void* stream);
int gvi_mixture_grad_full_h16_f32(const float* X, const float* tileinf, int N, int D, const float* means,
                                  const float* minf, const void* p_hi, const void* p_lo, const float* tmaxp,
                                  const float* lq, const float* logw, const float* logq, int K, float* grad, void* ws,
                                  size_t ws_bytes, void* stream);

/* ---- MMD evaluation (experiments/evaluation/mmd.py:41-60) -------------------------------------------------
 * sum_{i<n1, j<n2} exp(-sum_d w[d] (X[i,d] - Y[j,d])^2): compute_ustat (Y = X) and kernel_mix of the reference with
 * the diagonal bandwidth w = 1 / (alpha sigma).  The result is the sum of partial[0 .. gvi_gauss_kernel_sum_partials). */
size_t gvi_gauss_kernel_sum_partials(int n1, int n2);
int gvi_gauss_kernel_sum_f32(const float* X, int n1, const float* Y, int n2, int D, const float* w, double* partial,
                             void* stream);

/* ---- direct / iBLR update for a NON-SYMMETRIC -E[H] (ng_based_component_updater.py:97-141, 160-223) ----------------------
 * The Stein estimator does not symmetrise its Hessian estimate with standard importance weights (ng_estimator.py:168 vs
 * :186); the reference then inverts the general matrix P' = P + s R (iBLR: + s^2/2 R Sigma R) with an LU factorisation
 * (tf.linalg.inv / tf.linalg.solve, :116-117, :199) and factors the LOWER triangle of the inverse (tf.linalg.cholesky, :118,
 * :200).  Restated literally: Gauss-Jordan with partial pivoting + Cholesky per component.  mode 1 = direct, 2 = iBLR;
 * prec[K,D,D] = old precisions; success[k] = 0 keeps the old parameters (singular P' or non-positive pivot). */
size_t gvi_update_full_general_workspace(int K, int D);
int gvi_update_full_general_f32(int mode, const float* means, const float* chols, const float* prec, const float* Hneg,
                                const float* gneg, const float* stepsizes, const float* num_updates, int K, int D,
                                float* out_means, float* out_chols, int32_t* success, void* ws, size_t ws_bytes,
                                void* stream);

/* ---- construction-time Cholesky (models/full_cov_gmm.py:23, :67: tf.linalg.cholesky(covs) in the constructor and in
 * add_component) -- A[K,D,D] symmetric (lower triangle read) -> L[K,D,D] lower, fp64 arithmetic rounded to fp32; a matrix
 * that is not positive definite gives a NaN-filled factor (TensorFlow's behaviour) and ok[k] = 0 (ok nullable). */
size_t gvi_cholesky_workspace(int K, int D);
int gvi_cholesky_f32(const float* A, int K, int D, float* L, int32_t* ok, void* ws, size_t ws_bytes, void* stream);

/* ---- planar-robot target (experiments/target_distributions/planar_robot.py:29-66; BASELINE config C2) ----------------
 * theta[N,D] joint angles (D = number of links <= 64) -> lnpdf[n] = N(theta_n; 0, diag(prior_stds^2)) + max over the G <= 8
 * goals[G,2] of N(forward_kinematics(theta_n); goal, likelihood_std^2 I)  (:49-53, :57-66), and, when grad != NULL,
 * grad[N,D] = its gradient through the arg-max goal (replaces the tf.GradientTape of sample_selector.py:73-77).
 * link_lengths[D] may be NULL (all ones, :35). */
int gvi_planar_robot_f32(const float* theta, int N, int D, const float* prior_stds, const float* link_lengths,
                         const float* goals, int G, float likelihood_std, float* lnpdf, float* grad, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GMMVI_B200_H */
