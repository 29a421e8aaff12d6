// Direct / iBLR component update for a NON-SYMMETRIC -E[H] (ng_based_component_updater.py:97-141, 160-223).
//
// The Stein estimator symmetrises its Hessian estimate only in the self-normalised branch (ng_estimator.py:186, not :168),
// so with standard importance weights the direct and iBLR updaters receive a non-symmetric R.  The reference then forms
//     P' = P + s R                     (direct)        P' = P + s (R + s/2 R Sigma R)     (iBLR)
// as a GENERAL matrix, inverts it with an LU factorisation (tf.linalg.inv, :117 / :199; the direct mean is
// tf.linalg.solve(P', q'), :116), and takes tf.linalg.cholesky of that inverse, which reads only its lower triangle.
// The whitened-frame kernels (update.cu, update_blocked.cu) assume a symmetric step and cannot reproduce this; this file
// restates the reference's sequence literally: Gauss-Jordan inversion with partial pivoting, Cholesky of the lower
// triangle of the inverse, one CTA per component, matrices in a global (L2-resident) workspace.  It is a corner path
// (no default configuration of the reference combines standard importance weights with these updaters) and is written
// for fidelity, not speed: ~1 ms per component at D = 256.
#include "common.cuh"

namespace gvi {

int launch_gemm_auto(int transA, int transB, int batch, int M, int N, int Kd, float alpha, const float* A, int lda,
                     long long strideA, const float* B, int ldb, long long strideB, float* C, int ldc,
                     long long strideC, float* ws, size_t ws_floats, cudaStream_t st);
size_t tc_gemm_workspace_floats(int batch, int M, int N, int Kd);

constexpr int UG_THREADS = 256;

// A (global, D x D row-major) <- inverse of A by Gauss-Jordan elimination with partial (row) pivoting.  Returns false
// (uniformly) when a pivot is zero or not finite.  perm: shared int[D].
__device__ bool gauss_jordan_inverse(float* __restrict__ A, int D, int* perm, float* red_val, int* red_idx) {
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  for (int k = 0; k < D; ++k) {
    // pivot search over rows >= k of column k (first maximum wins, like LAPACK's isamax)
    float best = -1.f;
    int bi = k;
    for (int i = k + tid; i < D; i += nt) {
      const float v = fabsf(A[(long long)i * D + k]);
      if (v > best || !(v == v)) { best = (v == v) ? v : INFINITY; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { red_val[warp] = best; red_idx[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      best = lane < nw ? red_val[lane] : -1.f;
      bi = lane < nw ? red_idx[lane] : D;
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (lane == 0) { red_val[0] = best; red_idx[0] = bi; }
    }
    __syncthreads();
    const int p = red_idx[0];
    const float pv = red_val[0];
    if (!(pv > 0.f) || !isfinite(pv)) return false;
    if (tid == 0) perm[k] = p;
    if (p != k) {
      for (int j = tid; j < D; j += nt) {
        const float a = A[(long long)k * D + j], b = A[(long long)p * D + j];
        A[(long long)k * D + j] = b;
        A[(long long)p * D + j] = a;
      }
    }
    __syncthreads();
    const float piv = A[(long long)k * D + k];
    __syncthreads();
    const float ipiv = 1.f / piv;
    for (int j = tid; j < D; j += nt) A[(long long)k * D + j] = (j == k) ? ipiv : A[(long long)k * D + j] * ipiv;
    __syncthreads();
    // eliminate column k from every other row: a warp per row, lanes over the columns (coalesced)
    for (int i = warp; i < D; i += nw) {
      if (i == k) continue;
      float* ri = A + (long long)i * D;
      const float f = ri[k];
      __syncwarp();
      const float* rk = A + (long long)k * D;
      for (int j = lane; j < D; j += 32) ri[j] = (j == k) ? -f * rk[k] : fmaf(-f, rk[j], ri[j]);
    }
    __syncthreads();
  }
  // undo the row interchanges as column interchanges, last to first
  for (int k = D - 1; k >= 0; --k) {
    const int p = perm[k];
    if (p != k) {
      for (int i = tid; i < D; i += nt) {
        const float a = A[(long long)i * D + k], b = A[(long long)i * D + p];
        A[(long long)i * D + k] = b;
        A[(long long)i * D + p] = a;
      }
    }
    __syncthreads();
  }
  return true;
}

// In-place Cholesky of the LOWER triangle of A (global, row-major), left-looking, one column at a time.
__device__ bool cholesky_lower_global(float* __restrict__ A, int D) {
  const int tid = threadIdx.x, nt = blockDim.x;
  __shared__ float s_diag;
  for (int j = 0; j < D; ++j) {
    const float* rj = A + (long long)j * D;
    for (int i = j + tid; i < D; i += nt) {
      float* ri = A + (long long)i * D;
      float s = ri[j];
      for (int m = 0; m < j; ++m) s = fmaf(-ri[m], rj[m], s);
      ri[j] = s;                                  // row j (i == j) only reads columns < j of itself: no hazard
    }
    __syncthreads();
    if (tid == 0) s_diag = A[(long long)j * D + j];
    __syncthreads();
    const float d = s_diag;
    if (!(d > 0.f) || !isfinite(d)) return false;
    const float r = sqrtf(d), ir = 1.f / r;
    for (int i = j + tid; i < D; i += nt) A[(long long)i * D + j] = (i == j) ? r : A[(long long)i * D + j] * ir;
    __syncthreads();
  }
  return true;
}

__global__ void __launch_bounds__(UG_THREADS)
update_general_kernel(int mode, const float* __restrict__ means, const float* __restrict__ chols,
                      const float* __restrict__ prec, const float* __restrict__ R, const float* __restrict__ RSR,
                      const float* __restrict__ gneg, const float* __restrict__ stepsizes,
                      const float* __restrict__ num_updates, int D, float* __restrict__ work,
                      float* __restrict__ out_means, float* __restrict__ out_chols, int32_t* __restrict__ success) {
  extern __shared__ float sm[];
  __shared__ float red_val[32];
  __shared__ int red_idx[32];
  const int k = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const long long DD = (long long)D * D;
  float* A = work + k * DD;
  const float* P = prec + k * DD;
  const float* Rk = R + k * DD;
  const float* L = chols + k * DD;
  const float* mu = means + (long long)k * D;
  const float* g = gneg + (long long)k * D;
  float* q = sm;                     // [D] new_lin (direct) / Sigma g (iBLR)
  int* perm = reinterpret_cast<int*>(sm + D);
  const float s = stepsizes[k];
  // new_precision (general)
  for (long long e = tid; e < DD; e += nt) {
    float d = Rk[e];
    if (mode == 2) d = fmaf(0.5f * s, RSR[k * DD + e], d);      // R + s/2 R Sigma R  (:176-181)
    A[e] = fmaf(s, d, P[e]);
  }
  if (mode == 1) {
    // new_lin = P mu + s (R mu - g)   (:109-113)
    for (int i = tid; i < D; i += nt) {
      float a = 0.f, b = 0.f;
      for (int j = 0; j < D; ++j) {
        a = fmaf(P[(long long)i * D + j], mu[j], a);
        b = fmaf(Rk[(long long)i * D + j], mu[j], b);
      }
      q[i] = fmaf(s, b - g[i], a);
    }
  } else {
    // Sigma g = L (L^T g)
    float* t = reinterpret_cast<float*>(perm) + D;      // [D] scratch after perm
    for (int j = tid; j < D; j += nt) {
      float a = 0.f;
      for (int i = j; i < D; ++i) a = fmaf(L[(long long)i * D + j], g[i], a);
      t[j] = a;
    }
    __syncthreads();
    for (int i = tid; i < D; i += nt) {
      float a = 0.f;
      for (int j = 0; j <= i; ++j) a = fmaf(L[(long long)i * D + j], t[j], a);
      q[i] = a;
    }
  }
  __syncthreads();
  bool ok = gauss_jordan_inverse(A, D, perm, red_val, red_idx);          // A = new_cov (general)
  float* om = out_means + (long long)k * D;
  float* oc = out_chols + k * DD;
  if (ok) {
    if (mode == 1) {
      for (int i = tid; i < D; i += nt) {        // new_mean = new_cov new_lin (the reference solves P' x = q', :116)
        float a = 0.f;
        for (int j = 0; j < D; ++j) a = fmaf(A[(long long)i * D + j], q[j], a);
        om[i] = a;
      }
    } else {
      const bool first = num_updates[k] == 0.f;  // iBLR: no mean update on a component's first update (:184-186)
      for (int i = tid; i < D; i += nt) om[i] = first ? mu[i] : fmaf(-s, q[i], mu[i]);
    }
    __syncthreads();
    ok = cholesky_lower_global(A, D);
  }
  bool finite = true;
  if (ok) {
    for (long long e = tid; e < DD; e += nt) {
      const int i = (int)(e / D), j = (int)(e % D);
      const float v = (j <= i) ? A[e] : 0.f;
      finite &= isfinite(v);
      oc[e] = v;
    }
    for (int i = tid; i < D; i += nt) finite &= isfinite(om[i]);
  }
  ok = ok && !__syncthreads_or(!finite);
  if (!ok) {                                     // NaN in the new factor => keep the old parameters (:120-123, 202-205)
    for (int i = tid; i < D; i += nt) om[i] = mu[i];
    for (long long e = tid; e < DD; e += nt) oc[e] = L[e];
  }
  if (tid == 0) success[k] = ok ? 1 : 0;
}

}  // namespace gvi

using namespace gvi;

extern "C" size_t gvi_update_full_general_workspace(int K, int D) {
  if (K <= 0) return 0;
  const size_t kdd = (size_t)K * D * D;
  return (3 * kdd + tc_gemm_workspace_floats(K, D, D, D) + 64) * sizeof(float);
}

extern "C" int gvi_update_full_general_f32(int mode, const float* means, const float* chols, const float* prec,
                                           const float* Hneg, const float* gneg, const float* stepsizes,
                                           const float* num_updates, int K, int D, float* out_means, float* out_chols,
                                           int32_t* success, void* ws, size_t ws_bytes, void* stream) {
  GVI_REQUIRE(mode == 1 || mode == 2, "gvi_update_full_general_f32: mode %d (1 = direct, 2 = iBLR)", mode);
  GVI_REQUIRE(K >= 0 && D > 0, "gvi_update_full_general_f32: bad sizes");
  if (K == 0) return GVI_OK;
  GVI_REQUIRE(means && chols && prec && Hneg && gneg && stepsizes && out_means && out_chols && success && ws,
              "gvi_update_full_general_f32: null pointer");
  GVI_REQUIRE(mode != 2 || num_updates, "gvi_update_full_general_f32: num_updates required for iBLR");
  GVI_REQUIRE(K <= 65535 && D <= 2048, "gvi_update_full_general_f32: K or D too large");
  if (ws_bytes < gvi_update_full_general_workspace(K, D)) {
    set_last_error("gvi_update_full_general_f32: workspace %zu < %zu", ws_bytes, gvi_update_full_general_workspace(K, D));
    return GVI_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const long long DD = (long long)D * D;
  float* work = (float*)ws;
  float* T = work + (size_t)K * DD;
  float* RSR = T + (size_t)K * DD;
  float* tcws = RSR + (size_t)K * DD;
  const size_t tcws_floats = tc_gemm_workspace_floats(K, D, D, D);
  if (mode == 2) {
    // R Sigma R = (R L)(L^T R): T = R L, work = L^T R, RSR = T work
    int rc = launch_gemm_auto(0, 0, K, D, D, D, 1.f, Hneg, D, DD, chols, D, DD, T, D, DD, tcws, tcws_floats, st);
    if (rc) return rc;
    rc = launch_gemm_auto(1, 0, K, D, D, D, 1.f, chols, D, DD, Hneg, D, DD, work, D, DD, tcws, tcws_floats, st);
    if (rc) return rc;
    rc = launch_gemm_auto(0, 0, K, D, D, D, 1.f, T, D, DD, work, D, DD, RSR, D, DD, tcws, tcws_floats, st);
    if (rc) return rc;
  }
  const size_t smem = (size_t)3 * D * sizeof(float);
  update_general_kernel<<<K, UG_THREADS, smem, st>>>(mode, means, chols, prec, Hneg, mode == 2 ? RSR : nullptr, gneg,
                                                     stepsizes, num_updates, D, work, out_means, out_chols, success);
  return check_launch("update_general_kernel");
}

// ---- construction-time Cholesky (models/full_cov_gmm.py:23, :67: tf.linalg.cholesky of user-supplied covariances) --------
// One CTA per matrix, fp64 arithmetic on a global (L2-resident) scratch copy, left-looking, result rounded to fp32 -- the
// same factor a host LAPACK dpotrf would give, without the device -> host -> device round trip the adaptive runs paid on
// every component addition.  ok[k] = 0 and a NaN-filled factor (TensorFlow's behaviour) when a pivot is not positive.
namespace gvi {
__global__ void __launch_bounds__(256)
cholesky_f64_kernel(const float* __restrict__ A, int D, double* __restrict__ work, float* __restrict__ L,
                    int32_t* __restrict__ ok) {
  const int k = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const long long DD = (long long)D * D;
  double* W = work + k * DD;
  const float* Ak = A + k * DD;
  for (long long e = tid; e < DD; e += nt) W[e] = (double)Ak[e];
  __syncthreads();
  __shared__ double s_diag;
  bool good = true;
  for (int j = 0; j < D && good; ++j) {
    const double* rj = W + (long long)j * D;
    for (int i = j + tid; i < D; i += nt) {
      double* ri = W + (long long)i * D;
      double s = ri[j];
      for (int m = 0; m < j; ++m) s -= ri[m] * rj[m];
      ri[j] = s;
    }
    __syncthreads();
    if (tid == 0) s_diag = W[(long long)j * D + j];
    __syncthreads();
    const double d = s_diag;
    if (!(d > 0.0) || !isfinite(d)) { good = false; break; }
    const double r = sqrt(d), ir = 1.0 / r;
    for (int i = j + tid; i < D; i += nt) W[(long long)i * D + j] = (i == j) ? r : W[(long long)i * D + j] * ir;
    __syncthreads();
  }
  float* Lk = L + k * DD;
  for (long long e = tid; e < DD; e += nt) {
    const int i = (int)(e / D), j = (int)(e % D);
    Lk[e] = good ? (j <= i ? (float)W[e] : 0.f) : __int_as_float(0x7fc00000);
  }
  if (tid == 0 && ok != nullptr) ok[k] = good ? 1 : 0;
}
}  // namespace gvi

extern "C" size_t gvi_cholesky_workspace(int K, int D) { return K > 0 ? (size_t)K * D * D * sizeof(double) : 0; }
extern "C" int gvi_cholesky_f32(const float* A, int K, int D, float* L, int32_t* ok, void* ws, size_t ws_bytes, void* stream) {
  GVI_REQUIRE(K >= 0 && D > 0, "gvi_cholesky_f32: bad sizes");
  if (K == 0) return GVI_OK;
  GVI_REQUIRE(A && L && ws, "gvi_cholesky_f32: null pointer");
  GVI_REQUIRE(K <= 65535, "gvi_cholesky_f32: K=%d exceeds 65535", K);
  if (ws_bytes < gvi_cholesky_workspace(K, D)) {
    set_last_error("gvi_cholesky_f32: workspace %zu < %zu", ws_bytes, gvi_cholesky_workspace(K, D));
    return GVI_ERR_WORKSPACE;
  }
  cholesky_f64_kernel<<<K, 256, 0, (cudaStream_t)stream>>>(A, D, (double*)ws, L, ok);
  return check_launch("cholesky_f64_kernel");
}
