// MORE natural-gradient estimator: weighted ridge regression on quadratic features, all on device.
//
// Reference: optimization/gmmvi_modules/ng_estimator.py:296-376 (MoreNgEstimator) and
// optimization/least_squares.py:34-76 (RegressionFunc.fit), :113-124 (QuadFunc._feature_fn), :126-191 (fit_quadratic).
//
// Per component k (processed in chunks of Kc components):
//   Z   = (X - mu_k) Linv_k^T                                  whitening (least_squares.py:170-173)
//   Phi = [ z_i z_j (i <= j, row-major) | z | 1 | y ]            [N, F+1]; the reward y is carried as an extra column
//   A'  = Phi^T diag(w_k) Phi  (+ lambda_k on the first F-1 diagonal entries; the bias is not regularised)
//         so A'[:F,:F] is the normal matrix and A'[F,:F] the right-hand side       (least_squares.py:60-75)
//   blocked Cholesky of A'[:F,:F] carried through row F gives  v = L^-1 b  in A'[F,:F];  L^T theta = v by back substitution
//   Q_z = -(T + T^T) with T = upper-triangular scatter of theta, r_z, then the un-whitening of :184-189.
// The reference solves the (symmetric positive definite) system with an LU (tf.linalg.solve); a Cholesky factorisation
// gives the same solution and reports a non-positive pivot through ok[k] instead of returning garbage.
#include "common.cuh"
#include "../../include/gmmvi_b200.h"

namespace gvi {

int launch_bgemm_ex(int transA, int transB, int batch, int M, int N, int Kd, float alpha, const float* A, int lda,
                    long long strideA, const float* B, int ldb, long long strideB, float* C, int ldc,
                    long long strideC, const float* scaleK, long long strideScale, float beta, int lower_only,
                    cudaStream_t st);

namespace more {

constexpr int NB = 128;       // Cholesky panel width

// Xc[kc][n][:] = X[n][:] - mu[k0 + kc][:]
__global__ void center_kernel(const float* __restrict__ X, const float* __restrict__ means, int N, int D, int k0,
                              float* __restrict__ Xc) {
  const int kc = blockIdx.y;
  const float* mu = means + (long long)(k0 + kc) * D;
  float* out = Xc + (long long)kc * N * D;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)N * D;
       e += (long long)gridDim.x * blockDim.x)
    out[e] = X[e] - mu[e % D];
}

// Phi[kc][n][f]: quadratic (row-major upper triangle), linear, constant, reward
__global__ void __launch_bounds__(256)
features_kernel(const float* __restrict__ Z, const float* __restrict__ y, int N, int D, int F, float* __restrict__ Phi) {
  extern __shared__ float zs[];      // [8][D]
  const int kc = blockIdx.y;
  const int n0 = blockIdx.x * 8;
  const float* Zk = Z + (long long)kc * N * D;
  for (int e = threadIdx.x; e < 8 * D; e += blockDim.x) {
    const int r = e / D, d = e % D;
    zs[e] = (n0 + r < N) ? Zk[(long long)(n0 + r) * D + d] : 0.f;
  }
  __syncthreads();
  const int nq = D * (D + 1) / 2;
  const int Fa = F + 1;
  float* out = Phi + (long long)kc * N * Fa;
  for (int f = threadIdx.x; f < Fa; f += blockDim.x) {
    int i = -1, j = -1;     // i == -1: not a quadratic feature
    if (f < nq) {
      // f = i*D - i(i-1)/2 + (j - i)
      const float b = 2.f * D + 1.f;
      i = (int)floorf((b - sqrtf(fmaxf(b * b - 8.f * f, 0.f))) * 0.5f);
      while (i > 0 && i * D - i * (i - 1) / 2 > f) --i;
      while ((i + 1) * D - (i + 1) * i / 2 <= f) ++i;
      j = i + (f - (i * D - i * (i - 1) / 2));
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int n = n0 + r;
      if (n >= N) break;
      float v;
      if (f < nq) v = zs[r * D + i] * zs[r * D + j];
      else if (f < nq + D) v = zs[r * D + (f - nq)];
      else if (f == F - 1) v = 1.f;
      else v = y[n];
      out[(long long)n * Fa + f] = v;
    }
  }
}

__global__ void ridge_kernel(float* __restrict__ A, int Fa, int F, const float* __restrict__ l2, int k0) {
  const int kc = blockIdx.y;
  float* Ak = A + (long long)kc * Fa * Fa;
  const float lam = l2[k0 + kc];
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < F - 1; f += gridDim.x * blockDim.x)
    Ak[(long long)f * Fa + f] += lam;
}

// Diagonal block [nb x nb] at (p, p): in-place Cholesky (fp64 accumulation) and its inverse into Sinv[kc][NB][NB].
__global__ void __launch_bounds__(256)
potrf_inv_kernel(float* __restrict__ A, int Fa, int p, int nb, float* __restrict__ Sinv, int32_t* __restrict__ ok,
                 int k0) {
  extern __shared__ float sm[];          // L[nb][NB+1], Y[nb][NB+1]
  float* L = sm;
  float* Y = sm + NB * (NB + 1);
  const int kc = blockIdx.x;
  float* Ak = A + (long long)kc * Fa * Fa + (long long)p * Fa + p;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < nb * nb; e += nt) {
    const int i = e / nb, j = e % nb;
    L[i * (NB + 1) + j] = (j <= i) ? Ak[(long long)i * Fa + j] : 0.f;
  }
  __syncthreads();
  __shared__ int bad;
  if (tid == 0) bad = 0;
  for (int j = 0; j < nb; ++j) {
    __syncthreads();
    for (int i = j + tid; i < nb; i += nt) {
      double s = (double)L[i * (NB + 1) + j];
      for (int m = 0; m < j; ++m) s -= (double)L[i * (NB + 1) + m] * (double)L[j * (NB + 1) + m];
      L[i * (NB + 1) + j] = (float)s;     // unscaled
    }
    __syncthreads();
    float piv = L[j * (NB + 1) + j];
    if (!(piv > 0.f) || !isfinite(piv)) {
      if (tid == 0) bad = 1;
      piv = 1.f;
    }
    const float c = sqrtf(piv);
    __syncthreads();
    for (int i = j + tid; i < nb; i += nt) L[i * (NB + 1) + j] = (i == j) ? c : L[i * (NB + 1) + j] / c;
  }
  __syncthreads();
  // inverse, thread per column
  for (int c = tid; c < nb; c += nt) {
    for (int i = 0; i < c; ++i) Y[i * (NB + 1) + c] = 0.f;
    Y[c * (NB + 1) + c] = 1.f / L[c * (NB + 1) + c];
    for (int i = c + 1; i < nb; ++i) {
      double s = 0.0;
      for (int m = c; m < i; ++m) s += (double)L[i * (NB + 1) + m] * (double)Y[m * (NB + 1) + c];
      Y[i * (NB + 1) + c] = (float)(-s / (double)L[i * (NB + 1) + i]);
    }
  }
  __syncthreads();
  float* So = Sinv + (long long)kc * NB * NB;
  for (int e = tid; e < nb * nb; e += nt) {
    const int i = e / nb, j = e % nb;
    Ak[(long long)i * Fa + j] = L[i * (NB + 1) + j];
    So[i * NB + j] = Y[i * (NB + 1) + j];
  }
  if (tid == 0 && bad && ok) ok[k0 + kc] = 0;
}

// A[kc][r0 + r][p + c] = T[kc][r][c]
__global__ void copy_panel_kernel(const float* __restrict__ T, int ldt, long long strideT, float* __restrict__ A, int Fa,
                                  int r0, int p, int rows, int nb) {
  const int kc = blockIdx.y;
  const float* Tk = T + kc * strideT;
  float* Ak = A + (long long)kc * Fa * Fa;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)rows * nb;
       e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / nb), c = (int)(e % nb);
    Ak[(long long)(r0 + r) * Fa + p + c] = Tk[(long long)r * ldt + c];
  }
}

// Back substitution L^T theta = v (v = row F of the factored augmented matrix), then the coefficients are unpacked:
// Qz = -(T + T^T), rz.  One CTA per component.
__global__ void __launch_bounds__(1024)
backsolve_unpack_kernel(const float* __restrict__ A, int Fa, int F, int D, float* __restrict__ theta_ws,
                        float* __restrict__ Qz, float* __restrict__ rz) {
  extern __shared__ float v[];       // [F]
  __shared__ float ts;
  const int kc = blockIdx.x;
  const float* Ak = A + (long long)kc * Fa * Fa;
  float* theta = theta_ws + (long long)kc * F;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int f = tid; f < F; f += nt) v[f] = Ak[(long long)F * Fa + f];
  __syncthreads();
  for (int i = F - 1; i >= 0; --i) {
    const float* Li = Ak + (long long)i * Fa;
    if (tid == 0) {
      ts = v[i] / Li[i];
      theta[i] = ts;
    }
    __syncthreads();
    const float t = ts;
    for (int m = tid; m < i; m += nt) v[m] = fmaf(-Li[m], t, v[m]);
    __syncthreads();
  }
  const int nq = D * (D + 1) / 2;
  float* Q = Qz + (long long)kc * D * D;
  for (int e = tid; e < D * D; e += nt) {
    const int a = e / D, b = e % D;
    const int i = min(a, b), j = max(a, b);
    const float t = theta[i * D - i * (i - 1) / 2 + (j - i)];
    Q[e] = (a == b) ? -2.f * t : -t;          // -(T + T^T) doubles the diagonal (quirk 9)
  }
  for (int d = tid; d < D; d += nt) rz[(long long)kc * D + d] = theta[nq + d];
}

// lin = Linv^T rz + quad mu
__global__ void __launch_bounds__(256)
unwhiten_lin_kernel(const float* __restrict__ linv, const float* __restrict__ means, const float* __restrict__ quad,
                    const float* __restrict__ rz, int D, int k0, float* __restrict__ lin) {
  const int kc = blockIdx.x, k = k0 + kc;
  const float* Li = linv + (long long)k * D * D;
  const float* mu = means + (long long)k * D;
  const float* Q = quad + (long long)k * D * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float t1 = 0.f, t2 = 0.f;
    for (int i = d; i < D; ++i) t1 = fmaf(Li[(long long)i * D + d], rz[(long long)kc * D + i], t1);
    for (int j = 0; j < D; ++j) t2 = fmaf(Q[(long long)d * D + j], mu[j], t2);
    lin[(long long)k * D + d] = t1 + t2;
  }
}

struct Layout {
  size_t xc, z, phi, a, sinv, t21, theta, qz, rz, t1, total;
};
static Layout layout(int Kc, int N, int D) {
  const size_t F = (size_t)D * (D + 1) / 2 + D + 1, Fa = F + 1;
  Layout l;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 63) / 64 * 64; return r; };
  l.xc = take((size_t)Kc * N * D);
  l.z = take((size_t)Kc * N * D);
  l.phi = take((size_t)Kc * N * Fa);
  l.a = take((size_t)Kc * Fa * Fa);
  l.sinv = take((size_t)Kc * NB * NB);
  l.t21 = take((size_t)Kc * Fa * NB);
  l.theta = take((size_t)Kc * F);
  l.qz = take((size_t)Kc * D * D);
  l.rz = take((size_t)Kc * D);
  l.t1 = take((size_t)Kc * D * D);
  l.total = o;
  return l;
}

}  // namespace more
}  // namespace gvi

using namespace gvi;

extern "C" size_t gvi_more_workspace(int chunk, int N, int D) {
  if (chunk <= 0 || N <= 0 || D <= 0) return 0;
  return more::layout(chunk, N, D).total * sizeof(float);
}

extern "C" int gvi_more_fit_f32(const float* X, int N, int D, const float* means, const float* linv, const float* W,
                                const float* y, const float* l2reg, int K, int chunk, float* quad, float* lin,
                                int32_t* ok, void* ws, size_t ws_bytes, void* stream) {
  GVI_REQUIRE(N > 0 && D > 0 && K >= 0 && chunk > 0, "gvi_more_fit_f32: bad sizes");
  if (K == 0) return GVI_OK;
  GVI_REQUIRE(X && means && linv && W && y && l2reg && quad && lin && ws, "gvi_more_fit_f32: null pointer");
  if (ws_bytes < gvi_more_workspace(chunk, N, D)) {
    set_last_error("gvi_more_fit_f32: workspace %zu < %zu", ws_bytes, gvi_more_workspace(chunk, N, D));
    return GVI_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int F = D * (D + 1) / 2 + D + 1, Fa = F + 1;
  const more::Layout l = more::layout(chunk, N, D);
  float* base = (float*)ws;
  float *Xc = base + l.xc, *Z = base + l.z, *Phi = base + l.phi, *A = base + l.a, *Sinv = base + l.sinv,
        *T21 = base + l.t21, *theta = base + l.theta, *Qz = base + l.qz, *rz = base + l.rz, *T1 = base + l.t1;
  const long long DD = (long long)D * D, ND = (long long)N * D;
  static bool attr_done = false;
  const size_t potrf_smem = (size_t)2 * more::NB * (more::NB + 1) * sizeof(float);
  if (!attr_done) {
    cudaFuncSetAttribute(more::potrf_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)potrf_smem);
    cudaFuncSetAttribute(more::backsolve_unpack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done = true;
  }
  if ((size_t)F * sizeof(float) > 200 * 1024) {
    set_last_error("gvi_more_fit_f32: D=%d gives F=%d features, too many for the back-substitution kernel", D, F);
    return GVI_ERR_UNSUPPORTED;
  }
  int rc;
  for (int k0 = 0; k0 < K; k0 += chunk) {
    const int Kc = min(chunk, K - k0);
    dim3 g1((unsigned)min((long long)2048, (ND + 255) / 256), Kc);
    more::center_kernel<<<g1, 256, 0, st>>>(X, means, N, D, k0, Xc);
    if ((rc = check_launch("more::center_kernel"))) return rc;
    // Z = Xc Linv^T
    rc = launch_bgemm_ex(0, 1, Kc, N, D, D, 1.f, Xc, D, ND, linv + (long long)k0 * DD, D, DD, Z, D, ND, nullptr, 0, 0.f,
                         0, st);
    if (rc) return rc;
    dim3 g2(ceil_div(N, 8), Kc);
    more::features_kernel<<<g2, 256, 8 * D * sizeof(float), st>>>(Z, y, N, D, F, Phi);
    if ((rc = check_launch("more::features_kernel"))) return rc;
    // A' = Phi^T diag(w) Phi (lower triangle)
    rc = launch_bgemm_ex(1, 0, Kc, Fa, Fa, N, 1.f, Phi, Fa, (long long)N * Fa, Phi, Fa, (long long)N * Fa, A, Fa,
                         (long long)Fa * Fa, W + (long long)k0 * N, N, 0.f, 1, st);
    if (rc) return rc;
    dim3 g3(ceil_div(F, 256), Kc);
    more::ridge_kernel<<<g3, 256, 0, st>>>(A, Fa, F, l2reg, k0);
    if ((rc = check_launch("more::ridge_kernel"))) return rc;
    // blocked Cholesky of A[:F,:F], carried through row F
    for (int p = 0; p < F; p += more::NB) {
      const int nb = min(more::NB, F - p);
      more::potrf_inv_kernel<<<Kc, 256, potrf_smem, st>>>(A, Fa, p, nb, Sinv, ok, k0);
      if ((rc = check_launch("more::potrf_inv_kernel"))) return rc;
      const int r0 = p + nb, rows = Fa - r0;
      if (rows <= 0) continue;
      // T21 = A21 Linv11^T
      rc = launch_bgemm_ex(0, 1, Kc, rows, nb, nb, 1.f, A + (long long)r0 * Fa + p, Fa, (long long)Fa * Fa, Sinv,
                           more::NB, (long long)more::NB * more::NB, T21, more::NB, (long long)Fa * more::NB, nullptr, 0,
                           0.f, 0, st);
      if (rc) return rc;
      dim3 g4((unsigned)min((long long)1024, ((long long)rows * nb + 255) / 256), Kc);
      more::copy_panel_kernel<<<g4, 256, 0, st>>>(T21, more::NB, (long long)Fa * more::NB, A, Fa, r0, p, rows, nb);
      if ((rc = check_launch("more::copy_panel_kernel"))) return rc;
      // A22 -= T21 T21^T (lower triangle)
      rc = launch_bgemm_ex(0, 1, Kc, rows, rows, nb, -1.f, T21, more::NB, (long long)Fa * more::NB, T21, more::NB,
                           (long long)Fa * more::NB, A + (long long)r0 * Fa + r0, Fa, (long long)Fa * Fa, nullptr, 0,
                           1.f, 1, st);
      if (rc) return rc;
    }
    more::backsolve_unpack_kernel<<<Kc, 1024, F * sizeof(float), st>>>(A, Fa, F, D, theta, Qz, rz);
    if ((rc = check_launch("more::backsolve_unpack_kernel"))) return rc;
    // quad = Linv^T Qz Linv
    rc = launch_bgemm_ex(0, 0, Kc, D, D, D, 1.f, Qz, D, DD, linv + (long long)k0 * DD, D, DD, T1, D, DD, nullptr, 0, 0.f,
                         0, st);
    if (rc) return rc;
    rc = launch_bgemm_ex(1, 0, Kc, D, D, D, 1.f, linv + (long long)k0 * DD, D, DD, T1, D, DD, quad + (long long)k0 * DD, D,
                         DD, nullptr, 0, 0.f, 0, st);
    if (rc) return rc;
    more::unwhiten_lin_kernel<<<Kc, 256, 0, st>>>(linv, means, quad, rz, D, k0, lin);
    if ((rc = check_launch("more::unwhiten_lin_kernel"))) return rc;
  }
  return GVI_OK;
}
