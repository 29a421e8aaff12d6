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
//
// Tensor-core route (default, gvi_more_tensor_cores()): the two O(F^2 N) / O(F^3) products run on tcgen05
// (tc_bgemm.cu) in split precision, 2 x fp16 by default (3xTF32 with GMMVI_B200_MORE_TC=tf32).  The feature kernel
// writes S = sqrt(w) Phi directly as the operand the MMA wants - transposed [F+1][N] (reduction dimension
// contiguous) and split into hi / lo parts - so A' = S^T S needs one operand; the reduction over the samples is cut
// into 512-sample segments added with round-to-nearest adds; only tiles of the lower triangle are computed.  The
// trailing update A22 -= T21 T21^T of the blocked Cholesky takes the hi / lo split of the panel from the kernel that
// copies the panel back.  fp16 parts are taken under one power-of-two scale per component: for S from the bound
// max_n sqrt(w_n) max(|z_n|_inf^2, |z_n|_inf, 1, |y_n|), for the panels from sqrt(max_i A'_ii) >= |T21_ic| (Cholesky rows
// have norm sqrt(A'_ii); the right-hand-side row is covered because the augmented matrix is a Gram matrix too).
#include "tc_common.cuh"
#include "../../include/gmmvi_b200.h"
#include <stdlib.h>
#include <cuda_fp16.h>

namespace gvi {

int launch_bgemm_ex(int transA, int transB, int batch, int M, int N, int Kd, float alpha, const float* A, int lda,
                    long long strideA, const float* B, int ldb, long long strideB, float* C, int ldc,
                    long long strideC, const float* scaleK, long long strideScale, float beta, int lower_only,
                    cudaStream_t st);
int launch_tc_bgemm_ex(int batch, int M, int N, int Kd, float alpha, const float* Ah, const float* Al, const float* Bh,
                       const float* Bl, float* C, int ldc, long long strideC, float beta, int kseg_kblocks,
                       int lower_only, cudaStream_t st);
int launch_tc_bgemm_h16_ex(int batch, int M, int N, int Kd, float alpha, const float* alpha_b, const void* Ah,
                           const void* Al, const void* Bh, const void* Bl, float* C, int ldc, long long strideC,
                           float beta, int kseg_kblocks, int lower_only, cudaStream_t st);
int launch_tc_bgemm_h16_strided(int batch, int M, int N, int Kd, float alpha, const float* alpha_b, const void* Ah,
                                const void* Al, long long opStrideA, const void* Bh, const void* Bl,
                                long long opStrideB, float* C, int ldc, long long strideC, float beta,
                                int kseg_kblocks, int lower_only, cudaStream_t st);
bool tc_gemm_enabled();

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

// Tensor-core operand: S[kc][f][n] = sqrt(w_kn) Phi[n][f] as TF32 hi / lo, [Fa][Np] with the samples contiguous
// (Np = N rounded up to 4, the padding is zero).  One CTA = 32 samples, a warp walks the features (lane = sample),
// so every store instruction writes 128 contiguous bytes.
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// smax[kc] = max_n sqrt(w_kn) max(|z_n|_inf^2, |z_n|_inf, 1, |y_n|) >= every entry of S (as uint bits; zeroed by the caller)
__global__ void __launch_bounds__(256)
feature_bound_kernel(const float* __restrict__ Z, const float* __restrict__ y, const float* __restrict__ W, int N, int D,
                     int k0, unsigned* __restrict__ smax) {
  const int kc = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  float b = 0.f;
  if (n < N) {
    const float* z = Z + ((long long)kc * N + n) * D;
    float zm = 0.f;
    for (int d = lane; d < D; d += 32) zm = fmaxf(zm, fabsf(z[d]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) zm = fmaxf(zm, __shfl_xor_sync(0xffffffffu, zm, o));
    b = sqrtf(fmaxf(W[(long long)(k0 + kc) * N + n], 0.f)) * fmaxf(fmaxf(zm * zm, zm), fmaxf(1.f, fabsf(y[n])));
  }
  __shared__ float red[8];
  if (lane == 0) red[warp] = b;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    if (isfinite(m)) atomicMax(smax + kc, __float_as_uint(m));
  }
}

template <bool H16>
__global__ void __launch_bounds__(256)
features_split_kernel(const float* __restrict__ Z, const float* __restrict__ y, const float* __restrict__ W, int N,
                      int Np, int D, int F, int k0, void* __restrict__ Sh_, void* __restrict__ Sl_,
                      const unsigned* __restrict__ smax, float* __restrict__ alphaS) {
  extern __shared__ float zs[];      // [32][D + 1]
  __shared__ float sw[32], sy[32];
  const int kc = blockIdx.y;
  const int n0 = blockIdx.x * 32;
  const int Dp = D + 1;
  const float scale = H16 ? tcx::h16_scale_of(__uint_as_float(smax[kc])) : 1.f;
  if (H16 && blockIdx.x == 0 && threadIdx.x == 0) alphaS[kc] = 1.f / (scale * scale);
  const float* Zk = Z + (long long)kc * N * D;
  for (int e = threadIdx.x; e < 32 * D; e += blockDim.x) {
    const int r = e / D, d = e % D;
    zs[r * Dp + d] = (n0 + r < N) ? Zk[(long long)(n0 + r) * D + d] : 0.f;
  }
  if (threadIdx.x < 32) {
    const int n = n0 + threadIdx.x;
    sw[threadIdx.x] = n < N ? scale * sqrtf(fmaxf(W[(long long)(k0 + kc) * N + n], 0.f)) : 0.f;
    sy[threadIdx.x] = n < N ? y[n] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int nq = D * (D + 1) / 2, Fa = F + 1;
  const float* zr = zs + lane * Dp;
  const float w = sw[lane];
  const bool live = n0 + lane < Np;
  const long long o0 = (long long)kc * Fa * Np + n0 + lane;
  // (i, j) of quadratic feature f = warp, advanced by nwarps per step
  int i = 0, j = warp;
  while (i < D && j >= D) { j = j - D + i + 1; ++i; }
  for (int f = warp; f < Fa; f += nwarps) {
    float v;
    if (f < nq) {
      v = zr[i] * zr[j];
      j += nwarps;
      while (i < D && j >= D) { j = j - D + i + 1; ++i; }
    } else if (f < nq + D) v = zr[f - nq];
    else if (f == F - 1) v = 1.f;
    else v = sy[lane];
    v *= w;
    if (live) {
      if (H16) {
        const __half h = __float2half_rn(v);
        reinterpret_cast<__half*>(Sh_)[o0 + (long long)f * Np] = h;
        reinterpret_cast<__half*>(Sl_)[o0 + (long long)f * Np] = __float2half_rn(v - __half2float(h));
      } else {
        const float h = tf32_rna(v);
        reinterpret_cast<float*>(Sh_)[o0 + (long long)f * Np] = h;
        reinterpret_cast<float*>(Sl_)[o0 + (long long)f * Np] = tf32_rna(v - h);
      }
    }
  }
}

// (Fa = row pitch of A, sA = distance between the matrices of two components, here and below)
__global__ void ridge_kernel(float* __restrict__ A, int Fa, long long sA, int F, const float* __restrict__ l2, int k0,
                             unsigned* __restrict__ dmax) {
  const int kc = blockIdx.y;
  float* Ak = A + kc * sA;
  const float lam = l2[k0 + kc];
  float m = 0.f;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f <= F; f += gridDim.x * blockDim.x) {
    float d = Ak[(long long)f * Fa + f];
    if (f < F - 1) Ak[(long long)f * Fa + f] = d = d + lam;
    m = fmaxf(m, d);
  }
  if (dmax) {       // largest diagonal entry of the augmented matrix (uint bits of a non-negative float)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && isfinite(m)) atomicMax(dmax + kc, __float_as_uint(m));
  }
}
// scale of the fp16 split of the Cholesky panels: every entry of T21 is at most sqrt(max_i A'_ii)
__global__ void trail_scale_kernel(const unsigned* __restrict__ dmax, int Kc, float* __restrict__ sT,
                                   float* __restrict__ alphaT) {
  const int kc = blockIdx.x * blockDim.x + threadIdx.x;
  if (kc < Kc) {
    const float s = tcx::h16_scale_of(sqrtf(__uint_as_float(dmax[kc])));
    sT[kc] = s;
    alphaT[kc] = 1.f / (s * s);
  }
}

// Diagonal block [nb x nb] at (p, p): in-place Cholesky and its inverse into Sinv[kc][NB][NB].
__global__ void __launch_bounds__(256)
potrf_inv_kernel(float* __restrict__ A, int Fa, long long sA, int p, int nb, float* __restrict__ Sinv,
                 int32_t* __restrict__ ok, int k0) {
  extern __shared__ float sm[];          // L[nb][NB+1], Y[nb][NB+1]
  float* L = sm;
  float* Y = sm + NB * (NB + 1);
  const int kc = blockIdx.x;
  float* Ak = A + kc * sA + (long long)p * Fa + p;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int e = tid; e < nb * nb; e += nt) {
    const int i = e / nb, j = e % nb;
    L[i * (NB + 1) + j] = (j <= i) ? Ak[(long long)i * Fa + j] : 0.f;
  }
  __syncthreads();
  __shared__ int bad;
  if (tid == 0) bad = 0;
  // Two threads per row (even / odd columns of the dot product), four fp32 accumulators each, combined with a shuffle
  // inside the pair: dot products of at most 128 terms, as accurate as the fp32-grade trailing updates around them (an
  // fp64 chain with its fp32 -> fp64 conversions made this kernel 17 % of the C3 iteration).
  const int half = tid & 1, pr = tid >> 1;
  for (int j = 0; j < nb; ++j) {
    __syncthreads();
    {
      const int i = j + pr;                       // blockDim = 256 covers the nb <= 128 rows below the diagonal
      const float* Li = L + min(i, nb - 1) * (NB + 1);
      const float* Lj = L + j * (NB + 1);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      int m = half;
      for (; m + 6 < j; m += 8) {
        s0 = fmaf(Li[m], Lj[m], s0);
        s1 = fmaf(Li[m + 2], Lj[m + 2], s1);
        s2 = fmaf(Li[m + 4], Lj[m + 4], s2);
        s3 = fmaf(Li[m + 6], Lj[m + 6], s3);
      }
      for (; m < j; m += 2) s0 = fmaf(Li[m], Lj[m], s0);
      float s = (s0 + s1) + (s2 + s3);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (half == 0 && i < nb) L[i * (NB + 1) + j] = Li[j] - s;     // unscaled
    }
    __syncthreads();
    float piv = L[j * (NB + 1) + j];
    if (!(piv > 0.f) || !isfinite(piv)) {
      if (tid == 0) bad = 1;
      piv = 1.f;
    }
    const float c = sqrtf(piv), rc = 1.f / c;
    __syncthreads();
    for (int i = j + tid; i < nb; i += nt) L[i * (NB + 1) + j] = (i == j) ? c : L[i * (NB + 1) + j] * rc;
  }
  __syncthreads();
  // inverse: a thread pair per column
  for (int c = pr; c < nb; c += nt / 2) {
    const unsigned pm = 3u << ((tid & 31) & ~1);
    if (half == 0) {
      for (int i = 0; i < c; ++i) Y[i * (NB + 1) + c] = 0.f;
      Y[c * (NB + 1) + c] = 1.f / L[c * (NB + 1) + c];
    }
    __syncwarp(pm);
    for (int i = c + 1; i < nb; ++i) {
      const float* Li = L + i * (NB + 1);
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
      int m = c + half;
      for (; m + 6 < i; m += 8) {
        s0 = fmaf(Li[m], Y[m * (NB + 1) + c], s0);
        s1 = fmaf(Li[m + 2], Y[(m + 2) * (NB + 1) + c], s1);
        s2 = fmaf(Li[m + 4], Y[(m + 4) * (NB + 1) + c], s2);
        s3 = fmaf(Li[m + 6], Y[(m + 6) * (NB + 1) + c], s3);
      }
      for (; m < i; m += 2) s0 = fmaf(Li[m], Y[m * (NB + 1) + c], s0);
      float s = (s0 + s1) + (s2 + s3);
      s += __shfl_xor_sync(pm, s, 1);
      if (half == 0) Y[i * (NB + 1) + c] = -s / Li[i];
      __syncwarp(pm);
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

// The same factorisation and inverse, blocked by 32: the default (0.26 -> 0.13 ms per launch at C3; the column-by-column
// kernel above spends 3 block barriers per column and runs at ~10 cycles per instruction; it stays selectable with
// GMMVI_B200_MORE_POTRF=columns and as cross-check in the tests).  Per 32-column block: warp 0 factors the 32 x 32 diagonal block in registers (lane = row, the column
// being eliminated is passed round by shuffles) and inverts it (lane = column of the inverse); all warps solve the
// panel below against that inverse (lane = column, its row of the inverse in registers) and apply the rank-32 update to
// the trailing lower triangle (warp = row, lane = column).  The off-diagonal blocks of the inverse follow from
// Y_ij = -Y_ii sum_k L_ik Y_kj, block diagonal by block diagonal, the intermediate product parked in the unused upper
// triangle of L.  Blocks beyond nb are padded with the identity.
__global__ void __launch_bounds__(256)
potrf_inv_blocked_kernel(float* __restrict__ A, int Fa, long long sA, int p, int nb, float* __restrict__ Sinv,
                         int32_t* __restrict__ ok, int k0) {
  extern __shared__ float sm[];          // L[NB][NB+1], Y[NB][NB+1]
  constexpr int P = NB + 1;
  float* L = sm;
  float* Y = sm + NB * P;
  const int kc = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* Ak = A + kc * sA + (long long)p * Fa + p;
  __shared__ int bad;
  if (tid == 0) bad = 0;
  for (int e = tid; e < NB * NB; e += 256) {
    const int i = e / NB, j = e % NB;
    float v = 0.f;
    if (i < nb && j <= i) v = Ak[(long long)i * Fa + j];
    else if (i >= nb && i == j) v = 1.f;
    L[i * P + j] = v;
    Y[i * P + j] = 0.f;
  }
  __syncthreads();
  for (int jb = 0; jb < NB / 32; ++jb) {
    const int j0 = 32 * jb;
    if (warp == 0) {
      // a[c] = entry (lane, k + c) of the block while column k is eliminated: the array is shifted down by one after
      // every column, so all register indices are static and the loop over k stays rolled
      float a[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) a[c] = L[(j0 + lane) * P + j0 + c];
      int mybad = 0;
      for (int k = 0; k < 32; ++k) {
        float akk = __shfl_sync(0xffffffffu, a[0], k);
        if (!(akk > 0.f) || !isfinite(akk)) {
          mybad = 1;
          akk = 1.f;
        }
        const float d = sqrtf(akk);
        const float lk = lane == k ? d : a[0] * (1.f / d);          // L[lane][k] (zero above the diagonal)
        L[(j0 + lane) * P + j0 + k] = lane >= k ? lk : 0.f;
#pragma unroll
        for (int c = 1; c < 32; ++c) {
          const float lck = __shfl_sync(0xffffffffu, lk, (k + c) & 31);        // L[k + c][k]
          if (k + c < 32 && lane >= k + c) a[c] = fmaf(-lk, lck, a[c]);
        }
#pragma unroll
        for (int c = 0; c < 31; ++c) a[c] = a[c + 1];
        a[31] = 0.f;
      }
      if (mybad) bad = 1;
      __syncwarp();
      // inverse of the diagonal block, lane = column: forward substitution on the unit vector, right-hand side shifted
      // the same way (b[r] = entry i + r while row i is solved)
      float b[32];
#pragma unroll
      for (int r = 0; r < 32; ++r) b[r] = r == lane ? 1.f : 0.f;
      for (int i = 0; i < 32; ++i) {
        const float yi = b[0] / L[(j0 + i) * P + j0 + i];
        Y[(j0 + i) * P + j0 + lane] = yi;
#pragma unroll
        for (int r = 1; r < 32; ++r)
          if (i + r < 32) b[r] = fmaf(-L[(j0 + i + r) * P + j0 + i], yi, b[r]);
#pragma unroll
        for (int r = 0; r < 31; ++r) b[r] = b[r + 1];
        b[31] = 0.f;
      }
    }
    __syncthreads();
    if (j0 + 32 < NB) {
      // panel: X[i][c] = sum_m A[i][j0 + m] Yjj[c][m], in place
      float yr[32];
#pragma unroll
      for (int m = 0; m < 32; ++m) yr[m] = Y[(j0 + lane) * P + j0 + m];
      for (int i = j0 + 32 + warp; i < NB; i += 8) {
        float x = 0.f;
#pragma unroll
        for (int m = 0; m < 32; ++m) x = fmaf(L[i * P + j0 + m], yr[m], x);
        __syncwarp();
        L[i * P + j0 + lane] = x;
      }
      __syncthreads();
      // trailing lower triangle: A[i][c] -= sum_m X[i][m] X[c][m]
      for (int i = j0 + 32 + warp; i < NB; i += 8) {
        float xi[32];
#pragma unroll
        for (int m = 0; m < 32; ++m) xi[m] = L[i * P + j0 + m];
        for (int c = j0 + 32 + lane; c <= i; c += 32) {
          float s = 0.f;
#pragma unroll
          for (int m = 0; m < 32; ++m) s = fmaf(xi[m], L[c * P + j0 + m], s);
          L[i * P + c] -= s;
        }
      }
      __syncthreads();
    }
  }
  for (int d = 1; d < NB / 32; ++d) {
    const int ntask = 32 * (NB / 32 - d);
    for (int t = warp; t < ntask; t += 8) {          // T = sum_k L_ik Y_kj, parked at block (jb, ib) of L
      const int jb = t >> 5, r = t & 31, ib = jb + d;
      float s = 0.f;
      for (int k = 32 * jb; k < 32 * ib; ++k) s = fmaf(L[(32 * ib + r) * P + k], Y[k * P + 32 * jb + lane], s);
      L[(32 * jb + r) * P + 32 * ib + lane] = s;
    }
    __syncthreads();
    for (int t = warp; t < ntask; t += 8) {          // Y_ij = -Y_ii T
      const int jb = t >> 5, r = t & 31, ib = jb + d;
      float s = 0.f;
      for (int m = 0; m <= r; ++m)
        s = fmaf(Y[(32 * ib + r) * P + 32 * ib + m], L[(32 * jb + m) * P + 32 * ib + lane], s);
      Y[(32 * ib + r) * P + 32 * jb + lane] = -s;
    }
    __syncthreads();
  }
  float* So = Sinv + (long long)kc * NB * NB;
  for (int e = tid; e < nb * nb; e += 256) {
    const int i = e / nb, j = e % nb;
    Ak[(long long)i * Fa + j] = j <= i ? L[i * P + j] : 0.f;
    So[i * NB + j] = Y[i * P + j];
  }
  if (tid == 0 && bad && ok) ok[k0 + kc] = 0;
}

// A[kc][r0 + r][p + c] = T[kc][r][c]; with Th / Tl also the dense hi / lo split [kc][rows][nb] of the panel
// (TF32 floats, or with sT fp16 halves of sT[kc] T); with Ch / Cl (fp16 route, full-width panels) the rows from cat_row0
// on also go to columns cat_col0 .. cat_col0 + nb of the two-panel operand [kc][cat_rows][2 NB] of the paired update
__global__ void copy_panel_kernel(const float* __restrict__ T, int ldt, long long strideT, float* __restrict__ A, int Fa,
                                  long long strideA, int r0, int p, int rows, int nb, void* __restrict__ Th,
                                  void* __restrict__ Tl, const float* __restrict__ sT, __half* __restrict__ Ch,
                                  __half* __restrict__ Cl, int cat_row0, int cat_col0, int cat_rows) {
  const int kc = blockIdx.y;
  const float* Tk = T + kc * strideT;
  float* Ak = A + kc * strideA;
  const long long ob = (long long)kc * rows * nb;
  const float s = sT ? sT[kc] : 1.f;
  if (nb % 4 == 0 && ldt % 4 == 0 && Fa % 4 == 0 && p % 4 == 0 && strideT % 4 == 0 && strideA % 4 == 0) {
    // 16 bytes per thread and step, 32-bit index arithmetic (all panels but a ragged last one)
    const int q4 = nb / 4, total = rows * q4;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
      const int r = e / q4, c = (e - r * q4) * 4;
      const float4 v = *reinterpret_cast<const float4*>(Tk + (long long)r * ldt + c);
      *reinterpret_cast<float4*>(Ak + (long long)(r0 + r) * Fa + p + c) = v;
      const long long o = ob + (long long)r * nb + c;
      if (sT && (Th || Ch)) {
        const float x[4] = {v.x * s, v.y * s, v.z * s, v.w * s};
        __half2 h01 = __floats2half2_rn(x[0], x[1]), h23 = __floats2half2_rn(x[2], x[3]);
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        __half2 l01 = __floats2half2_rn(x[0] - f01.x, x[1] - f01.y), l23 = __floats2half2_rn(x[2] - f23.x, x[3] - f23.y);
        uint2 hv, lv;
        hv.x = *reinterpret_cast<uint32_t*>(&h01); hv.y = *reinterpret_cast<uint32_t*>(&h23);
        lv.x = *reinterpret_cast<uint32_t*>(&l01); lv.y = *reinterpret_cast<uint32_t*>(&l23);
        if (Th) {
          *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(Th) + o) = hv;
          *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(Tl) + o) = lv;
        }
        if (Ch && r >= cat_row0) {
          const long long oc = ((long long)kc * cat_rows + (r - cat_row0)) * (2 * NB) + cat_col0 + c;
          *reinterpret_cast<uint2*>(Ch + oc) = hv;
          *reinterpret_cast<uint2*>(Cl + oc) = lv;
        }
      } else if (Th) {
        float4 h, l;
        h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
        l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(Th) + o) = h;
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(Tl) + o) = l;
      }
    }
    return;
  }
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)rows * nb;
       e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / nb), c = (int)(e % nb);
    const float v = Tk[(long long)r * ldt + c];
    Ak[(long long)(r0 + r) * Fa + p + c] = v;
    if (Th && sT) {
      const __half h = __float2half_rn(v * s);
      reinterpret_cast<__half*>(Th)[ob + e] = h;
      reinterpret_cast<__half*>(Tl)[ob + e] = __float2half_rn(v * s - __half2float(h));
    } else if (Th) {
      const float h = tf32_rna(v);
      reinterpret_cast<float*>(Th)[ob + e] = h;
      reinterpret_cast<float*>(Tl)[ob + e] = tf32_rna(v - h);
    }
  }
}

// Back substitution L^T theta = v (v = row F of the factored augmented matrix), then the coefficients are unpacked:
// Qz = -(T + T^T), rz.  One CTA per component.  Blocked by BS rows from the bottom: warp 0 solves the BS x BS diagonal
// block held in shared memory (lane = two unknowns, theta broadcast by shuffle), then all threads subtract the block's
// contribution from the remaining right-hand side, thread = column, rows of L read coalesced.
constexpr int BS = 64;
__global__ void __launch_bounds__(1024)
backsolve_unpack_kernel(const float* __restrict__ A, int Fa, long long sA, int F, int D, float* __restrict__ theta_ws,
                        float* __restrict__ Qz, float* __restrict__ rz) {
  extern __shared__ float v[];       // [F]
  __shared__ float Ld[BS][BS + 1];
  __shared__ float th[BS];
  const int kc = blockIdx.x;
  const float* Ak = A + kc * sA;
  float* theta = theta_ws + (long long)kc * F;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
  for (int f = tid; f < F; f += nt) v[f] = Ak[(long long)F * Fa + f];
  for (int p = ((F - 1) / BS) * BS; p >= 0; p -= BS) {
    const int nb = min(BS, F - p);
    for (int e = tid; e < nb * nb; e += nt) {
      const int i = e / nb, j = e % nb;
      Ld[i][j] = Ak[(long long)(p + i) * Fa + p + j];
    }
    __syncthreads();        // also orders the v updates of the previous block
    if (tid < 32) {
      float v0 = lane < nb ? v[p + lane] : 0.f, v1 = lane + 32 < nb ? v[p + lane + 32] : 0.f;
      for (int i = nb - 1; i >= 0; --i) {
        const float vi = __shfl_sync(0xffffffffu, i < 32 ? v0 : v1, i & 31);
        const float t = vi / Ld[i][i];
        if (lane == 0) th[i] = t;
        if (lane < i) v0 = fmaf(-Ld[i][lane], t, v0);
        if (lane + 32 < i) v1 = fmaf(-Ld[i][lane + 32], t, v1);
      }
    }
    __syncthreads();
    for (int i = tid; i < nb; i += nt) theta[p + i] = th[i];
    // rows come from HBM and one CTA per component has to cover the latency with its own requests: 16 bytes per
    // thread and load, four loads in flight (p is a multiple of BS, the pitch Fa a multiple of 4)
    for (int m = 4 * tid; m < p; m += 4 * nt) {
      const float* col = Ak + (long long)p * Fa + m;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int i = 0;
      for (; i + 3 < nb; i += 4) {
        const float4 x0 = __ldg(reinterpret_cast<const float4*>(col + (long long)i * Fa));
        const float4 x1 = __ldg(reinterpret_cast<const float4*>(col + (long long)(i + 1) * Fa));
        const float4 x2 = __ldg(reinterpret_cast<const float4*>(col + (long long)(i + 2) * Fa));
        const float4 x3 = __ldg(reinterpret_cast<const float4*>(col + (long long)(i + 3) * Fa));
        const float t0 = th[i], t1 = th[i + 1], t2 = th[i + 2], t3 = th[i + 3];
        acc.x = fmaf(x0.x, t0, acc.x); acc.y = fmaf(x0.y, t0, acc.y); acc.z = fmaf(x0.z, t0, acc.z); acc.w = fmaf(x0.w, t0, acc.w);
        acc.x = fmaf(x1.x, t1, acc.x); acc.y = fmaf(x1.y, t1, acc.y); acc.z = fmaf(x1.z, t1, acc.z); acc.w = fmaf(x1.w, t1, acc.w);
        acc.x = fmaf(x2.x, t2, acc.x); acc.y = fmaf(x2.y, t2, acc.y); acc.z = fmaf(x2.z, t2, acc.z); acc.w = fmaf(x2.w, t2, acc.w);
        acc.x = fmaf(x3.x, t3, acc.x); acc.y = fmaf(x3.y, t3, acc.y); acc.z = fmaf(x3.z, t3, acc.z); acc.w = fmaf(x3.w, t3, acc.w);
      }
      for (; i < nb; ++i) {
        const float4 x0 = __ldg(reinterpret_cast<const float4*>(col + (long long)i * Fa));
        const float t0 = th[i];
        acc.x = fmaf(x0.x, t0, acc.x); acc.y = fmaf(x0.y, t0, acc.y); acc.z = fmaf(x0.z, t0, acc.z); acc.w = fmaf(x0.w, t0, acc.w);
      }
      v[m] -= acc.x; v[m + 1] -= acc.y; v[m + 2] -= acc.z; v[m + 3] -= acc.w;
    }
  }
  __syncthreads();
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
  size_t xc, z, phi, sh, sl, a, sinv, t21, t21h, t21l, tcath, tcatl, theta, qz, rz, t1, small, total;
};
static inline int pitch_of(int Fa) { return (Fa + 3) & ~3; }      // 16-byte aligned rows of the normal matrix
// route: 0 = SIMT fp32, 1 = tensor cores 3xTF32, 2 = tensor cores 2 x fp16
static Layout layout(int Kc, int N, int D, int route) {
  const size_t F = (size_t)D * (D + 1) / 2 + D + 1, Fa = F + 1, ld = pitch_of((int)Fa), Np = ((size_t)N + 7) & ~(size_t)7;
  const size_t opnd = route == 2 ? 2 : 1;        // operand elements per float of workspace
  Layout l;
  size_t o = 0;
  auto take = [&](size_t n) { size_t r = o; o += (n + 63) / 64 * 64; return r; };
  l.xc = take((size_t)Kc * N * D);
  l.z = take((size_t)Kc * N * D);
  l.phi = take(route ? 0 : (size_t)Kc * N * Fa);
  l.sh = take(route ? (size_t)Kc * Fa * Np / opnd : 0);
  l.sl = take(route ? (size_t)Kc * Fa * Np / opnd : 0);
  l.a = take((size_t)Kc * Fa * ld);
  l.sinv = take((size_t)Kc * NB * NB);
  l.t21 = take((size_t)Kc * Fa * NB);
  l.t21h = take(route ? (size_t)Kc * Fa * NB / opnd : 0);
  l.t21l = take(route ? (size_t)Kc * Fa * NB / opnd : 0);
  l.tcath = take(route == 2 ? (size_t)Kc * Fa * NB : 0);      // [Kc][rows][2 NB] halves: two panels side by side
  l.tcatl = take(route == 2 ? (size_t)Kc * Fa * NB : 0);
  l.theta = take((size_t)Kc * F);
  l.qz = take((size_t)Kc * D * D);
  l.rz = take((size_t)Kc * D);
  l.t1 = take((size_t)Kc * D * D);
  l.small = take((size_t)5 * Kc);               // smax, dmax (uint), alphaS, sT, alphaT
  l.total = o;
  return l;
}

static bool more_pairs() {      // trailing updates applied two panels at a time (fp16 route)
  const char* e = getenv("GMMVI_B200_MORE_PAIRS");
  return !(e != nullptr && e[0] == '0');
}
static bool more_potrf_blocked() {      // default; GMMVI_B200_MORE_POTRF=columns selects the column-by-column kernel
  const char* e = getenv("GMMVI_B200_MORE_POTRF");
  return !(e != nullptr && e[0] == 'c');
}
static int more_route() {      // read per call: the tests switch routes inside one process
  const char* e = getenv("GMMVI_B200_MORE_TC");
  if ((e != nullptr && e[0] == '0') || !tc_gemm_enabled()) return 0;
  if (e != nullptr && e[0] == 't') return 1;
  return 2;
}

constexpr int SEG_SAMPLES = 512;      // samples per accumulator segment of the normal-equation build

}  // namespace more
}  // namespace gvi

using namespace gvi;

extern "C" int gvi_more_tensor_cores(void) { return more::more_route(); }

extern "C" size_t gvi_more_workspace(int chunk, int N, int D) {
  if (chunk <= 0 || N <= 0 || D <= 0) return 0;
  return more::layout(chunk, N, D, more::more_route()).total * sizeof(float);
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
  const int route = more::more_route();
  if (route && reinterpret_cast<uintptr_t>(ws) % 16 != 0) {
    set_last_error("gvi_more_fit_f32: workspace must be 16-byte aligned for the tensor-core route");
    return GVI_ERR_INVALID;
  }
  const bool tc = route != 0, h16 = route == 2;
  constexpr int NBc = more::NB;
  const int F = D * (D + 1) / 2 + D + 1, Fa = F + 1, ld = more::pitch_of(Fa), Np = (N + 7) & ~7;
  const long long sA = (long long)Fa * ld;
  const more::Layout l = more::layout(chunk, N, D, route);
  float* base = (float*)ws;
  float *Xc = base + l.xc, *Z = base + l.z, *Phi = base + l.phi, *Sh = base + l.sh, *Sl = base + l.sl, *A = base + l.a,
        *Sinv = base + l.sinv, *T21 = base + l.t21, *T21h = base + l.t21h, *T21l = base + l.t21l,
        *theta = base + l.theta, *Qz = base + l.qz, *rz = base + l.rz, *T1 = base + l.t1;
  __half *Tch = (__half*)(base + l.tcath), *Tcl = (__half*)(base + l.tcatl);
  const bool pairs = h16 && more::more_pairs();
  unsigned* smax = (unsigned*)(base + l.small);
  unsigned* dmax = smax + chunk;
  float *alphaS = (float*)(dmax + chunk), *sT = alphaS + chunk, *alphaT = sT + chunk;
  const long long DD = (long long)D * D, ND = (long long)N * D;
  static unsigned long long attr_done_mask = 0;
  const size_t potrf_smem = (size_t)2 * more::NB * (more::NB + 1) * sizeof(float);
  if (first_call_on_device(attr_done_mask)) {
    cudaFuncSetAttribute(more::potrf_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)potrf_smem);
    cudaFuncSetAttribute(more::potrf_inv_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)potrf_smem);
    cudaFuncSetAttribute(more::backsolve_unpack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(more::features_split_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(more::features_split_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  if ((size_t)F * sizeof(float) > 200 * 1024) {
    set_last_error("gvi_more_fit_f32: D=%d gives F=%d features, too many for the back-substitution kernel", D, F);
    return GVI_ERR_UNSUPPORTED;
  }
  int rc;
  auto potrf = more::more_potrf_blocked() ? more::potrf_inv_blocked_kernel : more::potrf_inv_kernel;
  for (int k0 = 0; k0 < K; k0 += chunk) {
    const int Kc = min(chunk, K - k0);
    dim3 g1((unsigned)min((long long)2048, (ND + 255) / 256), Kc);
    more::center_kernel<<<g1, 256, 0, st>>>(X, means, N, D, k0, Xc);
    if ((rc = check_launch("more::center_kernel"))) return rc;
    // Z = Xc Linv^T
    rc = launch_bgemm_ex(0, 1, Kc, N, D, D, 1.f, Xc, D, ND, linv + (long long)k0 * DD, D, DD, Z, D, ND, nullptr, 0, 0.f,
                         0, st);
    if (rc) return rc;
    if (tc) {
      // S = sqrt(w) Phi, transposed and split; A' = S^T S (lower triangle)
      const size_t fsm = (size_t)32 * (D + 1) * sizeof(float);
      dim3 g2(ceil_div(Np, 32), Kc);
      if (h16) {
        cudaMemsetAsync(smax, 0, (size_t)2 * chunk * sizeof(unsigned), st);        // smax and dmax
        dim3 gb(ceil_div(N, 8), Kc);
        more::feature_bound_kernel<<<gb, 256, 0, st>>>(Z, y, W, N, D, k0, smax);
        if ((rc = check_launch("more::feature_bound_kernel"))) return rc;
        more::features_split_kernel<true><<<g2, 256, fsm, st>>>(Z, y, W, N, Np, D, F, k0, Sh, Sl, smax, alphaS);
        if ((rc = check_launch("more::features_split_kernel"))) return rc;
        rc = launch_tc_bgemm_h16_ex(Kc, Fa, Fa, Np, 1.f, alphaS, Sh, Sl, Sh, Sl, A, ld, sA, 0.f, more::SEG_SAMPLES / 64,
                                    1, st);
      } else {
        more::features_split_kernel<false><<<g2, 256, fsm, st>>>(Z, y, W, N, Np, D, F, k0, Sh, Sl, nullptr, nullptr);
        if ((rc = check_launch("more::features_split_kernel"))) return rc;
        rc = launch_tc_bgemm_ex(Kc, Fa, Fa, Np, 1.f, Sh, Sl, Sh, Sl, A, ld, sA, 0.f, more::SEG_SAMPLES / 32, 1, st);
      }
      if (rc) return rc;
    } else {
      dim3 g2(ceil_div(N, 8), Kc);
      more::features_kernel<<<g2, 256, 8 * D * sizeof(float), st>>>(Z, y, N, D, F, Phi);
      if ((rc = check_launch("more::features_kernel"))) return rc;
      // A' = Phi^T diag(w) Phi (lower triangle)
      rc = launch_bgemm_ex(1, 0, Kc, Fa, Fa, N, 1.f, Phi, Fa, (long long)N * Fa, Phi, Fa, (long long)N * Fa, A, ld, sA,
                           W + (long long)k0 * N, N, 0.f, 1, st);
      if (rc) return rc;
    }
    dim3 g3(ceil_div(Fa, 256), Kc);
    more::ridge_kernel<<<g3, 256, 0, st>>>(A, ld, sA, F, l2reg, k0, h16 ? dmax : nullptr);
    if ((rc = check_launch("more::ridge_kernel"))) return rc;
    if (h16) {
      more::trail_scale_kernel<<<ceil_div(Kc, 128), 128, 0, st>>>(dmax, Kc, sT, alphaT);
      if ((rc = check_launch("more::trail_scale_kernel"))) return rc;
    }
    // blocked Cholesky of A[:F,:F], carried through row F
    for (int p = 0; p < F; p += more::NB) {
      const int nb = min(more::NB, F - p);
      if (pairs && F - p >= 2 * NBc && Fa - (p + 2 * NBc) >= 64) {
        // Two full panels at a time: the trailing matrix is read and written once per PAIR (its read-modify-write from
        // HBM bounds the update, not the MMAs).  Panel p is factored and applied only to the columns of panel p + NB;
        // panel p + NB is factored; both are applied to the rest as one product with a reduction length of 2 NB.
        const int p1 = p + NBc, rW = p + 2 * NBc, rows0 = Fa - p1, rowsW = Fa - rW;
        dim3 gp((unsigned)min((long long)1024, ((long long)rows0 * (NBc / 4) + 255) / 256), Kc);
        potrf<<<Kc, 256, potrf_smem, st>>>(A, ld, sA, p, NBc, Sinv, ok, k0);
        if ((rc = check_launch("more::potrf_inv_kernel"))) return rc;
        rc = launch_bgemm_ex(0, 1, Kc, rows0, NBc, NBc, 1.f, A + (long long)p1 * ld + p, ld, sA, Sinv, NBc,
                             (long long)NBc * NBc, T21, NBc, (long long)Fa * NBc, nullptr, 0, 0.f, 0, st);
        if (rc) return rc;
        more::copy_panel_kernel<<<gp, 256, 0, st>>>(T21, NBc, (long long)Fa * NBc, A, ld, sA, p1, p, rows0, NBc, T21h, T21l,
                                                    sT, Tch, Tcl, NBc, 0, rowsW);
        if ((rc = check_launch("more::copy_panel_kernel"))) return rc;
        // A[p1:, p1 : p1 + NB] -= T0 T0[0:NB]^T
        rc = launch_tc_bgemm_h16_strided(Kc, rows0, NBc, NBc, -1.f, alphaT, T21h, T21l, (long long)rows0 * NBc, T21h, T21l,
                                         (long long)rows0 * NBc, A + (long long)p1 * ld + p1, ld, sA, 1.f, 0, 0, st);
        if (rc) return rc;
        potrf<<<Kc, 256, potrf_smem, st>>>(A, ld, sA, p1, NBc, Sinv, ok, k0);
        if ((rc = check_launch("more::potrf_inv_kernel"))) return rc;
        rc = launch_bgemm_ex(0, 1, Kc, rowsW, NBc, NBc, 1.f, A + (long long)rW * ld + p1, ld, sA, Sinv, NBc,
                             (long long)NBc * NBc, T21, NBc, (long long)Fa * NBc, nullptr, 0, 0.f, 0, st);
        if (rc) return rc;
        more::copy_panel_kernel<<<gp, 256, 0, st>>>(T21, NBc, (long long)Fa * NBc, A, ld, sA, rW, p1, rowsW, NBc, nullptr,
                                                    nullptr, sT, Tch, Tcl, 0, NBc, rowsW);
        if ((rc = check_launch("more::copy_panel_kernel"))) return rc;
        // A[rW:, rW:] -= [T0 T1] [T0 T1]^T (lower triangle)
        rc = launch_tc_bgemm_h16_ex(Kc, rowsW, rowsW, 2 * NBc, -1.f, alphaT, Tch, Tcl, Tch, Tcl,
                                    A + (long long)rW * ld + rW, ld, sA, 1.f, 0, 1, st);
        if (rc) return rc;
        p += NBc;         // two panels done
        continue;
      }
      potrf<<<Kc, 256, potrf_smem, st>>>(A, ld, sA, p, nb, Sinv, ok, k0);
      if ((rc = check_launch("more::potrf_inv_kernel"))) return rc;
      const int r0 = p + nb, rows = Fa - r0;
      if (rows <= 0) continue;
      // T21 = A21 Linv11^T
      rc = launch_bgemm_ex(0, 1, Kc, rows, nb, nb, 1.f, A + (long long)r0 * ld + p, ld, sA, Sinv, more::NB,
                           (long long)more::NB * more::NB, T21, more::NB, (long long)Fa * more::NB, nullptr, 0, 0.f, 0,
                           st);
      if (rc) return rc;
      // the tensor-core trailing update wants a reduction length that is a multiple of 4 (8 for fp16) and enough
      // rows to fill a tile
      const bool tc_trail = tc && nb % 8 == 0 && rows >= 64;
      dim3 g4((unsigned)min((long long)1024, ((long long)rows * nb + 255) / 256), Kc);
      more::copy_panel_kernel<<<g4, 256, 0, st>>>(T21, more::NB, (long long)Fa * more::NB, A, ld, sA, r0, p, rows, nb,
                                                  tc_trail ? T21h : nullptr, tc_trail ? T21l : nullptr,
                                                  tc_trail && h16 ? sT : nullptr, nullptr, nullptr, 0, 0, 0);
      if ((rc = check_launch("more::copy_panel_kernel"))) return rc;
      // A22 -= T21 T21^T (lower triangle)
      if (tc_trail && h16)
        rc = launch_tc_bgemm_h16_ex(Kc, rows, rows, nb, -1.f, alphaT, T21h, T21l, T21h, T21l,
                                    A + (long long)r0 * ld + r0, ld, sA, 1.f, 0, 1, st);
      else if (tc_trail)
        rc = launch_tc_bgemm_ex(Kc, rows, rows, nb, -1.f, T21h, T21l, T21h, T21l, A + (long long)r0 * ld + r0, ld, sA,
                                1.f, 0, 1, st);
      else
        rc = launch_bgemm_ex(0, 1, Kc, rows, rows, nb, -1.f, T21, more::NB, (long long)Fa * more::NB, T21, more::NB,
                             (long long)Fa * more::NB, A + (long long)r0 * ld + r0, ld, sA, nullptr, 0, 1.f, 1, st);
      if (rc) return rc;
    }
    more::backsolve_unpack_kernel<<<Kc, 1024, F * sizeof(float), st>>>(A, ld, sA, F, D, theta, Qz, rz);
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
