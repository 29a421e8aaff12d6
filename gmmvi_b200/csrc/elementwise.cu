// Parameter preparation, mixture reductions, importance weights, diagonal-covariance kernels,
// weight updates and the counter-based normal generator.  Everything here is bandwidth- or
// latency-bound; no tensor-core work.
#include "common.cuh"
#include <stdlib.h>
#include "../../include/gmmvi_b200.h"
#include <stdarg.h>
#include <stdio.h>

namespace gvi {

// ---- error plumbing -------------------------------------------------------------------------
static thread_local char g_last_error[512] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("%s: %s", what, cudaGetErrorString(e));
    return GVI_ERR_CUDA;
  }
  return GVI_OK;
}
int launch_bgemm(int transA, int transB, int batch, int M, int N, int Kd, float alpha, const float* A, int lda,
                 long long strideA, const float* B, int ldb, long long strideB, float* C, int ldc,
                 long long strideC, cudaStream_t st);
int launch_stein_stats_full(const float* X, int N, int D, const float* means, const float* W,
                            const uint8_t* active, const float* G, int K, float* M, cudaStream_t st);
int launch_gemm_auto(int transA, int transB, int batch, int M, int N, int Kd, float alpha, const float* A, int lda,
                     long long strideA, const float* B, int ldb, long long strideB, float* C, int ldc,
                     long long strideC, float* ws, size_t ws_floats, cudaStream_t st);
size_t tc_gemm_workspace_floats(int batch, int M, int N, int Kd);
bool stein_tc_supported(int N, int D);
int launch_logdens_diag2(const float* X, int N, int D, const float* means, const float* stds, int K, float* lq,
                         cudaStream_t st);
size_t stein_diag_workspace_floats(int N, int K, int D);
int launch_stein_diag2(const float* X, int N, int D, const float* means, const float* stds, const float* W, const float* G,
                       int K, float* Hneg, float* gneg, float* ws, cudaStream_t st);
size_t stein_tc_workspace_floats(int N, int K, int D);
int launch_stein_stats_tc(const float* X, int N, int D, const float* means, const float* W, const uint8_t* active,
                          const float* G, int K, float* M, float* ws, cudaStream_t st);

// =================================================================================================
// prepare_full: one CTA per component, fp64 arithmetic, thread-per-column forward substitution.
// =================================================================================================
__global__ void __launch_bounds__(256)
prepare_full_kernel(const float* __restrict__ chol, int D, float* __restrict__ linv, float* __restrict__ prec,
                    float* __restrict__ cst, int32_t* __restrict__ ok, double* __restrict__ ws) {
  const int k = blockIdx.x;
  const float* L = chol + (long long)k * D * D;
  double* Y = ws + (long long)k * D * D;
  __shared__ double red[33];
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  // const part and diagonal check
  double ls = 0.0;
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    const float d = L[(long long)i * D + i];
    if (!(d > 0.f) || !isfinite(d)) bad = 1;
    ls += log((double)d);
  }
  ls = block_sum(ls, red);
  if (threadIdx.x == 0) {
    cst[k] = (float)(-ls - 0.5 * D * kLog2PiD);
    if (ok) ok[k] = bad ? 0 : 1;
  }
  // Y = L^-1, column c per thread.  Rows are produced four at a time so that every Y[m][c] fetched from L2 feeds
  // four accumulators (the column-oriented substitution is bound by the re-reads of Y, not by the flops).
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    double* Yc = Y + c;
    for (int i = 0; i < c; ++i) Yc[(long long)i * D] = 0.0;
    Yc[(long long)c * D] = 1.0 / (double)L[(long long)c * D + c];
    for (int i0 = c + 1; i0 < D; i0 += 4) {
      const int nb = min(4, D - i0);
      const float* L0 = L + (long long)i0 * D;
      const float* L1 = L + (long long)min(i0 + 1, D - 1) * D;
      const float* L2 = L + (long long)min(i0 + 2, D - 1) * D;
      const float* L3 = L + (long long)min(i0 + 3, D - 1) * D;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int m = c;
      for (; m + 1 < i0; m += 2) {
        const double ya = Yc[(long long)m * D], yb = Yc[(long long)(m + 1) * D];
        s0 = fma((double)L0[m], ya, s0); s1 = fma((double)L1[m], ya, s1);
        s2 = fma((double)L2[m], ya, s2); s3 = fma((double)L3[m], ya, s3);
        s0 = fma((double)L0[m + 1], yb, s0); s1 = fma((double)L1[m + 1], yb, s1);
        s2 = fma((double)L2[m + 1], yb, s2); s3 = fma((double)L3[m + 1], yb, s3);
      }
      if (m < i0) {
        const double ya = Yc[(long long)m * D];
        s0 = fma((double)L0[m], ya, s0); s1 = fma((double)L1[m], ya, s1);
        s2 = fma((double)L2[m], ya, s2); s3 = fma((double)L3[m], ya, s3);
      }
      const double y0 = -s0 / (double)L0[i0];
      Yc[(long long)i0 * D] = y0;
      if (nb > 1) {
        s1 = fma((double)L1[i0], y0, s1);
        const double y1 = -s1 / (double)L1[i0 + 1];
        Yc[(long long)(i0 + 1) * D] = y1;
        if (nb > 2) {
          s2 = fma((double)L2[i0], y0, s2); s2 = fma((double)L2[i0 + 1], y1, s2);
          const double y2 = -s2 / (double)L2[i0 + 2];
          Yc[(long long)(i0 + 2) * D] = y2;
          if (nb > 3) {
            s3 = fma((double)L3[i0], y0, s3); s3 = fma((double)L3[i0 + 1], y1, s3); s3 = fma((double)L3[i0 + 2], y2, s3);
            Yc[(long long)(i0 + 3) * D] = -s3 / (double)L3[i0 + 3];
          }
        }
      }
    }
  }
  __syncthreads();
  float* Lo = linv + (long long)k * D * D;
  for (long long e = threadIdx.x; e < (long long)D * D; e += blockDim.x) Lo[e] = (float)Y[e];
  if (prec == nullptr) return;
  // P = Y^T Y, upper part per thread-column then mirrored
  float* P = prec + (long long)k * D * D;
  for (int b = threadIdx.x; b < D; b += blockDim.x) {
    for (int a0 = 0; a0 <= b; a0 += 4) {       // four columns a per pass over column b
      const int a1 = min(a0 + 1, D - 1), a2 = min(a0 + 2, D - 1), a3 = min(a0 + 3, D - 1);
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      for (int i = b; i < D; ++i) {
        const double* yr = Y + (long long)i * D;
        const double yb = yr[b];
        s0 = fma(yr[a0], yb, s0); s1 = fma(yr[a1], yb, s1);
        s2 = fma(yr[a2], yb, s2); s3 = fma(yr[a3], yb, s3);
      }
      const double sv[4] = {s0, s1, s2, s3};
      for (int r = 0; r < 4 && a0 + r <= b; ++r) {
        const float v = (float)sv[r];
        P[(long long)(a0 + r) * D + b] = v;
        P[(long long)b * D + a0 + r] = v;
      }
    }
  }
}

// =================================================================================================
// prepare_blocked: Linv = L^-1 by block forward substitution, one CTA per (component, 32-column block).
//   X_jj = inv(L_jj);   X_ij = -inv(L_ii) * sum_{k=j}^{i-1} L_ik X_kj   (i > j)
// Column blocks of the inverse are independent, so the grid is K * ceil(D/32) CTAs; all arithmetic and the
// running column block are fp64 (shared memory), only the final store rounds to fp32 -- the same accuracy as the
// column-per-thread kernel above at ~1/20 of its time (that kernel re-reads its fp64 scratch from L2).
// D <= 256.
// =================================================================================================
constexpr int PB = 32;          // block size
constexpr int PP = 34;          // shared-memory pitch in doubles (16-byte aligned rows for double2 loads)

// acc[2][2] += sum_m At[m * PP + 2ty..2ty+1] * Bm[m * PP + 2tx..2tx+1]  for m in [m0, m1)
__device__ __forceinline__ void pb_acc(double (&acc)[2][2], const double* At, const double* Bm, int ty, int tx, int m0,
                                       int m1) {
#pragma unroll 8
  for (int m = m0; m < m1; ++m) {
    const double2 a = *reinterpret_cast<const double2*>(At + m * PP + 2 * ty);
    const double2 b = *reinterpret_cast<const double2*>(Bm + m * PP + 2 * tx);
    acc[0][0] = fma(a.x, b.x, acc[0][0]);
    acc[0][1] = fma(a.x, b.y, acc[0][1]);
    acc[1][0] = fma(a.y, b.x, acc[1][0]);
    acc[1][1] = fma(a.y, b.y, acc[1][1]);
  }
}

// Loads the 32x32 block (bi, bj) of L transposed: Lt[m * PP + r] = L[bi*32 + r][bj*32 + m] (identity padding past D).
__device__ __forceinline__ void pb_load_T(double* Lt, const float* __restrict__ L, int D, int bi, int bj) {
  for (int e = threadIdx.x; e < PB * PB; e += blockDim.x) {
    const int r = e >> 5, m = e & 31;
    const int gr = bi * PB + r, gc = bj * PB + m;
    double v = (gr == gc) ? 1.0 : 0.0;
    if (gr < D && gc < D) v = (double)L[(long long)gr * D + gc];
    Lt[m * PP + r] = v;
  }
}

// inverse of a diagonal block by ONE WARP (lane = column c of the inverse = row c of Dt):
// Dt[m * PP + r] = inv(L_bb)[r][m], from Lt[m * PP + r] = L_bb[r][m]
__device__ __forceinline__ void pb_diag_inverse_warp(const double* Lt, double* Dt) {
  const int c = threadIdx.x & 31;
  double* y = Dt + c * PP;
  for (int r = 0; r < c; ++r) y[r] = 0.0;
  y[c] = 1.0 / Lt[c * PP + c];
  for (int r = c + 1; r < PB; ++r) {
    double sacc = 0.0;
    for (int m = c; m < r; ++m) sacc = fma(Lt[m * PP + r], y[m], sacc);
    y[r] = -sacc / Lt[r * PP + r];
  }
}

// Inverses of all diagonal 32 x 32 blocks, one warp per block: dinv[b][m * 32 + r] = inv(L_ii)[r][m] for b = k * nb + i.
// Every column-block CTA of prepare_blocked_kernel needs inv(L_ii) for all i >= j; computing them there (one warp busy
// for ~7 k cycles, seven idle, once per (i, j) pair) was most of that kernel's critical path.
constexpr int PD_WARPS = 4;
__global__ void __launch_bounds__(32 * PD_WARPS)
prepare_diag_kernel(const float* __restrict__ chol, int D, int nb, int nblocks, double* __restrict__ dinv) {
  extern __shared__ __align__(16) double pd_sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * PD_WARPS + warp;
  if (b >= nblocks) return;
  double* Lt = pd_sm + (size_t)warp * 2 * PB * PP;
  double* Dt = Lt + PB * PP;
  const int k = b / nb, i = b - k * nb;
  const float* L = chol + (long long)k * D * D;
  for (int e = lane; e < PB * PB; e += 32) {
    const int r = e >> 5, m = e & 31;
    const int gr = i * PB + r, gc = i * PB + m;
    double v = (gr == gc) ? 1.0 : 0.0;
    if (gr < D && gc < D) v = (double)L[(long long)gr * D + gc];
    Lt[m * PP + r] = v;
  }
  __syncwarp();
  pb_diag_inverse_warp(Lt, Dt);
  __syncwarp();
  double* out = dinv + (size_t)b * PB * PB;
  for (int e = lane; e < PB * PB; e += 32) out[e] = Dt[(e >> 5) * PP + (e & 31)];
}

__global__ void __launch_bounds__(256)
prepare_blocked_kernel(const float* __restrict__ chol, int D, int nb, const double* __restrict__ dinv,
                       float* __restrict__ linv, float* __restrict__ cst, int32_t* __restrict__ ok) {
  extern __shared__ __align__(16) double pb_sm[];
  constexpr int BLK = PB * PP;                      // doubles per padded 32 x 32 block
  double* Xs = pb_sm;                               // [nb] blocks: the column block of X being built
  double* Lt = pb_sm + (size_t)nb * BLK;            // current block of L, transposed
  double* Dt = Lt + BLK;                            // inverse of a diagonal block, transposed
  double* Ts = Dt + BLK;                            // T = sum_k L_ik X_kj
  const int k = blockIdx.x / nb, j = blockIdx.x % nb;
  const float* L = chol + (long long)k * D * D;
  float* Xo = linv + (long long)k * D * D;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  __shared__ double red[33];
  __shared__ int bad;

  if (j == 0) {     // log-normaliser and diagonal check once per component
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    double ls = 0.0;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      const float d = L[(long long)i * D + i];
      if (!(d > 0.f) || !isfinite(d)) bad = 1;
      ls += log((double)d);
    }
    ls = block_sum(ls, red);
    if (threadIdx.x == 0) {
      cst[k] = (float)(-ls - 0.5 * D * kLog2PiD);
      if (ok) ok[k] = bad ? 0 : 1;
    }
    __syncthreads();
  }
  const double* dk = dinv + (size_t)k * nb * PB * PB;

  // ---- X_jj = inv(L_jj): dinv holds it transposed ----
  for (int e = threadIdx.x; e < PB * PB; e += blockDim.x) {
    const int c = e >> 5, r = e & 31;
    Xs[r * PP + c] = dk[(size_t)j * PB * PB + e];
  }
  __syncthreads();

  // ---- X_ij, i > j.  The blocks L_ik of the running sum are loaded one step ahead into registers (four elements per
  // thread), so the L2 latency of a block overlaps the product with the previous one. ----
  float pre[4];
  auto fetch = [&](int bi, int bj) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = threadIdx.x + 256 * q;
      const int r = e >> 5, m = e & 31;
      const int gr = bi * PB + r, gc = bj * PB + m;
      float v = (gr == gc) ? 1.f : 0.f;
      if (gr < D && gc < D) v = __ldg(L + (long long)gr * D + gc);
      pre[q] = v;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = threadIdx.x + 256 * q;
      Lt[(e & 31) * PP + (e >> 5)] = (double)pre[q];
    }
  };
  if (j + 1 < nb) fetch(j + 1, j);
  for (int i = j + 1; i < nb; ++i) {
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    for (int kk = j; kk < i; ++kk) {
      stash();
      __syncthreads();
      if (kk + 1 < i) fetch(i, kk + 1);
      else if (i + 1 < nb) fetch(i + 1, j);
      pb_acc(acc, Lt, Xs + (size_t)(kk - j) * BLK, ty, tx, 0, PB);
      __syncthreads();
    }
    Ts[(2 * ty) * PP + 2 * tx] = acc[0][0];
    Ts[(2 * ty) * PP + 2 * tx + 1] = acc[0][1];
    Ts[(2 * ty + 1) * PP + 2 * tx] = acc[1][0];
    Ts[(2 * ty + 1) * PP + 2 * tx + 1] = acc[1][1];
    for (int e = threadIdx.x; e < PB * PB; e += blockDim.x) Dt[(e >> 5) * PP + (e & 31)] = dk[(size_t)i * PB * PB + e];
    __syncthreads();
    double out[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    pb_acc(out, Dt, Ts, ty, tx, 0, 2 * ty + 2);      // inv(L_ii) is lower triangular: m <= row
    double* Xi = Xs + (size_t)(i - j) * BLK;
    Xi[(2 * ty) * PP + 2 * tx] = -out[0][0];
    Xi[(2 * ty) * PP + 2 * tx + 1] = -out[0][1];
    Xi[(2 * ty + 1) * PP + 2 * tx] = -out[1][0];
    Xi[(2 * ty + 1) * PP + 2 * tx + 1] = -out[1][1];
    __syncthreads();
  }

  // ---- store column block j (zeros above the diagonal block) ----
  for (int bi = 0; bi < nb; ++bi) {
    for (int e = threadIdx.x; e < PB * PB; e += blockDim.x) {
      const int r = e >> 5, c = e & 31;
      const int gr = bi * PB + r, gc = j * PB + c;
      if (gr < D && gc < D) Xo[(long long)gr * D + gc] = bi < j ? 0.f : (float)Xs[(size_t)(bi - j) * BLK + r * PP + c];
    }
  }
}

// =================================================================================================
// mixture logsumexp over components: out[n] = LSE_k(lq[k,n] + logw[k])
// =================================================================================================
// 64 samples per CTA x 4 interleaved slices of the component range (online max / sum per slice, merged through
// shared memory): 4 x the parallelism of a thread per sample, which matters for sharded runs with few samples per GPU.
__global__ void __launch_bounds__(256)
mixture_lse_kernel(const float* __restrict__ lq, const float* __restrict__ logw, int K, int N, float* __restrict__ out) {
  __shared__ float sm_m[4][64], sm_s[4][64];
  const int sx = threadIdx.x & 63, sl = threadIdx.x >> 6;
  const int n = blockIdx.x * 64 + sx;
  float m = -INFINITY, s = 0.f;
  if (n < N) {
    for (int k = sl; k < K; k += 4) {
      const float v = __ldg(lq + (long long)k * N + n) + __ldg(logw + k);
      if (v > m) {
        s = s * expf(m - v) + 1.f;   // m == -inf -> s*0
        m = v;
      } else if (v > -INFINITY) {
        s += expf(v - m);
      }
    }
  }
  sm_m[sl][sx] = m;
  sm_s[sl][sx] = s;
  __syncthreads();
  if (sl == 0 && n < N) {
    float mm = fmaxf(fmaxf(sm_m[0][sx], sm_m[1][sx]), fmaxf(sm_m[2][sx], sm_m[3][sx]));
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (sm_m[i][sx] > -INFINITY) ss += sm_s[i][sx] * expf(sm_m[i][sx] - mm);
    out[n] = (mm > -INFINITY) ? mm + logf(ss) : -INFINITY;
  }
}

// =================================================================================================
// diagonal log densities: CTA = 32 samples (lanes) x all components (warps stride over k)
// =================================================================================================
__global__ void __launch_bounds__(256)
logdens_diag_kernel(const float* __restrict__ X, int N, int D, const float* __restrict__ means,
                    const float* __restrict__ stds, int K, float* __restrict__ lq) {
  extern __shared__ float Xs[];   // [D][32]
  const int n0 = blockIdx.x * 32;
  for (int e = threadIdx.x; e < 32 * D; e += blockDim.x) {
    const int r = e / D, d = e % D;
    Xs[d * 32 + r] = (n0 + r < N) ? X[(long long)(n0 + r) * D + d] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int k = w + blockIdx.y * nw; k < K; k += nw * gridDim.y) {
    const float* mu = means + (long long)k * D;
    const float* sg = stds + (long long)k * D;
    float ls = 0.f;
    for (int d = lane; d < D; d += 32) ls += logf(sg[d]);
    ls = warp_sum(ls);
    float acc = 0.f;
    for (int d = 0; d < D; ++d) {
      const float t = (1.f / __ldg(sg + d)) * (__ldg(mu + d) - Xs[d * 32 + lane]);
      acc = fmaf(t, t, acc);
    }
    if (n0 + lane < N) lq[(long long)k * N + n0 + lane] = (-0.5f * D * kLog2Pi - ls) - 0.5f * acc;
  }
}

// grad[n,d] = - sum_k r_kn (x_nd - mu_kd) / std_kd^2;  CTA = 32 samples, threads over d
__global__ void __launch_bounds__(256)
mixture_grad_diag_kernel(const float* __restrict__ X, int N, int D, const float* __restrict__ means,
                         const float* __restrict__ stds, const float* __restrict__ lq,
                         const float* __restrict__ logw, const float* __restrict__ logq, int K,
                         float* __restrict__ grad) {
  __shared__ float rs[64][33];
  const int n0 = blockIdx.x * 32;
  for (int d0 = 0; d0 < D; d0 += blockDim.x) {
    const int d = d0 + threadIdx.x;
    float x[32], acc[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      x[r] = (d < D && n0 + r < N) ? X[(long long)(n0 + r) * D + d] : 0.f;
      acc[r] = 0.f;
    }
    for (int k0 = 0; k0 < K; k0 += 64) {
      __syncthreads();
      bool any = false;
      for (int e = threadIdx.x; e < 64 * 32; e += blockDim.x) {
        const int kk = e >> 5, r = e & 31;
        const int k = k0 + kk, n = n0 + r;
        float v = 0.f;
        if (k < K && n < N) {
          const float a = lq[(long long)k * N + n] + logw[k] - logq[n];
          v = a > -60.f ? expf(a) : 0.f;
        }
        rs[kk][r] = v;
        any |= v > 0.f;
      }
      if (!__syncthreads_or(any)) continue;
      if (d < D) {
        for (int kk = 0; kk < 64 && k0 + kk < K; ++kk) {
          const float mu = means[(long long)(k0 + kk) * D + d];
          const float sg = stds[(long long)(k0 + kk) * D + d];
          const float iv = 1.f / (sg * sg);
#pragma unroll
          for (int r = 0; r < 32; ++r) acc[r] = fmaf(rs[kk][r] * iv, x[r] - mu, acc[r]);
        }
      }
    }
    if (d < D) {
#pragma unroll
      for (int r = 0; r < 32; ++r)
        if (n0 + r < N) grad[(long long)(n0 + r) * D + d] = -acc[r];
    }
  }
}

// =================================================================================================
// importance weights: one CTA per component row
// =================================================================================================
__global__ void __launch_bounds__(512)
importance_weights_kernel(const float* __restrict__ lq, const float* __restrict__ bg,
                          const int32_t* __restrict__ rel_map, int K, int N, int self_normalized,
                          const float* __restrict__ rho, float* __restrict__ W, float* __restrict__ dot,
                          float* __restrict__ ess, uint8_t* __restrict__ active) {
  __shared__ float red[33];
  const int k = blockIdx.x;
  const float* row = lq + (long long)k * N;
  const int nblk = ceil_div(N, 128);
  if (rel_map != nullptr) {   // only_use_own_samples: uniform weights over the component's own samples
    float cnt = 0.f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) cnt += (rel_map[n] == k) ? 1.f : 0.f;
    cnt = block_sum(cnt, red);
    const float w1 = cnt > 0.f ? 1.f / cnt : 0.f;
    float d = 0.f;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      const bool own = rel_map[n] == k;
      if (W) W[(long long)k * N + n] = own ? w1 : 0.f;
      if (own && rho) d += w1 * rho[n];
      if (own && active) active[(long long)k * nblk + (n >> 7)] = 1;
    }
    if (dot) {
      d = block_sum(d, red);
      if (threadIdx.x == 0) dot[k] = d;
    }
    if (ess && threadIdx.x == 0) ess[k] = cnt;
    return;
  }
  // pass 1: max and sum exp
  float m = -INFINITY;
  for (int n = threadIdx.x; n < N; n += blockDim.x) m = fmaxf(m, row[n] - bg[n]);
  m = block_max(m, red);
  if (!(m > -INFINITY) || !isfinite(m)) m = 0.f;   // tf.reduce_logsumexp: non-finite max -> 0
  float s = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) s += expf(row[n] - bg[n] - m);
  s = block_sum(s, red);
  const float lse = m + logf(s);
  // pass 2: second normaliser and sum of squares of the singly-normalised weights
  float s2 = 0.f, sq = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float w = expf(row[n] - bg[n] - lse);
    s2 += w;
    sq = fmaf(w, w, sq);
  }
  s2 = block_sum(s2, red);
  sq = block_sum(sq, red);
  if (ess && threadIdx.x == 0) ess[k] = 1.f / sq;
  if (W == nullptr && dot == nullptr && active == nullptr) return;
  const float inv_s2 = 1.f / s2;
  const float logN = self_normalized == 2 ? 0.f : logf((float)N);     // mode 2: plain exp(lw) (MORE, ng_estimator.py:356)
  float d = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float lw = row[n] - bg[n];
    const float w = self_normalized == 1 ? expf(lw - lse) * inv_s2 : expf(lw - logN);
    if (W) W[(long long)k * N + n] = w;
    if (rho) d = fmaf(w, rho[n], d);
    if (active && (lw - m) > -60.f) active[(long long)k * nblk + (n >> 7)] = 1;
  }
  if (dot) {
    d = block_sum(d, red);
    if (threadIdx.x == 0) dot[k] = d;
  }
}

// Same computation with the row lw[n] = lq[k, n] - bg[n] held ON CHIP between the passes: 48 K floats in shared memory
// plus 16 per thread in registers (1024 threads), i.e. rows of up to 65536 samples.  The kernel above re-reads the row
// from L2 / HBM for each of its four passes (0.28 ms per call at C5, 2.4 TB/s of mostly L2 traffic); here every row is
// read once and W written once.  Persistent: one CTA per SM walks the rows k = blockIdx.x, + gridDim.x, ...
constexpr int IWC_THREADS = 1024;
constexpr int IWC_SMEM4 = 12288;                 // float4 slots in shared memory (192 KB)
constexpr int IWC_REG4 = 4;                      // float4 per thread in registers
constexpr int IWC_MAX_N = 4 * (IWC_SMEM4 + IWC_REG4 * IWC_THREADS);

__global__ void __launch_bounds__(IWC_THREADS, 1)
importance_weights_cached_kernel(const float* __restrict__ lq, const float* __restrict__ bg, int K, int N,
                                 int self_normalized, const float* __restrict__ rho, float* __restrict__ W,
                                 float* __restrict__ dot, float* __restrict__ ess, uint8_t* __restrict__ active) {
  extern __shared__ __align__(16) float4 iw_row[];
  __shared__ float red[33];
  const int tid = threadIdx.x;
  const int N4 = N >> 2;                           // N % 4 == 0 (checked by the host)
  const int ns4 = min(N4, IWC_SMEM4);
  const int nblk = ceil_div(N, 128);
  const float4* __restrict__ bg4 = reinterpret_cast<const float4*>(bg);
  const float4* __restrict__ rho4 = reinterpret_cast<const float4*>(rho);
  const float logN = self_normalized == 2 ? 0.f : logf((float)N);      // mode 2: plain exp(lw) (MORE)
  const bool sn = self_normalized == 1;
  for (int k = blockIdx.x; k < K; k += gridDim.x) {
    const float4* __restrict__ row4 = reinterpret_cast<const float4*>(lq + (long long)k * N);
    float4 r[IWC_REG4];
    float m = -INFINITY;
    for (int i = tid; i < ns4; i += IWC_THREADS) {
      const float4 a = __ldcs(row4 + i), b = __ldg(bg4 + i);
      const float4 v = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
      iw_row[i] = v;
      m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    }
#pragma unroll
    for (int j = 0; j < IWC_REG4; ++j) {
      const int i = IWC_SMEM4 + tid + j * IWC_THREADS;
      r[j] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);     // past the row: weight exp(-inf) = 0
      if (i < N4) {
        const float4 a = __ldcs(row4 + i), b = __ldg(bg4 + i);
        r[j] = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
        m = fmaxf(fmaxf(m, fmaxf(r[j].x, r[j].y)), fmaxf(r[j].z, r[j].w));
      }
    }
    m = block_max(m, red);
    if (!(m > -INFINITY) || !isfinite(m)) m = 0.f;   // tf.reduce_logsumexp: non-finite max -> 0
    float s = 0.f;
    for (int i = tid; i < ns4; i += IWC_THREADS) {
      const float4 v = iw_row[i];
      s += expf(v.x - m) + expf(v.y - m) + expf(v.z - m) + expf(v.w - m);
    }
#pragma unroll
    for (int j = 0; j < IWC_REG4; ++j) s += expf(r[j].x - m) + expf(r[j].y - m) + expf(r[j].z - m) + expf(r[j].w - m);
    s = block_sum(s, red);
    const float lse = m + logf(s);
    float s2 = 0.f, sq = 0.f;
    auto acc2 = [&](float v) {
      const float w = expf(v - lse);
      s2 += w;
      sq = fmaf(w, w, sq);
    };
    for (int i = tid; i < ns4; i += IWC_THREADS) {
      const float4 v = iw_row[i];
      acc2(v.x); acc2(v.y); acc2(v.z); acc2(v.w);
    }
#pragma unroll
    for (int j = 0; j < IWC_REG4; ++j) { acc2(r[j].x); acc2(r[j].y); acc2(r[j].z); acc2(r[j].w); }
    s2 = block_sum(s2, red);
    sq = block_sum(sq, red);
    if (ess && tid == 0) ess[k] = 1.f / sq;
    if (W == nullptr && dot == nullptr && active == nullptr) continue;
    const float inv_s2 = 1.f / s2;
    float d = 0.f;
    float4* __restrict__ W4 = W ? reinterpret_cast<float4*>(W + (long long)k * N) : nullptr;
    auto fin = [&](const float4 v, int i) {
      float4 w;
      w.x = sn ? expf(v.x - lse) * inv_s2 : expf(v.x - logN);
      w.y = sn ? expf(v.y - lse) * inv_s2 : expf(v.y - logN);
      w.z = sn ? expf(v.z - lse) * inv_s2 : expf(v.z - logN);
      w.w = sn ? expf(v.w - lse) * inv_s2 : expf(v.w - logN);
      if (W4) __stcs(W4 + i, w);
      if (rho) {
        const float4 q = __ldg(rho4 + i);
        d = fmaf(w.x, q.x, d); d = fmaf(w.y, q.y, d); d = fmaf(w.z, q.z, d); d = fmaf(w.w, q.w, d);
      }
      if (active && (fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)) - m) > -60.f) active[(long long)k * nblk + (i >> 5)] = 1;
    };
    for (int i = tid; i < ns4; i += IWC_THREADS) fin(iw_row[i], i);
#pragma unroll
    for (int j = 0; j < IWC_REG4; ++j) {
      const int i = IWC_SMEM4 + tid + j * IWC_THREADS;
      if (i < N4) fin(r[j], i);
    }
    if (dot) {
      d = block_sum(d, red);
      if (tid == 0) dot[k] = d;
    }
    __syncthreads();                                 // the row buffer is reused by the next row
  }
}

// ---- pieces of the same computation for sample-sharded (multi-GPU) runs: the row statistics are reduced
// across ranks between the calls (gmmvi_b200/distributed.py) ----------------------------------------------
__global__ void __launch_bounds__(512)
row_max_kernel(const float* __restrict__ lq, const float* __restrict__ bg, int N, float* __restrict__ out) {
  __shared__ float red[33];
  const int k = blockIdx.x;
  const float* row = lq + (long long)k * N;
  float m = -INFINITY;
  for (int n = threadIdx.x; n < N; n += blockDim.x) m = fmaxf(m, row[n] - bg[n]);
  m = block_max(m, red);
  if (threadIdx.x == 0) out[k] = m;
}

__global__ void __launch_bounds__(512)
row_sumexp_kernel(const float* __restrict__ lq, const float* __restrict__ bg, int N,
                  const float* __restrict__ shift, float* __restrict__ out) {
  __shared__ float red[33];
  const int k = blockIdx.x;
  const float* row = lq + (long long)k * N;
  const float sh = shift[k];
  float s = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) s += expf(row[n] - bg[n] - sh);
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[k] = s;
}

// w = exp(lw - lse[k]) * scale[k]; dot[k] = sum_n w rho[n] (local partial sum)
__global__ void __launch_bounds__(512)
importance_weights_ext_kernel(const float* __restrict__ lq, const float* __restrict__ bg, int N,
                              const float* __restrict__ lse, const float* __restrict__ scale,
                              const float* __restrict__ rowmax, const float* __restrict__ rho,
                              float* __restrict__ W, float* __restrict__ dot, uint8_t* __restrict__ active) {
  __shared__ float red[33];
  const int k = blockIdx.x;
  const float* row = lq + (long long)k * N;
  const int nblk = ceil_div(N, 128);
  const float l = lse[k], sc = scale ? scale[k] : 1.f, m = rowmax ? rowmax[k] : l;
  float d = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float lw = row[n] - bg[n];
    const float w = expf(lw - l) * sc;
    if (W) W[(long long)k * N + n] = w;
    if (rho) d = fmaf(w, rho[n], d);
    if (active && (lw - m) > -60.f) active[(long long)k * nblk + (n >> 7)] = 1;
  }
  if (dot) {
    d = block_sum(d, red);
    if (threadIdx.x == 0) dot[k] = d;
  }
}

// =================================================================================================
// Stein finalisation and diagonal Stein
// =================================================================================================
// gneg[k][d] = - sum_n W[k][n] G[n][d], skipping the 128-sample blocks that carry no weight
__global__ void __launch_bounds__(256)
stein_gsum_kernel(const float* __restrict__ W, const uint8_t* __restrict__ active, const float* __restrict__ G, int N,
                  int D, float* __restrict__ gneg, const int* __restrict__ dense_flag) {
  if (dense_flag != nullptr && *dense_flag != 0) return;       // the tiled kernel below takes the dense case
  const int k = blockIdx.y;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const float* Wk = W + (long long)k * N;
  const int nblk = ceil_div(N, 128);
  float a0 = 0.f, a1 = 0.f;
  for (int b = 0; b < nblk; ++b) {
    if (active != nullptr && active[(long long)k * nblk + b] == 0) continue;
    const int n1 = min(N, (b + 1) * 128);
    if (d < D) {
      int n = b * 128;
      for (; n + 1 < n1; n += 2) {
        a0 = fmaf(__ldg(Wk + n), G[(long long)n * D + d], a0);
        a1 = fmaf(__ldg(Wk + n + 1), G[(long long)(n + 1) * D + d], a1);
      }
      if (n < n1) a0 = fmaf(__ldg(Wk + n), G[(long long)n * D + d], a0);
    }
  }
  if (d < D) gneg[(long long)k * D + d] = -(a0 + a1);
}

// Second generation of the gradient sums (the kernel above re-reads all of G once per component: 34 GB of L2 traffic and
// 7.3 ms at C5 when no block can be skipped).  One CTA = 32 components x a range of 128-sample blocks, one thread = one
// dimension with 32 accumulators: G streams through shared memory 32 samples at a time and is read K / 32 times in
// total; a block is skipped when none of the CTA's components carries weight there.  part[s][k][d], added in a fixed
// order by stein_gsum_reduce_kernel.
constexpr int GS_KT = 32, GS_NS = 32;
// Which of the two gradient-sum kernels runs is decided ON THE DEVICE from the block mask (no host read): with well
// separated components (< 1/8 of the (component, block) pairs carry weight) the per-component kernel above skips almost
// everything and wins (0.10 ms against 0.6 ms at C5); with overlapping components the tiled kernel below wins (0.9 ms
// against 7.3 ms).  Both are launched; the one that is not selected returns at once.
__global__ void stein_density_kernel(const uint8_t* __restrict__ active, long long n, int* __restrict__ dense_flag) {
  __shared__ int red[33];
  int c = 0;
  if (active != nullptr)
    for (long long i = threadIdx.x; i < n; i += blockDim.x) c += active[i] != 0;
  c = block_sum(c, red);
  if (threadIdx.x == 0) *dense_flag = (active == nullptr || (long long)c * 8 > n) ? 1 : 0;
}
__global__ void __launch_bounds__(256)
stein_gsum2_kernel(const float* __restrict__ W, const uint8_t* __restrict__ active, const float* __restrict__ G, int N,
                   int D, int K, int S, float* __restrict__ part, const int* __restrict__ dense_flag) {
  if (*dense_flag == 0) return;
  __shared__ float gs[GS_NS][256];
  __shared__ __align__(16) float ws[GS_NS][GS_KT];
  __shared__ int any_active;
  const int k0 = blockIdx.x * GS_KT, s = blockIdx.y;
  const int d0 = blockIdx.z * 256, d = d0 + threadIdx.x;
  const int nblk = ceil_div(N, 128);
  const int b0 = (int)((long long)nblk * s / S), b1 = (int)((long long)nblk * (s + 1) / S);
  const int nk = min(GS_KT, K - k0);
  float acc[GS_KT];
#pragma unroll
  for (int kk = 0; kk < GS_KT; ++kk) acc[kk] = 0.f;
  for (int b = b0; b < b1; ++b) {
    __syncthreads();
    if (threadIdx.x == 0) any_active = (active == nullptr);
    __syncthreads();
    if (active != nullptr && threadIdx.x < nk && active[(long long)(k0 + threadIdx.x) * nblk + b] != 0) any_active = 1;
    __syncthreads();
    if (!any_active) continue;
    for (int n0 = b * 128; n0 < min(N, (b + 1) * 128); n0 += GS_NS) {
      const int cnt = min(GS_NS, N - n0);
      __syncthreads();
      for (int r = 0; r < GS_NS; ++r) gs[r][threadIdx.x] = (r < cnt && d < D) ? G[(long long)(n0 + r) * D + d] : 0.f;
      for (int e = threadIdx.x; e < GS_NS * GS_KT; e += 256) {
        const int kk = e / GS_NS, r = e % GS_NS;          // consecutive threads read consecutive samples of one row of W
        ws[r][kk] = (kk < nk && r < cnt) ? __ldg(W + (long long)(k0 + kk) * N + n0 + r) : 0.f;
      }
      __syncthreads();
#pragma unroll 4
      for (int r = 0; r < GS_NS; ++r) {
        const float g = gs[r][threadIdx.x];
#pragma unroll
        for (int q = 0; q < GS_KT / 4; ++q) {
          const float4 w = *reinterpret_cast<const float4*>(&ws[r][4 * q]);
          acc[4 * q] = fmaf(w.x, g, acc[4 * q]);
          acc[4 * q + 1] = fmaf(w.y, g, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(w.z, g, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(w.w, g, acc[4 * q + 3]);
        }
      }
    }
  }
  if (d < D) {
#pragma unroll
    for (int kk = 0; kk < GS_KT; ++kk)
      if (kk < nk) part[((long long)s * K + k0 + kk) * D + d] = acc[kk];
  }
}
__global__ void stein_gsum_reduce_kernel(const float* __restrict__ part, int S, long long kd, float* __restrict__ gneg,
                                         const int* __restrict__ dense_flag) {
  if (*dense_flag == 0) return;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < kd; e += (long long)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int s = 0; s < S; ++s) a += part[(long long)s * kd + e];
    gneg[e] = -a;
  }
}

__global__ void stein_finalize_kernel(const float* __restrict__ T, int D, int symmetrize, float* __restrict__ H) {
  const int k = blockIdx.y;
  const float* Tk = T + (long long)k * D * D;
  float* Hk = H + (long long)k * D * D;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)D * D;
       e += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(e / D), b = (int)(e % D);
    const float tab = Tk[(long long)a * D + b], tba = Tk[(long long)b * D + a];
    Hk[e] = symmetrize ? -0.5f * (tab + tba) : -tba;
  }
}

__global__ void __launch_bounds__(256)
stein_diag_kernel(const float* __restrict__ X, int N, int D, const float* __restrict__ means,
                  const float* __restrict__ stds, const float* __restrict__ W, const float* __restrict__ G,
                  float* __restrict__ Hneg, float* __restrict__ gneg) {
  const int k = blockIdx.y;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const float* Wk = W + (long long)k * N;
  const float mu = d < D ? means[(long long)k * D + d] : 0.f;
  const float sg = d < D ? stds[(long long)k * D + d] : 1.f;
  const float iv = 1.f / (sg * sg);
  float h = 0.f, g = 0.f;
  for (int n = 0; n < N; ++n) {
    const float w = __ldg(Wk + n);
    if (w == 0.f) continue;
    if (d < D) {
      const float wg = w * G[(long long)n * D + d];
      g += wg;
      h = fmaf(iv * (X[(long long)n * D + d] - mu), wg, h);
    }
  }
  if (d < D) {
    Hneg[(long long)k * D + d] = -h;
    gneg[(long long)k * D + d] = -g;
  }
}

// =================================================================================================
// weight updates (single CTA; fp32 like the reference so that the bisection takes the same branches)
// =================================================================================================
__device__ float block_lse(const float* v, int K, float* red) {
  float m = -INFINITY;
  for (int k = threadIdx.x; k < K; k += blockDim.x) m = fmaxf(m, v[k]);
  m = block_max(m, red);
  if (!isfinite(m)) m = 0.f;
  float s = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) s += expf(v[k] - m);
  s = block_sum(s, red);
  return m + logf(s);
}

// new_lw <- normalise(floor(normalise(v)));  returns KL(new || old)
__device__ float weight_kl_eval(float eta, float T, const float* logw, const float* elr, int K, float* nl,
                                float* red) {
  const float a = (eta + 1.f) / (T + eta), b = 1.f / (T + eta);
  for (int k = threadIdx.x; k < K; k += blockDim.x) nl[k] = a * logw[k] + b * elr[k];
  __syncthreads();
  float l = block_lse(nl, K, red);
  for (int k = threadIdx.x; k < K; k += blockDim.x) nl[k] = fmaxf(nl[k] - l, -69.07f);
  __syncthreads();
  l = block_lse(nl, K, red);
  float kl = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float v = nl[k] - l;
    nl[k] = v;
    kl += expf(v) * (v - logw[k]);
  }
  kl = block_sum(kl, red);
  return kl;
}

__global__ void __launch_bounds__(1024)
weight_update_kernel(int trust_region, const float* __restrict__ logw, const float* __restrict__ elr, int K,
                     const float* __restrict__ stepsize, float T, float* __restrict__ out, float* __restrict__ info) {
  __shared__ float red[33];
  const float step = stepsize[0];
  if (K <= 1) {   // weight_updater.py:136,275: nothing happens for a single component
    for (int k = threadIdx.x; k < K; k += blockDim.x) out[k] = logw[k];
    if (info && threadIdx.x == 0) { info[0] = -1.f; info[1] = -1.f; }
    return;
  }
  if (!trust_region) {   // weight_updater.py:136-141
    for (int k = threadIdx.x; k < K; k += blockDim.x) out[k] = logw[k] + step / T * elr[k];
    __syncthreads();
    float l = block_lse(out, K, red);
    for (int k = threadIdx.x; k < K; k += blockDim.x) out[k] = fmaxf(out[k] - l, -69.07f);
    __syncthreads();
    l = block_lse(out, K, red);
    for (int k = threadIdx.x; k < K; k += blockDim.x) out[k] -= l;
    if (info && threadIdx.x == 0) { info[0] = -1.f; info[1] = -1.f; }
    return;
  }
  // trust region, weight_updater.py:193-279
  const float kl_bound = step;
  float lower = -45.f, upper = 45.f;
  float log_eta = 0.5f * (upper + lower);
  bool feasible = false;
  float kl = -1.f, eta = -1.f;
  bool evaluated = false;
  for (int it = 0; it < 50; ++it) {
    eta = expf(log_eta);
    const float diff = fabsf(expf(upper) - expf(lower));
    if (diff < 1e-1f) break;
    kl = weight_kl_eval(eta, T, logw, elr, K, out, red);
    evaluated = true;
    if (fabsf(kl_bound - kl) < 1e-1f * kl_bound) {
      lower = upper;
      break;
    }
    if (kl_bound > kl) {
      upper = log_eta;
      feasible = true;
    } else {
      lower = log_eta;
    }
    log_eta = 0.5f * (upper + lower);
  }
  if (lower == upper && evaluated) {
    // keep the weights of the last evaluation
  } else if (feasible) {
    eta = expf(upper);
    kl = weight_kl_eval(eta, T, logw, elr, K, out, red);
  } else {
    kl = -1.f;
    eta = -1.f;
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) out[k] = logw[k];
  }
  if (info && threadIdx.x == 0) { info[0] = kl; info[1] = eta; }
}

// =================================================================================================
// Philox4x32-10 + Box-Muller
// =================================================================================================
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// subseq_dev (nullable): the draw counter lives in device memory (CUDA-graph replays must not bake it into the launch);
// subseq_add is added to it.
__global__ void fill_normal_kernel(float* __restrict__ out, long long rows, int D, unsigned long long seed,
                                   unsigned long long subseq, long long row_offset,
                                   const unsigned long long* __restrict__ subseq_dev) {
  if (subseq_dev != nullptr) subseq += *subseq_dev;
  const int groups = ceil_div(D, 4);
  const long long total = rows * groups;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / groups;
    const int g = (int)(e % groups);
    const unsigned long long grow = (unsigned long long)(r + row_offset);
    uint32_t c[4] = {(uint32_t)grow, (uint32_t)(grow >> 32), (uint32_t)g, (uint32_t)subseq};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32) ^ (uint32_t)(subseq >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      philox_round(c, k0, k1);
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    float z[4];
    {
      const float r0 = sqrtf(-2.f * logf(u01(c[0])));
      float s, co;
      sincospif(2.f * u01(c[1]), &s, &co);
      z[0] = r0 * co; z[1] = r0 * s;
      const float r1 = sqrtf(-2.f * logf(u01(c[2])));
      sincospif(2.f * u01(c[3]), &s, &co);
      z[2] = r1 * co; z[3] = r1 * s;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int d = g * 4 + q;
      if (d < D) out[r * D + d] = z[q];
    }
  }
}

}  // namespace gvi

using namespace gvi;

extern "C" int gvi_version(void) { return 100; }
extern "C" const char* gvi_last_error(void) { return g_last_error; }

extern "C" size_t gvi_prepare_full_workspace(int K, int D) {
  if (K <= 0) return 0;
  if (D <= 256) {   // blocked path: the inverses of the diagonal blocks, then (same memory) the GEMM scratch
    const size_t gemm = tc_gemm_workspace_floats(K, D, D, D) * sizeof(float);
    const size_t diag = (size_t)K * ceil_div(D, PB) * PB * PB * sizeof(double);
    return (gemm > diag ? gemm : diag) + 16;
  }
  return (size_t)K * D * D * sizeof(double);
}
extern "C" int gvi_prepare_full_f32(const float* chol, int K, int D, float* linv, float* prec, float* cst,
                                    int32_t* ok, void* ws, size_t ws_bytes, void* stream) {
  GVI_REQUIRE(K >= 0 && D > 0, "gvi_prepare_full_f32: bad sizes K=%d D=%d", K, D);
  if (K == 0) return GVI_OK;
  GVI_REQUIRE(chol && linv && cst && ws, "gvi_prepare_full_f32: null pointer");
  if (ws_bytes < gvi_prepare_full_workspace(K, D)) {
    set_last_error("gvi_prepare_full_f32: workspace %zu < %zu", ws_bytes, gvi_prepare_full_workspace(K, D));
    return GVI_ERR_WORKSPACE;
  }
  if (D > 256) {
    prepare_full_kernel<<<K, 256, 0, (cudaStream_t)stream>>>(chol, D, linv, prec, cst, ok, (double*)ws);
    return check_launch("prepare_full_kernel");
  }
  const int nb = ceil_div(D, PB);
  const size_t smem = (size_t)(nb + 3) * PB * PP * sizeof(double);
  static unsigned long long attr_set_mask = 0;
  if (first_call_on_device(attr_set_mask)) {
    cudaError_t e = cudaFuncSetAttribute(prepare_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)((size_t)(8 + 3) * PB * PP * sizeof(double)));
    if (e != cudaSuccess) {
      set_last_error("gvi_prepare_full_f32: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return GVI_ERR_CUDA;
    }
  }
  static unsigned long long attr2_set_mask = 0;
  const size_t smem_diag = (size_t)PD_WARPS * 2 * PB * PP * sizeof(double);
  if (first_call_on_device(attr2_set_mask)) {
    cudaError_t e = cudaFuncSetAttribute(prepare_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_diag);
    if (e != cudaSuccess) {
      set_last_error("gvi_prepare_full_f32: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return GVI_ERR_CUDA;
    }
  }
  double* dinv = (double*)ws;
  prepare_diag_kernel<<<ceil_div(K * nb, PD_WARPS), 32 * PD_WARPS, smem_diag, (cudaStream_t)stream>>>(chol, D, nb, K * nb, dinv);
  int rc = check_launch("prepare_diag_kernel");
  if (rc) return rc;
  prepare_blocked_kernel<<<K * nb, 256, smem, (cudaStream_t)stream>>>(chol, D, nb, dinv, linv, cst, ok);
  rc = check_launch("prepare_blocked_kernel");
  if (rc || prec == nullptr) return rc;
  // prec = linv^T linv (batched, tensor cores in 3xTF32 when the shape allows)
  return launch_gemm_auto(1, 0, K, D, D, D, 1.0f, linv, D, (long long)D * D, linv, D, (long long)D * D, prec, D,
                          (long long)D * D, (float*)ws, ws_bytes / sizeof(float), (cudaStream_t)stream);
}

extern "C" int gvi_mixture_lse_f32(const float* lq, const float* logw, int K, int N, float* out, void* stream) {
  GVI_REQUIRE(K >= 0 && N >= 0, "gvi_mixture_lse_f32: bad sizes");
  if (N == 0) return GVI_OK;
  GVI_REQUIRE(lq && logw && out, "gvi_mixture_lse_f32: null pointer");
  mixture_lse_kernel<<<ceil_div(N, 64), 256, 0, (cudaStream_t)stream>>>(lq, logw, K, N, out);
  return check_launch("mixture_lse_kernel");
}

extern "C" int gvi_logdens_diag_f32(const float* X, int N, int D, const float* means, const float* stds, int K,
                                    float* lq, void* stream) {
  GVI_REQUIRE(N >= 0 && D > 0 && K >= 0, "gvi_logdens_diag_f32: bad sizes");
  if (N == 0 || K == 0) return GVI_OK;
  GVI_REQUIRE(X && means && stds && lq, "gvi_logdens_diag_f32: null pointer");
  if (!(getenv("GMMVI_B200_DIAG_V1"))) {              // second-generation kernel (diag.cu); 1 = shape not supported
    const int rc2 = launch_logdens_diag2(X, N, D, means, stds, K, lq, (cudaStream_t)stream);
    if (rc2 != 1) return rc2;
  }
  const size_t smem = (size_t)32 * D * sizeof(float);
  if (smem > 200 * 1024) {
    set_last_error("gvi_logdens_diag_f32: D=%d too large for the sample tile", D);
    return GVI_ERR_UNSUPPORTED;
  }
  static unsigned long long attr_set_mask = 0;
  if (smem > 48 * 1024 && first_call_on_device(attr_set_mask)) {
    cudaFuncSetAttribute(logdens_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  dim3 grid(ceil_div(N, 32), min(ceil_div(K, 8), 8));
  logdens_diag_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(X, N, D, means, stds, K, lq);
  return check_launch("logdens_diag_kernel");
}

extern "C" int gvi_mixture_grad_diag_f32(const float* X, int N, int D, const float* means, const float* stds,
                                         const float* lq, const float* logw, const float* logq, int K, float* grad,
                                         void* stream) {
  GVI_REQUIRE(N >= 0 && D > 0 && K >= 0, "gvi_mixture_grad_diag_f32: bad sizes");
  if (N == 0) return GVI_OK;
  GVI_REQUIRE(X && means && stds && lq && logw && logq && grad, "gvi_mixture_grad_diag_f32: null pointer");
  mixture_grad_diag_kernel<<<ceil_div(N, 32), 256, 0, (cudaStream_t)stream>>>(X, N, D, means, stds, lq, logw, logq,
                                                                             K, grad);
  return check_launch("mixture_grad_diag_kernel");
}

extern "C" int gvi_importance_weights_f32(const float* lq, const float* bg, const int32_t* rel_map, int K, int N,
                                          int self_normalized, const float* rho, float* W, float* dot, float* ess,
                                          uint8_t* active, void* stream) {
  GVI_REQUIRE(K >= 0 && N >= 0, "gvi_importance_weights_f32: bad sizes");
  if (K == 0) return GVI_OK;
  GVI_REQUIRE(lq && (bg || rel_map), "gvi_importance_weights_f32: null pointer");
  GVI_REQUIRE(!dot || rho, "gvi_importance_weights_f32: dot requested without rho");
  cudaStream_t st = (cudaStream_t)stream;
  if (active) {
    cudaError_t e = cudaMemsetAsync(active, 0, (size_t)K * ceil_div(N, 128), st);
    if (e != cudaSuccess) {
      set_last_error("gvi_importance_weights_f32: memset: %s", cudaGetErrorString(e));
      return GVI_ERR_CUDA;
    }
  }
  const bool aligned = (reinterpret_cast<uintptr_t>(lq) | reinterpret_cast<uintptr_t>(bg) | reinterpret_cast<uintptr_t>(rho) |
                        reinterpret_cast<uintptr_t>(W)) % 16 == 0;
  if (rel_map == nullptr && N % 4 == 0 && N >= 4096 && N <= IWC_MAX_N && aligned) {
    static int sms_of_device[64] = {0};             // per device ordinal: SM count, and the smem attribute is set
    int dev = 0;
    cudaGetDevice(&dev);
    int num_sms = (dev >= 0 && dev < 64) ? sms_of_device[dev] : 0;
    if (num_sms == 0) {
      cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
      cudaError_t e = cudaFuncSetAttribute(importance_weights_cached_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           IWC_SMEM4 * 16);
      if (e != cudaSuccess) {
        set_last_error("gvi_importance_weights_f32: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return GVI_ERR_CUDA;
      }
      if (dev >= 0 && dev < 64) sms_of_device[dev] = num_sms;
    }
    importance_weights_cached_kernel<<<min(K, num_sms), IWC_THREADS, IWC_SMEM4 * 16, st>>>(lq, bg, K, N, self_normalized,
                                                                                           rho, W, dot, ess, active);
    return check_launch("importance_weights_cached_kernel");
  }
  importance_weights_kernel<<<K, 512, 0, st>>>(lq, bg, rel_map, K, N, self_normalized, rho, W, dot, ess, active);
  return check_launch("importance_weights_kernel");
}

extern "C" int gvi_row_max_f32(const float* lq, const float* bg, int K, int N, float* out, void* stream) {
  GVI_REQUIRE(K >= 0 && N >= 0, "gvi_row_max_f32: bad sizes");
  if (K == 0) return GVI_OK;
  GVI_REQUIRE(lq && bg && out, "gvi_row_max_f32: null pointer");
  row_max_kernel<<<K, 512, 0, (cudaStream_t)stream>>>(lq, bg, N, out);
  return check_launch("row_max_kernel");
}

extern "C" int gvi_row_sumexp_f32(const float* lq, const float* bg, int K, int N, const float* shift, float* out,
                                  void* stream) {
  GVI_REQUIRE(K >= 0 && N >= 0, "gvi_row_sumexp_f32: bad sizes");
  if (K == 0) return GVI_OK;
  GVI_REQUIRE(lq && bg && shift && out, "gvi_row_sumexp_f32: null pointer");
  row_sumexp_kernel<<<K, 512, 0, (cudaStream_t)stream>>>(lq, bg, N, shift, out);
  return check_launch("row_sumexp_kernel");
}

extern "C" int gvi_importance_weights_ext_f32(const float* lq, const float* bg, int K, int N, const float* lse,
                                              const float* scale, const float* rowmax, const float* rho, float* W,
                                              float* dot, uint8_t* active, void* stream) {
  GVI_REQUIRE(K >= 0 && N >= 0, "gvi_importance_weights_ext_f32: bad sizes");
  if (K == 0) return GVI_OK;
  GVI_REQUIRE(lq && bg && lse, "gvi_importance_weights_ext_f32: null pointer");
  GVI_REQUIRE(!dot || rho, "gvi_importance_weights_ext_f32: dot requested without rho");
  cudaStream_t st = (cudaStream_t)stream;
  if (active) {
    cudaError_t e = cudaMemsetAsync(active, 0, (size_t)K * ceil_div(N, 128), st);
    if (e != cudaSuccess) {
      set_last_error("gvi_importance_weights_ext_f32: memset: %s", cudaGetErrorString(e));
      return GVI_ERR_CUDA;
    }
  }
  importance_weights_ext_kernel<<<K, 512, 0, st>>>(lq, bg, N, lse, scale, rowmax, rho, W, dot, active);
  return check_launch("importance_weights_ext_kernel");
}

static bool stein_tc_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GMMVI_B200_TC_STEIN");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
namespace gvi {
bool small_dim_supported(int D);
size_t stein_small_workspace_floats(int N, int K, int D);
int launch_stein_small(const float* X, int N, int D, const float* means, const float* W, const uint8_t* active,
                       const float* G, int K, float* M, float* gneg, float* ws, cudaStream_t st);
}  // namespace gvi
// splits of 16 blocks (2048 samples): with well separated components the blocks that carry weight for a tile of 32
// components are few and contiguous, so coarse splits would leave all of a tile's work to one CTA (measured: 0.82 ms at C5
// with 18 splits against 0.10 ms for the per-component kernel)
static int stein_gsum_splits(int N, int K) {
  (void)K;
  return max(1, min(64, ceil_div(ceil_div(N, 128), 16)));
}
static size_t stein_gsum_floats(int N, int K, int D) {
  return (N > 0 && K > 0 && D > 32) ? (size_t)stein_gsum_splits(N, K) * K * D + 64 : 0;
}
// scratch of the statistics kernels that precedes the gradient-sum partials in a stein_stats workspace
static size_t stein_stats_kernel_floats(int N, int K, int D) {
  if (N <= 0 || K <= 0) return 0;
  if (D <= 32) return stein_small_workspace_floats(N, K, D);
  return stein_tc_supported(N, D) ? (stein_tc_workspace_floats(N, K, D) + 63) / 64 * 64 : 0;
}
static size_t stein_base_floats(int K, int D) {
  return ((size_t)2 * K * D * D + tc_gemm_workspace_floats(K, D, D, D) + 63) / 64 * 64;
}
extern "C" size_t gvi_stein_full_workspace(int N, int K, int D) {
  if (K <= 0) return 0;
  size_t f = stein_base_floats(K, D);
  f += stein_stats_kernel_floats(N, K, D) + stein_gsum_floats(N, K, D);
  return f * sizeof(float);
}
// raw statistics: M_k = sum_n w_kn (x_n - mu_k) g_n^T and gneg_k = -sum_n w_kn g_n; `tcws` = scratch of the tensor-core kernel
static int stein_stats(const float* X, int N, int D, const float* means, const float* W, const uint8_t* active,
                       const float* G, int K, float* M, float* gneg, float* tcws, cudaStream_t st) {
  float* gsum_ws = (tcws != nullptr && D > 32) ? tcws + stein_stats_kernel_floats(N, K, D) : nullptr;
  int rc;
  if (N > 0 && small_dim_supported(D))      // D <= 32: statistics and gradient sums in one pass (small_dim.cu)
    return launch_stein_small(X, N, D, means, W, active, G, K, M, gneg, tcws, st);
  if (N > 0 && stein_tc_supported(N, D) && stein_tc_enabled())
    rc = launch_stein_stats_tc(X, N, D, means, W, active, G, K, M, tcws, st);
  else
    rc = launch_stein_stats_full(X, N, D, means, W, active, G, K, M, st);
  if (rc) return rc;
  dim3 gg(ceil_div(D, 256), K);
  if (gsum_ws != nullptr && N > 0) {
    int* flag = reinterpret_cast<int*>(gsum_ws);
    float* part = gsum_ws + 64;
    stein_density_kernel<<<1, 1024, 0, st>>>(active, (long long)K * ceil_div(N, 128), flag);
    if ((rc = check_launch("stein_density_kernel"))) return rc;
    stein_gsum_kernel<<<gg, 256, 0, st>>>(W, active, G, N, D, gneg, flag);
    if ((rc = check_launch("stein_gsum_kernel"))) return rc;
    const int S = stein_gsum_splits(N, K);
    dim3 g2(ceil_div(K, GS_KT), S, ceil_div(D, 256));
    stein_gsum2_kernel<<<g2, 256, 0, st>>>(W, active, G, N, D, K, S, part, flag);
    if ((rc = check_launch("stein_gsum2_kernel"))) return rc;
    const long long kd = (long long)K * D;
    stein_gsum_reduce_kernel<<<(int)min((long long)1024, (kd + 255) / 256), 256, 0, st>>>(part, S, kd, gneg, flag);
    return check_launch("stein_gsum_reduce_kernel");
  }
  stein_gsum_kernel<<<gg, 256, 0, st>>>(W, active, G, N, D, gneg, nullptr);
  return check_launch("stein_gsum_kernel");
}
// Hneg_k = -(P_k M_k) (symmetrised when asked); T: K D^2 floats followed by the batched-GEMM scratch
static int stein_finalize(const float* prec, const float* M, int K, int D, int symmetrize, float* Hneg, float* T,
                          cudaStream_t st) {
  int rc = launch_gemm_auto(0, 0, K, D, D, D, 1.f, prec, D, (long long)D * D, M, D, (long long)D * D, T, D,
                            (long long)D * D, T + (size_t)K * D * D, tc_gemm_workspace_floats(K, D, D, D), st);
  if (rc) return rc;
  dim3 grid(min(ceil_div(D * D, 256), 1024), K);
  stein_finalize_kernel<<<grid, 256, 0, st>>>(T, D, symmetrize, Hneg);
  return check_launch("stein_finalize_kernel");
}

extern "C" int gvi_stein_full_f32(const float* X, int N, int D, const float* means, const float* prec,
                                  const float* W, const uint8_t* active, const float* G, int K, int symmetrize,
                                  float* Hneg, float* gneg, void* ws, size_t ws_bytes, void* stream) {
  GVI_REQUIRE(N >= 0 && D > 0 && K >= 0, "gvi_stein_full_f32: bad sizes");
  if (K == 0) return GVI_OK;
  GVI_REQUIRE(X && means && prec && W && G && Hneg && gneg && ws, "gvi_stein_full_f32: null pointer");
  GVI_REQUIRE(K <= 65535, "gvi_stein_full_f32: K=%d exceeds 65535", K);
  if (ws_bytes < gvi_stein_full_workspace(N, K, D)) {
    set_last_error("gvi_stein_full_f32: workspace %zu < %zu", ws_bytes, gvi_stein_full_workspace(N, K, D));
    return GVI_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* M = (float*)ws;
  float* T = M + (size_t)K * D * D;
  int rc = stein_stats(X, N, D, means, W, active, G, K, M, gneg, (float*)ws + stein_base_floats(K, D), st);
  if (rc) return rc;
  return stein_finalize(prec, M, K, D, symmetrize, Hneg, T, st);
}

// The two halves of gvi_stein_full_f32 for sample-sharded runs: the raw statistics are linear in the samples, so the
// ranks reduce-scatter M / gneg by component and every rank finalises only the components it updates.
extern "C" size_t gvi_stein_stats_full_workspace(int N, int K, int D) {
  if (K <= 0 || N <= 0) return 0;
  return (stein_stats_kernel_floats(N, K, D) + stein_gsum_floats(N, K, D)) * sizeof(float);
}
extern "C" int gvi_stein_stats_full_f32(const float* X, int N, int D, const float* means, const float* W,
                                        const uint8_t* active, const float* G, int K, float* M, float* gneg, void* ws,
                                        size_t ws_bytes, void* stream) {
  GVI_REQUIRE(N >= 0 && D > 0 && K >= 0, "gvi_stein_stats_full_f32: bad sizes");
  if (K == 0) return GVI_OK;
  GVI_REQUIRE(X && means && W && G && M && gneg, "gvi_stein_stats_full_f32: null pointer");
  GVI_REQUIRE(K <= 65535, "gvi_stein_stats_full_f32: K=%d exceeds 65535", K);
  if (ws_bytes < gvi_stein_stats_full_workspace(N, K, D) || (ws_bytes > 0 && ws == nullptr)) {
    set_last_error("gvi_stein_stats_full_f32: workspace %zu < %zu", ws_bytes, gvi_stein_stats_full_workspace(N, K, D));
    return GVI_ERR_WORKSPACE;
  }
  return stein_stats(X, N, D, means, W, active, G, K, M, gneg, (float*)ws, (cudaStream_t)stream);
}
extern "C" size_t gvi_stein_finalize_full_workspace(int K, int D) {
  if (K <= 0) return 0;
  return ((size_t)K * D * D + tc_gemm_workspace_floats(K, D, D, D)) * sizeof(float);
}
extern "C" int gvi_stein_finalize_full_f32(const float* prec, const float* M, int K, int D, int symmetrize, float* Hneg,
                                           void* ws, size_t ws_bytes, void* stream) {
  GVI_REQUIRE(D > 0 && K >= 0, "gvi_stein_finalize_full_f32: bad sizes");
  if (K == 0) return GVI_OK;
  GVI_REQUIRE(prec && M && Hneg && ws, "gvi_stein_finalize_full_f32: null pointer");
  GVI_REQUIRE(K <= 65535, "gvi_stein_finalize_full_f32: K=%d exceeds 65535", K);
  if (ws_bytes < gvi_stein_finalize_full_workspace(K, D)) {
    set_last_error("gvi_stein_finalize_full_f32: workspace %zu < %zu", ws_bytes, gvi_stein_finalize_full_workspace(K, D));
    return GVI_ERR_WORKSPACE;
  }
  return stein_finalize(prec, M, K, D, symmetrize, Hneg, (float*)ws, (cudaStream_t)stream);
}

extern "C" size_t gvi_stein_diag_workspace(int N, int K, int D) { return stein_diag_workspace_floats(N, K, D) * sizeof(float); }
extern "C" int gvi_stein_diag_f32(const float* X, int N, int D, const float* means, const float* stds,
                                  const float* W, const float* G, int K, float* Hneg, float* gneg, void* ws,
                                  size_t ws_bytes, void* stream) {
  GVI_REQUIRE(N >= 0 && D > 0 && K >= 0, "gvi_stein_diag_f32: bad sizes");
  if (K == 0) return GVI_OK;
  GVI_REQUIRE(X && means && stds && W && G && Hneg && gneg, "gvi_stein_diag_f32: null pointer");
  // with a workspace: two matrix products W [X o G | G] on the GEMM engines (diag.cu); without: the serial kernel
  if (ws != nullptr && N > 0 && ws_bytes >= gvi_stein_diag_workspace(N, K, D) && !getenv("GMMVI_B200_DIAG_V1"))
    return launch_stein_diag2(X, N, D, means, stds, W, G, K, Hneg, gneg, (float*)ws, (cudaStream_t)stream);
  dim3 grid(ceil_div(D, 256), K);
  stein_diag_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, N, D, means, stds, W, G, Hneg, gneg);
  return check_launch("stein_diag_kernel");
}

extern "C" int gvi_weight_update_f32(int trust_region, const float* logw, const float* elr, int K,
                                     const float* stepsize, float temperature, float* out_logw, float* info,
                                     void* stream) {
  GVI_REQUIRE(K >= 0, "gvi_weight_update_f32: bad K");
  if (K == 0) return GVI_OK;
  GVI_REQUIRE(logw && elr && stepsize && out_logw, "gvi_weight_update_f32: null pointer");
  GVI_REQUIRE(out_logw != logw, "gvi_weight_update_f32: in-place update not supported");
  weight_update_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(trust_region, logw, elr, K, stepsize, temperature,
                                                            out_logw, info);
  return check_launch("weight_update_kernel");
}

extern "C" int gvi_fill_normal_f32(float* out, long long rows, int D, unsigned long long seed,
                                   unsigned long long subsequence, long long row_offset, void* stream) {
  GVI_REQUIRE(rows >= 0 && D > 0, "gvi_fill_normal_f32: bad sizes");
  if (rows == 0) return GVI_OK;
  GVI_REQUIRE(out, "gvi_fill_normal_f32: null pointer");
  const long long total = rows * ceil_div(D, 4);
  const int blocks = (int)min((long long)148 * 16, (total + 255) / 256);
  fill_normal_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, rows, D, seed, subsequence, row_offset, nullptr);
  return check_launch("fill_normal_kernel");
}

extern "C" int gvi_fill_normal_dev_f32(float* out, long long rows, int D, unsigned long long seed,
                                       const unsigned long long* subsequence_dev, unsigned long long subsequence_add,
                                       long long row_offset, void* stream) {
  GVI_REQUIRE(rows >= 0 && D > 0, "gvi_fill_normal_dev_f32: bad sizes");
  if (rows == 0) return GVI_OK;
  GVI_REQUIRE(out && subsequence_dev, "gvi_fill_normal_dev_f32: null pointer");
  const long long total = rows * ceil_div(D, 4);
  const int blocks = (int)min((long long)148 * 16, (total + 255) / 256);
  fill_normal_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, rows, D, seed, subsequence_add, row_offset,
                                                               subsequence_dev);
  return check_launch("fill_normal_kernel");
}

// =================================================================================================
// MMD evaluation (experiments/evaluation/mmd.py:41-60): sum_{i,j} exp(-sum_d w_d (x_id - y_jd)^2), the Gaussian
// kernel with the diagonal bandwidth w_d = 1 / (alpha sigma_d).  compute_ustat is the call with Y = X, kernel_mix the
// call with X = ground truth, Y = model sample.  The differences are formed explicitly in fp32 like the reference
// does (a norm expansion would cancel for near-by points).  One CTA = 64 x 64 pairs, thread = 4 x 4 pairs, the
// coordinates stream through shared memory 32 at a time; the CTA's sum goes to partial[blockIdx] in double and the
// host adds the partials in a fixed order (deterministic).
// =================================================================================================
namespace gvi {
constexpr int MMD_T = 64, MMD_DK = 32;
__global__ void __launch_bounds__(256)
gauss_kernel_sum_kernel(const float* __restrict__ X, int n1, const float* __restrict__ Y, int n2, int D,
                        const float* __restrict__ w, double* __restrict__ partial) {
  __shared__ float xs[MMD_DK][MMD_T + 1], ys[MMD_DK][MMD_T + 1], ws[MMD_DK];
  __shared__ float red[33];
  const int i0 = blockIdx.y * MMD_T, j0 = blockIdx.x * MMD_T;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  for (int d0 = 0; d0 < D; d0 += MMD_DK) {
    for (int e = threadIdx.x; e < MMD_T * MMD_DK; e += 256) {
      const int r = e / MMD_DK, c = e - r * MMD_DK;
      const bool dv = d0 + c < D;
      xs[c][r] = (dv && i0 + r < n1) ? __ldg(X + (long long)(i0 + r) * D + d0 + c) : 0.f;
      ys[c][r] = (dv && j0 + r < n2) ? __ldg(Y + (long long)(j0 + r) * D + d0 + c) : 0.f;
    }
    if (threadIdx.x < MMD_DK) ws[threadIdx.x] = d0 + threadIdx.x < D ? __ldg(w + d0 + threadIdx.x) : 0.f;
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < MMD_DK; ++c) {
      const float wc = ws[c];
      float xv[4], yv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        xv[a] = xs[c][ty + 16 * a];
        yv[a] = ys[c][tx + 16 * a];
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const float df = xv[a] - yv[b];
          acc[a][b] = fmaf(wc * df, df, acc[a][b]);
        }
    }
    __syncthreads();
  }
  float s = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (i0 + ty + 16 * a < n1 && j0 + tx + 16 * b < n2) s += expf(-acc[a][b]);
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.y * gridDim.x + blockIdx.x] = (double)s;
}
}  // namespace gvi

extern "C" size_t gvi_gauss_kernel_sum_partials(int n1, int n2) {
  if (n1 <= 0 || n2 <= 0) return 0;
  return (size_t)gvi::ceil_div(n1, gvi::MMD_T) * gvi::ceil_div(n2, gvi::MMD_T);
}
extern "C" int gvi_gauss_kernel_sum_f32(const float* X, int n1, const float* Y, int n2, int D, const float* w,
                                        double* partial, void* stream) {
  GVI_REQUIRE(n1 >= 0 && n2 >= 0 && D > 0, "gvi_gauss_kernel_sum_f32: bad sizes");
  if (n1 == 0 || n2 == 0) return GVI_OK;
  GVI_REQUIRE(X && Y && w && partial, "gvi_gauss_kernel_sum_f32: null pointer");
  GVI_REQUIRE(gvi::ceil_div(n1, gvi::MMD_T) <= 65535, "gvi_gauss_kernel_sum_f32: n1 too large");
  dim3 grid(gvi::ceil_div(n2, gvi::MMD_T), gvi::ceil_div(n1, gvi::MMD_T));
  gvi::gauss_kernel_sum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, n1, Y, n2, D, w, partial);
  return gvi::check_launch("gauss_kernel_sum_kernel");
}
