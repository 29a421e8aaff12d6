// SIMT fp32 tile-engine kernels: full-covariance log-density, mixture gradient, Stein statistics,
// component sampling and the batched GEMM used by the estimators / updaters.
//
// These are the exact-fp32 kernels of the path: they serve every shape (any D, K, N) and are the
// GPU-side reference the tcgen05 kernels in tc_logdens.cu are validated against.
#include "common.cuh"
#include "../../include/gmmvi_b200.h"

namespace gvi {

__host__ inline bool ptr_vec_ok(const void* p, long long ld) {
  return (reinterpret_cast<uintptr_t>(p) % 16 == 0) && (ld % 4 == 0);
}

// ---- loader building blocks -------------------------------------------------------------------
// A tile (128 rows), source contiguous along k.  out[4*h+q] = S[(row0 + t/4 + 64h) * ld + k0 + (t%4)*4 + q]
__device__ __forceinline__ void fetchA_kcontig(const float* __restrict__ S, long long ld, int nrows, int kdim,
                                               int row0, int k0, bool vec, float (&r)[8]) {
  const int t = threadIdx.x, rr = t >> 2, kq = (t & 3) * 4;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int row = row0 + rr + 64 * h;
    const int valid = (row < nrows) ? (kdim - (k0 + kq)) : 0;
    const float4 v = load4(S + (long long)row * ld + k0 + kq, valid, vec);
    r[4 * h + 0] = v.x; r[4 * h + 1] = v.y; r[4 * h + 2] = v.z; r[4 * h + 3] = v.w;
  }
}
__device__ __forceinline__ void storeA_kcontig(float (*As)[AS_LD], const float (&r)[8]) {
  const int t = threadIdx.x, rr = t >> 2, kq = (t & 3) * 4;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int q = 0; q < 4; ++q) As[kq + q][rr + 64 * h] = r[4 * h + q];
}
// A tile, source contiguous along rows: out[4*h+q] = S[(k0 + t/16) * ld + row0 + (t%16)*4 + 64h + q]
__device__ __forceinline__ void fetchA_rcontig(const float* __restrict__ S, long long ld, int nrows, int kdim,
                                               int row0, int k0, bool vec, float (&r)[8]) {
  const int t = threadIdx.x, kk = t >> 4, rq = (t & 15) * 4;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int row = row0 + rq + 64 * h;
    const int valid = (k0 + kk < kdim) ? (nrows - row) : 0;
    const float4 v = load4(S + (long long)(k0 + kk) * ld + row, valid, vec);
    r[4 * h + 0] = v.x; r[4 * h + 1] = v.y; r[4 * h + 2] = v.z; r[4 * h + 3] = v.w;
  }
}
__device__ __forceinline__ void storeA_rcontig(float (*As)[AS_LD], const float (&r)[8]) {
  const int t = threadIdx.x, kk = t >> 4, rq = (t & 15) * 4;
#pragma unroll
  for (int h = 0; h < 2; ++h)
    *reinterpret_cast<float4*>(&As[kk][rq + 64 * h]) = make_float4(r[4 * h], r[4 * h + 1], r[4 * h + 2], r[4 * h + 3]);
}
// B tile (64 rows)
__device__ __forceinline__ void fetchB_kcontig(const float* __restrict__ S, long long ld, int nrows, int kdim,
                                               int row0, int k0, bool vec, float (&r)[4]) {
  const int t = threadIdx.x, rr = t >> 2, kq = (t & 3) * 4;
  const int row = row0 + rr;
  const int valid = (row < nrows) ? (kdim - (k0 + kq)) : 0;
  const float4 v = load4(S + (long long)row * ld + k0 + kq, valid, vec);
  r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
}
__device__ __forceinline__ void storeB_kcontig(float (*Bs)[BS_LD], const float (&r)[4]) {
  const int t = threadIdx.x, rr = t >> 2, kq = (t & 3) * 4;
#pragma unroll
  for (int q = 0; q < 4; ++q) Bs[kq + q][rr] = r[q];
}
__device__ __forceinline__ void fetchB_rcontig(const float* __restrict__ S, long long ld, int nrows, int kdim,
                                               int row0, int k0, bool vec, float (&r)[4]) {
  const int t = threadIdx.x, kk = t >> 4, rq = (t & 15) * 4;
  const int row = row0 + rq;
  const int valid = (k0 + kk < kdim) ? (nrows - row) : 0;
  const float4 v = load4(S + (long long)(k0 + kk) * ld + row, valid, vec);
  r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
}
__device__ __forceinline__ void storeB_rcontig(float (*Bs)[BS_LD], const float (&r)[4]) {
  const int t = threadIdx.x, kk = t >> 4, rq = (t & 15) * 4;
  *reinterpret_cast<float4*>(&Bs[kk][rq]) = make_float4(r[0], r[1], r[2], r[3]);
}

__device__ __forceinline__ void zero_acc(float (&acc)[TM][TN]) {
#pragma unroll
  for (int r = 0; r < TM; ++r)
#pragma unroll
    for (int c = 0; c < TN; ++c) acc[r][c] = 0.f;
}

// =================================================================================================
// Full-covariance component log densities (reference: models/full_cov_gmm.py:56-62)
//   grid = (ceil(N/128), K); CTA = 128 samples x one component; z = Linv_k (x - mu_k) is produced
//   64 output dims at a time, only over the non-zero (lower-triangular) part of Linv_k.
// =================================================================================================
__global__ void __launch_bounds__(NTHREADS)
logdens_full_kernel(const float* __restrict__ X, int N, int D, const float* __restrict__ means,
                    const float* __restrict__ linv, const float* __restrict__ cst, float* __restrict__ lq,
                    bool vecX, bool vecL) {
  __shared__ SmemTiles sm;
  extern __shared__ float mu_s[];
  const int k = blockIdx.y, n0 = blockIdx.x * BM;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  for (int d = threadIdx.x; d < D; d += NTHREADS) mu_s[d] = means[(long long)k * D + d];
  __syncthreads();
  const float* Lk = linv + (long long)k * D * D;
  const int kq = (threadIdx.x & 3) * 4;

  float sumsq[TM];
#pragma unroll
  for (int r = 0; r < TM; ++r) sumsq[r] = 0.f;

  for (int i0 = 0; i0 < D; i0 += BN) {
    float acc[TM][TN];
    zero_acc(acc);
    const int jend = min(D, i0 + BN);   // Linv[i][j] == 0 for j > i
    auto fA = [&](int c, float (&r)[8]) {
      fetchA_kcontig(X, D, N, D, n0, c * BK, vecX, r);
      const int j = c * BK + kq;
      const int rr = threadIdx.x >> 2;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const bool rowok = (n0 + rr + 64 * h) < N;
#pragma unroll
        for (int q = 0; q < 4; ++q) r[4 * h + q] = (rowok && j + q < D) ? r[4 * h + q] - mu_s[j + q] : 0.f;
      }
    };
    auto sA = [&](float (*As)[AS_LD], const float (&r)[8]) { storeA_kcontig(As, r); };
    auto fB = [&](int c, float (&r)[4]) { fetchB_kcontig(Lk, D, D, D, i0, c * BK, vecL, r); };
    auto sB = [&](float (*Bs)[BS_LD], const float (&r)[4]) { storeB_kcontig(Bs, r); };
    tile_mainloop(acc, sm, 0, ceil_div(jend, BK), ty, tx, fA, sA, fB, sB);
#pragma unroll
    for (int r = 0; r < TM; ++r)
#pragma unroll
      for (int c = 0; c < TN; ++c) sumsq[r] = fmaf(acc[r][c], acc[r][c], sumsq[r]);
  }
  // reduce over the 16 threads (tx) that share the same 8 samples
#pragma unroll
  for (int r = 0; r < TM; ++r) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) sumsq[r] += __shfl_xor_sync(0xffffffffu, sumsq[r], o);
  }
  if (tx == 0) {
    const float c = cst[k];
#pragma unroll
    for (int r = 0; r < TM; ++r) {
      const int n = n0 + ty * TM + r;
      if (n < N) lq[(long long)k * N + n] = c - 0.5f * sumsq[r];
    }
  }
}

// =================================================================================================
// Mixture gradient (analytic form of models/gmm.py:294-300):
//   grad[n, i] = - sum_k sum_j r_kn (x_nj - mu_kj) P_k[j, i];   grid = (ceil(N/128), ceil(D/64)).
//   Components whose responsibility is below e^-60 for the whole 128-sample block are skipped.
// =================================================================================================
// mask[b][k / 32] bit (k % 32): some sample of the 128-sample block b has responsibility > e^-60 for component k.
// One warp per 32 samples, warp ballots instead of block-wide barriers; lq is read exactly once, coalesced.
__global__ void __launch_bounds__(128)
resp_mask_kernel(const float* __restrict__ lq, const float* __restrict__ logw, const float* __restrict__ logq, int K,
                 int N, uint32_t* __restrict__ mask) {
  extern __shared__ uint32_t smask[];
  const int words = ceil_div(K, 32);
  for (int i = threadIdx.x; i < words; i += blockDim.x) smask[i] = 0u;
  __syncthreads();
  const int n = blockIdx.x * BM + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const float lqn = n < N ? logq[n] : 0.f;
  for (int k0 = 0; k0 < K; k0 += 8) {
    float a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = k0 + u;
      a[u] = (n < N && k < K) ? __ldg(lq + (long long)k * N + n) + __ldg(logw + k) - lqn : -INFINITY;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const unsigned any = __ballot_sync(0xffffffffu, a[u] > -60.f);
      if (any && lane == 0) atomicOr(&smask[(k0 + u) >> 5], 1u << ((k0 + u) & 31));
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < words; i += blockDim.x) mask[(long long)blockIdx.x * words + i] = smask[i];
}

__global__ void __launch_bounds__(NTHREADS)
mixture_grad_full_kernel(const float* __restrict__ X, int N, int D, const float* __restrict__ means,
                         const float* __restrict__ prec, const float* __restrict__ lq,
                         const float* __restrict__ logw, const float* __restrict__ logq, int K,
                         const uint32_t* __restrict__ mask, float* __restrict__ grad, bool vecX, bool vecP) {
  __shared__ SmemTiles sm;
  const int n0 = blockIdx.x * BM, i0 = blockIdx.y * BN;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int rr = threadIdx.x >> 2, kq = (threadIdx.x & 3) * 4;
  float acc[TM][TN];
  zero_acc(acc);
  float lqn[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int n = n0 + rr + 64 * h;
    lqn[h] = n < N ? logq[n] : 0.f;
  }
  const int words = ceil_div(K, 32);
  const uint32_t* bmask = mask + (long long)blockIdx.x * words;
  for (int wd = 0; wd < words; ++wd) {
   uint32_t bits = __ldg(bmask + wd);          // same word in every thread: the loop below is uniform
   while (bits) {
    const int k = wd * 32 + (__ffs(bits) - 1);
    bits &= bits - 1;
    float resp[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + rr + 64 * h;
      const float a = n < N ? (lq[(long long)k * N + n] + logw[k] - lqn[h]) : -INFINITY;
      resp[h] = a > -60.f ? expf(a) : 0.f;
    }
    const float* mu = means + (long long)k * D;
    const float* Pk = prec + (long long)k * D * D;
    auto fA = [&](int c, float (&r)[8]) {
      fetchA_kcontig(X, D, N, D, n0, c * BK, vecX, r);
      const int j = c * BK + kq;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const bool rowok = (n0 + rr + 64 * h) < N;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          r[4 * h + q] = (rowok && j + q < D) ? resp[h] * (r[4 * h + q] - __ldg(mu + j + q)) : 0.f;
      }
    };
    auto sA = [&](float (*As)[AS_LD], const float (&r)[8]) { storeA_kcontig(As, r); };
    auto fB = [&](int c, float (&r)[4]) { fetchB_rcontig(Pk, D, D, D, i0, c * BK, vecP, r); };
    auto sB = [&](float (*Bs)[BS_LD], const float (&r)[4]) { storeB_rcontig(Bs, r); };
    tile_mainloop(acc, sm, 0, ceil_div(D, BK), ty, tx, fA, sA, fB, sB);
   }
  }
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    const int n = n0 + ty * TM + r;
    if (n >= N) continue;
#pragma unroll
    for (int c = 0; c < TN; ++c) {
      const int i = i0 + tx * TN + c;
      if (i < D) grad[(long long)n * D + i] = -acc[r][c];
    }
  }
}

// =================================================================================================
// Stein statistics (ng_estimator.py:173-188):  M[k][j][i] = sum_n W[k,n] (x_nj - mu_kj) G[n,i]
//   grid = (ceil(D/128) * ceil(D/64), K); reduction over samples, 128-sample blocks with no weight skipped.
// =================================================================================================
__global__ void __launch_bounds__(NTHREADS)
stein_stats_full_kernel(const float* __restrict__ X, int N, int D, const float* __restrict__ means,
                        const float* __restrict__ W, const uint8_t* __restrict__ active,
                        const float* __restrict__ G, float* __restrict__ M, bool vecX, bool vecG) {
  __shared__ SmemTiles sm;
  const int k = blockIdx.y;
  const int nti = ceil_div(D, BN);
  const int j0 = (blockIdx.x / nti) * BM, i0 = (blockIdx.x % nti) * BN;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const int kk = threadIdx.x >> 4, rq = (threadIdx.x & 15) * 4;
  const float* mu = means + (long long)k * D;
  const float* Wk = W + (long long)k * N;
  const int nblk = ceil_div(N, 128);
  float muv[8];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int j = j0 + rq + 64 * h + q;
      muv[4 * h + q] = j < D ? mu[j] : 0.f;
    }
  float acc[TM][TN];
  zero_acc(acc);
  auto fA = [&](int c, float (&r)[8]) {
    fetchA_rcontig(X, D, D, N, j0, c * BK, vecX, r);   // rows = dims j, k-dim = samples
    const int n = c * BK + kk;
    const float w = n < N ? __ldg(Wk + n) : 0.f;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = j0 + rq + 64 * h + q;
        r[4 * h + q] = (n < N && j < D) ? w * (r[4 * h + q] - muv[4 * h + q]) : 0.f;
      }
  };
  auto sA = [&](float (*As)[AS_LD], const float (&r)[8]) { storeA_rcontig(As, r); };
  auto fB = [&](int c, float (&r)[4]) { fetchB_rcontig(G, D, D, N, i0, c * BK, vecG, r); };
  auto sB = [&](float (*Bs)[BS_LD], const float (&r)[4]) { storeB_rcontig(Bs, r); };
  for (int b = 0; b < nblk; ++b) {
    if (active != nullptr && active[(long long)k * nblk + b] == 0) continue;
    const int c0 = b * (128 / BK);
    const int c1 = min(ceil_div(N, BK), c0 + 128 / BK);
    tile_mainloop(acc, sm, c0, c1, ty, tx, fA, sA, fB, sB);
  }
  float* Mk = M + (long long)k * D * D;
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    const int j = j0 + ty * TM + r;
    if (j >= D) continue;
#pragma unroll
    for (int c = 0; c < TN; ++c) {
      const int i = i0 + tx * TN + c;
      if (i < D) Mk[(long long)j * D + i] = acc[r][c];
    }
  }
}

// =================================================================================================
// Sampling (models/gmm.py:361-386, full_cov_gmm.py:36-39):  x_n = mu_k + L_k eps_n for the rows of
// component k.  grid = (ceil(max_rows/128), ceil(D/64), K).
// =================================================================================================
__global__ void __launch_bounds__(NTHREADS)
sample_full_kernel(const float* __restrict__ eps, const int32_t* __restrict__ offsets,
                   const float* __restrict__ means, const float* __restrict__ chols, int D,
                   float* __restrict__ X, int32_t* __restrict__ mapping, bool vecE, bool vecL) {
  __shared__ SmemTiles sm;
  const int k = blockIdx.z;
  const int row_begin = offsets[k], row_end = offsets[k + 1];
  const int n0 = row_begin + blockIdx.x * BM;
  if (n0 >= row_end) return;
  const int i0 = blockIdx.y * BN;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const float* Lk = chols + (long long)k * D * D;
  float acc[TM][TN];
  zero_acc(acc);
  auto fA = [&](int c, float (&r)[8]) { fetchA_kcontig(eps, D, row_end, D, n0, c * BK, vecE, r); };
  auto sA = [&](float (*As)[AS_LD], const float (&r)[8]) { storeA_kcontig(As, r); };
  auto fB = [&](int c, float (&r)[4]) { fetchB_kcontig(Lk, D, D, D, i0, c * BK, vecL, r); };
  auto sB = [&](float (*Bs)[BS_LD], const float (&r)[4]) { storeB_kcontig(Bs, r); };
  tile_mainloop(acc, sm, 0, ceil_div(min(D, i0 + BN), BK), ty, tx, fA, sA, fB, sB);
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    const int n = n0 + ty * TM + r;
    if (n >= row_end) continue;
#pragma unroll
    for (int c = 0; c < TN; ++c) {
      const int i = i0 + tx * TN + c;
      if (i < D) X[(long long)n * D + i] = means[(long long)k * D + i] + acc[r][c];
    }
    if (blockIdx.y == 0 && tx == 0) mapping[n] = k;
  }
}

__global__ void sample_diag_kernel(const float* __restrict__ eps, const int32_t* __restrict__ offsets,
                                   const float* __restrict__ means, const float* __restrict__ stds, int K, int D,
                                   float* __restrict__ X, int32_t* __restrict__ mapping) {
  const int k = blockIdx.y;
  const int row_begin = offsets[k], row_end = offsets[k + 1];
  const long long total = (long long)(row_end - row_begin) * D;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int n = row_begin + (int)(e / D), d = (int)(e % D);
    X[(long long)n * D + d] = means[(long long)k * D + d] + stds[(long long)k * D + d] * eps[(long long)n * D + d];
    if (d == 0) mapping[n] = k;
  }
}

// =================================================================================================
// Batched GEMM  C[b] = alpha * opA(A[b]) opB(B[b])
// =================================================================================================
// C[b] = alpha * opA(A[b]) diag(scaleK[b]) opB(B[b]) + beta * C[b];  lower_only skips tiles strictly above the diagonal
// LONGK: reductions longer than 2048 are accumulated in 512-element segments that are folded into a second
// accumulator set (two-level summation), which keeps the fp32 rounding error of e.g. Phi^T W Phi over thousands of
// samples at the level of a blocked BLAS instead of growing linearly with the reduction length.
template <bool LONGK>
__global__ void __launch_bounds__(NTHREADS)
bgemm_kernel(int transA, int transB, int M, int Nn, int Kd, float alpha, const float* __restrict__ A, int lda,
             long long strideA, const float* __restrict__ B, int ldb, long long strideB, float* __restrict__ C,
             int ldc, long long strideC, bool vecA, bool vecB, const float* __restrict__ scaleK,
             long long strideScale, float beta, int lower_only) {
  __shared__ SmemTiles sm;
  const int b = blockIdx.y;
  const int ntn = ceil_div(Nn, BN);
  const int m0 = (blockIdx.x / ntn) * BM, n0 = (blockIdx.x % ntn) * BN;
  if (lower_only && n0 > m0 + BM - 1) return;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  const float* Ab = A + b * strideA;
  const float* Bb = B + b * strideB;
  const float* sc = scaleK ? scaleK + b * strideScale : nullptr;
  float acc[TM][TN];
  zero_acc(acc);
  auto fA = [&](int c, float (&r)[8]) {
    if (transA) {
      fetchA_rcontig(Ab, lda, M, Kd, m0, c * BK, vecA, r);
      if (sc) {
        const int kidx = c * BK + (threadIdx.x >> 4);
        const float w = kidx < Kd ? __ldg(sc + kidx) : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) r[q] *= w;
      }
    } else {
      fetchA_kcontig(Ab, lda, M, Kd, m0, c * BK, vecA, r);
      if (sc) {
        const int kq = c * BK + (threadIdx.x & 3) * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float w = kq + q < Kd ? __ldg(sc + kq + q) : 0.f;
          r[q] *= w;
          r[4 + q] *= w;
        }
      }
    }
  };
  auto sA = [&](float (*As)[AS_LD], const float (&r)[8]) {
    if (transA) storeA_rcontig(As, r); else storeA_kcontig(As, r);
  };
  auto fB = [&](int c, float (&r)[4]) {
    if (transB) fetchB_kcontig(Bb, ldb, Nn, Kd, n0, c * BK, vecB, r);
    else        fetchB_rcontig(Bb, ldb, Nn, Kd, n0, c * BK, vecB, r);
  };
  auto sB = [&](float (*Bs)[BS_LD], const float (&r)[4]) {
    if (transB) storeB_kcontig(Bs, r); else storeB_rcontig(Bs, r);
  };
  if constexpr (LONGK) {
    float tot[TM][TN];
    zero_acc(tot);
    const int nchunks = ceil_div(Kd, BK);
    for (int c0 = 0; c0 < nchunks; c0 += 32) {
      tile_mainloop(acc, sm, c0, min(nchunks, c0 + 32), ty, tx, fA, sA, fB, sB);
#pragma unroll
      for (int r = 0; r < TM; ++r)
#pragma unroll
        for (int c = 0; c < TN; ++c) {
          tot[r][c] += acc[r][c];
          acc[r][c] = 0.f;
        }
    }
#pragma unroll
    for (int r = 0; r < TM; ++r)
#pragma unroll
      for (int c = 0; c < TN; ++c) acc[r][c] = tot[r][c];
  } else {
    tile_mainloop(acc, sm, 0, ceil_div(Kd, BK), ty, tx, fA, sA, fB, sB);
  }
  float* Cb = C + b * strideC;
#pragma unroll
  for (int r = 0; r < TM; ++r) {
    const int m = m0 + ty * TM + r;
    if (m >= M) continue;
#pragma unroll
    for (int c = 0; c < TN; ++c) {
      const int n = n0 + tx * TN + c;
      if (n < Nn) {
        float* dst = Cb + (long long)m * ldc + n;
        *dst = beta != 0.f ? fmaf(beta, *dst, alpha * acc[r][c]) : alpha * acc[r][c];
      }
    }
  }
}

int launch_bgemm_ex(int transA, int transB, int batch, int M, int N, int Kd, float alpha, const float* A, int lda,
                    long long strideA, const float* B, int ldb, long long strideB, float* C, int ldc,
                    long long strideC, const float* scaleK, long long strideScale, float beta, int lower_only,
                    cudaStream_t st) {
  if (batch <= 0 || M <= 0 || N <= 0) return GVI_OK;
  const bool vecA = ptr_vec_ok(A, lda) && (strideA % 4 == 0);
  const bool vecB = ptr_vec_ok(B, ldb) && (strideB % 4 == 0);
  dim3 grid(ceil_div(M, BM) * ceil_div(N, BN), batch);
  if (Kd > 2048)
    bgemm_kernel<true><<<grid, NTHREADS, 0, st>>>(transA, transB, M, N, Kd, alpha, A, lda, strideA, B, ldb, strideB, C,
                                                  ldc, strideC, vecA, vecB, scaleK, strideScale, beta, lower_only);
  else
    bgemm_kernel<false><<<grid, NTHREADS, 0, st>>>(transA, transB, M, N, Kd, alpha, A, lda, strideA, B, ldb, strideB, C,
                                                   ldc, strideC, vecA, vecB, scaleK, strideScale, beta, lower_only);
  return check_launch("bgemm_kernel");
}

int launch_bgemm(int transA, int transB, int batch, int M, int N, int Kd, float alpha, const float* A, int lda,
                 long long strideA, const float* B, int ldb, long long strideB, float* C, int ldc,
                 long long strideC, cudaStream_t st) {
  return launch_bgemm_ex(transA, transB, batch, M, N, Kd, alpha, A, lda, strideA, B, ldb, strideB, C, ldc, strideC,
                         nullptr, 0, 0.f, 0, st);
}

int launch_logdens_full(const float* X, int N, int D, const float* means, const float* linv, const float* cst,
                        int K, float* lq, cudaStream_t st) {
  dim3 grid(ceil_div(N, BM), K);
  logdens_full_kernel<<<grid, NTHREADS, D * sizeof(float), st>>>(X, N, D, means, linv, cst, lq, ptr_vec_ok(X, D),
                                                                 ptr_vec_ok(linv, D));
  return check_launch("logdens_full_kernel");
}

int launch_stein_stats_full(const float* X, int N, int D, const float* means, const float* W,
                            const uint8_t* active, const float* G, int K, float* M, cudaStream_t st) {
  dim3 grid(ceil_div(D, BM) * ceil_div(D, BN), K);
  stein_stats_full_kernel<<<grid, NTHREADS, 0, st>>>(X, N, D, means, W, active, G, M, ptr_vec_ok(X, D),
                                                     ptr_vec_ok(G, D));
  return check_launch("stein_stats_full_kernel");
}

}  // namespace gvi

namespace gvi {
bool small_dim_supported(int D);
int launch_logdens_small(const float* X, int N, int D, const float* means, const float* linv, const float* cst, int K,
                         float* lq, cudaStream_t st);
int launch_mixgrad_small(const float* X, int N, int D, const float* means, const float* prec, const float* lq,
                         const float* logw, const float* logq, int K, float* grad, cudaStream_t st);
}  // namespace gvi

using namespace gvi;

extern "C" int gvi_logdens_full_f32(const float* X, int N, int D, const float* means, const float* linv,
                                    const float* cst, int K, float* lq, void* stream) {
  GVI_REQUIRE(N >= 0 && D > 0 && K >= 0, "gvi_logdens_full_f32: bad sizes N=%d D=%d K=%d", N, D, K);
  GVI_REQUIRE(K <= 65535, "gvi_logdens_full_f32: K=%d exceeds 65535", K);
  if (N == 0 || K == 0) return GVI_OK;
  GVI_REQUIRE(X && means && linv && cst && lq, "gvi_logdens_full_f32: null pointer");
  if (small_dim_supported(D)) return launch_logdens_small(X, N, D, means, linv, cst, K, lq, (cudaStream_t)stream);
  return launch_logdens_full(X, N, D, means, linv, cst, K, lq, (cudaStream_t)stream);
}

namespace gvi {
int launch_resp_mask(const float* lq, const float* logw, const float* logq, int K, int N, uint32_t* mask, cudaStream_t st) {
  resp_mask_kernel<<<ceil_div(N, BM), 128, ceil_div(K, 32) * sizeof(uint32_t), st>>>(lq, logw, logq, K, N, mask);
  return check_launch("resp_mask_kernel");
}
}  // namespace gvi

extern "C" size_t gvi_mixture_grad_full_workspace(int N, int K) {
  if (N <= 0 || K <= 0) return 0;
  return (size_t)ceil_div(N, BM) * ceil_div(K, 32) * sizeof(uint32_t);
}

extern "C" int gvi_mixture_grad_full_f32(const float* X, int N, int D, const float* means, const float* prec,
                                         const float* lq, const float* logw, const float* logq, int K, float* grad,
                                         void* ws, size_t ws_bytes, void* stream) {
  GVI_REQUIRE(N >= 0 && D > 0 && K >= 0, "gvi_mixture_grad_full_f32: bad sizes");
  if (N == 0) return GVI_OK;
  GVI_REQUIRE(X && means && prec && lq && logw && logq && grad && ws, "gvi_mixture_grad_full_f32: null pointer");
  if (ws_bytes < gvi_mixture_grad_full_workspace(N, K)) {
    set_last_error("gvi_mixture_grad_full_f32: workspace %zu < %zu", ws_bytes, gvi_mixture_grad_full_workspace(N, K));
    return GVI_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (small_dim_supported(D)) return launch_mixgrad_small(X, N, D, means, prec, lq, logw, logq, K, grad, st);
  uint32_t* mask = (uint32_t*)ws;
  resp_mask_kernel<<<ceil_div(N, BM), 128, ceil_div(K, 32) * sizeof(uint32_t), st>>>(lq, logw, logq, K, N, mask);
  int rc = check_launch("resp_mask_kernel");
  if (rc) return rc;
  dim3 grid(ceil_div(N, BM), ceil_div(D, BN));
  mixture_grad_full_kernel<<<grid, NTHREADS, 0, st>>>(X, N, D, means, prec, lq, logw, logq, K, mask, grad,
                                                      ptr_vec_ok(X, D), ptr_vec_ok(prec, D));
  return check_launch("mixture_grad_full_kernel");
}

extern "C" int gvi_sample_f32(int diagonal, const float* eps, const int32_t* offsets, const float* means,
                              const float* chols, int K, int D, int max_rows_per_component, float* X,
                              int32_t* mapping, void* stream) {
  GVI_REQUIRE(K >= 0 && D > 0 && max_rows_per_component >= 0, "gvi_sample_f32: bad sizes");
  if (K == 0 || max_rows_per_component == 0) return GVI_OK;
  GVI_REQUIRE(eps && offsets && means && chols && X && mapping, "gvi_sample_f32: null pointer");
  GVI_REQUIRE(K <= 65535, "gvi_sample_f32: K=%d exceeds 65535", K);
  cudaStream_t st = (cudaStream_t)stream;
  if (diagonal) {
    const long long per = (long long)max_rows_per_component * D;
    dim3 grid((unsigned)min((long long)4096, (per + 255) / 256), K);
    sample_diag_kernel<<<grid, 256, 0, st>>>(eps, offsets, means, chols, K, D, X, mapping);
    return check_launch("sample_diag_kernel");
  }
  dim3 grid(ceil_div(max_rows_per_component, BM), ceil_div(D, BN), K);
  sample_full_kernel<<<grid, NTHREADS, 0, st>>>(eps, offsets, means, chols, D, X, mapping, ptr_vec_ok(eps, D),
                                                ptr_vec_ok(chols, D));
  return check_launch("sample_full_kernel");
}

extern "C" int gvi_bgemm_f32(int transA, int transB, int batch, int M, int N, int Kd, float alpha, const float* A,
                             int lda, long long strideA, const float* B, int ldb, long long strideB, float* C,
                             int ldc, long long strideC, void* stream) {
  GVI_REQUIRE(batch >= 0 && M >= 0 && N >= 0 && Kd >= 0, "gvi_bgemm_f32: bad sizes");
  GVI_REQUIRE(batch <= 65535, "gvi_bgemm_f32: batch=%d exceeds 65535", batch);
  if (batch == 0 || M == 0 || N == 0) return GVI_OK;
  GVI_REQUIRE(A && B && C, "gvi_bgemm_f32: null pointer");
  return launch_bgemm(transA, transB, batch, M, N, Kd, alpha, A, lda, strideA, B, ldb, strideB, C, ldc, strideC,
                      (cudaStream_t)stream);
}
