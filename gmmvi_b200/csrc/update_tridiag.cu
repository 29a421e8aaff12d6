// Householder tridiagonalisation of the whitened natural-gradient matrix, one CTA per component (D <= 256).
//
// The KL-constrained update (ng_based_component_updater.py:244-524) bisects on eta with
//     KL(eta) = 1/2 [ log det M - D + tr M^-1 + |M^-1 h|^2 / eta^2 ],   M = I + B / eta,   B = L^T R L,  h = L^T g
// (DESIGN.md section 5).  Evaluating that through a Cholesky factorisation and a triangular inverse of M costs
// 2/3 D^3 flops per eta (245 k cycles of one SM at D = 256, 5.5 evaluations per component at C5).  All of those
// terms are invariant under an orthogonal change of basis, so this kernel reduces B ONCE to tridiagonal form
//     T = P^T B P,   h' = P^T h          (P = H_0 H_1 ... H_{D-3}, Householder reflectors)
// and the bisection in update_blocked.cu evaluates KL(eta) from (T, h') with O(D) recurrences (pivots of the LDL^T
// factorisation from both ends) -- every candidate eta of five bisection levels at once, one per lane.  Only the
// final eta is factored the expensive way, because the new Cholesky factor L' = L U^-T needs the real thing.
// STATUS: opt-in (GMMVI_B200_UPDATE_TRIDIAG=1).  The search itself is 2.7 x faster this way (update kernel 3.99 -> 1.48 ms
// at C5), but this reduction kernel takes 6.4 ms: 254 dependent Householder steps of two passes over a packed triangle,
// ~20 % useful FMA lanes in the short rows and ~9 block barriers per step.  It has to come down to ~2 ms to pay off.
//
// The symmetric matrix is the one the update kernel builds: S[i][j] = S[j][i] = B[min(i,j)][max(i,j)] (assemble()
// reads B[D-1-a][D-1-b] for b <= a, i.e. the upper triangle).  It is held as a packed lower triangle in shared memory
// in a layout whose addresses are closed-form and conflict free for row-wise and column-wise access (rfp() below).
#include "common.cuh"

namespace gvi {
namespace td {

constexpr int THREADS = 512;
constexpr int NWARPS = THREADS / 32;
constexpr int MAXD = 256;

// Packed storage of the lower triangle ("rectangular full packed" for order 256): line l of 257 floats holds the long
// row 128 + l in positions 0 .. 128 + l and the short row 127 - l, reversed, in positions 256 .. 129 + l.  The address
// of (i, c), c <= i, is a compare and a multiply-add (no table), consecutive columns of a row are consecutive (or
// reverse-consecutive) words, and consecutive rows at a fixed column are 257 = 1 (mod 32) words apart: "lane = column"
// and "lane = row" accesses are both free of bank conflicts.
constexpr int HALF = MAXD / 2;
constexpr int PITCH = MAXD + 1;
constexpr int A_FLOATS = HALF * PITCH;
__device__ __forceinline__ int rfp(int i, int c) {
  return i >= HALF ? (i - HALF) * PITCH + c : (HALF - 1 - i) * PITCH + (MAXD - c);
}

// sums of two values over the block; every thread gets both results.  red holds >= 64 floats.
__device__ __forceinline__ void block_sum2(float& a, float& b, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  __syncthreads();
  if (lane == 0) {
    red[w] = a;
    red[32 + w] = b;
  }
  __syncthreads();
  float x = lane < NWARPS ? red[lane] : 0.f, y = lane < NWARPS ? red[32 + lane] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    x += __shfl_xor_sync(0xffffffffu, x, o);
    y += __shfl_xor_sync(0xffffffffu, y, o);
  }
  a = x;
  b = y;
}

__global__ void __launch_bounds__(THREADS, 1)
tridiag_kernel(const float* __restrict__ Bmat, const float* __restrict__ hvec, int D, float* __restrict__ dout,
               float* __restrict__ eout, float* __restrict__ hout) {
  extern __shared__ __align__(16) float td_smem[];
  __shared__ float red[64];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int comp = blockIdx.x;
  const float* B = Bmat + (long long)comp * D * D;
  float* A = td_smem;
  float* v = A + A_FLOATS;
  float* prow = v + MAXD;                          // row parts of S22 v
  float* w = prow + MAXD;
  float* h = w + MAXD;
  float* pcol = h + MAXD;                          // [NWARPS][MAXD] column parts of S22 v, one slice per warp
  // S[c][r] = B[r][c] for r <= c (coalesced along c)
  for (int e = tid; e < D * D; e += THREADS) {
    const int r = e / D, c = e - r * D;
    if (c >= r) A[rfp(c, r)] = __ldg(B + e);
  }
  for (int i = tid; i < D; i += THREADS) h[i] = hvec[(long long)comp * D + i];
  __syncthreads();

  float* dk = dout + (long long)comp * D;
  float* ek = eout + (long long)comp * D;
  for (int k = 0; k + 2 < D; ++k) {
    const int m = D - k - 1;                        // order of the trailing matrix, rows / columns k+1 .. D-1
    const int g = k + 1 + tid;                      // my row / column of it (tid < m)
    float xv = 0.f;
    if (tid < m) {
      xv = A[rfp(g, k)];
      v[g] = xv;
    }
    float xn2 = (tid >= 1 && tid < m) ? xv * xv : 0.f, dummy = 0.f;
    block_sum2(xn2, dummy, red);                    // (its barriers also publish v)
    const float alpha = v[k + 1];
    if (tid == 0) dk[k] = A[rfp(k, k)];
    if (xn2 == 0.f) {                               // column already tridiagonal: H = I
      if (tid == 0) ek[k] = alpha;
      __syncthreads();
      continue;
    }
    const float beta = -copysignf(sqrtf(fmaf(alpha, alpha, xn2)), alpha);
    const float tau = (beta - alpha) / beta;
    const float scale = 1.f / (alpha - beta);
    if (tid == 0) ek[k] = beta;
    __syncthreads();                                // everyone has read alpha = v[k+1]
    if (tid < m) v[g] = tid == 0 ? 1.f : xv * scale;
    __syncthreads();
    // ---- p = S22 v in one pass over the packed triangle: warp per row r, lane per column c = cb + lane + 32 j.
    // Element (r, c) adds a v[c] to the row sum of r and, off the diagonal, a v[r] to the column sum of c, which the
    // lane keeps in a register (colacc[j]) until all rows of the warp are done. ----
    const int cb = (k + 1) & ~31;
    float colacc[MAXD / 32];
#pragma unroll
    for (int j = 0; j < MAXD / 32; ++j) colacc[j] = 0.f;
    for (int r = k + 1 + warp; r < D; r += NWARPS) {
      const float vr = v[r];
      const int base = r >= HALF ? (r - HALF) * PITCH : (HALF - 1 - r) * PITCH + MAXD;
      const int sgn = r >= HALF ? 1 : -1;
      float rowacc = 0.f;
#pragma unroll
      for (int j = 0; j < MAXD / 32; ++j) {
        const int c = cb + 32 * j + lane;
        if (cb + 32 * j <= r) {                     // warp-uniform
          if (c > k && c <= r) {
            const float a = A[base + sgn * c];
            rowacc = fmaf(a, v[c], rowacc);
            if (c < r) colacc[j] = fmaf(a, vr, colacc[j]);
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rowacc += __shfl_xor_sync(0xffffffffu, rowacc, o);
      if (lane == 0) prow[r] = rowacc;
    }
#pragma unroll
    for (int j = 0; j < MAXD / 32; ++j) pcol[warp * MAXD + ((cb + 32 * j + lane) & (MAXD - 1))] = colacc[j];
    __syncthreads();
    float pg = 0.f, vg = 0.f, s1 = 0.f, s2 = 0.f;
    if (tid < m) {
      vg = v[g];
      float acc = prow[g];
#pragma unroll
      for (int ww = 0; ww < NWARPS; ++ww) acc += pcol[ww * MAXD + g];
      pg = tau * acc;
      s1 = pg * vg;
      s2 = vg * h[g];
    }
    block_sum2(s1, s2, red);
    if (tid < m) {
      w[g] = fmaf(-0.5f * tau * s1, vg, pg);
      h[g] = fmaf(-tau * s2, vg, h[g]);
    }
    __syncthreads();
    // ---- S22 -= v w^T + w v^T on the packed lower triangle: one warp per row ----
    for (int r = k + 1 + warp; r < D; r += NWARPS) {
      const float vr = v[r], wr = w[r];
      const int base = r >= HALF ? (r - HALF) * PITCH : (HALF - 1 - r) * PITCH + MAXD;
      const int sgn = r >= HALF ? 1 : -1;
      for (int c = k + 1 + lane; c <= r; c += 32) {
        float* a = A + base + sgn * c;
        *a -= fmaf(vr, w[c], wr * v[c]);
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    if (D >= 2) {
      dk[D - 2] = A[rfp(D - 2, D - 2)];
      ek[D - 2] = A[rfp(D - 1, D - 2)];
    }
    dk[D - 1] = A[rfp(D - 1, D - 1)];
    ek[D - 1] = 0.f;
  }
  for (int i = tid; i < D; i += THREADS) hout[(long long)comp * D + i] = h[i];
}

size_t tridiag_smem_bytes(int) { return ((size_t)A_FLOATS + (4 + NWARPS) * MAXD) * sizeof(float); }

}  // namespace td

bool tridiag_supported(int D) { return D >= 1 && D <= td::MAXD; }

// d[K, D], e[K, D] (e[k][D-1] = 0), hp[K, D] = P^T h
int launch_tridiag(const float* Bm, const float* hv, int K, int D, float* d, float* e, float* hp, cudaStream_t st) {
  if (K <= 0) return GVI_OK;
  static unsigned long long attr_set_mask = 0;
  if (first_call_on_device(attr_set_mask)) {
    cudaError_t err = cudaFuncSetAttribute(td::tridiag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)td::tridiag_smem_bytes(td::MAXD));
    if (err != cudaSuccess) {
      set_last_error("tridiag: cudaFuncSetAttribute: %s", cudaGetErrorString(err));
      return GVI_ERR_CUDA;
    }
  }
  td::tridiag_kernel<<<K, td::THREADS, td::tridiag_smem_bytes(D), st>>>(Bm, hv, D, d, e, hp);
  return check_launch("tridiag_kernel");
}

}  // namespace gvi

using namespace gvi;

// Stand-alone entry point (used by the tests): T = P^T S P and h' = P^T h for S[i][j] = B[min(i,j)][max(i,j)].
extern "C" int gvi_tridiag_f32(const float* B, const float* h, int K, int D, float* d, float* e, float* hp,
                               void* stream) {
  GVI_REQUIRE(K >= 0 && D >= 1, "gvi_tridiag_f32: bad sizes");
  if (K == 0) return GVI_OK;
  GVI_REQUIRE(B && h && d && e && hp, "gvi_tridiag_f32: null pointer");
  if (!tridiag_supported(D)) {
    set_last_error("gvi_tridiag_f32: D=%d unsupported (D <= 256)", D);
    return GVI_ERR_UNSUPPORTED;
  }
  return launch_tridiag(B, h, K, D, d, e, hp, (cudaStream_t)stream);
}
