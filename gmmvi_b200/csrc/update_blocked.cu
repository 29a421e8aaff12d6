// Full-covariance component update, blocked version (D <= 256): one CTA per component with the whitened matrix
// M~ (index-reversed I + a1 B + a2 B^2, see update.cu / DESIGN.md section 5) held in shared memory as 32 x 32
// blocks of its lower triangle.  Each KL evaluation of the eta bisection factors M~ = C C^T and inverts C in place;
// both are organised as block algorithms whose bulk is register-tiled 32 x 32 x 32 products:
//   * right-looking Cholesky: diagonal block by one warp (rows in registers, row broadcast by shuffles), panel
//     solve with one thread per row, trailing update as warp tasks C_ij -= P_i P_j^T with 4 x 8 register tiles fed
//     by 128-bit loads from a transposed copy of the panel (3 shared-memory wavefronts per 32 FMAs, against 5 loads
//     per 4 FMAs in the column-panel kernel of update.cu that this replaces);
//   * inverse: diagonal blocks in parallel, then column blocks from the right, X_ip = -(sum_k X_ik C_kp) X_pp with
//     one thread per row holding its row of the left factor in registers and the right factor broadcast.
// Blocks are stored row major with the 16-byte column chunks XOR-swizzled by (row & 7), which makes both
// "lane = row" 128-bit accesses and "lane = column" accesses bank-conflict free.
// The diagonal of M~ is carried as (diag - 1) (dm1) exactly as in update.cu, so log det and tr(M^-1) - D keep
// their relative accuracy for large eta.  The bisection itself (ng_based_component_updater.py:335-429) is unchanged.
#include "common.cuh"

namespace gvi {
namespace ub {

constexpr int NB = 32;
constexpr int BS = NB * NB;          // floats per block
constexpr int THREADS = 512;
constexpr int NWARPS = THREADS / 32;
constexpr int MAXBLK = 8;            // D <= 256

__host__ __device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }
__device__ __forceinline__ int sw(int r, int c) { return r * NB + ((((c >> 2) ^ (r & 7)) << 2) | (c & 3)); }
__device__ __forceinline__ int sw4(int r, int q) { return r * NB + ((q ^ (r & 7)) << 2); }     // float offset of chunk q
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float comp(const float4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

struct Smem {
  float* A;       // blocks of the lower triangle, block (bi, bj) at (tri(bi) + bj) * BS
  float* PT;      // (MAXBLK - 1) blocks of scratch: transposed panel (Cholesky) / T rows (inverse)
  float* hrev;    // [Dp]
  float* v;       // [Dp]
  float* u;       // [Dp]
  float* dm1;     // [Dp] diagonal minus one / pivot deltas
  float* idiag;   // [32] reciprocal diagonal of the current diagonal block
  float* red;     // [33]
  int* flag;
  const int* blk_bi;  // [36] row block of packed block index
};

// ---- assembly: A~[a][b] = a1 B[D-1-a][D-1-b] (+ a2 B2[..]) for b < a, dm1[a] = the diagonal value ----------------
// One 16-byte chunk (row a, columns 4q..4q+3 of a block) per thread and step; the four source elements are
// contiguous (in reverse order) in row D-1-a of B.
__device__ __noinline__ void assemble(const Smem s, const float* __restrict__ B, const float* __restrict__ B2, float a1, float a2,
                         int D, int nbk) {
  const int tid = threadIdx.x;
  const int nchunk = tri(nbk) * (BS / 4);
  const bool vec = (D & 3) == 0;
  for (int e = tid; e < nchunk; e += THREADS) {
    const int blk = e >> 8, r = (e >> 3) & 31, q = e & 7;
    const int bi = s.blk_bi[blk], bj = blk - tri(bi);
    const int a = bi * NB + r, b0 = bj * NB + 4 * q;
    float vals[4] = {0.f, 0.f, 0.f, 0.f};
    if (a < D && b0 <= a) {
      const long long g = (long long)(D - 1 - a) * D + (D - 1 - b0 - 3);     // element for b0 + 3; b0 is at g + 3
      if (vec) {      // D % 4 == 0 and b0 % 4 == 0: aligned, and b0 + 3 < D because a < D
        const float4 t = __ldg(reinterpret_cast<const float4*>(B + g));
        vals[0] = a1 * t.w; vals[1] = a1 * t.z; vals[2] = a1 * t.y; vals[3] = a1 * t.x;
        if (B2) {
          const float4 t2 = __ldg(reinterpret_cast<const float4*>(B2 + g));
          vals[0] = fmaf(a2, t2.w, vals[0]); vals[1] = fmaf(a2, t2.z, vals[1]);
          vals[2] = fmaf(a2, t2.y, vals[2]); vals[3] = fmaf(a2, t2.x, vals[3]);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (b0 + c < D) {
            vals[c] = a1 * __ldg(B + g + 3 - c);
            if (B2) vals[c] = fmaf(a2, __ldg(B2 + g + 3 - c), vals[c]);
          }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int b = b0 + c;
      if (b == a) {
        s.dm1[a] = vals[c];
        vals[c] = 1.f;
      } else if (b > a) {
        vals[c] = 0.f;
      }
    }
    st4(s.A + blk * BS + sw4(r, q), make_float4(vals[0], vals[1], vals[2], vals[3]));
  }
}

// ---- Cholesky of one 32 x 32 diagonal block by one warp (lane = row) ------------------------------------------
// In: lower part of the block + dm1 (diagonal - 1).  Out: factor (diagonal entries = l_jj), dm1 = pivot - 1,
// idiag = 1 / l_jj.  Returns false on a non-positive / non-finite pivot (same value in every lane).
__device__ __noinline__ bool chol_diag(float* blk, float* dm1p, float* idiag) {
  const int i = threadIdx.x & 31;
  float a[NB];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 t = ld4(blk + sw4(i, q));
    a[4 * q] = t.x; a[4 * q + 1] = t.y; a[4 * q + 2] = t.z; a[4 * q + 3] = t.w;
  }
  float dd = dm1p[i];
  float myinv = 1.f;
  bool good = true;
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    // column j: l_ij = (a_ij - sum_{m<j} l_im l_jm) / l_jj.  Row j (m < j) is already final in shared memory
    // (lane j stored its entries as they were produced) and is read as broadcast 128-bit loads.
    float s0 = a[j], s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int q = 0; q < (j + 3) / 4; ++q) {
      const float4 cv = ld4(blk + sw4(j, q));
      if (4 * q < j) s0 = fmaf(-a[4 * q], cv.x, s0);
      if (4 * q + 1 < j) s1 = fmaf(-a[4 * q + 1], cv.y, s1);
      if (4 * q + 2 < j) s2 = fmaf(-a[4 * q + 2], cv.z, s2);
      if (4 * q + 3 < j) s3 = fmaf(-a[4 * q + 3], cv.w, s3);
    }
    const float ddj = __shfl_sync(0xffffffffu, dd, j);
    const float piv = 1.f + ddj;
    if (!(piv > 0.f) || !isfinite(piv)) good = false;
    const float inv = rsqrtf(piv);
    const float lij = ((s0 + s1) + (s2 + s3)) * inv;
    float val;
    if (i > j) {
      val = lij;
      dd = fmaf(-lij, lij, dd);
    } else if (i == j) {
      val = piv * inv;
      myinv = inv;
    } else {
      val = 0.f;
    }
    a[j] = val;
    blk[sw(i, j)] = val;
    __syncwarp();
  }
  dm1p[i] = dd;
  idiag[i] = myinv;
  __syncwarp();
  return good;
}

// ---- panel solve: row g of the panel below diagonal block p, x C_pp^T = a (one thread per row) -----------------
__device__ __noinline__ void panel_solve(const Smem s, int p, int n) {
  const int tid = threadIdx.x;
  const float* Cpp = s.A + (tri(p) + p) * BS;
  for (int g = tid; g < n * NB; g += THREADS) {
    const int ib = g >> 5, r = g & 31;
    float* blk = s.A + (tri(p + 1 + ib) + p) * BS;
    float x[NB];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 t = ld4(blk + sw4(r, q));
      x[4 * q] = t.x; x[4 * q + 1] = t.y; x[4 * q + 2] = t.z; x[4 * q + 3] = t.w;
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      float acc0 = x[j], acc1 = 0.f;
#pragma unroll
      for (int q = 0; q < (j + 3) / 4; ++q) {
        const float4 cv = ld4(Cpp + sw4(j, q));       // C_pp[j][4q .. 4q+3], broadcast
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (4 * q + e < j) {
            if (e & 1) acc1 = fmaf(-x[4 * q + e], comp(cv, e), acc1);
            else acc0 = fmaf(-x[4 * q + e], comp(cv, e), acc0);
          }
      }
      x[j] = (acc0 + acc1) * s.idiag[j];
    }
    float* pt = s.PT + ib * BS;
#pragma unroll
    for (int j = 0; j < NB; ++j) pt[j * NB + r] = x[j];
#pragma unroll
    for (int q = 0; q < 8; ++q) st4(blk + sw4(r, q), make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]));
  }
}

// ---- one trailing-update task: C(p+1+ii, p+1+jj) -= P_ii P_jj^T, 4 x 8 register tile per lane ---------------
__device__ __noinline__ void trailing_task(const Smem s, int p, int ii, int jj) {
  const int lane = threadIdx.x & 31;
  const int ry = lane >> 2, cx = lane & 3;
  const float* Pi = s.PT + ii * BS + 4 * ry;
  const float* Pj = s.PT + jj * BS + 8 * cx;
  float acc[4][8];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
#pragma unroll 8
  for (int k = 0; k < NB; ++k) {
    const float4 av = ld4(Pi + k * NB);
    const float4 b0 = ld4(Pj + k * NB), b1 = ld4(Pj + k * NB + 4);
    const float ar[4] = {av.x, av.y, av.z, av.w};
    const float bc[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(ar[r], bc[c], acc[r][c]);
  }
  const int bi = p + 1 + ii, bj = p + 1 + jj;
  float* blk = s.A + (tri(bi) + bj) * BS;
  const bool diag = (bi == bj);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int row = 4 * ry + r;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float* ptr = blk + sw4(row, 2 * cx + h);
      float4 cur = ld4(ptr);
      float cv[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = 8 * cx + 4 * h + e;
        if (!diag || col < row) cv[e] -= acc[r][4 * h + e];
        else if (col == row) s.dm1[bi * NB + row] -= acc[r][4 * h + e];
      }
      st4(ptr, make_float4(cv[0], cv[1], cv[2], cv[3]));
    }
  }
}

// ---- blocked Cholesky of the whole matrix ----------------------------------------------------------------------
__device__ __noinline__ bool chol_blocked(const Smem s, int nbk) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    const bool good = chol_diag(s.A, s.dm1, s.idiag);
    if (lane == 0) *s.flag = good ? 0 : 1;
  }
  __syncthreads();
  for (int p = 0; p < nbk; ++p) {
    // the diagonal block (p, p) has been factored (by warp 0, overlapped with the previous trailing update)
    if (*s.flag) return false;
    const int n = nbk - 1 - p;                 // row blocks below the diagonal block
    if (n == 0) break;
    panel_solve(s, p, n);
    __syncthreads();
    // ---- trailing update: C(bi, bj) -= P_bi P_bj^T, one warp per block, 4 x 8 register tile per lane.
    // Warp 0 takes the next diagonal block (task 0) and factors it right away while the other warps work
    // through the remaining tasks.
    const int ntask = tri(n);
    for (int t = (warp == 0 ? 0 : warp); t < ntask; t += (warp == 0 ? ntask : NWARPS - 1)) {
      int ii = 0;
      while (tri(ii + 1) <= t) ++ii;
      trailing_task(s, p, ii, t - tri(ii));
    }
    if (warp == 0) {
      __syncwarp();
      const bool good = chol_diag(s.A + (tri(p + 1) + p + 1) * BS, s.dm1 + (p + 1) * NB, s.idiag);
      if (lane == 0) *s.flag = good ? 0 : 1;
    }
    __syncthreads();
  }
  return true;
}

// ---- in-place inverse of the blocked lower-triangular factor --------------------------------------------------
// Returns this thread's share of sum_{i>j} X_ij^2 (the strictly lower part of the inverse).
__device__ __noinline__ float inv_blocked(const Smem s, int nbk) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float sq = 0.f;
  // diagonal blocks, one warp each: lane c solves for column c of the inverse
  for (int b = warp; b < nbk; b += NWARPS) {
    float* blk = s.A + (tri(b) + b) * BS;
    float y[NB];
#pragma unroll
    for (int r = 0; r < NB; ++r) {
      float acc = (r == lane) ? 1.f : 0.f;
      float4 dv = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q <= r / 4; ++q) {
        const float4 cv = ld4(blk + sw4(r, q));
        if (q == r / 4) dv = cv;
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (4 * q + e < r) acc = fmaf(-comp(cv, e), y[4 * q + e], acc);
      }
      y[r] = acc / comp(dv, r & 3);
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < NB; ++r) {
      blk[sw(r, lane)] = y[r];
      if (r != lane) sq = fmaf(y[r], y[r], sq);      // y[r] = 0 above the diagonal
    }
  }
  __syncthreads();
  // column blocks from the right: X_ip = -(sum_{k=p+1..i} X_ik C_kp) X_pp
  for (int p = nbk - 2; p >= 0; --p) {
    const int n = nbk - 1 - p;
    const int ntask = n * 4;                         // (row block, quarter of the 32 output columns)
    // phase 1: T_i = sum_k X_ik C_kp  -> PT (row major, swizzled)
    for (int t = warp; t < ntask; t += NWARPS) {
      const int ii = t >> 2, cq = t & 3;
      const int bi = p + 1 + ii;
      float acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = 0.f;
      for (int bk = p + 1; bk <= bi; ++bk) {
        const float* Xb = s.A + (tri(bi) + bk) * BS;
        const float* Cb = s.A + (tri(bk) + p) * BS;
        float x[NB];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 tv = ld4(Xb + sw4(lane, q));
          x[4 * q] = tv.x; x[4 * q + 1] = tv.y; x[4 * q + 2] = tv.z; x[4 * q + 3] = tv.w;
        }
#pragma unroll
        for (int k = 0; k < NB; ++k) {
          const float4 c0 = ld4(Cb + sw4(k, 2 * cq)), c1 = ld4(Cb + sw4(k, 2 * cq + 1));
          acc[0] = fmaf(x[k], c0.x, acc[0]); acc[1] = fmaf(x[k], c0.y, acc[1]);
          acc[2] = fmaf(x[k], c0.z, acc[2]); acc[3] = fmaf(x[k], c0.w, acc[3]);
          acc[4] = fmaf(x[k], c1.x, acc[4]); acc[5] = fmaf(x[k], c1.y, acc[5]);
          acc[6] = fmaf(x[k], c1.z, acc[6]); acc[7] = fmaf(x[k], c1.w, acc[7]);
        }
      }
      float* T = s.PT + ii * BS;
      st4(T + sw4(lane, 2 * cq), make_float4(acc[0], acc[1], acc[2], acc[3]));
      st4(T + sw4(lane, 2 * cq + 1), make_float4(acc[4], acc[5], acc[6], acc[7]));
    }
    __syncthreads();
    // phase 2: X_ip = -T_i X_pp
    const float* Xpp = s.A + (tri(p) + p) * BS;
    for (int t = warp; t < ntask; t += NWARPS) {
      const int ii = t >> 2, cq = t & 3;
      const int bi = p + 1 + ii;
      const float* T = s.PT + ii * BS;
      float x[NB];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 tv = ld4(T + sw4(lane, q));
        x[4 * q] = tv.x; x[4 * q + 1] = tv.y; x[4 * q + 2] = tv.z; x[4 * q + 3] = tv.w;
      }
      float acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = 0.f;
#pragma unroll
      for (int k = 0; k < NB; ++k) {
        const float4 c0 = ld4(Xpp + sw4(k, 2 * cq)), c1 = ld4(Xpp + sw4(k, 2 * cq + 1));
        acc[0] = fmaf(x[k], c0.x, acc[0]); acc[1] = fmaf(x[k], c0.y, acc[1]);
        acc[2] = fmaf(x[k], c0.z, acc[2]); acc[3] = fmaf(x[k], c0.w, acc[3]);
        acc[4] = fmaf(x[k], c1.x, acc[4]); acc[5] = fmaf(x[k], c1.y, acc[5]);
        acc[6] = fmaf(x[k], c1.z, acc[6]); acc[7] = fmaf(x[k], c1.w, acc[7]);
      }
      float* Xo = s.A + (tri(bi) + p) * BS;
      st4(Xo + sw4(lane, 2 * cq), make_float4(-acc[0], -acc[1], -acc[2], -acc[3]));
      st4(Xo + sw4(lane, 2 * cq + 1), make_float4(-acc[4], -acc[5], -acc[6], -acc[7]));
#pragma unroll
      for (int c = 0; c < 8; ++c) sq = fmaf(acc[c], acc[c], sq);
    }
    __syncthreads();
  }
  return sq;
}

struct KlTerms {
  float kl;
  bool ok;
};

// Assemble M~, factor, invert, evaluate the KL terms.  On return (ok): A holds X = chol(M~)^-1 and u = M~^-1 h~.
__device__ __noinline__ KlTerms eval_whitened(const Smem s, const float* __restrict__ B, const float* __restrict__ B2, float a1,
                                 float a2, int D, int nbk, float inv_eta) {
  const int tid = threadIdx.x;
  const int Dp = nbk * NB;
  assemble(s, B, B2, a1, a2, D, nbk);
  __syncthreads();
  KlTerms out;
  out.ok = chol_blocked(s, nbk);
  if (!out.ok) {
    out.kl = FLT_MAX;
    return out;
  }
  float ld = 0.f;
  for (int j = tid; j < Dp; j += THREADS) ld += log1pf(s.dm1[j]);
  ld = block_sum(ld, s.red);
  // tr(M^-1) - D = sum_j (-delta_j / (1 + delta_j)) + sum_{i>j} X_ij^2
  float tr = inv_blocked(s, nbk);
  for (int j = tid; j < Dp; j += THREADS) tr -= s.dm1[j] / (1.f + s.dm1[j]);
  tr = block_sum(tr, s.red);
  // v = X h~ (thread per row), u = X^T v (thread per column)
  for (int g = tid; g < Dp; g += THREADS) {
    const int bi = g >> 5, r = g & 31;
    float acc = 0.f;
    for (int bj = 0; bj <= bi; ++bj) {
      const float* blk = s.A + (tri(bi) + bj) * BS;
      const float* hh = s.hrev + bj * NB;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 xv = ld4(blk + sw4(r, q));
        acc = fmaf(xv.x, hh[4 * q], acc); acc = fmaf(xv.y, hh[4 * q + 1], acc);
        acc = fmaf(xv.z, hh[4 * q + 2], acc); acc = fmaf(xv.w, hh[4 * q + 3], acc);
      }
    }
    s.v[g] = acc;
  }
  __syncthreads();
  float mh = 0.f;
  for (int m = tid; m < Dp; m += THREADS) {
    const int bj = m >> 5, c = m & 31;
    float acc = 0.f;
    for (int bi = bj; bi < nbk; ++bi) {
      const float* blk = s.A + (tri(bi) + bj) * BS;
      const float* vv = s.v + bi * NB;
#pragma unroll 8
      for (int r = 0; r < NB; ++r) acc = fmaf(blk[sw(r, c)], vv[r], acc);
    }
    s.u[m] = acc;
    mh = fmaf(acc, acc, mh);
  }
  mh = block_sum(mh, s.red);
  out.kl = 0.5f * (ld + tr + mh * inv_eta * inv_eta);
  return out;
}

// KL(new || old) for M = I + a1 T with T tridiagonal (diagonal d, sub-diagonal e) and h' = hp, evaluated by ONE
// thread (update_tridiag.cu explains the change of basis).  With the pivots of the LDL^T factorisation from the top,
// q_i = 1 + r_i, and from the bottom, q'_i = 1 + r'_i (both carried "minus one", like dm1 in the Cholesky path):
//   log det M = sum log1p(r_i),   (M^-1)_ii = 1 / (q_i + q'_i - m_ii)  =>  tr M^-1 - D = sum -s_i / (1 + s_i) with
//   s_i = r_i + r'_i - a1 d_i,    M x = h' by the same pivots (Thomas algorithm),   KL = 1/2 (logdet + tr + |x|^2 a1^2).
// rq / ys: per-lane scratch, element i of lane l at [32 i + l].  Not positive definite (a pivot <= 0) -> FLT_MAX,
// the value kl() returns for a failed Cholesky (:320-324).
__device__ __noinline__ float kl_tridiag_lane(float a1, const float* __restrict__ d, const float* __restrict__ e,
                                              const float* __restrict__ hp, int D, float* __restrict__ rq,
                                              float* __restrict__ ys) {
  const int lane = threadIdx.x & 31;
  float r_prev = 0.f, y_prev = 0.f, ld = 0.f;
  bool pd = true;
  for (int i = 0; i < D; ++i) {
    const float delta = a1 * d[i];
    float r = delta, y = hp[i];
    if (i > 0) {
      const float b = a1 * e[i - 1];
      const float l = b / (1.f + r_prev);
      r = fmaf(-b, l, delta);
      y = fmaf(-l, y_prev, y);
    }
    pd = pd && (1.f + r > 0.f);
    ld += log1pf(r);
    rq[32 * i + lane] = r;
    ys[32 * i + lane] = y;
    r_prev = r;
    y_prev = y;
  }
  float rb_next = 0.f, x_next = 0.f, tr = 0.f, mh = 0.f;
  for (int i = D - 1; i >= 0; --i) {
    const float delta = a1 * d[i];
    const float r = rq[32 * i + lane];
    const float iq = 1.f / (1.f + r);
    float rb = delta, x = ys[32 * i + lane] * iq;
    if (i < D - 1) {
      const float b = a1 * e[i];
      rb = fmaf(-b, b / (1.f + rb_next), delta);
      x = fmaf(-(b * iq), x_next, x);
    }
    const float sdev = r + rb - delta;
    tr -= sdev / (1.f + sdev);
    mh = fmaf(x, x, mh);
    rb_next = rb;
    x_next = x;
  }
  const float kl = 0.5f * (ld + tr + mh * a1 * a1);
  return (pd && isfinite(kl)) ? kl : FLT_MAX;
}

__global__ void __launch_bounds__(THREADS, 1)
update_full_blocked_kernel(int mode, const float* __restrict__ means, const float* __restrict__ chols,
                           const float* __restrict__ Bmat, const float* __restrict__ B2mat,
                           const float* __restrict__ hvec, const float* __restrict__ stepsizes,
                           const float* __restrict__ last_etas, const float* __restrict__ num_updates, int D,
                           float temperature, float* __restrict__ out_means, float* __restrict__ out_chols,
                           int32_t* __restrict__ success, float* __restrict__ etas, float* __restrict__ kls,
                           int32_t* __restrict__ evals, const float* __restrict__ tdiag,
                           const float* __restrict__ toff, const float* __restrict__ thp) {
  extern __shared__ __align__(16) float ub_smem[];
  __shared__ float search_out[4];
  __shared__ float red[33];
  __shared__ float idiag[NB];
  __shared__ int flag;
  __shared__ int blk_bi[MAXBLK * (MAXBLK + 1) / 2];
  const int k = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < MAXBLK * (MAXBLK + 1) / 2) {
    int bi = 0;
    while (tri(bi + 1) <= tid) ++bi;
    blk_bi[tid] = bi;
  }
  const int nbk = (D + NB - 1) / NB, Dp = nbk * NB;
  Smem s;
  s.A = ub_smem;
  s.PT = s.A + tri(nbk) * BS;
  s.hrev = s.PT + (nbk > 1 ? nbk - 1 : 1) * BS;
  s.v = s.hrev + Dp;
  s.u = s.v + Dp;
  s.dm1 = s.u + Dp;
  s.idiag = idiag;
  s.red = red;
  s.flag = &flag;
  s.blk_bi = blk_bi;
  const float* B = Bmat + (long long)k * D * D;
  const float* B2 = (mode == 2) ? B2mat + (long long)k * D * D : nullptr;
  const float* L = chols + (long long)k * D * D;
  const float* mu = means + (long long)k * D;
  for (int i = tid; i < Dp; i += THREADS) s.hrev[i] = i < D ? hvec[(long long)k * D + (D - 1 - i)] : 0.f;
  __syncthreads();

  const float step = stepsizes[k];
  bool ok = true;
  float eta = -1.f, kl = -1.f;
  int n_evals = 0;
  if (mode == 0) {
    // ---- bracketing search in log space (:335-429), cold / warm bracket (:462-471) ----
    const float last = last_etas[k];
    float lower, upper;
    if (last < 0.f) { lower = -20.f; upper = 80.f; }
    else { lower = fmaxf(0.f, logf(last) - 3.f); upper = logf(last) + 3.f; }
    float leta = 0.5f * (upper + lower);
    bool feasible = false;
    float last_e = -1.f;                           // eta of the most recent Cholesky evaluation (its results are in smem)
    KlTerms last_t;
    last_t.ok = false;
    last_t.kl = FLT_MAX;
    if (tdiag != nullptr) {
      // ---- the same search with KL(eta) evaluated from the tridiagonal form: warp 0 evaluates the 31 candidate etas
      // of the next five bisection levels at once (lane = node of the decision tree, heap numbering: left child =
      // "feasible, upper = eta", right child = "lower = eta") and then walks the tree with the reference's rules ----
      float* sd = s.A;
      float* se = sd + Dp;
      float* sh = se + Dp;
      float* rq = sh + Dp;
      float* ys = rq + 32 * Dp;
      for (int i = tid; i < D; i += THREADS) {
        sd[i] = tdiag[(long long)k * D + i];
        se[i] = toff[(long long)k * D + i];
        sh[i] = thp[(long long)k * D + i];
      }
      __syncthreads();
      if (warp == 0) {
        bool done = false;
        for (int round = 0; round < 200 && !done; ++round) {
          float lo = lower, up = upper;
          const int depth = 31 - __clz(max(lane, 1));
          for (int b = depth - 1; b >= 0; --b) {
            const float mid = 0.5f * (up + lo);
            if ((lane >> b) & 1) lo = mid; else up = mid;
          }
          const float my_kl = kl_tridiag_lane(1.f / expf(0.5f * (up + lo)), sd, se, sh, D, rq, ys);
          int node = 1;
          for (int lvl = 0; lvl < 5; ++lvl) {
            leta = 0.5f * (upper + lower);
            const float diff = fminf(expf(upper) - expf(leta), expf(leta) - expf(lower));
            if (diff < 1e-1f) { done = true; break; }
            const float klv = __shfl_sync(0xffffffffu, my_kl, node);
            ++n_evals;
            if (fabsf(step - klv) < 1e-1f * step) { lower = upper = leta; done = true; break; }
            if (step > klv) { upper = leta; feasible = true; node = 2 * node; }
            else { lower = leta; node = 2 * node + 1; }
          }
        }
        if (lane == 0) {
          search_out[0] = lower;
          search_out[1] = upper;
          search_out[2] = feasible ? 1.f : 0.f;
          search_out[3] = (float)n_evals;
        }
      }
      __syncthreads();
      lower = search_out[0];
      upper = search_out[1];
      feasible = search_out[2] != 0.f;
      n_evals = (int)search_out[3];
      // the scratch of the search may reach into the vectors behind the blocks when D <= 64: restore h~
      for (int i = tid; i < Dp; i += THREADS) s.hrev[i] = i < D ? hvec[(long long)k * D + (D - 1 - i)] : 0.f;
      __syncthreads();
    } else
    for (int it = 0; it < 1000; ++it) {
      const float diff = fminf(expf(upper) - expf(leta), expf(leta) - expf(lower));
      if (diff < 1e-1f) break;
      const float e = expf(leta);
      const KlTerms t = eval_whitened(s, B, nullptr, 1.f / e, 0.f, D, nbk, 1.f / e);
      ++n_evals;
      last_e = e;
      last_t = t;
      if (fabsf(step - t.kl) < 1e-1f * step) { lower = upper = leta; break; }
      if (step > t.kl) { upper = leta; feasible = true; }
      else lower = leta;
      leta = 0.5f * (upper + lower);
    }
    if (feasible) lower = upper;
    const float new_lower = expf(lower), new_upper = expf(upper);
    eta = fmaxf(new_lower, temperature);
    ok = (new_lower == new_upper);
    if (ok) {
      // The reference evaluates kl() once more at the eta it settled on (:478-497).  When that is the eta of the search's
      // last evaluation (the usual exit: |eps - KL| < 0.1 eps, or the bracket closing right after a feasible step), the
      // factor, its inverse and M^-1 h are still in shared memory and the second, bit-identical evaluation is skipped.
      const KlTerms t = (eta == last_e) ? last_t : eval_whitened(s, B, nullptr, 1.f / eta, 0.f, D, nbk, 1.f / eta);
      ++n_evals;                                   // counted like the reference does
      ok = t.ok && (t.kl < FLT_MAX) && isfinite(t.kl);
      kl = t.kl;
    }
  } else if (mode == 1) {
    eta = 1.f / step;
    const KlTerms t = eval_whitened(s, B, nullptr, step, 0.f, D, nbk, step);
    ok = t.ok && isfinite(t.kl);
    kl = t.kl;
  } else {
    const KlTerms t = eval_whitened(s, B, B2, step, 0.5f * step * step, D, nbk, step);
    ok = t.ok && isfinite(t.kl);
    kl = t.kl;
  }

  float* om = out_means + (long long)k * D;
  float* oc = out_chols + (long long)k * D * D;
  if (ok) {
    // L'[i][j] = sum_{m=j..i} L[i][m] X~[D-1-j][D-1-m]   (L' = L U^-T, U^-T[m][j] = X~[D-1-j][D-1-m]),
    // and the new mean mu - scale * L w with w[m] = u~[D-1-m] (KL / direct) or h~[D-1-m] (iBLR).
    // Warp task = (row block bi, column block bj, quarter cq of the 32 output columns); lane = row, holding its
    // 32 entries of L(bi, bm) in registers while X~ is broadcast.
    const bool first = (mode == 2) && (num_updates[k] == 0.f);
    const float scale = (mode == 0) ? 1.f / eta : step;
    const float* wrev = (mode == 2) ? s.hrev : s.u;
    bool finite = true;
    const int ntask = tri(nbk) * 4;
    for (int t = warp; t < ntask; t += NWARPS) {
      const int b = t >> 2, cq = t & 3;
      int bi = 0;
      while (tri(bi + 1) <= b) ++bi;
      const int bj = b - tri(bi);
      const int i = bi * NB + lane;                          // output row (original index)
      float acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = 0.f;
      float macc = 0.f;
      const bool do_mean = (bj == 0 && cq == 0);
      for (int bm = bj; bm <= bi; ++bm) {
        float l[NB];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int m0 = bm * NB + 4 * q;
          float4 tv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (i < D) {
            const float* src = L + (long long)i * D + m0;
            if (m0 + 3 < D && (D & 3) == 0) tv = __ldg(reinterpret_cast<const float4*>(src));
            else {
              if (m0 < D) tv.x = __ldg(src);
              if (m0 + 1 < D) tv.y = __ldg(src + 1);
              if (m0 + 2 < D) tv.z = __ldg(src + 2);
              if (m0 + 3 < D) tv.w = __ldg(src + 3);
            }
          }
          l[4 * q] = tv.x; l[4 * q + 1] = tv.y; l[4 * q + 2] = tv.z; l[4 * q + 3] = tv.w;
        }
        if (bm == bi) {       // L is lower triangular: ignore whatever the caller left above the diagonal
#pragma unroll
          for (int kk = 0; kk < NB; ++kk)
            if (kk > lane) l[kk] = 0.f;
        }
        if (do_mean) {
#pragma unroll
          for (int kk = 0; kk < NB; ++kk) {
            const int m = bm * NB + kk;
            if (m < D) macc = fmaf(l[kk], wrev[D - 1 - m], macc);
          }
        }
        if ((D & 31) == 0) {
          // aligned case: X~[a][D-1-m] for m = 32 bm + kk is column 31 - kk of block column nbk-1-bm, so four
          // consecutive kk are one (reversed) 16-byte chunk; entries with m < j are the stored zeros above the diagonal
          const int cb = nbk - 1 - bm;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int a = D - 1 - (bj * NB + 8 * cq + c);
            const float* xrow = s.A + (tri(a >> 5) + cb) * BS;
            const int ar = a & 31;
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
              const float4 xv = ld4(xrow + sw4(ar, 7 - k4));
              s0 = fmaf(l[4 * k4], xv.w, s0);
              s1 = fmaf(l[4 * k4 + 1], xv.z, s1);
              s0 = fmaf(l[4 * k4 + 2], xv.y, s0);
              s1 = fmaf(l[4 * k4 + 3], xv.x, s1);
            }
            acc[c] += s0 + s1;
          }
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int j = bj * NB + 8 * cq + c;               // output column (original index)
            if (j >= D) continue;
            const int a = D - 1 - j;                          // row of X~
            const int ab = a >> 5, ar = a & 31;
            float sacc = 0.f;
#pragma unroll
            for (int kk = 0; kk < NB; ++kk) {
              const int m = bm * NB + kk;
              if (m < D && m >= j) {
                const int bb = D - 1 - m;                     // column of X~ (bb <= a)
                sacc = fmaf(l[kk], s.A[(tri(ab) + (bb >> 5)) * BS + sw(ar, bb & 31)], sacc);
              }
            }
            acc[c] += sacc;
          }
        }
      }
      if (i < D) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int j = bj * NB + 8 * cq + c;
          if (j < D) {
            const float val = (j <= i) ? acc[c] : 0.f;
            finite &= isfinite(val);
            oc[(long long)i * D + j] = val;
          }
        }
        if (do_mean) om[i] = first ? mu[i] : mu[i] - scale * macc;
      }
    }
    // strictly upper blocks are zero
    for (long long e = tid; e < (long long)D * D; e += THREADS) {
      const int i = (int)(e / D), j = (int)(e % D);
      if ((j >> 5) > (i >> 5)) oc[e] = 0.f;
    }
    ok = !__syncthreads_or(!finite);
  }
  if (!ok) {
    __syncthreads();
    for (int i = tid; i < D; i += THREADS) om[i] = mu[i];
    for (long long e = tid; e < (long long)D * D; e += THREADS) oc[e] = L[e];
    eta = -1.f;
    kl = -1.f;
  }
  if (tid == 0) {
    success[k] = ok ? 1 : 0;
    if (etas) etas[k] = eta;
    if (kls) kls[k] = kl;
    if (evals) evals[k] = n_evals;
  }
}

size_t blocked_smem_bytes(int D) {
  const int nbk = (D + NB - 1) / NB;
  return ((size_t)(tri(nbk) + (nbk > 1 ? nbk - 1 : 1)) * BS + 4 * (size_t)nbk * NB) * sizeof(float);
}

}  // namespace ub

bool update_blocked_supported(int D) { return D <= ub::MAXBLK * ub::NB; }

int launch_update_full_blocked(int mode, const float* means, const float* chols, const float* Bm, const float* B2,
                               const float* hv, const float* stepsizes, const float* last_etas,
                               const float* num_updates, int K, int D, float temperature, float* out_means,
                               float* out_chols, int32_t* success, float* etas, float* kls, int32_t* evals,
                               const float* tdiag, const float* toff, const float* thp, cudaStream_t st) {
  const size_t smem = ub::blocked_smem_bytes(D);
  static unsigned long long attr_set_mask = 0;
  if (first_call_on_device(attr_set_mask)) {
    cudaError_t e = cudaFuncSetAttribute(ub::update_full_blocked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)ub::blocked_smem_bytes(ub::MAXBLK * ub::NB));
    if (e != cudaSuccess) {
      set_last_error("update_full_blocked: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return GVI_ERR_CUDA;
    }
  }
  ub::update_full_blocked_kernel<<<K, ub::THREADS, smem, st>>>(mode, means, chols, Bm, B2, hv, stepsizes, last_etas,
                                                               num_updates, D, temperature, out_means, out_chols,
                                                               success, etas, kls, evals, tdiag, toff, thp);
  return check_launch("update_full_blocked_kernel");
}

}  // namespace gvi
