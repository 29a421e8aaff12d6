// Shared helpers for the gmmvi_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <float.h>
#include "../../include/gmmvi_b200.h"

namespace gvi {

// ---- error plumbing (C ABI never throws; see include/gmmvi_b200.h) -------------------------
void set_last_error(const char* fmt, ...);
int  check_launch(const char* what);          // cudaGetLastError -> GVI_ERR_CUDA


#define GVI_REQUIRE(cond, ...)                                  \
  do {                                                          \
    if (!(cond)) {                                              \
      ::gvi::set_last_error(__VA_ARGS__);                       \
      return GVI_ERR_INVALID;                            \
    }                                                           \
  } while (0)

// One-time-per-DEVICE initialisation (cudaFuncSetAttribute, SM count): `mask` is a function-local static; returns true the
// first time it is called on the current device.  (One process normally drives one GPU; a process that drives several
// must not inherit the first device's setup.)
inline bool first_call_on_device(unsigned long long& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return true;
  const unsigned long long bit = 1ull << dev;
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

constexpr float kLog2Pi = 1.8378770664093454835606594728112f;
constexpr double kLog2PiD = 1.8378770664093454835606594728112;

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- warp / block reductions ----------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T u = __shfl_xor_sync(0xffffffffu, v, o);
    v = u > v ? u : v;
  }
  return v;
}
// Block-wide sum; every thread gets the result.  `scratch` holds >= 33 T.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  if (w == 0) {
    T t = lane < nw ? scratch[lane] : T(0);
    t = warp_sum(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}
template <typename T>
__device__ __forceinline__ T block_max(T v, T* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  if (w == 0) {
    T t = lane < nw ? scratch[lane] : scratch[0];
    t = warp_max(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

// ---- SIMT fp32 tile engine -------------------------------------------------------------------
// C[BM x BN] += A[BM x k] * B[BN x k]^T, 256 threads, each thread an 8x4 micro tile.
// Operands are staged in shared memory k-major ("As[kk][m]") by per-kernel loader functors, so
// the same engine serves the log-density, mixture-gradient, Stein, sampling and batched-GEMM
// kernels, each with its own fused prologue (x - mu, weights, ...).
constexpr int BM = 128, BN = 64, BK = 16, NTHREADS = 256, TM = 8, TN = 4;
constexpr int AS_LD = BM + 4, BS_LD = BN + 4;

struct __align__(16) SmemTiles {
  float As[2][BK][AS_LD];
  float Bs[2][BK][BS_LD];
};

__device__ __forceinline__ void tile_compute(float (&acc)[TM][TN], const float (*As)[AS_LD],
                                             const float (*Bs)[BS_LD], int ty, int tx) {
#pragma unroll
  for (int kk = 0; kk < BK; ++kk) {
    const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * TM]);
    const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * TM + 4]);
    const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN]);
    const float a[TM] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float b[TN] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
    for (int r = 0; r < TM; ++r)
#pragma unroll
      for (int c = 0; c < TN; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
  }
}

// Loader functors: fetch(chunk, regs) reads global memory into registers, store(smem, regs)
// writes them k-major into shared memory.  A tile: 8 floats / thread, B tile: 4 floats / thread.
template <class FetchA, class StoreA, class FetchB, class StoreB>
__device__ __forceinline__ void tile_mainloop(float (&acc)[TM][TN], SmemTiles& sm, int chunk_begin,
                                              int chunk_end, int ty, int tx, FetchA fetchA,
                                              StoreA storeA, FetchB fetchB, StoreB storeB) {
  if (chunk_end <= chunk_begin) return;
  float ra[8], rb[4];
  fetchA(chunk_begin, ra);
  fetchB(chunk_begin, rb);
  storeA(sm.As[0], ra);
  storeB(sm.Bs[0], rb);
  __syncthreads();
  int cur = 0;
  for (int c = chunk_begin; c < chunk_end; ++c) {
    const bool more = (c + 1 < chunk_end);
    if (more) {
      fetchA(c + 1, ra);
      fetchB(c + 1, rb);
    }
    tile_compute(acc, sm.As[cur], sm.Bs[cur], ty, tx);
    if (more) {
      storeA(sm.As[cur ^ 1], ra);
      storeB(sm.Bs[cur ^ 1], rb);
    }
    __syncthreads();
    cur ^= 1;
  }
}

// Read 4 consecutive floats p[0..3] where only the first `valid` (<=4, may be <=0) are in range.
__device__ __forceinline__ float4 load4(const float* __restrict__ p, int valid, bool vec_ok) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid >= 4 && vec_ok) {
    v = __ldg(reinterpret_cast<const float4*>(p));
  } else {
    if (valid > 0) v.x = __ldg(p);
    if (valid > 1) v.y = __ldg(p + 1);
    if (valid > 2) v.z = __ldg(p + 2);
    if (valid > 3) v.w = __ldg(p + 3);
  }
  return v;
}

// "k-contiguous" source (row r holds its k values contiguously, e.g. X[n][j]): thread t loads rows
// t/4 (and t/4+64 for the 128-row A tile), 4 consecutive k at (t%4)*4; stored transposed.
// "row-contiguous" source (for fixed k the rows are contiguous, e.g. P[j][i]): thread t loads k = t/16,
// rows (t%16)*4.. (and +64 for A); stored as float4.

}  // namespace gvi
