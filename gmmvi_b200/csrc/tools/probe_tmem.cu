// Stand-alone probe (developer tool, not part of libgmmvi_b200.so): checks on a real sm_100a device
//   T1  which (lane, column) each register of tcgen05.st.16x256b lands in (read back with tcgen05.ld.32x32b)
//   T2  the operand layout tcgen05.mma.kind::f16 expects for an A matrix that lives in TMEM
//   T3  the issue rate of SS (A from shared memory) and TS (A from TMEM) MMAs for several N
// build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a probe_tmem.cu -o probe_tmem
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 2000000000LL) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st_16x256b_x1(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B descriptor: 8-row x 128-byte atoms, 1024 B between atoms
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

struct Shared {
  uint64_t bar;
  uint32_t tmem_base;
};

// out1[128][8]: T1 read-back; out2[128][16]: T2 result; out3[16]: T3 cycle counts
__global__ void __launch_bounds__(128) probe_kernel(uint32_t* out1, float* out2, long long* out3) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ Shared sh;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&sh.bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&sh.tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = sh.tmem_base;
  const uint32_t quad = tb + ((uint32_t)(32 * warp) << 16);

  // ---------------- T1 ----------------
  for (int half = 0; half < 2; ++half) {
    const uint32_t tag = ((uint32_t)half << 16) | ((uint32_t)lane << 8);
    tmem_st_16x256b_x1(quad + ((uint32_t)(16 * half) << 16), tag | 0, tag | 1, tag | 2, tag | 3);
  }
  tmem_st_wait();
  {
    uint32_t v[8];
    tmem_ld8(quad, v);
    for (int c = 0; c < 8; ++c) out1[(32 * warp + lane) * 8 + c] = v[c];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---------------- T2 ----------------
  // A[r][k] = 16 r + k (exact in fp16), written with the layout under test: lane = row, column c packs (2c, 2c+1),
  // low half = even k.  With the 16x256b store: thread t of a 16-lane half holds row t/4 (+8), columns 2 (t%4), +1.
  for (int half = 0; half < 2; ++half) {
    uint32_t r[4];
    for (int q = 0; q < 2; ++q) {
      const int row = 32 * warp + 16 * half + (lane >> 2) + 8 * q;
      for (int cc = 0; cc < 2; ++cc) {
        const int col = 2 * (lane & 3) + cc;
        const __half2 h = __floats2half2_rn((float)(16 * row + 2 * col), (float)(16 * row + 2 * col + 1));
        r[2 * q + cc] = *reinterpret_cast<const uint32_t*>(&h);
      }
    }
    tmem_st_16x256b_x1(quad + ((uint32_t)(16 * half) << 16) + 64u, r[0], r[1], r[2], r[3]);
  }
  tmem_st_wait();
  // B[n][k] = (n == k), N = 16 rows of 128 bytes, SWIZZLE_128B: 16-byte chunk c of row n stored at chunk c ^ (n & 7)
  for (int i = threadIdx.x; i < 16 * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  if (threadIdx.x < 16) {
    const int n = threadIdx.x, k = n;
    const int chunk = (k >> 3) ^ (n & 7);
    reinterpret_cast<__half*>(smem + n * 128 + chunk * 16)[k & 7] = __float2half(1.0f);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    umma_ts(tb + 128u, tb + 64u, make_desc(smem_u32(smem)), make_idesc_f16(16), 0u);
    umma_commit(&sh.bar);
  }
  mbar_wait(&sh.bar, 0);
  tc_fence_after();
  for (int cb = 0; cb < 2; ++cb) {
    uint32_t v[8];
    tmem_ld8(quad + 128u + 8u * cb, v);
    for (int c = 0; c < 8; ++c) out2[(32 * warp + lane) * 16 + cb * 8 + c] = __uint_as_float(v[c]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---------------- T3: MMA issue rate, SS vs TS, K = 16 per instruction ----------------
  // (operand contents are whatever is in shared memory / TMEM; only the timing matters).  Warp 0 runs the loop with
  // warp-uniform operands and elects one lane per MMA, like the production kernels.
  uint32_t phase = 1;
  const int NS[4] = {256, 128, 64, 16};
  for (int mode = 0; mode < 2; ++mode) {
    for (int ni = 0; ni < 5; ++ni) {
      long long t0 = 0;
      if (warp == 0) {
        const uint64_t adesc = make_desc(smem_u32(smem + 65536));
        const uint64_t bdesc = make_desc(smem_u32(smem));
        t0 = clock64();
        if (ni < 4) {
          const uint32_t idesc = make_idesc_f16(NS[ni]);
          for (int i = 0; i < 1024; i += 4) {
            if (elect_one()) {
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                if (mode == 0) umma_ss(tb + 256u, adesc + (uint64_t)(u * 2), bdesc + (uint64_t)(u * 2), idesc, 1u);
                else umma_ts(tb + 256u, tb + (uint32_t)(u * 8), bdesc + (uint64_t)(u * 2), idesc, 1u);
              }
            }
            __syncwarp();
          }
        } else {
          // 16 work items of the log-density kernel: 16 steps x 3 MMAs with N = 256 - 16 s
          for (int item = 0; item < 16; ++item) {
            for (int s4 = 0; s4 < 16; s4 += 2) {
              if (elect_one()) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  const int st = s4 + u;
                  const uint32_t idesc = make_idesc_f16(256 - 16 * st);
                  const uint32_t d = tb + 256u + (uint32_t)(16 * st);
                  const uint64_t bd = bdesc + (uint64_t)(u * 2);
                  if (mode == 0) {
                    umma_ss(d, adesc + (uint64_t)(u * 2), bd, idesc, 1u);
                    umma_ss(d, adesc + (uint64_t)(u * 2 + 512), bd, idesc, 1u);
                    umma_ss(d, adesc + (uint64_t)(u * 2), bd + 1024, idesc, 1u);
                  } else {
                    umma_ts(d, tb + (uint32_t)(u * 8), bd, idesc, 1u);
                    umma_ts(d, tb + (uint32_t)(u * 8 + 16), bd, idesc, 1u);
                    umma_ts(d, tb + (uint32_t)(u * 8), bd + 1024, idesc, 1u);
                  }
                }
              }
              __syncwarp();
            }
          }
        }
        if (lane == 0) umma_commit(&sh.bar);
        __syncwarp();
      }
      mbar_wait(&sh.bar, phase);
      phase ^= 1;
      tc_fence_after();
      if (threadIdx.x == 0) out3[mode * 5 + ni] = clock64() - t0;
      __syncthreads();
    }
  }
  // ---------------- T4: TS MMAs, one accumulator vs two alternating accumulators; issue time vs completion time ----
  for (int var = 0; var < 6; ++var) {
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
      const uint64_t bdesc = make_desc(smem_u32(smem));
      const int n = (var % 3 == 0) ? 16 : ((var % 3 == 1) ? 64 : 256);
      const bool alt = var >= 3;
      const uint32_t idesc = make_idesc_f16(n);
      t0 = clock64();
      for (int i = 0; i < 256; i += 4) {
        if (elect_one()) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t d = tb + ((alt && (u & 1)) ? 0u : 256u);   // A lives in columns 0..63 -> use 64.. for the 2nd acc
            umma_ts(alt && (u & 1) ? tb + 64u : d, tb + (uint32_t)(u * 8), bdesc + (uint64_t)(u * 2), idesc, 1u);
          }
        }
        __syncwarp();
      }
      t1 = clock64();
      if (lane == 0) umma_commit(&sh.bar);
      __syncwarp();
    }
    mbar_wait(&sh.bar, phase);
    phase ^= 1;
    tc_fence_after();
    if (threadIdx.x == 0) {
      out3[10 + 2 * var] = t1 - t0;
      out3[11 + 2 * var] = clock64() - t0;
    }
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  uint32_t* d1;
  float* d2;
  long long* d3;
  cudaMalloc(&d1, 128 * 8 * 4);
  cudaMalloc(&d2, 128 * 16 * 4);
  cudaMalloc(&d3, 32 * 8);
  cudaMemset(d3, 0, 32 * 8);
  const int smem = 1024 + 65536 + 65536;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem>>>(d1, d2, d3);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<uint32_t> h1(128 * 8);
  std::vector<float> h2(128 * 16);
  long long h3[32];
  cudaMemcpy(h1.data(), d1, h1.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h2.data(), d2, h2.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h3, d3, sizeof(h3), cudaMemcpyDeviceToHost);
  // T1: expected mapping row = 16 half + lane/4 + 8 (reg/2), col = 2 (lane%4) + reg%2
  int bad1 = 0;
  for (int row = 0; row < 32; ++row)
    for (int c = 0; c < 8; ++c) {
      const uint32_t v = h1[row * 8 + c];
      const int half = (v >> 16) & 1, lane = (v >> 8) & 31, reg = v & 3;
      const int erow = 16 * half + lane / 4 + 8 * (reg / 2), ecol = 2 * (lane % 4) + reg % 2;
      if (erow != row || ecol != c) {
        if (bad1 < 16) printf("T1 row %d col %d holds half %d lane %d reg %d\n", row, c, half, lane, reg);
        ++bad1;
      }
    }
  printf("T1 (tcgen05.st.16x256b mapping): %s\n", bad1 ? "DIFFERENT from the assumed mapping" : "as assumed");
  int bad2 = 0;
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < 16; ++n)
      if (h2[r * 16 + n] != (float)(16 * r + n)) {
        if (bad2 < 16) printf("T2 D[%d][%d] = %g, expected %d\n", r, n, h2[r * 16 + n], 16 * r + n);
        ++bad2;
      }
  printf("T2 (A operand from TMEM, lane = row, column packs k = 2c, 2c+1): %s\n", bad2 ? "MISMATCH" : "ok");
  const int NS[4] = {256, 128, 64, 16};
  for (int mode = 0; mode < 2; ++mode) {
    for (int ni = 0; ni < 4; ++ni)
      printf("T3 %s M=128 N=%3d K=16: %.1f cycles / MMA  (%.0f MAC/clk)\n", mode ? "TS" : "SS", NS[ni], h3[mode * 5 + ni] / 1024.0,
             128.0 * NS[ni] * 16 * 1024.0 / h3[mode * 5 + ni]);
    printf("T3 %s log-density item (48 MMAs, N = 256 - 16 s): %.0f cycles / item\n", mode ? "TS" : "SS", h3[mode * 5 + 4] / 16.0);
  }
  for (int var = 0; var < 6; ++var)
    printf("T4 TS N=%3d %s: issue loop %.1f cycles / MMA, until completion %.1f cycles / MMA\n",
           (var % 3 == 0) ? 16 : ((var % 3 == 1) ? 64 : 256), var >= 3 ? "two accumulators" : "one accumulator ",
           h3[10 + 2 * var] / 256.0, h3[11 + 2 * var] / 256.0);
  return (bad1 || bad2) ? 2 : 0;
}
