// Microbenchmark: tcgen05.mma rate on every SM while a second warp streams cp.async.bulk loads
// (L2 -> shared memory) into the same CTA, as conv_tc_kernel's producer does.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_tma_contention tools/mma_tma_contention.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}

struct Cfg { int N; int iters; int copy_bytes; int src_mode; int pace_cycles; };   // src_mode 0 = none, 1 = hot (same for all CTAs), 2 = per-CTA

constexpr int SLOTS = 4;

__global__ void __launch_bounds__(128, 1) k(Cfg c, const uint8_t* src, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];   // [0,96K) A, [96K,128K) B, [128K, 128K+64K) copy ring
  __shared__ uint64_t bar, cbar[SLOTS];
  __shared__ uint32_t tmem_ptr;
  __shared__ volatile int done;
  __shared__ long long copied;
  for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    done = 0; copied = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    for (int s = 0; s < SLOTS; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&cbar[s])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = __shfl_sync(0xffffffffu, tmem_ptr, 0);
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | (8u << 24);
    const uint32_t sb = smem_u32(smem);
    const uint32_t hi = 8u | (1u << 14);
    const uint32_t a_lo = ((sb >> 4) & 0x3FFF) | (((16384u + 64u) >> 4) << 16);
    const uint32_t b_lo = (((sb + 96 * 1024) >> 4) & 0x3FFF) | ((((uint32_t)c.N * 16u) >> 4) << 16);
    const int n_acc = 512 / c.N < 4 ? 512 / c.N : 4;
    const long long t0 = clock64();
    uint32_t a = a_lo;
    int acc_i = 0;
    for (int i = 0; i < c.iters; ++i) {
      if (leader) mma(tm + (uint32_t)(acc_i * c.N), ((uint64_t)hi << 32) | a, ((uint64_t)hi << 32) | b_lo, idesc, 1u);
      a += 1; if ((i & 31) == 31) a = a_lo;
      if (++acc_i == n_acc) acc_i = 0;
    }
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    while (!try_wait(smem_u32(&bar), 0)) {}
    const long long t2 = clock64();
    done = 1;
    if (threadIdx.x == 0) { out[blockIdx.x * 2] = t2 - t0; }
  } else if (warp == 1 && c.src_mode != 0) {
    if (threadIdx.x == 32) {
      const uint8_t* s = src + (c.src_mode == 2 ? (size_t)blockIdx.x * (1u << 20) : 0);
      const uint32_t ring = smem_u32(smem) + 128 * 1024;
      uint32_t ph[SLOTS] = {0, 0, 0, 0};
      long long n = 0;
      int slot = 0, inflight = 0;
      uint32_t off = 0;
      long long next_t = clock64();
      while (!done) {
        if (inflight == SLOTS || 0) {
          while (!try_wait(smem_u32(&cbar[slot]), ph[slot])) {}
          ph[slot] ^= 1; --inflight; n += c.copy_bytes;
        }
        if (c.pace_cycles) { while (clock64() < next_t) {} next_t += c.pace_cycles; }
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&cbar[slot])), "r"(c.copy_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ring + slot * 16384),
                     "l"(s + off), "r"(c.copy_bytes), "r"(smem_u32(&cbar[slot])) : "memory");
        off = (off + c.copy_bytes) & ((1u << 19) - 1);
        ++inflight;
        if (++slot == SLOTS) slot = 0;
      }
      // drain
      for (int j = 0; j < inflight; ++j) {
        while (!try_wait(smem_u32(&cbar[slot]), ph[slot])) {}
        ph[slot] ^= 1;
        if (++slot == SLOTS) slot = 0;
      }
      out[blockIdx.x * 2 + 1] = n;
    }
  } else if (warp == 1 && threadIdx.x == 32) {
    out[blockIdx.x * 2 + 1] = 0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 16);
  uint8_t* src; cudaMalloc(&src, (size_t)160 << 20); cudaMemset(src, 0, (size_t)160 << 20);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("N src copyB pace  cyc/mma  copy_B/cyc/SM  chip_TB/s@1.9GHz\n");
  const int Ns[] = {64, 128, 256};
  for (int ni = 0; ni < 3; ++ni)
    for (int mode = 0; mode < 3; ++mode)
      for (int cb = 4096; cb <= 16384; cb *= 4)
        for (int pace = 0; pace <= 2048; pace = pace ? pace * 4 : 128) {
          if (mode == 0 && (cb != 4096 || pace != 0)) continue;
          Cfg c{Ns[ni], 8192, cb, mode, pace};
          k<<<148, 128, 200 * 1024>>>(c, src, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[296]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
          double cyc = 0, by = 0;
          for (int b = 0; b < 148; ++b) { cyc += h[2 * b]; by += h[2 * b + 1]; }
          cyc /= 148; by /= 148;
          printf("%3d %d %5d %4d  %7.1f  %7.2f  %6.2f\n", c.N, mode, cb, pace, cyc / c.iters, by / cyc, by / cyc * 148 * 1.9e9 / 1e12);
        }
  return 0;
}
