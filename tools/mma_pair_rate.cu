// Microbenchmark: tcgen05.mma.cta_group::2 (M = 256 over a CTA pair) issue/execute rate vs N, no-swizzle K-major
// operands, B N-split across the pair.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_pair_rate tools/mma_pair_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
struct Cfg { int N; int iters; int nw; };

template <int NW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  __shared__ long long tmax[4];
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(NW));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;"); asm volatile("barrier.cluster.wait.acquire.aligned;");
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = __shfl_sync(0xffffffffu, tmem_ptr, 0);
  if (rank == 0 && warp < NW) {
    const bool leader = elect_one();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | (16u << 24);   // M = 256
    const uint32_t sb = smem_u32(smem);
    const uint32_t hi = 8u | (1u << 14);
    const uint32_t a_lo = ((sb >> 4) & 0x3FFF) | (((16384u + 64u) >> 4) << 16);
    const uint32_t nh = (uint32_t)c.N / 2;
    const uint32_t b_lo = (((sb + 96 * 1024) >> 4) & 0x3FFF) | (((nh * 16u) >> 4) << 16);
    const int n_acc = 512 / c.N < 4 ? 512 / c.N : 4;
    const long long t0 = clock64();
    uint32_t a = a_lo;
    for (int i = 0; i < c.iters; ++i) {
      const int acc_i = (i * NW + warp) % n_acc;
      if (leader)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tm + (uint32_t)(acc_i * c.N)), "l"(((uint64_t)hi << 32) | a), "l"(((uint64_t)hi << 32) | b_lo), "r"(idesc), "r"(1u) : "memory");
      a += 1; if ((i & 31) == 31) a = a_lo;
    }
    if (leader) asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
    while (!try_wait(smem_u32(&bar), 0)) {}
    const long long t2 = clock64();
    if ((threadIdx.x & 31) == 0) tmax[warp] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;"); asm volatile("barrier.cluster.wait.acquire.aligned;");
  if (rank == 0 && threadIdx.x == 0) { long long m = 0; for (int w = 0; w < NW; ++w) if (tmax[w] > m) m = tmax[w]; out[blockIdx.x / 2] = m; }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  printf("pair MMA (M=256)  N  warps  cyc per MMA instruction (ideal math N/2; per-SM smem model 32 + N/8)\n");
  const int Ns[] = {32, 64, 128, 256};
  for (int ni = 0; ni < 4; ++ni)
    for (int nw = 1; nw <= 4; nw *= 2) {
      Cfg c{Ns[ni], 4096, nw};
      void (*fn)(Cfg, long long*) = nw == 4 ? k<4> : nw == 2 ? k<2> : k<1>;
      cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      fn<<<148, 128, 200 * 1024>>>(c, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[74]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      double cyc = 0; for (int b = 0; b < 74; ++b) cyc += h[b]; cyc /= 74;
      printf("N=%3d warps=%d : %6.1f cyc/MMA  (math %d, smem model %d)\n", c.N, nw, cyc / (c.iters * nw), c.N / 2, 32 + c.N / 8);
    }
  return 0;
}
