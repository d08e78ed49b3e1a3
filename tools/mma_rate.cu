// Microbenchmark: sustained tcgen05.mma rate on one SM for the operand layouts conv_tc_kernel uses.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_rate tools/mma_rate.cu
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

struct Cfg { int N; int layout; int vary_a; int n_acc; int iters; int a_step16; };

__global__ void __launch_bounds__(128, 1) rate_kernel(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
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
    uint32_t a_lo, b_lo, hi;
    if (c.layout == 0) {  // no swizzle: SBO 128 B, LBO 16 KB (A) / N*16 (B)
      hi = 8u | (1u << 14);
      a_lo = ((sb >> 4) & 0x3FFF) | (((16384u + 64u) >> 4) << 16);
      b_lo = (((sb + 96 * 1024) >> 4) & 0x3FFF) | ((((uint32_t)c.N * 16u) >> 4) << 16);
    } else {              // 128B swizzle: SBO 1024 B, layout type 2 (bits 61-63 -> hi bits 29-31)
      hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      a_lo = ((sb >> 4) & 0x3FFF) | (1u << 16);
      b_lo = (((sb + 96 * 1024) >> 4) & 0x3FFF) | (1u << 16);
    }
    const long long t0 = clock64();
    uint32_t a = a_lo;
    int acc_i = 0;
    for (int i = 0; i < c.iters; ++i) {
      if (leader) mma(tm + (uint32_t)(acc_i * c.N), ((uint64_t)hi << 32) | a, ((uint64_t)hi << 32) | b_lo, idesc, 1u);
      if (c.vary_a) { a += (uint32_t)c.a_step16; if ((i & 31) == 31) a = a_lo; }
      if (++acc_i == c.n_acc) acc_i = 0;
    }
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    const long long t2 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int Ns[] = {16, 32, 64, 128, 256};
  printf("layout N vary_a n_acc  issue_cyc/mma  total_cyc/mma  (ideal math N/2)\n");
  for (int layout = 0; layout < 2; ++layout)
    for (int ni = 0; ni < 5; ++ni)
      for (int vary = 0; vary < 2; ++vary)
        for (int nacc = 1; nacc <= 4; nacc += 3) {
          if (Ns[ni] * nacc > 512) continue;
          Cfg c{Ns[ni], layout, vary, nacc, 4096, layout == 0 ? 1 : 64};
          rate_kernel<<<148, 128, 200 * 1024>>>(c, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("%s %3d %d %d  %8.1f  %8.1f  (%d)\n", layout ? "sw128" : "none ", c.N, vary, nacc, h[0] / 4096.0, h[1] / 4096.0, c.N / 2);
        }
  return 0;
}
