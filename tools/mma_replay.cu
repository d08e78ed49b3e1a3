// Microbenchmark: replay conv_tc_kernel's exact tcgen05.mma descriptor sequence (window shifts,
// tiles, K-steps, weight slots, commits) on static shared memory, without producer or epilogue,
// to separate the operand-fetch cost of the instruction stream from pipeline effects.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_replay tools/mma_replay.cu
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

struct Cfg {
  int N, ksteps, tiles, k, n_dx, pitch_y, items;
  int win_bytes;      // A plane stride inside a stage (LBO of A)
  int tile_rows;      // rows between tiles (128)
  int commit_mode;    // 0 none, 1 per tap, 2 per dx
  int b_vary;         // 1: B start moves per tap (weight slot walk)
  int tap_bytes;
  int nw;            // issuing warps (tiles split round-robin)
  int fill;          // 0: zero operands, 1: pseudo-random bf16 in [-1,1)
};

template <int KSTEPS, int TILES, int NW>
__global__ void __launch_bounds__(128, 1) k(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, dummy;
  __shared__ long long tmax[4];
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < 220 * 1024 / 4; i += blockDim.x) {
    uint32_t v = 0;
    if (c.fill) {
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
      // two bf16 with exponent in [0x3c,0x3f] (|x| in [2^-6, 1)), random sign and mantissa
      const uint32_t lo = ((h & 0x8000u) | (0x3c00u + ((h >> 3) & 0x3ffu))) & 0xffffu;
      const uint32_t hi = (((h >> 16) & 0x8000u) | (0x3c00u + ((h >> 19) & 0x3ffu))) & 0xffffu;
      v = lo | (hi << 16);
    }
    ((uint32_t*)smem)[i] = v;
  }
  const int warp = threadIdx.x >> 5;
  const bool lane0 = (threadIdx.x & 31) == 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(c.nw));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&dummy)));
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
  if (warp < c.nw) {
    const bool leader = elect_one();
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(c.N >> 3) << 17) | (8u << 24);
    const uint32_t sb = smem_u32(smem);
    const uint64_t desc_hi = (uint64_t)(8u | (1u << 14)) << 32;
    const uint32_t a_lo_flags = (((uint32_t)c.win_bytes >> 4) & 0x3FFFu) << 16;
    const uint32_t b_lo_flags = (((uint32_t)c.N * 16u >> 4) & 0x3FFFu) << 16;
    const uint32_t a_ks_step = (2u * (uint32_t)c.win_bytes) >> 4;
    const uint32_t b_ks_step = (2u * (uint32_t)c.N * 16u) >> 4;
    const uint32_t stage_bytes = (uint32_t)c.win_bytes * 2 * c.ksteps;
    const uint32_t w_base = sb + 2 * stage_bytes;
    const uint32_t halo = (uint32_t)(c.k / 2) * (c.pitch_y + 1);
    long long n_mma = 0;
    const long long t0 = clock64();
    for (int it = 0; it < c.items; ++it) {
      const uint32_t d0 = tm + (2 * c.tiles * c.N <= 512 ? (uint32_t)(it & 1) * (uint32_t)(c.tiles * c.N) : 0u);
      uint32_t acc = 0;
      for (int dx = 0; dx < c.n_dx; ++dx) {
        const uint32_t win = sb + (uint32_t)(dx & 1) * stage_bytes;
        uint32_t b_lo = ((w_base >> 4) & 0x3FFFu) | b_lo_flags;
        for (int dy = 0; dy < c.k; ++dy)
          for (int dz = 0; dz < c.k; ++dz) {
            const uint32_t a_lo = (((win >> 4) + halo + (uint32_t)((dy - c.k / 2) * c.pitch_y + (dz - c.k / 2))) & 0x3FFFu) | a_lo_flags;
            if (leader) {
#pragma unroll
              for (int t = 0; t < TILES; ++t)
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks)
                  if (NW == 1 || (t & (NW - 1)) == warp)
                  mma(d0 + (uint32_t)(t * c.N), desc_hi | (a_lo + (uint32_t)(t * c.tile_rows) + (uint32_t)ks * a_ks_step),
                      desc_hi | (b_lo + (uint32_t)ks * b_ks_step), idesc, ks == 0 ? acc : 1u);
            }
            n_mma += c.tiles * c.ksteps / c.nw;
            acc = 1;
            if (c.b_vary) { b_lo += (uint32_t)c.tap_bytes >> 4; if (((b_lo & 0x3FFFu) << 4) + 2 * c.tap_bytes > 220 * 1024) b_lo = ((w_base >> 4) & 0x3FFFu) | b_lo_flags; }
            if (c.commit_mode == 1 && leader)
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy)) : "memory");
          }
        if (c.commit_mode == 2 && leader)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&dummy)) : "memory");
      }
    }
    if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    while (!try_wait(smem_u32(&bar), 0)) {}
    const long long t2 = clock64();
    if (lane0) { tmax[warp] = t2 - t0; if (warp == 0) out[blockIdx.x * 2 + 1] = n_mma * c.nw; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x == 0) { long long m = tmax[0]; for (int w = 1; w < c.nw; ++w) if (tmax[w] > m) m = tmax[w]; out[blockIdx.x * 2] = m; }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

static void run(const char* name, Cfg c, long long* d) {
  void (*fn)(Cfg, long long*) = nullptr;
#define CASE(K, T) if (c.ksteps == K && c.tiles == T) fn = c.nw == 4 ? k<K, T, 4> : c.nw == 2 ? k<K, T, 2> : k<K, T, 1>;
  CASE(2, 4) CASE(2, 2) CASE(2, 1) CASE(2, 8) CASE(3, 4) CASE(4, 4) CASE(8, 2)
  if (!fn) { printf("no instantiation\n"); return; }
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  fn<<<148, 128, 220 * 1024>>>(c, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: error %s\n", name, cudaGetErrorString(e)); exit(1); }
  long long h[296]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  double cyc = 0; for (int b = 0; b < 148; ++b) cyc += h[2 * b]; cyc /= 148;
  printf("%-46s N=%3d ks=%d tiles=%d LBO=%6d commit=%d bvary=%d nw=%d fill=%d : %6.1f cyc/mma\n", name, c.N, c.ksteps, c.tiles, c.win_bytes, c.commit_mode, c.b_vary, c.nw, c.fill, cyc / h[1]);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 16);
  for (int nw = 1; nw <= 4; nw *= 2) {
    run("c32 xs2", Cfg{64, 2, 4, 3, 4, 65, 64, 10304, 128, 1, 1, 4096, nw, 1}, d);
    run("c32 xs2 commit/dx", Cfg{64, 2, 4, 3, 4, 65, 64, 10304, 128, 2, 1, 4096, nw, 1}, d);
    run("c32 xs4 N=128 tiles=4 (single-buffered)", Cfg{128, 2, 4, 3, 6, 65, 64, 10304, 128, 1, 1, 8192, nw, 1}, d);
    run("stem xs4", Cfg{64, 3, 4, 7, 10, 67, 8, 14720, 128, 1, 1, 6144, nw, 1}, d);
    run("stem 2x2x2 N=128 ks=2 tiles=4", Cfg{128, 2, 4, 7, 8, 34, 8, 10432, 128, 1, 1, 8192, nw, 1}, d);
    run("c64 S32", Cfg{64, 4, 4, 3, 3, 33, 64, 9280, 128, 1, 1, 8192, nw, 1}, d);
  }
  return 0;
}
