// a1: process_features[0] = Conv2d(256, 32, 1) on the backbone's (B,256,h,w) f32 NCHW map
// (network/voxel_net_depth.py:58-63) -> (B,h,w,32) f32 channel-last, on tcgen05 with fp32-grade accuracy.
//
// The op is a GEMM (M = pixels, N = 32, K = 256) that the CUDA-core kernel (csrc/geometry.cu) runs at 0.2 of the HBM
// roofline: 33.5 MFMA per frame on the fp32 pipe behind two barriers per 32 channels.  The parity bar of lift()
// wants fp32 accuracy, so the operands are SPLIT: x = xh + xl, w = wh + wl with xh = bf16(x), xl = bf16(x - xh)
// (likewise w), and x.w ~ xh.wh + xl.wh + xh.wl in three bf16 MMAs with fp32 accumulation -- the dropped xl.wl term
// and the rounding of the low parts are 2^-17 of each product (measured against torch fp32: ~2e-6 of the output range,
// the summation-order noise of the fp32 kernels themselves is 1e-6).  Always bf16 parts, whatever the library's
// activation storage type: an fp16 high part would overflow at 65504.
//
// One CTA = 128 consecutive pixels of one frame, 128 threads, thread p owns pixel p:
//   per quarter of K (64 channels): 64 coalesced loads per thread (512 B per channel and CTA), all in flight before
//   the first is used and before the wait for the previous quarter's MMAs; split, eight 16-byte stores per part into
//   the K-major core-matrix image [k-chunk][128 rows][8] of the A operand; thread 0 issues 4 K-steps x 3 MMAs
//   (M128 N32 K16) into 32 tensor-memory columns and commits; the weights (32 KB as hi / lo images) are split once per
//   CTA.  64 KB of shared memory: three CTAs per SM, ~96 KB of loads in flight per SM -- the kernel is a stream of
//   4-byte gathers and bytes in flight are what it runs on (3.6 us per frame on CUDA cores, 3.0 with 16 loads in
//   flight per thread, 2.3 with 64).  Epilogue: tcgen05.ld, + bias, 128 contiguous bytes per pixel.
#include "tc_common.cuh"
#include <stdlib.h>    // getenv

namespace sceneego {

constexpr int FT_PIX = 128, FT_CIN = 256, FT_CO = 32, FT_KQ = 64;     // K per pipeline stage
constexpr uint32_t FT_A_PART = (FT_KQ / 8) * FT_PIX * 16;             // one part (hi or lo) of one stage: 16 KB
constexpr uint32_t FT_W_PART = (FT_CIN / 8) * FT_CO * 16;             // hi or lo image of the weights: 16 KB
constexpr uint32_t FT_OFF_W = 2 * FT_A_PART;                          // [A hi][A lo][W hi][W lo]: 64 KB, three CTAs per SM
constexpr uint32_t FT_OFF_BIAS = FT_OFF_W + 2 * FT_W_PART;
constexpr uint32_t FT_OFF_BAR = FT_OFF_BIAS + FT_CO * 4;
constexpr uint32_t FT_SMEM = FT_OFF_BAR + 64;

__device__ __forceinline__ void split_bf16(float v, float& hi_f, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16(v);
  hi_f = __bfloat162float(hi);
  lo = __float2bfloat16(v - hi_f);
}
__device__ __forceinline__ uint32_t pack_bf(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

__global__ void __launch_bounds__(FT_PIX, 3) feature_conv1x1_tc_kernel(const float* __restrict__ feat,
                                                                      const float* __restrict__ weight,
                                                                      const float* __restrict__ bias,
                                                                      float* __restrict__ out, int hw) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.y;
  const int pix = blockIdx.x * FT_PIX + tid;
  const bool in = pix < hw;
  float* s_bias = reinterpret_cast<float*>(smem + FT_OFF_BIAS);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + FT_OFF_BAR + 32);
  auto BAR = [&](int i) { return sbase + FT_OFF_BAR + 8u * (uint32_t)i; };      // 0, 1: stage consumed; 2: all MMAs done
  if (tid < FT_CO) s_bias[tid] = bias[tid];
  if (tid == 0) {
    mbar_init(BAR(0), 1); mbar_init(BAR(1), 1); mbar_init(BAR(2), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // weights (32, 256) f32 -> hi / lo images [k-chunk 32][32 rows][8]: 1024 cells, 8 per thread
  for (int cell = tid; cell < (FT_CIN / 8) * FT_CO; cell += FT_PIX) {
    const int kc = cell / FT_CO, co = cell % FT_CO;
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(weight + (size_t)co * FT_CIN + kc * 8));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(weight + (size_t)co * FT_CIN + kc * 8 + 4));
    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    __nv_bfloat16 h[8], l[8];
    float hf;
#pragma unroll
    for (int i = 0; i < 8; ++i) split_bf16(wv[i], hf, h[i], l[i]);
    *reinterpret_cast<uint4*>(smem + FT_OFF_W + (uint32_t)cell * 16u) =
        make_uint4(pack_bf(h[0], h[1]), pack_bf(h[2], h[3]), pack_bf(h[4], h[5]), pack_bf(h[6], h[7]));
    *reinterpret_cast<uint4*>(smem + FT_OFF_W + FT_W_PART + (uint32_t)cell * 16u) =
        make_uint4(pack_bf(l[0], l[1]), pack_bf(l[2], l[3]), pack_bf(l[4], l[5]), pack_bf(l[6], l[7]));
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;
  const float* src = feat + (size_t)b * FT_CIN * hw + (in ? pix : 0);
  // descriptors: K-major no-swizzle, SBO = 128 B between 8-row groups, LBO = distance between the two k-chunks of a K-step
  constexpr uint32_t DHI = 8u | (1u << 14);
  auto DESC = [](uint32_t lo) { return ((uint64_t)DHI << 32) | (uint64_t)lo; };
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((32u >> 3) << 17) | (8u << 24);   // f32 += bf16 x bf16, M128 N32
  constexpr uint32_t A_LBO = ((FT_PIX * 16u) >> 4) << 16, B_LBO = ((FT_CO * 16u) >> 4) << 16;
  for (int q = 0; q < FT_CIN / FT_KQ; ++q) {
    // all 64 loads of the quarter in flight before the first is used (the kernel is a stream of 4-byte gathers: what
    // it needs is bytes in flight -- 64 KB per SM this way), and before waiting for the stage to be free
    float v[FT_KQ];
#pragma unroll
    for (int i = 0; i < FT_KQ; ++i) v[i] = in ? __ldcs(src + (size_t)(q * FT_KQ + i) * hw) : 0.f;
    if (q >= 1) mbar_wait(BAR(0), (uint32_t)(q - 1) & 1u);              // the previous quarter's MMAs have read the A image
    uint8_t* a_hi = smem;
#pragma unroll
    for (int g = 0; g < FT_KQ / 8; ++g) {
      __nv_bfloat16 h[8], l[8];
      float hf;
#pragma unroll
      for (int i = 0; i < 8; ++i) split_bf16(v[g * 8 + i], hf, h[i], l[i]);
      *reinterpret_cast<uint4*>(a_hi + (uint32_t)g * (FT_PIX * 16u) + (uint32_t)tid * 16u) =
          make_uint4(pack_bf(h[0], h[1]), pack_bf(h[2], h[3]), pack_bf(h[4], h[5]), pack_bf(h[6], h[7]));
      *reinterpret_cast<uint4*>(a_hi + FT_A_PART + (uint32_t)g * (FT_PIX * 16u) + (uint32_t)tid * 16u) =
          make_uint4(pack_bf(l[0], l[1]), pack_bf(l[2], l[3]), pack_bf(l[4], l[5]), pack_bf(l[6], l[7]));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> visible to the tensor core
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t ah = ((sbase >> 4) & 0x3FFFu) | A_LBO;
      const uint32_t al = ah + (FT_A_PART >> 4);
      const uint32_t wh = (((sbase + FT_OFF_W) >> 4) & 0x3FFFu) | B_LBO;
      const uint32_t wl = wh + (FT_W_PART >> 4);
#pragma unroll
      for (int ks = 0; ks < FT_KQ / 16; ++ks) {
        const uint32_t ao = (uint32_t)ks * ((2u * FT_PIX * 16u) >> 4);                        // two k-chunks per K-step
        const uint32_t bo = (uint32_t)(q * (FT_KQ / 16) + ks) * ((2u * FT_CO * 16u) >> 4);
        tc_mma_bf16(tmem, DESC(ah + ao), DESC(wh + bo), IDESC, (q | ks) ? 1u : 0u);
        tc_mma_bf16(tmem, DESC(al + ao), DESC(wh + bo), IDESC, 1u);
        tc_mma_bf16(tmem, DESC(ah + ao), DESC(wl + bo), IDESC, 1u);
      }
      tc_commit(BAR(0));
      if (q == FT_CIN / FT_KQ - 1) tc_commit(BAR(2));
    }
  }
  mbar_wait(BAR(2), 0u);
  tc_fence_after();
  uint32_t r[2][16];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  tc_ld16(taddr, r[0]);
  tc_ld16(taddr + 16u, r[1]);
  tc_wait_ld();
  if (in) {
    float4* o = reinterpret_cast<float4*>(out + ((size_t)b * hw + pix) * FT_CO);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o[j] = make_float4(__uint_as_float(r[j >> 2][(j & 3) * 4 + 0]) + s_bias[4 * j + 0],
                         __uint_as_float(r[j >> 2][(j & 3) * 4 + 1]) + s_bias[4 * j + 1],
                         __uint_as_float(r[j >> 2][(j & 3) * 4 + 2]) + s_bias[4 * j + 2],
                         __uint_as_float(r[j >> 2][(j & 3) * 4 + 3]) + s_bias[4 * j + 3]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

}  // namespace sceneego

using namespace sceneego;

extern "C" int sceneego_feature_conv1x1_f32(const float* d_feat, const float* d_weight, const float* d_bias,
                                            float* d_out, int batch, int cin, int cout, int h, int w, void* stream) {
  SE_REQUIRE(d_feat && d_weight && d_bias && d_out, "feature_conv1x1: null argument");
  SE_REQUIRE(batch > 0 && h > 0 && w > 0, "feature_conv1x1: empty input");
  cudaStream_t st = (cudaStream_t)stream;
  const char* e = getenv("SCENEEGO_FEATURE_CONV_SIMT");
  const bool aligned = (((uintptr_t)d_weight) & 15) == 0 && (((uintptr_t)d_out) & 15) == 0;
  if (cin != FT_CIN || cout != FT_CO || !aligned || (e && atoi(e) != 0))
    return launch_feature_conv1x1_simt(d_feat, d_weight, d_bias, d_out, batch, cin, cout, h, w, st);
  if (int rc = ensure_max_dynamic_smem((const void*)feature_conv1x1_tc_kernel, (int)FT_SMEM)) return rc;
  dim3 grid((unsigned)((h * w + FT_PIX - 1) / FT_PIX), (unsigned)batch);
  feature_conv1x1_tc_kernel<<<grid, FT_PIX, FT_SMEM, st>>>(d_feat, d_weight, d_bias, d_out, h * w);
  SE_CUDA_LAUNCH_CHECK("feature_conv1x1_tc");
  return SCENEEGO_OK;
}
