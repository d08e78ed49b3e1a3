// SURVEY section 8(f) row 1 -- the backbone hand-off: the LAST stage of pose_resnet's deconvolution head
// (ConvTranspose2d(256,256,k4,s2,p1,bias=False) + BatchNorm2d + ReLU, network/pose_resnet.py:203-226,238-240) fused
// with `process_features[0]` (Conv2d(256,32,1), network/voxel_net_depth.py:58-63) on tcgen05: the (B,256,64,64) f32
// feature map of the reference is never written, the volumetric stage receives its (B,64,64,32) channel-last input
// directly and `feature_conv1x1_kernel` (268 MB of f32 reads per 64 frames) drops out of `forward()`.
//
// A stride-2 k4 transposed conv is four 2x2 convolutions, one per output parity (a,b): output pixel (2Y+a, 2X+b)
// reads input rows {Y, Y-1} (a = 0: ky = 1, 3) or {Y+1, Y} (a = 1: ky = 0, 2), columns alike.  With the input in a
// zero-bordered planar layout (cells of 8 channels, pitch w+2) every tap is a constant position offset, so GEMM rows
// are 128 consecutive positions and the A operand of a tap is the staged window at a shifted start address -- the
// same construction as the 3-D kernels.  Per 128-row tile and parity:
//     GEMM 1: 4 taps x 16 K-steps of N256 K16 (weights streamed from L2, BN scale folded in)   -> 256 TMEM columns
//     epilogue 1: + BN shift, ReLU, bf16 -> the 128 x 256 hidden tile, packed back into the tensor-memory columns just read
//     GEMM 2: 16 K-steps of N32 K16, A operand FROM TENSOR MEMORY, against the resident 1x1 weights -> 32 drained columns
//     epilogue 2: + bias, f32 -> out[b][2Y+a][2X+b][0..31] (128 contiguous bytes per pixel)
// What the round-2 form does, and why (the first form -- hidden tile in shared memory, five 8 KB weight slots, the four
// phases of a parity back to back -- ran at 0.38 of the cuBLAS burst peak; history in profiles/r02_ncu_handoff_b64.txt):
//   * the hidden tile never touches shared memory (tcgen05.st over the accumulator half just drained; GEMM 2 reads it
//     as its A operand); the 64 KB this frees go to the weight ring (6 x 16 KB in flight against a ~1.7k-cycle L2
//     round trip at 64 B/clk per SM);
//   * the two 256-column halves of tensor memory alternate between parities, and GEMM 2 of parity k-1 is issued in
//     the middle of GEMM 1 of parity k, so both epilogues run under the next parity's MMAs;
//   * CTAs run in clusters of two and each loads HALF of every weight chunk, multicast into both (cp.async.bulk
//     .multicast::cluster): half the L2 reads, same speed (SCENEEGO_HANDOFF_MULTICAST=0 turns it off);
//   * the issuing warp's loop is registers only -- no kernel-parameter loads, K unrolled, one wait and one commit per
//     two MMAs: at ~55 instructions per MMA it, not the memory system, paced the first forms at ~350 cycles per MMA.
//   (CTA-pair MMAs -- cta_group::2, weights N-split -- do not work here: every pair instruction, MMA or commit, keeps
//   its issuing warp for ~222 cycles, and one accumulator per CTA leaves room for one issuer only.)
//   * the window is loaded in two K halves and the K loop of a parity runs half-major, so the next tile's window
//     lands under MMAs of the current one (no bubble between tiles).
// warp 0 = producer (cp.async.bulk: the 32-plane window once per tile, weight chunks), warp 1 = MMA issuer,
// warps 2..5 = epilogue (the four tensor-memory lane quarters).
#include "tc_common.cuh"
#include <math.h>
#include <stdlib.h>    // getenv

namespace sceneego {

constexpr int HO_CIN = 256, HO_MID = 256, HO_OUT = 32;
constexpr int HO_PLANES = HO_CIN / 8;                        // 32 channel-group planes
constexpr int HO_THREADS = 32 * 6;
constexpr int HO_KSTEPS = HO_CIN / 16;
constexpr int HO_KSTEP_BYTES = 2 * HO_MID * 16;              // one K-step of GEMM 1: [2 k-chunks][256 rows][8] = 8 KB
constexpr int HO_CHUNK_KSTEPS = 2;                           // K-steps per ring slot: one wait + one commit per two MMAs
constexpr int HO_CHUNK_BYTES = HO_CHUNK_KSTEPS * HO_KSTEP_BYTES;         // 16 KB
constexpr int HO_CHUNKS = 4 * 4 * HO_KSTEPS / HO_CHUNK_KSTEPS;           // parity x tap x K-step pairs
constexpr int HO_W_SLOTS = 6;                                // 96 KB in flight
constexpr int HO_W1_BYTES = HO_CHUNKS * HO_CHUNK_BYTES;                  // 2 MB
constexpr int HO_W2_BYTES = (HO_MID / 16) * 2 * HO_OUT * 16;             // [K-step 16][2][32 rows][8] = 16 KB
constexpr int HO_GUARD = 40;
constexpr int HO_BARRIERS = 4 + 2 * HO_W_SLOTS + 2 + 2 + 1 + 1;          // window halves, weight ring, accumulator halves, hidden tile, GEMM 2

struct HandoffParams {
  const __nv_bfloat16* x;     // planar zero-bordered input, HO_PLANES planes of plane_cells cells
  const uint8_t* w;           // [W1 2 MB][W2 16 KB]
  const float* bias;          // [256 BN shift][32 conv bias]
  float* out;                 // (B, 2h, 2w, 32) f32
  int batch, h, w_in, pitch, frame_cells, n_items, halo, win_cells;
  int64_t plane_cells;
  uint32_t win_plane_bytes;
  uint32_t off_w2, off_ring, off_bias, off_bar;
  int multicast;              // 1: clusters of two CTAs, each loads half of every weight chunk for both
  FastDiv fd_frame, fd_pitch;
};

__host__ __device__ inline int ho_tap_offset(int parity_bit, int t) {   // input offset of tap t (0,1) for output parity bit
  return parity_bit == 0 ? (t == 0 ? 0 : -1) : (t == 0 ? 1 : 0);
}
__host__ __device__ inline int ho_tap_k(int parity_bit, int t) {        // kernel index of that tap
  return parity_bit == 0 ? (t == 0 ? 1 : 3) : (t == 0 ? 0 : 2);
}

// (B,256,h,w) f32 NCHW -> planar bf16 with a zero border (and zero guards): one thread per cell
__global__ void __launch_bounds__(256) handoff_pack_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int batch,
                                                          int h, int w, int pitch, int frame_cells, int64_t plane_cells) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int g = blockIdx.y;
  if (q >= plane_cells) return;
  __align__(16) __nv_bfloat16 cell[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) cell[i] = act_from_float(0.f);
  const int64_t r = q - HO_GUARD;
  if (r >= 0) {
    const int b = (int)(r / frame_cells);
    const int rem = (int)(r - (int64_t)b * frame_cells);
    const int yp = rem / pitch, xp = rem - yp * pitch;
    if (b < batch && yp >= 1 && yp <= h && xp >= 1 && xp <= w) {
      const float* src = x + (((size_t)b * HO_CIN + 8 * g) * h + (yp - 1)) * w + (xp - 1);
#pragma unroll
      for (int i = 0; i < 8; ++i) cell[i] = act_from_float(__ldg(src + (size_t)i * h * w));
    }
  }
  *reinterpret_cast<uint4*>(out + ((int64_t)g * plane_cells + q) * 8) = *reinterpret_cast<const uint4*>(cell);
}

template <bool MC>
__global__ void __launch_bounds__(HO_THREADS, 1) handoff_tc_kernel(const __grid_constant__ HandoffParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  float* s_bias = reinterpret_cast<float*>(smem + p.off_bias);
  const uint32_t bar0 = sbase + p.off_bar;
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int B_WIN_FULL = 0, B_WIN_EMPTY = 2, B_W_FULL = 4, B_W_EMPTY = B_W_FULL + HO_W_SLOTS, B_TM1_FULL = B_W_EMPTY + HO_W_SLOTS,
                B_TM_EMPTY = B_TM1_FULL + 2, B_H_FULL = B_TM_EMPTY + 2, B_TM2_FULL = B_H_FULL + 1, B_COUNT = B_TM2_FULL + 1;
  static_assert(B_COUNT == HO_BARRIERS, "barrier map and shared-memory plan disagree");
  uint32_t* s_tmem_ptr = reinterpret_cast<uint32_t*>(smem + p.off_bar + 8 * B_COUNT);
  constexpr bool mc = MC;
  const uint32_t cta_rank = cluster_ctarank();                     // the launch is always in clusters of two

  for (int i = threadIdx.x; i < HO_MID + HO_OUT; i += HO_THREADS) s_bias[i] = p.bias[i];
  for (int i = threadIdx.x; i < HO_W2_BYTES / 16; i += HO_THREADS)      // the 1x1 weights stay resident
    reinterpret_cast<uint4*>(smem + p.off_w2)[i] = reinterpret_cast<const uint4*>(p.w + HO_W1_BYTES)[i];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(B_WIN_FULL + i), 1); mbar_init(BAR(B_WIN_EMPTY + i), 1); }
    // multicast: a slot is free when the issuers of BOTH CTAs are done with it (each commit arrives in both CTAs)
    for (int i = 0; i < HO_W_SLOTS; ++i) { mbar_init(BAR(B_W_FULL + i), 1); mbar_init(BAR(B_W_EMPTY + i), mc ? 2 : 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(B_TM1_FULL + i), 1); mbar_init(BAR(B_TM_EMPTY + i), 4); }
    mbar_init(BAR(B_H_FULL), 4);
    mbar_init(BAR(B_TM2_FULL), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // W2 was written with generic stores
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                               // the peer's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem_ptr;
  // work split: cluster c of n_cl takes item pairs c, c + n_cl, ...; CTA r takes item 2 * pair + r.  Both CTAs of a
  // cluster run the same number of items (the weight ring couples them); an odd item count repeats the last item in
  // CTA 1, which then skips its stores.
  const int n_cl = (int)gridDim.x / 2, cl = (int)blockIdx.x / 2;
  const int n_groups = (p.n_items + 1) / 2;
  const int my_items = (n_groups - cl + n_cl - 1) / n_cl;
  auto raw_item = [&](int it) -> int { return (cl + it * n_cl) * 2 + (int)cta_rank; };

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      int sl = 0, sph = 0;
      for (int it = 0; it < my_items; ++it) {
        const int item = min(raw_item(it), p.n_items - 1);
        const int64_t q0 = (int64_t)HO_GUARD + (int64_t)item * 128 - p.halo;
        // the window in two K halves (channel-group planes 0..15 / 16..31): half 0 is free 32 MMAs before the tile
        // ends, so the next tile's half 0 lands under them; half 1 follows behind the first weight chunks of the new
        // tile and lands under its first 32 MMAs, which read half 0 only -- no window bubble between tiles
        auto load_half = [&](int hf) {
          mbar_wait(BAR(B_WIN_EMPTY + hf), (it & 1) ^ 1);
          mbar_expect_tx(BAR(B_WIN_FULL + hf), p.win_plane_bytes * (uint32_t)(HO_PLANES / 2));
          for (int g = hf * (HO_PLANES / 2); g < (hf + 1) * (HO_PLANES / 2); ++g)
            bulk_g2s(sbase + (uint32_t)g * p.win_plane_bytes, p.x + ((int64_t)g * p.plane_cells + q0) * 8, p.win_plane_bytes, BAR(B_WIN_FULL + hf));
        };
        load_half(0);
        for (int c = 0; c < HO_CHUNKS; ++c) {
          if (c == 8) load_half(1);
          mbar_wait(BAR(B_W_EMPTY + sl), sph ^ 1);
          mbar_expect_tx(BAR(B_W_FULL + sl), HO_CHUNK_BYTES);      // my half and the peer's half both land here
          const uint32_t dst = sbase + p.off_ring + (uint32_t)sl * HO_CHUNK_BYTES;
          const uint8_t* src = p.w + (size_t)c * HO_CHUNK_BYTES;
          if (mc) bulk_g2s_mc(dst + cta_rank * (HO_CHUNK_BYTES / 2), src + cta_rank * (HO_CHUNK_BYTES / 2), HO_CHUNK_BYTES / 2, BAR(B_W_FULL + sl), 3);
          else bulk_g2s(dst, src, HO_CHUNK_BYTES, BAR(B_W_FULL + sl));
          if (++sl == HO_W_SLOTS) { sl = 0; sph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const bool leader = elect_one();
    const uint64_t desc_hi = (uint64_t)(8u | (1u << 14)) << 32;            // SBO = 128 B, descriptor version 1
    constexpr uint32_t idesc0 = (1u << 4) | kIdescAB | (8u << 24);
    constexpr uint32_t ID256 = idesc0 | ((256u >> 3) << 17), ID32 = idesc0 | ((32u >> 3) << 17);
    const uint32_t win_plane16 = p.win_plane_bytes >> 4;
    const uint32_t a_lbo = (win_plane16 & 0x3FFFu) << 16;                  // K chunk 1 = the next channel-group plane
    constexpr uint32_t b1_lbo = (uint32_t)HO_MID << 16;                    // [k-chunk][256 rows][8]
    constexpr uint32_t b2_lbo = (uint32_t)HO_OUT << 16;                    // [k-chunk][32 rows][8]
    const uint32_t w216 = ((sbase + p.off_w2) >> 4) & 0x3FFFu;
    // GEMM 2 of step j: A = the bf16 hidden tile the epilogue packed into the first 128 columns of step j's (drained)
    // accumulator half, in tensor memory; B = the resident 1x1 weights; D = columns 128..159 of the same half
    auto gemm2 = [&](uint32_t j) {
      mbar_wait_warp(BAR(B_H_FULL), j & 1u);
      tc_fence_after();
      if (leader) {
        const uint32_t tb = tmem_u + (j & 1u) * 256u;
#pragma unroll
        for (int ks = 0; ks < HO_MID / 16; ++ks)
          tc_mma_bf16_ts(tb + 128u, tb + (uint32_t)(8 * ks), desc_hi | (uint64_t)((w216 + (uint32_t)ks * (2u * HO_OUT)) | b2_lbo), ID32,
                         ks == 0 ? 0u : 1u);
        tc_commit(BAR(B_TM2_FULL));
      }
    };
    // The loop below runs once per two MMAs and is all this warp does: everything in it is a register (no kernel-
    // parameter loads, no divisions), the K loop is unrolled, descriptors advance by adds on their low words.
    const uint32_t pitch = (uint32_t)p.pitch;
    const uint32_t a_base = (((sbase >> 4) + (uint32_t)p.halo) & 0x3FFFu) | a_lbo;         // + tap offset + 2 * ks * plane stride
    const uint32_t ring16 = (((sbase + p.off_ring) >> 4) & 0x3FFFu) | b1_lbo;
    const uint32_t w_full0 = BAR(B_W_FULL), w_empty0 = BAR(B_W_EMPTY);
    const uint32_t a_kstep = 2u * win_plane16;
    uint32_t sl = 0, sph = 0, k = 0;
    for (int it = 0; it < my_items; ++it) {
      for (int par = 0; par < 4; ++par, ++k) {
        const uint32_t buf = k & 1u, use = k >> 1;
        if (k >= 2) mbar_wait_warp(BAR(B_TM_EMPTY + (int)buf), (use - 1u) & 1u);
        tc_fence_after();
        const int pa = par >> 1, pb = par & 1;
        const uint32_t d_acc = tmem_u + buf * 256u;
        // K order of a parity: window half 0 (K-steps 0..7) for the four taps, then half 1 (weights packed alike)
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
          if (hf == 1 && k > 0) gemm2(k - 1);                              // under this parity's MMAs
          if (par == 0) mbar_wait_warp(BAR(B_WIN_FULL + hf), (uint32_t)(it & 1));
#pragma unroll 1
          for (int t = 0; t < 4; ++t) {
            const uint32_t a_tap = a_base + (uint32_t)(ho_tap_offset(pa, t >> 1) * (int)pitch + ho_tap_offset(pb, t & 1)) +
                                   (uint32_t)hf * (uint32_t)(HO_KSTEPS / 2) * a_kstep;
#pragma unroll
            for (int c = 0; c < HO_KSTEPS / 2 / HO_CHUNK_KSTEPS; ++c) {
              mbar_wait_warp(w_full0 + 8u * sl, sph);
              tc_fence_after();
              if (leader) {
                const uint32_t b_c = ring16 + sl * (uint32_t)(HO_CHUNK_BYTES >> 4);
#pragma unroll
                for (int j = 0; j < HO_CHUNK_KSTEPS; ++j)
                  tc_mma_bf16(d_acc, desc_hi | (uint64_t)(a_tap + (uint32_t)(c * HO_CHUNK_KSTEPS + j) * a_kstep),
                              desc_hi | (uint64_t)(b_c + (uint32_t)j * (uint32_t)(HO_KSTEP_BYTES >> 4)), ID256, (hf | t | c | j) ? 1u : 0u);
                if (mc) tc_commit_mc(w_empty0 + 8u * sl, 3); else tc_commit(w_empty0 + 8u * sl);
              }
              if (++sl == (uint32_t)HO_W_SLOTS) { sl = 0; sph ^= 1u; }
            }
          }
          if (par == 3 && leader) tc_commit(BAR(B_WIN_EMPTY + hf));          // this half of the window is done for the tile
        }
        if (leader) tc_commit(BAR(B_TM1_FULL + (int)buf));
      }
    }
    if (k > 0) gemm2(k - 1);
    __syncwarp();
  } else {
    // ===================== epilogue: warps 2..5 = tensor-memory lane quarters 2, 3, 0, 1 =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t k = 0;
    for (int it = 0; it < my_items; ++it) {
      const int item = raw_item(it);
      const uint32_t r = (uint32_t)(min(item, p.n_items - 1) * 128 + row);    // position relative to the first frame
      const uint32_t b = fdiv(r, p.fd_frame);
      const uint32_t rem = r - b * (uint32_t)p.frame_cells;
      const int yp = (int)fdiv(rem, p.fd_pitch), xp = (int)rem - yp * p.pitch;
      const bool valid = item < p.n_items && (int)b < p.batch && yp >= 1 && yp <= p.h && xp >= 1 && xp <= p.w_in;
      for (int par = 0; par < 4; ++par, ++k) {
        const uint32_t buf = k & 1u, use = k >> 1;
        const uint32_t tbuf = taddr + buf * 256u;
        mbar_wait(BAR(B_TM1_FULL + (int)buf), use & 1u);
        tc_fence_after();
        // epilogue 1: 32 accumulator columns -> + BN shift, ReLU -> 16 packed columns written back over columns already read
#pragma unroll 1
        for (int c = 0; c < HO_MID / 32; ++c) {
          uint32_t raw[2][16];
          tc_ld16(tbuf + (uint32_t)(32 * c), raw[0]);
          tc_ld16(tbuf + (uint32_t)(32 * c + 16), raw[1]);
          tc_wait_ld();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float lo = fmaxf(__uint_as_float(raw[j >> 3][2 * (j & 7)]) + s_bias[32 * c + 2 * j], 0.f);
            const float hi = fmaxf(__uint_as_float(raw[j >> 3][2 * (j & 7) + 1]) + s_bias[32 * c + 2 * j + 1], 0.f);
            pk[j] = act_pack2(lo, hi);
          }
          tc_st16(tbuf + (uint32_t)(16 * c), pk);
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_H_FULL));
        mbar_wait(BAR(B_TM2_FULL), k & 1u);
        tc_fence_after();
        uint32_t r2[2][16];
        tc_ld16(tbuf + 128u, r2[0]);
        tc_ld16(tbuf + 144u, r2[1]);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_TM_EMPTY + (int)buf));
        if (valid) {
          const int oy = 2 * (yp - 1) + (par >> 1), ox = 2 * (xp - 1) + (par & 1);
          float4* o = reinterpret_cast<float4*>(p.out + (((size_t)b * (2 * p.h) + oy) * (size_t)(2 * p.w_in) + ox) * HO_OUT);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            o[q] = make_float4(__uint_as_float(r2[q >> 2][(q & 3) * 4 + 0]) + s_bias[HO_MID + 4 * q + 0],
                               __uint_as_float(r2[q >> 2][(q & 3) * 4 + 1]) + s_bias[HO_MID + 4 * q + 1],
                               __uint_as_float(r2[q >> 2][(q & 3) * 4 + 2]) + s_bias[HO_MID + 4 * q + 2],
                               __uint_as_float(r2[q >> 2][(q & 3) * 4 + 3]) + s_bias[HO_MID + 4 * q + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                 // the peer may still multicast into this CTA's ring / arrive on its barriers until it is done too
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

static void handoff_geometry(int batch, int h, int w, int& pitch, int& frame_cells, int64_t& plane_cells) {
  pitch = w + 2;
  frame_cells = ((h + 2) * pitch + 7) / 8 * 8;
  plane_cells = (int64_t)HO_GUARD + (int64_t)batch * frame_cells + HO_GUARD + 256;      // trailing guard + the last tile's overhang
  plane_cells = (plane_cells + 7) / 8 * 8;
}

}  // namespace sceneego

using namespace sceneego;

extern "C" size_t sceneego_handoff_weight_bytes(void) { return (size_t)HO_W1_BYTES + HO_W2_BYTES; }

extern "C" size_t sceneego_handoff_workspace_bytes(int batch, int h, int w) {
  if (batch <= 0 || h <= 0 || w <= 0) return 0;
  int pitch, fc;
  int64_t pc;
  handoff_geometry(batch, h, w, pitch, fc, pc);
  return (size_t)pc * HO_PLANES * 16;
}

extern "C" int sceneego_handoff_pack(const float* h_deconv_w, const float* h_gamma, const float* h_beta, const float* h_mean,
                                     const float* h_var, double eps, const float* h_conv_w, const float* h_conv_b,
                                     uint16_t* h_w_out, float* h_b_out) {
  SE_REQUIRE(h_deconv_w && h_gamma && h_beta && h_mean && h_var && h_conv_w && h_w_out && h_b_out, "handoff_pack: null argument");
  // ConvTranspose2d weight (cin 256, cout 256, 4, 4); BatchNorm folded: scale into the weights, shift as the bias
  double scale[HO_MID];
  for (int co = 0; co < HO_MID; ++co) {
    scale[co] = (double)h_gamma[co] / sqrt((double)h_var[co] + eps);
    h_b_out[co] = (float)((double)h_beta[co] - (double)h_mean[co] * scale[co]);
  }
  for (int o = 0; o < HO_OUT; ++o) h_b_out[HO_MID + o] = h_conv_b ? h_conv_b[o] : 0.f;
  // GEMM 1 weights in the order the kernel consumes them: [parity][window half][tap][K-step of the half][k-chunk 2][256 rows][8]
  size_t kstep_slot = 0;
  for (int par = 0; par < 4; ++par)
    for (int hf = 0; hf < 2; ++hf)
      for (int t = 0; t < 4; ++t) {
        const int ky = ho_tap_k(par >> 1, t >> 1), kx = ho_tap_k(par & 1, t & 1);
        for (int kh = 0; kh < HO_KSTEPS / 2; ++kh, ++kstep_slot) {
          const int ks = hf * (HO_KSTEPS / 2) + kh;
          uint16_t* dst = h_w_out + kstep_slot * (HO_KSTEP_BYTES / 2);
          for (int c = 0; c < 2; ++c)
            for (int co = 0; co < HO_MID; ++co)
              for (int e = 0; e < 8; ++e) {
                const int ci = ks * 16 + c * 8 + e;
                dst[((size_t)c * HO_MID + co) * 8 + e] =
                    f2bf((float)((double)h_deconv_w[(((size_t)ci * HO_MID + co) * 4 + ky) * 4 + kx] * scale[co]));
              }
        }
      }
  uint16_t* w2 = h_w_out + HO_W1_BYTES / 2;                  // Conv2d weight (32, 256): [K-step][k-chunk][32 rows][8]
  for (int ks = 0; ks < HO_MID / 16; ++ks)
    for (int c = 0; c < 2; ++c)
      for (int o = 0; o < HO_OUT; ++o)
        for (int e = 0; e < 8; ++e)
          w2[(((size_t)ks * 2 + c) * HO_OUT + o) * 8 + e] = f2bf(h_conv_w[(size_t)o * HO_MID + ks * 16 + c * 8 + e]);
  return SCENEEGO_OK;
}

extern "C" int sceneego_backbone_handoff_f32(const float* d_x, int batch, int cin, int h, int w, const void* d_weights,
                                             const float* d_bias, void* d_workspace, float* d_out, void* stream) {
  SE_REQUIRE(d_x && d_weights && d_bias && d_workspace && d_out, "backbone_handoff: null argument");
  SE_REQUIRE(cin == HO_CIN && batch > 0 && h > 0 && w > 0, "backbone_handoff: expects (B,256,h,w) input");
  HandoffParams p;
  memset(&p, 0, sizeof(p));
  handoff_geometry(batch, h, w, p.pitch, p.frame_cells, p.plane_cells);
  p.x = (const __nv_bfloat16*)d_workspace;
  p.w = (const uint8_t*)d_weights;
  p.bias = d_bias;
  p.out = d_out;
  p.batch = batch; p.h = h; p.w_in = w;
  { const char* e = getenv("SCENEEGO_HANDOFF_MULTICAST"); p.multicast = e ? (atoi(e) != 0) : 1; }
  p.halo = p.pitch + 1;
  p.win_cells = (128 + 2 * p.halo + 7) / 8 * 8;
  p.win_plane_bytes = (uint32_t)p.win_cells * 16u;
  SE_REQUIRE(p.halo <= HO_GUARD, "backbone_handoff: input wider than 38 columns needs a larger guard");
  p.n_items = (int)(((int64_t)batch * p.frame_cells + 127) / 128);
  SE_REQUIRE((int64_t)batch * p.frame_cells + 4096 < (1ll << 31), "backbone_handoff: batch too large for one launch");
  p.off_w2 = (uint32_t)HO_PLANES * p.win_plane_bytes;
  p.off_ring = p.off_w2 + HO_W2_BYTES;
  p.off_bias = p.off_ring + HO_W_SLOTS * HO_CHUNK_BYTES;
  p.off_bar = p.off_bias + (HO_MID + HO_OUT) * 4;
  const size_t smem_bytes = (size_t)p.off_bar + 8 * HO_BARRIERS + 64;
  SE_REQUIRE(smem_bytes <= kMaxSmem, "backbone_handoff: shared-memory plan exceeds 227 KB");
  p.fd_frame = make_fastdiv((uint32_t)p.frame_cells);
  p.fd_pitch = make_fastdiv((uint32_t)p.pitch);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 pg((unsigned)((p.plane_cells + 255) / 256), HO_PLANES);
  handoff_pack_kernel<<<pg, 256, 0, st>>>(d_x, (__nv_bfloat16*)d_workspace, batch, h, w, p.pitch, p.frame_cells, p.plane_cells);
  SE_CUDA_LAUNCH_CHECK("handoff_pack");
  void (*fn)(const HandoffParams) = p.multicast ? handoff_tc_kernel<true> : handoff_tc_kernel<false>;
  if (int rc = ensure_max_dynamic_smem((const void*)fn, (int)kMaxSmem)) return rc;
  // clusters of two CTAs (148 = 2 x 74 packs the chip exactly)
  const int groups = (p.n_items + 1) / 2;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(2 * (groups < kNumSMs / 2 ? groups : kNumSMs / 2)));
  cfg.blockDim = dim3(HO_THREADS);
  cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, fn, p);
  if (e != cudaSuccess) { set_error("handoff_tc: %s", cudaGetErrorString(e)); return SCENEEGO_E_CUDA; }
  SE_CUDA_LAUNCH_CHECK("handoff_tc");
  return SCENEEGO_OK;
}
