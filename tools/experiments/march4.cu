// a7: the 3^3 convolutions with 32 output channels at full resolution (ten of V2V's layers, network/v2v.py:21-43,
// 147-156) as an x-marching banded GEMM with ROTATED weight rows -- the successor of csrc/march.cu, built like the
// marching stem (csrc/stem_march.cu).
//
// What round 2's ncu capture of march.cu showed (profiles/r02_ncu_march_b64.txt): its tensor pipe is 72 % active
// against a 74.6 % ceiling for its MMA shapes -- N = 96 costs 56 cycles for 48 of math, and on two planes of eight
// the band straddles the end of the 8-slot ring and is split into N = 32 + N = 64 (89.5 cycles); one issuing warp per
// CTA with only 18 MMAs per plane cannot hide anything.  Here:
//   * the ring is FOUR slots (32 columns each) per tile and every plane is issued as N = 128 over the whole ring with
//     the weight rows rotated to the plane's ring position: row block s holds W[dx = 2 - j], j = (s - r) mod 4, zeros
//     for the one block outside the band.  N = 128 runs at the tensor pipe's own rate (64 cycles, 75 % of them
//     useful MACs), never splits, and every MMA of the layer has the same shape;
//   * an item is four neighbouring tiles (512 cells) sharing the staged window and the streamed weight chunks, one
//     MMA-issuing warp per tile (their barrier bookkeeping overlaps the other tiles' MMAs), eight epilogue warps;
//   * the idle block accumulates zeros into the slot the next plane opens, so the issuer waits for the drain of
//     output x-2 before plane x (hidden behind the other three tiles);
//   * the four rotations of the weights (4 x 72 KB for 32 -> 32) stream from L2 in (dy) chunks of 3 x KSTEPS taps.
// The fused 1x1 projection shortcut (Res3DBlock.skip_con) is one N = 32 MMA per plane into the plane's own slot from a
// halo-free window of the second source.  Outputs -1 and S of a march are accumulated and drained without a store.
#include "tc_common.cuh"
#include <stdlib.h>
#include <math.h>

namespace sceneego {

constexpr int M4_TILES = 4;
constexpr int M4_L = 128 * M4_TILES;
constexpr int M4_RING = 4;
constexpr int M4_N0 = 32;                                  // output channels
constexpr int M4_THREADS = 32 * (2 + M4_TILES + 8);       // 448
constexpr int M4_MMA_B_BYTES = 2 * 128 * 16;              // [2 k-chunks][128 rows][8] bf16
constexpr int M4_MAX_PSLOTS = 4, M4_MAX_WSLOTS = 6;

struct March4Params {
  const __nv_bfloat16* src;
  const __nv_bfloat16* src2;   // fused shortcut source (2 planes) or nullptr
  const __nv_bfloat16* res;
  __nv_bfloat16* dst;
  const uint8_t* w;            // [rotation 4][dy 3][dz 3][k-step][2][128][8], then (shortcut) [2][32][8]
  const float* bias;
  sceneego_vol_layout_t ls, ld;
  int batch, flags;
  int cin_planes, cin2_planes;
  int groups_per_frame, n_items, cells_per_plane;
  int n_seg, seg_planes;
  int halo, win_cells;
  int p_slots, w_slots;
  uint32_t win_bytes, win2_bytes, stage_bytes, chunk_bytes, rot_bytes;
  uint32_t off_w, off_w2, off_bias, off_bar;
  FastDiv fd_gpf, fd_py, fd_seg;
};

template <int KSTEPS, int KSTEPS2>
__global__ void __launch_bounds__(M4_THREADS, 1) conv_march4_kernel(const __grid_constant__ March4Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  float* s_bias = reinterpret_cast<float*>(smem + p.off_bias);
  const uint32_t bar0 = sbase + p.off_bar;
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int B_PFULL = 0, B_PEMPTY = B_PFULL + M4_MAX_PSLOTS, B_WFULL = B_PEMPTY + M4_MAX_PSLOTS,
                B_WEMPTY = B_WFULL + M4_MAX_WSLOTS, B_ACC_FULL = B_WEMPTY + M4_MAX_WSLOTS,
                B_ACC_EMPTY = B_ACC_FULL + M4_TILES * M4_RING, B_COUNT = B_ACC_EMPTY + M4_TILES * M4_RING;
  uint32_t* s_tmem_ptr = reinterpret_cast<uint32_t*>(smem + p.off_bar + 8 * B_COUNT);
  constexpr uint32_t CHUNK_BYTES = 3u * KSTEPS * M4_MMA_B_BYTES;            // one dy row of taps

  if (threadIdx.x < M4_N0) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  if (KSTEPS2 > 0)                                                          // the shortcut's 1 KB of weights stay resident
    for (int i = threadIdx.x; i < KSTEPS2 * 2 * M4_N0; i += M4_THREADS)
      reinterpret_cast<uint4*>(smem + p.off_w2)[i] = reinterpret_cast<const uint4*>(p.w + 4u * p.rot_bytes)[i];
  if (threadIdx.x == 0) {
    for (int i = 0; i < M4_MAX_PSLOTS; ++i) { mbar_init(BAR(B_PFULL + i), 1); mbar_init(BAR(B_PEMPTY + i), M4_TILES); }
    for (int i = 0; i < M4_MAX_WSLOTS; ++i) { mbar_init(BAR(B_WFULL + i), 1); mbar_init(BAR(B_WEMPTY + i), M4_TILES); }
    for (int i = 0; i < M4_TILES * M4_RING; ++i) { mbar_init(BAR(B_ACC_FULL + i), 1); mbar_init(BAR(B_ACC_EMPTY + i), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem_ptr;
  if (warp >= 2 + M4_TILES && warp < 2 + M4_TILES + 4) {       // every accumulator slot starts cleared
    const uint32_t t0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (uint32_t c = 0; c < 512; c += 16) tc_st16_zero(t0 + c);
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int S = p.ls.side;
  const int my_items = ((int)p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  // item = (frame, group of four tiles, x-segment): stores outputs [x0, x1), marches over the input planes
  // [xa, xb] = [x0 - 1, x1] clipped to the volume, drains the outputs xa - 1 .. xb + 1 in order
  auto item_of = [&](int it, int& b, int& cell0, int& n_act, int& x0, int& x1, int& xa, int& xb) {
    const uint32_t item = blockIdx.x + (uint32_t)it * gridDim.x;
    const uint32_t bg = fdiv(item, p.fd_seg);
    const int seg = (int)(item - bg * (uint32_t)p.n_seg);
    b = (int)fdiv(bg, p.fd_gpf);
    cell0 = (int)(bg - (uint32_t)b * (uint32_t)p.groups_per_frame) * M4_L;
    const int left = p.cells_per_plane - cell0;
    n_act = left >= M4_L ? M4_TILES : (left + 127) / 128;
    x0 = seg * p.seg_planes;
    x1 = x0 + p.seg_planes < S ? x0 + p.seg_planes : S;
    xa = x0 - 1 > 0 ? x0 - 1 : 0;
    xb = x1 < S - 1 ? x1 : S - 1;
  };
  const int pitch_y = p.ls.pitch_y;

  if (warp == 0) {
    // ===================== weight producer =====================
    if (lane == 0) {
      int sl = 0, sph = 0;
      uint32_t G = 0;
      for (int it = 0; it < my_items; ++it) {
        int b, cell0, n_act, x0, x1, xa, xb;
        item_of(it, b, cell0, n_act, x0, x1, xa, xb);
        for (int x = xa; x <= xb; ++x) {
          const uint8_t* wr = p.w + (size_t)((G + (uint32_t)(x - xa)) & 3u) * p.rot_bytes;
          for (int c = 0; c < 3; ++c) {
            mbar_wait(BAR(B_WEMPTY + sl), sph ^ 1);
            mbar_expect_tx(BAR(B_WFULL + sl), CHUNK_BYTES);
            const uint32_t dst = sbase + p.off_w + (uint32_t)sl * CHUNK_BYTES;
            for (uint32_t o = 0; o < CHUNK_BYTES; o += 12288u)
              bulk_g2s(dst + o, wr + (size_t)c * CHUNK_BYTES + o, CHUNK_BYTES - o < 12288u ? CHUNK_BYTES - o : 12288u, BAR(B_WFULL + sl));
            if (++sl == p.w_slots) { sl = 0; sph ^= 1; }
          }
        }
        G += (uint32_t)(xb - xa + 1 + 2);
      }
    }
  } else if (warp == 1) {
    // ===================== window producer =====================
    if (lane == 0) {
      int ps = 0, pph = 0;
      for (int it = 0; it < my_items; ++it) {
        int b, cell0, n_act, x0, x1, xa, xb;
        item_of(it, b, cell0, n_act, x0, x1, xa, xb);
        const int64_t q0 = (int64_t)b * p.ls.frame_pitch + p.ls.guard + cell0;
        for (int x = xa; x <= xb; ++x) {
          const int64_t qx = q0 + (int64_t)x * p.ls.pitch_x;
          mbar_wait(BAR(B_PEMPTY + ps), pph ^ 1);
          mbar_expect_tx(BAR(B_PFULL + ps), p.stage_bytes);
          const uint32_t dst0 = sbase + (uint32_t)ps * p.stage_bytes;
#pragma unroll
          for (int g = 0; g < 2 * KSTEPS; ++g)
            bulk_g2s(dst0 + (uint32_t)g * p.win_bytes, p.src + ((int64_t)g * p.ls.plane_stride + qx - p.halo) * 8, p.win_bytes, BAR(B_PFULL + ps));
#pragma unroll
          for (int g = 0; g < 2 * KSTEPS2; ++g)
            bulk_g2s(dst0 + (uint32_t)(2 * KSTEPS) * p.win_bytes + (uint32_t)g * p.win2_bytes,
                     p.src2 + ((int64_t)g * p.ls.plane_stride + qx) * 8, p.win2_bytes, BAR(B_PFULL + ps));
          if (++ps == p.p_slots) { ps = 0; pph ^= 1; }
        }
      }
    }
  } else if (warp < 2 + M4_TILES) {
    // ===================== MMA issuers: warp 2 + t owns tile t (tensor-memory columns 128 t .. 128 t + 127) ====
    const int t = warp - 2;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const bool leader = elect_one();
    constexpr uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | (8u << 24);
    constexpr uint32_t ID128 = idesc0 | ((128u >> 3) << 17), ID32 = idesc0 | ((32u >> 3) << 17);
    const uint64_t desc_hi = (uint64_t)(8u | (1u << 14)) << 32;            // SBO = 128 B, descriptor version 1
    const uint32_t a_lbo = ((uint32_t)p.win_cells & 0x3FFFu) << 16;        // K chunk 1 = the next channel-group plane
    const uint32_t a2_lbo = ((uint32_t)M4_L & 0x3FFFu) << 16;              // shortcut window: planes 512 cells apart
    constexpr uint32_t b_lbo = 128u << 16, b2_lbo = (uint32_t)M4_N0 << 16;
    const uint32_t d_mine = tmem_u + (uint32_t)t * 128u;
    int ps = 0, sl = 0;
    uint32_t pph = 0, sph = 0;
    uint32_t G = 0;
    for (int it = 0; it < my_items; ++it) {
      int b_, cell0_, n_act, x0, x1, xa, xb;
      item_of(it, b_, cell0_, n_act, x0, x1, xa, xb);
      const bool active = t < n_act;       // tiles beyond the plane run the barrier protocol and skip only the MMAs
      if (it > 0)
        for (uint32_t k = 1; k <= M4_RING; ++k)   // all slots of the previous march have been drained and cleared
          mbar_wait_warp(BAR(B_ACC_EMPTY + t * M4_RING + (int)((G - k) & 3u)), ((G - k) >> 2) & 1u);
      for (int x = xa; x <= xb; ++x) {
        const uint32_t gx = G + (uint32_t)(x - xa);      // ring index of the first band block (output x - 1)
        if (x > xa)       // the slot this plane's idle block touches (and the next plane opens): output x-2 is gone
          mbar_wait_warp(BAR(B_ACC_EMPTY + t * M4_RING + (int)((gx - 1u) & 3u)), ((gx - 1u) >> 2) & 1u);
        mbar_wait_warp(BAR(B_PFULL + ps), pph);
        const uint32_t stage16 = (sbase + (uint32_t)ps * p.stage_bytes) >> 4;
        const uint32_t a_org = stage16 + (uint32_t)p.halo + (uint32_t)t * 128u;
        for (int dy = 0; dy < 3; ++dy) {
          mbar_wait_warp(BAR(B_WFULL + sl), sph);
          tc_fence_after();
          const uint32_t b_org = (((sbase + p.off_w + (uint32_t)sl * CHUNK_BYTES) >> 4) & 0x3FFFu) | b_lbo;
          const uint32_t a_row = a_org + (uint32_t)((dy - 1) * pitch_y - 1);
          if (leader) {
            if (active) {
#pragma unroll
              for (int dz = 0; dz < 3; ++dz)
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks)
                  tc_mma_bf16(d_mine, desc_hi | (uint64_t)(((a_row + (uint32_t)dz + (uint32_t)(2 * ks) * (uint32_t)p.win_cells) & 0x3FFFu) | a_lbo),
                              desc_hi | (uint64_t)(b_org + (uint32_t)(dz * KSTEPS + ks) * (M4_MMA_B_BYTES / 16)), ID128, 1u);
            }
            tc_commit(BAR(B_WEMPTY + sl));
          }
          if (++sl == p.w_slots) { sl = 0; sph ^= 1u; }
        }
        if (leader) {
          if constexpr (KSTEPS2 > 0) {
            // fused 1x1 shortcut into the plane's OWN output (band block 1), from the halo-free second window
            if (active) {
              const uint32_t a2 = ((stage16 + (uint32_t)(2 * KSTEPS) * (uint32_t)p.win_cells + (uint32_t)t * 128u) & 0x3FFFu) | a2_lbo;
              const uint32_t w2 = (((sbase + p.off_w2) >> 4) & 0x3FFFu) | b2_lbo;
              const uint32_t dc = d_mine + ((gx + 1u) & 3u) * M4_N0;
#pragma unroll
              for (int ks = 0; ks < KSTEPS2; ++ks)
                tc_mma_bf16(dc, desc_hi | (uint64_t)(a2 + (uint32_t)(2 * ks) * (uint32_t)M4_L), desc_hi | (uint64_t)(w2 + (uint32_t)ks * (2u * M4_N0)), ID32, 1u);
            }
          }
          tc_commit(BAR(B_PEMPTY + ps));
          tc_commit(BAR(B_ACC_FULL + t * M4_RING + (int)(gx & 3u)));                       // output x-1 is complete
          if (x == xb) {
            tc_commit(BAR(B_ACC_FULL + t * M4_RING + (int)((gx + 1u) & 3u)));              // and so are xb, xb+1
            tc_commit(BAR(B_ACC_FULL + t * M4_RING + (int)((gx + 2u) & 3u)));
          }
        }
        if (++ps == p.p_slots) { ps = 0; pph ^= 1u; }
      }
      G += (uint32_t)(xb - xa + 1 + 2);
    }
    __syncwarp();
  } else {
    // ===================== epilogue: 8 warps, two per tensor-memory lane quarter; warp pair h drains tiles 2h, 2h+1 ====
    const int quarter = warp & 3;
    const int half = (warp - (2 + M4_TILES)) >> 2;
    const bool has_res = (p.flags & (SCENEEGO_F_RESIDUAL | SCENEEGO_F_ADD_AFTER)) != 0;
    uint32_t G = 0;
    for (int it = 0; it < my_items; ++it) {
      int b, cell0, n_act, x0, x1, xa, xb;
      item_of(it, b, cell0, n_act, x0, x1, xa, xb);
      const int OUTS = xb - xa + 1 + 2;
      bool valid[2];
      int64_t dpos0[2];
#pragma unroll
      for (int tt = 0; tt < 2; ++tt) {
        const int t = 2 * half + tt;
        const int cell = cell0 + t * 128 + quarter * 32 + lane;
        const int y = (int)fdiv((uint32_t)cell, p.fd_py);
        const int z = cell - y * pitch_y;
        valid[tt] = t < n_act && y < S && z < S;          // pads keep their zeros: nothing is written there
        dpos0[tt] = (int64_t)b * p.ld.frame_pitch + p.ld.guard + (int64_t)y * p.ld.pitch_y + z;
      }
      for (int oi = 0; oi < OUTS; ++oi) {
        const int o = xa - 1 + oi;
        const uint32_t gi = G + (uint32_t)oi;
        const int slot = (int)(gi & 3u);
        const bool store_o = o >= x0 && o < x1;
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
          const int t = 2 * half + tt;
          const int64_t dpos = dpos0[tt] + (int64_t)o * p.ld.pitch_x;
          uint4 rr[4];
#pragma unroll
          for (int g = 0; g < 4; ++g) {                       // residual cells in flight while waiting for the accumulator
            rr[g] = make_uint4(0, 0, 0, 0);
            if (has_res && valid[tt] && store_o)
              rr[g] = *reinterpret_cast<const uint4*>(p.res + ((int64_t)g * p.ld.plane_stride + dpos) * 8);
          }
          mbar_wait(BAR(B_ACC_FULL + t * M4_RING + slot), (gi >> 2) & 1u);
          tc_fence_after();
          uint32_t raw[2][16];
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * 128 + slot * M4_N0);
          tc_ld16(taddr, raw[0]);
          tc_ld16(taddr + 16u, raw[1]);
          tc_wait_ld();
          tc_st16_zero(taddr);                                // hand the slot back cleared
          tc_st16_zero(taddr + 16u);
          tc_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_ACC_EMPTY + t * M4_RING + slot));
          if (valid[tt] && store_o) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float r8[8], ov[8];
              unpack8(rr[g], r8);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float v = __uint_as_float(raw[g >> 1][(g & 1) * 8 + j]) + s_bias[8 * g + j];
                if (p.flags & SCENEEGO_F_RESIDUAL) v += r8[j];
                if (p.flags & SCENEEGO_F_RELU) v = fmaxf(v, 0.f);
                if (p.flags & SCENEEGO_F_ADD_AFTER) v += r8[j];
                ov[j] = v;
              }
              *reinterpret_cast<uint4*>(p.dst + ((int64_t)g * p.ld.plane_stride + dpos) * 8) = pack8(ov);
            }
          }
        }
      }
      G += (uint32_t)OUTS;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// ---------------------------------------------------------------------------
// CUDA-core checker: same packed blob walked in the kernel's order (rotation = input plane & 3), one thread per
// output voxel.  op.impl = 1 / SCENEEGO_FORCE_SIMT.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) conv_march4_simt_kernel(const __grid_constant__ March4Params p, int ksteps, int ksteps2) {
  const int S = p.ld.side;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (n >= S * S * S) return;
  const int z = n % S, y = (n / S) % S, x = n / (S * S);
  float acc[M4_N0];
#pragma unroll
  for (int j = 0; j < M4_N0; ++j) acc[j] = 0.f;
  for (int xi = x - 1; xi <= x + 1; ++xi) {
    if (xi < 0 || xi >= S) continue;
    const int j = x - xi + 1;
    const int r = xi & 3, s = (r + j) & 3;
    const uint4* wr = reinterpret_cast<const uint4*>(p.w + (size_t)r * p.rot_bytes);
    const int64_t q = vol_pos(p.ls, b, xi, y, z);
    for (int dy = 0; dy < 3; ++dy)
      for (int dz = 0; dz < 3; ++dz)
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint4* wt = wr + ((size_t)((dy * 3 + dz) * ksteps + ks) * M4_MMA_B_BYTES) / 16;
          const int64_t qs = q + (int64_t)(dy - 1) * p.ls.pitch_y + (dz - 1);
          for (int c = 0; c < 2; ++c) {
            float a[8];
            unpack8(*reinterpret_cast<const uint4*>(p.src + ((int64_t)(2 * ks + c) * p.ls.plane_stride + qs) * 8), a);
            for (int co = 0; co < M4_N0; ++co) {
              float wv[8];
              unpack8(__ldg(wt + c * 128 + s * M4_N0 + co), wv);
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[co] = fmaf(a[i], wv[i], acc[co]);
            }
          }
        }
  }
  const int64_t q = vol_pos(p.ls, b, x, y, z);
  if (p.src2) {
    const uint4* w2 = reinterpret_cast<const uint4*>(p.w + 4u * (size_t)p.rot_bytes);
    for (int ks = 0; ks < ksteps2; ++ks)
      for (int c = 0; c < 2; ++c) {
        float a[8];
        unpack8(*reinterpret_cast<const uint4*>(p.src2 + ((int64_t)(2 * ks + c) * p.ls.plane_stride + q) * 8), a);
        for (int co = 0; co < M4_N0; ++co) {
          float wv[8];
          unpack8(__ldg(w2 + (ks * 2 + c) * M4_N0 + co), wv);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[co] = fmaf(a[i], wv[i], acc[co]);
        }
      }
  }
  const int64_t dpos = vol_pos(p.ld, b, x, y, z);
  for (int g = 0; g < 4; ++g) {
    float r8[8] = {0, 0, 0, 0, 0, 0, 0, 0}, o[8];
    if (p.flags & (SCENEEGO_F_RESIDUAL | SCENEEGO_F_ADD_AFTER))
      unpack8(*reinterpret_cast<const uint4*>(p.res + ((int64_t)g * p.ld.plane_stride + dpos) * 8), r8);
    for (int j = 0; j < 8; ++j) {
      float v = acc[8 * g + j] + p.bias[8 * g + j];
      if (p.flags & SCENEEGO_F_RESIDUAL) v += r8[j];
      if (p.flags & SCENEEGO_F_RELU) v = fmaxf(v, 0.f);
      if (p.flags & SCENEEGO_F_ADD_AFTER) v += r8[j];
      o[j] = v;
    }
    *reinterpret_cast<uint4*>(p.dst + ((int64_t)g * p.ld.plane_stride + dpos) * 8) = pack8(o);
  }
}

typedef void (*march4_fn)(const March4Params);
static march4_fn pick_march4(int ksteps, int ksteps2) {
  if (ksteps == 2 && ksteps2 == 0) return conv_march4_kernel<2, 0>;
  if (ksteps == 2 && ksteps2 == 1) return conv_march4_kernel<2, 1>;
  if (ksteps == 1 && ksteps2 == 0) return conv_march4_kernel<1, 0>;
  return nullptr;
}

// Called by sceneego_v2v_run for SCENEEGO_OP_CONV3_MARCH4.
int launch_conv_march4(const sceneego_v2v_op_t& op, void* const* d_buffers, const void* d_blob, int batch, int op_index,
                       bool simt, cudaStream_t st) {
  March4Params p;
  memset(&p, 0, sizeof(p));
  p.src = (const __nv_bfloat16*)d_buffers[op.src];
  p.dst = (__nv_bfloat16*)d_buffers[op.dst];
  p.res = op.res >= 0 ? (const __nv_bfloat16*)d_buffers[op.res] : nullptr;
  p.src2 = (op.src2 >= 0 && op.cin2 > 0) ? (const __nv_bfloat16*)d_buffers[op.src2] : nullptr;
  p.w = (const uint8_t*)d_blob + op.w_offset;
  p.bias = (const float*)((const char*)d_blob + op.b_offset);
  p.ls = op.lay_src; p.ld = op.lay_dst; p.batch = batch; p.flags = op.flags;
  const int S = p.ls.side;
  SE_REQUIRE(p.src && p.dst, "v2v_run: op %d has a null buffer", op_index);
  SE_REQUIRE(op.ksize == 3 && (op.cin == 16 || op.cin == 32) && op.cout == M4_N0 && op.cout_real == M4_N0 && S >= 2,
             "v2v_run: op %d: the rotated marching conv is 3^3, 16 or 32 -> 32 channels", op_index);
  SE_REQUIRE(!(op.flags & SCENEEGO_F_OUT_F32), "v2v_run: op %d: the marching conv writes planar bf16", op_index);
  SE_REQUIRE(p.ls.s2d == 0 && p.ld.s2d == 0 && p.ls.pad >= 1 && p.ld.side == S && p.ls.guard >= p.ls.pitch_y + 1,
             "v2v_run: op %d: layouts incompatible with the marching conv", op_index);
  SE_REQUIRE(!p.src2 || (op.cin2 == 16 && op.cin == 32 && op.res < 0), "v2v_run: op %d: bad fused shortcut", op_index);
  SE_REQUIRE(!(op.flags & (SCENEEGO_F_RESIDUAL | SCENEEGO_F_ADD_AFTER)) || p.res, "v2v_run: op %d needs a residual", op_index);
  const int ksteps = op.cin / 16, ksteps2 = p.src2 ? 1 : 0;
  p.cin_planes = op.cin / 8;
  p.cin2_planes = p.src2 ? 2 : 0;
  p.chunk_bytes = 3u * (uint32_t)ksteps * M4_MMA_B_BYTES;
  p.rot_bytes = 3u * p.chunk_bytes;
  if (simt) {
    dim3 grid((S * S * S + 127) / 128, batch);
    conv_march4_simt_kernel<<<grid, 128, 0, st>>>(p, ksteps, ksteps2);
    SE_CUDA_LAUNCH_CHECK("conv_march4_simt");
    return SCENEEGO_OK;
  }
  p.halo = p.ls.pitch_y + 1;
  p.win_cells = M4_L + 2 * p.halo;
  p.win_cells = (p.win_cells + 7) / 8 * 8;
  p.win_bytes = (uint32_t)p.win_cells * 16u;
  p.win2_bytes = (uint32_t)M4_L * 16u;
  p.stage_bytes = (uint32_t)p.cin_planes * p.win_bytes + (uint32_t)p.cin2_planes * p.win2_bytes;
  SE_REQUIRE(p.win_cells < 16384, "v2v_run: op %d: window too large (side %d)", op_index, S);
  p.cells_per_plane = (S - 1) * p.ls.pitch_y + S;
  p.groups_per_frame = (p.cells_per_plane + M4_L - 1) / M4_L;
  const int base_items = batch * p.groups_per_frame;
  int n_seg = kNumSMs / base_items;
  if (n_seg < 1) n_seg = 1;
  if (n_seg > 16) n_seg = 16;
  { const char* e = getenv("SCENEEGO_MARCH4_SEGMENTS"); if (e && atoi(e) >= 1 && atoi(e) <= 32) n_seg = atoi(e); }
  int seg_planes = (S + n_seg - 1) / n_seg;
  if (seg_planes < 4) seg_planes = 4 < S ? 4 : S;
  n_seg = (S + seg_planes - 1) / seg_planes;
  p.n_seg = n_seg; p.seg_planes = seg_planes;
  p.n_items = base_items * n_seg;
  SE_REQUIRE((int64_t)batch * p.ls.frame_pitch + 4096 < (1ll << 31), "v2v_run: op %d: batch * frame_pitch too large for one launch", op_index);
  const uint32_t bar_bytes = 8u * (2 * M4_MAX_PSLOTS + 2 * M4_MAX_WSLOTS + 2 * M4_TILES * M4_RING) + 64u;
  const uint32_t fixed = 128u + 1024u + bar_bytes;     // bias + shortcut weights + barriers
  int p_slots = 3, w_slots = 0;
  for (; p_slots >= 2; --p_slots) {
    const int64_t left = (int64_t)kMaxSmem - fixed - (int64_t)p_slots * p.stage_bytes;
    w_slots = (int)(left / (int64_t)p.chunk_bytes);
    if (w_slots >= 3) break;
  }
  if (w_slots > M4_MAX_WSLOTS) w_slots = M4_MAX_WSLOTS;
  SE_REQUIRE(p_slots >= 2 && w_slots >= 2, "v2v_run: op %d: marching conv windows do not fit shared memory (side %d)", op_index, S);
  p.p_slots = p_slots; p.w_slots = w_slots;
  p.off_w = (uint32_t)p_slots * p.stage_bytes;
  p.off_w2 = p.off_w + (uint32_t)w_slots * p.chunk_bytes;
  p.off_bias = p.off_w2 + 1024u;
  p.off_bar = p.off_bias + 128u;
  p.fd_gpf = make_fastdiv((uint32_t)p.groups_per_frame);
  p.fd_py = make_fastdiv((uint32_t)p.ls.pitch_y);
  p.fd_seg = make_fastdiv((uint32_t)p.n_seg);
  const size_t smem_bytes = (size_t)p.off_bar + bar_bytes;
  SE_REQUIRE(smem_bytes <= kMaxSmem, "v2v_run: op %d: marching conv shared memory plan exceeds 227 KB", op_index);
  march4_fn fn = pick_march4(ksteps, ksteps2);
  SE_REQUIRE(fn != nullptr, "v2v_run: op %d: no conv_march4 instantiation", op_index);
  if (int rc = ensure_max_dynamic_smem((const void*)fn, (int)kMaxSmem)) return rc;
  const int grid = p.n_items < kNumSMs ? p.n_items : kNumSMs;
  fn<<<grid, M4_THREADS, smem_bytes, st>>>(p);
  SE_CUDA_LAUNCH_CHECK("conv_march4");
  return SCENEEGO_OK;
}

}  // namespace sceneego

using namespace sceneego;

extern "C" size_t sceneego_v2v_march4_weight_elems(int cin_pad) { return (size_t)4 * 9 * (cin_pad / 16) * (M4_MMA_B_BYTES / 2); }

extern "C" int sceneego_v2v_pack_conv_march4(const float* h_weight, const float* h_bias, const float* h_gamma,
                                             const float* h_beta, const float* h_mean, const float* h_var, double eps,
                                             int cout, int cin, int cin_pad, uint16_t* h_w_out, float* h_b_out) {
  SE_REQUIRE(h_weight && h_w_out && h_b_out, "pack_conv_march4: null argument");
  SE_REQUIRE(cout == M4_N0 && cin <= cin_pad && cin_pad % 16 == 0, "pack_conv_march4: 32 output channels, cin padded to 16");
  const int ksteps = cin_pad / 16;
  const size_t rot_elems = (size_t)9 * ksteps * (M4_MMA_B_BYTES / 2);
  memset(h_w_out, 0, 4 * rot_elems * sizeof(uint16_t));
  double scale[M4_N0];
  for (int co = 0; co < cout; ++co) {
    double sc = 1.0, sh = 0.0;
    if (h_gamma) {
      sc = (double)h_gamma[co] / sqrt((double)h_var[co] + eps);
      sh = (double)h_beta[co] - (double)h_mean[co] * sc;
    }
    scale[co] = sc;
    h_b_out[co] = (float)((h_bias ? (double)h_bias[co] : 0.0) * sc + sh);
  }
  for (int r = 0; r < 4; ++r)
    for (int s = 0; s < 4; ++s) {
      const int j = (s - r) & 3;
      if (j == 3) continue;                                   // the slot outside the band: zeros
      const int dx = 2 - j;
      for (int dy = 0; dy < 3; ++dy)
        for (int dz = 0; dz < 3; ++dz)
          for (int ks = 0; ks < ksteps; ++ks) {
            uint16_t* tap = h_w_out + r * rot_elems + (size_t)((dy * 3 + dz) * ksteps + ks) * (M4_MMA_B_BYTES / 2);
            for (int c = 0; c < 2; ++c)
              for (int co = 0; co < cout; ++co)
                for (int e = 0; e < 8; ++e) {
                  const int ci = ks * 16 + c * 8 + e;
                  if (ci >= cin) continue;
                  tap[((size_t)c * 128 + s * M4_N0 + co) * 8 + e] =
                      f2bf((float)((double)h_weight[((((size_t)co * cin + ci) * 3 + dx) * 3 + dy) * 3 + dz] * scale[co]));
                }
          }
    }
  return SCENEEGO_OK;
}
