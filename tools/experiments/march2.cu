// a7: the 3^3 convolutions with 32 output channels at full resolution (ten of V2V's layers, network/v2v.py:21-43,
// 147-156) as an x-marching banded GEMM -- csrc/march.cu's formulation (input plane x feeds outputs x-1, x, x+1 through
// ONE dense N = 96 MMA per (dy, dz, k-step); resident weights; a ring of tensor-memory slots) on the CTA structure of the
// marching stem (csrc/stem_march.cu).
//
// What limited march.cu (profiles/r02_ncu_march_b64.txt and two negative experiments recorded in DESIGN.md): three
// things at about the same level, so that fixing any ONE of them changed nothing --
//   (a) pipe: on 2 planes of 8 the band straddled the end of the 8-slot ring and was split into N = 32 + N = 64
//       (89.5 cycles instead of 56);
//   (b) latency: two CTAs per SM, each with the 55 KB weight array resident, left room for only TWO window stages
//       per CTA -- one plane (~1000 cycles) of prefetch against an L2/DRAM latency of several thousand;
//   (c) issue: one MMA-issuing warp per CTA with 18 MMAs per plane between its barrier round trips.
// Here ONE CTA per SM marches TWO neighbouring tiles (256 cells) that share the staged window and ONE copy of the
// weights: 170 KB are left for a window ring of six or more stages (b); each tile has its own issuing warp (c); and each
// tile's 256 tensor-memory columns are a ring of SIX slots plus TWO MIRROR slots: a band that starts in ring slot 4 or
// 5 runs on into the mirrors of slots 0 / 1 instead of wrapping, and the epilogue adds a mirror to its slot when it
// drains outputs 0 / 1 of a revolution -- every plane is one dense N = 96 MMA sequence (a).
// Weight blob and CUDA-core checker are march.cu's (sceneego_v2v_pack_conv_march, conv_simt_kernel).
#include "tc_common.cuh"
#include <stdlib.h>
#include <math.h>

namespace sceneego {

constexpr int M2_TILES = 2;
constexpr int M2_L = 128 * M2_TILES;
constexpr int M2_RING = 6;                                 // + 2 mirror slots = 8 physical slots of 32 columns per tile
constexpr int M2_N0 = 32;                                  // output channels
constexpr int M2_THREADS = 32 * (2 + M2_TILES + 8);       // 384
constexpr int M2_MAX_PSLOTS = 8;

struct March2Params {
  const __nv_bfloat16* src;
  const __nv_bfloat16* src2;   // fused shortcut source (2 planes) or nullptr
  const __nv_bfloat16* res;
  __nv_bfloat16* dst;
  const uint8_t* w;            // [tap (dy,dz) 9][cin/8][96 rows][8] (march.cu's layout), then (shortcut) [2][32][8]
  const float* bias;
  sceneego_vol_layout_t ls, ld;
  int batch, flags;
  int cin_planes, cin2_planes;
  int groups_per_frame, n_items, cells_per_plane;
  int n_seg, seg_planes;
  int halo, win_cells;
  int p_slots;
  uint32_t win_bytes, win2_bytes, stage_bytes, w_bytes, w2_bytes;
  uint32_t off_win, off_bias, off_bar;   // weights live at offset 0
  FastDiv fd_gpf, fd_py, fd_seg;
};

template <int KSTEPS, int KSTEPS2>
__global__ void __launch_bounds__(M2_THREADS, 1) conv_march2_kernel(const __grid_constant__ March2Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  float* s_bias = reinterpret_cast<float*>(smem + p.off_bias);
  const uint32_t bar0 = sbase + p.off_bar;
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int B_PFULL = 0, B_PEMPTY = B_PFULL + M2_MAX_PSLOTS, B_W_FULL = B_PEMPTY + M2_MAX_PSLOTS, B_ACC_FULL = B_W_FULL + 1,
                B_ACC_EMPTY = B_ACC_FULL + M2_TILES * M2_RING, B_COUNT = B_ACC_EMPTY + M2_TILES * M2_RING;
  uint32_t* s_tmem_ptr = reinterpret_cast<uint32_t*>(smem + p.off_bar + 8 * B_COUNT);

  if (threadIdx.x < M2_N0) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  if (threadIdx.x == 0) {
    for (int i = 0; i < M2_MAX_PSLOTS; ++i) { mbar_init(BAR(B_PFULL + i), 1); mbar_init(BAR(B_PEMPTY + i), M2_TILES); }
    mbar_init(BAR(B_W_FULL), 1);
    for (int i = 0; i < M2_TILES * M2_RING; ++i) { mbar_init(BAR(B_ACC_FULL + i), 1); mbar_init(BAR(B_ACC_EMPTY + i), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem_ptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem_ptr;
  if (warp >= 2 + M2_TILES && warp < 2 + M2_TILES + 4) {       // every accumulator slot (and mirror) starts cleared
    const uint32_t t0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (uint32_t c = 0; c < 512; c += 16) tc_st16_zero(t0 + c);
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int S = p.ls.side;
  const int my_items = ((int)p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  // item = (frame, pair of tiles, x-segment): stores outputs [x0, x1), marches over the input planes
  // [xa, xb] = [x0 - 1, x1] clipped to the volume, drains the outputs xa - 1 .. xb + 1 in order (the two outside
  // [x0, x1) are accumulated like any other and dropped: every plane has the same MMA sequence)
  auto item_of = [&](int it, int& b, int& cell0, int& n_act, int& x0, int& x1, int& xa, int& xb) {
    const uint32_t item = blockIdx.x + (uint32_t)it * gridDim.x;
    const uint32_t bg = fdiv(item, p.fd_seg);
    const int seg = (int)(item - bg * (uint32_t)p.n_seg);
    b = (int)fdiv(bg, p.fd_gpf);
    cell0 = (int)(bg - (uint32_t)b * (uint32_t)p.groups_per_frame) * M2_L;
    const int left = p.cells_per_plane - cell0;
    n_act = left >= M2_L ? M2_TILES : (left + 127) / 128;
    x0 = seg * p.seg_planes;
    x1 = x0 + p.seg_planes < S ? x0 + p.seg_planes : S;
    xa = x0 - 1 > 0 ? x0 - 1 : 0;
    xb = x1 < S - 1 ? x1 : S - 1;
  };
  const int pitch_y = p.ls.pitch_y;
  auto SLOT = [](uint32_t g) { return g % (uint32_t)M2_RING; };
  auto PAR = [](uint32_t g) { return (g / (uint32_t)M2_RING) & 1u; };

  if (warp == 0) {
    // ===================== producer: the resident weights once, then one window stage per input plane =========
    if (lane == 0) {
      const uint32_t wtot = p.w_bytes + p.w2_bytes;
      mbar_expect_tx(BAR(B_W_FULL), wtot);
      for (uint32_t o = 0; o < wtot; o += 16384u)
        bulk_g2s(sbase + o, p.w + o, wtot - o < 16384u ? wtot - o : 16384u, BAR(B_W_FULL));
      int ps = 0, pph = 0;
      for (int it = 0; it < my_items; ++it) {
        int b, cell0, n_act, x0, x1, xa, xb;
        item_of(it, b, cell0, n_act, x0, x1, xa, xb);
        const int64_t q0 = (int64_t)b * p.ls.frame_pitch + p.ls.guard + cell0;
        for (int x = xa; x <= xb; ++x) {
          const int64_t qx = q0 + (int64_t)x * p.ls.pitch_x;
          mbar_wait(BAR(B_PEMPTY + ps), pph ^ 1);
          mbar_expect_tx(BAR(B_PFULL + ps), p.stage_bytes);
          const uint32_t dst0 = sbase + p.off_win + (uint32_t)ps * p.stage_bytes;
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
  } else if (warp == 1) {
    // (tensor-memory allocator only)
  } else if (warp < 2 + M2_TILES) {
    // ===================== MMA issuers: warp 2 + t owns tile t (tensor-memory columns 256 t .. 256 t + 255) ====
    const int t = warp - 2;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const bool leader = elect_one();
    constexpr uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | (8u << 24);
    constexpr uint32_t ID96 = idesc0 | ((96u >> 3) << 17), ID32 = idesc0 | ((32u >> 3) << 17);
    const uint64_t desc_hi = (uint64_t)(8u | (1u << 14)) << 32;            // SBO = 128 B, descriptor version 1
    const uint32_t a_lbo = ((uint32_t)p.win_cells & 0x3FFFu) << 16;        // K chunk 1 = the next channel-group plane
    const uint32_t a2_lbo = ((uint32_t)M2_L & 0x3FFFu) << 16;              // shortcut window: planes 256 cells apart
    constexpr uint32_t b_lbo = (uint32_t)(3 * M2_N0) << 16, b2_lbo = (uint32_t)M2_N0 << 16;
    constexpr uint32_t b_ks_step = 2u * 3u * M2_N0;                         // 16-byte units
    constexpr uint32_t tap_step = (uint32_t)(2 * KSTEPS) * 3u * M2_N0;
    const uint32_t w_b = ((sbase >> 4) & 0x3FFFu) | b_lbo;
    const uint32_t w2_b = (((sbase + p.w_bytes) >> 4) & 0x3FFFu) | b2_lbo;
    const uint32_t d_mine = tmem_u + (uint32_t)t * 256u;
    mbar_wait_warp(BAR(B_W_FULL), 0);
    int ps = 0;
    uint32_t pph = 0;
    uint32_t G = 0;
    for (int it = 0; it < my_items; ++it) {
      int b_, cell0_, n_act, x0, x1, xa, xb;
      item_of(it, b_, cell0_, n_act, x0, x1, xa, xb);
      const bool active = t < n_act;       // a tile beyond the plane runs the barrier protocol and skips only the MMAs
      for (int x = xa; x <= xb; ++x) {
        const uint32_t gx = G + (uint32_t)(x - xa);      // ring index of band block 0 (output x - 1)
        // the slots this plane OPENS have been drained (and cleared, mirrors included) by the epilogue
        for (uint32_t j = (x == xa ? 0u : 2u); j <= 2u; ++j)
          mbar_wait_warp(BAR(B_ACC_EMPTY + t * M2_RING + (int)SLOT(gx + j)), PAR(gx + j) ^ 1u);
        mbar_wait_warp(BAR(B_PFULL + ps), pph);
        tc_fence_after();
        const uint32_t stage16 = (sbase + p.off_win + (uint32_t)ps * p.stage_bytes) >> 4;
        const uint32_t a_org = stage16 + (uint32_t)p.halo + (uint32_t)t * 128u;
        if (leader) {
          if (active) {
            // consecutive PHYSICAL slots from the ring position of block 0: a band that starts in ring slot 4 / 5 runs
            // on into the mirror slots 6 / 7 instead of wrapping
            const uint32_t dA = d_mine + SLOT(gx) * M2_N0;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const uint32_t a_row = a_org + (uint32_t)((dy - 1) * pitch_y - 1);
#pragma unroll
              for (int dz = 0; dz < 3; ++dz)
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks)
                  tc_mma_bf16(dA, desc_hi | (uint64_t)(((a_row + (uint32_t)dz + (uint32_t)(2 * ks) * (uint32_t)p.win_cells) & 0x3FFFu) | a_lbo),
                              desc_hi | (uint64_t)(w_b + (uint32_t)(dy * 3 + dz) * tap_step + (uint32_t)ks * b_ks_step), ID96, 1u);
            }
            if constexpr (KSTEPS2 > 0) {
              // fused 1x1 shortcut into the plane's OWN output (band block 1), from the halo-free second window
              const uint32_t a2 = ((stage16 + (uint32_t)(2 * KSTEPS) * (uint32_t)p.win_cells + (uint32_t)t * 128u) & 0x3FFFu) | a2_lbo;
              const uint32_t dc = d_mine + SLOT(gx + 1u) * M2_N0;
#pragma unroll
              for (int ks = 0; ks < KSTEPS2; ++ks)
                tc_mma_bf16(dc, desc_hi | (uint64_t)(a2 + (uint32_t)(2 * ks) * (uint32_t)M2_L), desc_hi | (uint64_t)(w2_b + (uint32_t)ks * (2u * M2_N0)), ID32, 1u);
            }
          }
          tc_commit(BAR(B_PEMPTY + ps));
          tc_commit(BAR(B_ACC_FULL + t * M2_RING + (int)SLOT(gx)));                       // output x-1 is complete
          if (x == xb) {
            tc_commit(BAR(B_ACC_FULL + t * M2_RING + (int)SLOT(gx + 1u)));                // and so are xb, xb+1
            tc_commit(BAR(B_ACC_FULL + t * M2_RING + (int)SLOT(gx + 2u)));
          }
        }
        if (++ps == p.p_slots) { ps = 0; pph ^= 1u; }
      }
      G += (uint32_t)(xb - xa + 1 + 2);
    }
    __syncwarp();
  } else {
    // ===================== epilogue: 8 warps = 2 tiles x 4 tensor-memory lane quarters =====================
    const int quarter = warp & 3;
    const int t = (warp - (2 + M2_TILES)) >> 2;
    const bool has_res = (p.flags & (SCENEEGO_F_RESIDUAL | SCENEEGO_F_ADD_AFTER)) != 0;
    uint32_t G = 0;
    for (int it = 0; it < my_items; ++it) {
      int b, cell0, n_act, x0, x1, xa, xb;
      item_of(it, b, cell0, n_act, x0, x1, xa, xb);
      const int OUTS = xb - xa + 1 + 2;
      const int cell = cell0 + t * 128 + quarter * 32 + lane;
      const int y = (int)fdiv((uint32_t)cell, p.fd_py);
      const int z = cell - y * pitch_y;
      const bool valid = t < n_act && y < S && z < S;          // pads keep their zeros: nothing is written there
      const int64_t dpos0 = (int64_t)b * p.ld.frame_pitch + p.ld.guard + (int64_t)y * p.ld.pitch_y + z;
      uint4 rn[4];
      auto load_res = [&](int o) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          rn[g] = make_uint4(0, 0, 0, 0);
          if (has_res && valid && o >= x0 && o < x1)
            rn[g] = *reinterpret_cast<const uint4*>(p.res + ((int64_t)g * p.ld.plane_stride + dpos0 + (int64_t)o * p.ld.pitch_x) * 8);
        }
      };
      load_res(xa - 1);
      for (int oi = 0; oi < OUTS; ++oi) {
        const int o = xa - 1 + oi;
        const uint32_t gi = G + (uint32_t)oi;
        const int slot = (int)SLOT(gi);
        uint4 rc[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) rc[g] = rn[g];
        load_res(o + 1);                                      // in flight while this plane is drained
        mbar_wait(BAR(B_ACC_FULL + t * M2_RING + slot), PAR(gi));
        tc_fence_after();
        uint32_t raw[2][16];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * 256 + slot * M2_N0);
        tc_ld16(taddr, raw[0]);
        tc_ld16(taddr + 16u, raw[1]);
        tc_wait_ld();
        tc_st16_zero(taddr);                                  // hand the slot back cleared
        tc_st16_zero(taddr + 16u);
        if (slot < 2) {
          // ring slots 0 / 1: bands that started in ring slots 4 / 5 accumulated their share of this output in the mirror
          const uint32_t maddr = taddr + (uint32_t)M2_RING * M2_N0;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t mir[16];
            tc_ld16(maddr + (uint32_t)(16 * c), mir);
            tc_wait_ld();
            tc_st16_zero(maddr + (uint32_t)(16 * c));
#pragma unroll
            for (int j = 0; j < 16; ++j) raw[c][j] = __float_as_uint(__uint_as_float(raw[c][j]) + __uint_as_float(mir[j]));
          }
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_ACC_EMPTY + t * M2_RING + slot));
        if (valid && o >= x0 && o < x1) {
          const int64_t dpos = dpos0 + (int64_t)o * p.ld.pitch_x;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float r8[8], ov[8];
            unpack8(rc[g], r8);
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
      G += (uint32_t)OUTS;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

typedef void (*march2_fn)(const March2Params);
static march2_fn pick_march2(int ksteps, int ksteps2) {
  if (ksteps == 2 && ksteps2 == 0) return conv_march2_kernel<2, 0>;
  if (ksteps == 2 && ksteps2 == 1) return conv_march2_kernel<2, 1>;
  if (ksteps == 1 && ksteps2 == 0) return conv_march2_kernel<1, 0>;
  return nullptr;
}

// Tensor path of SCENEEGO_OP_CONV3_MARCH for 16 / 32 -> 32 channels (the CUDA-core checker lives in v2v.cu; op.impl = 3
// or another channel count selects march.cu's kernel).  Returns SCENEEGO_E_UNSUPPORTED when this kernel does not
// cover the op, so that the caller can fall back to launch_conv_march.
int launch_conv_march2(const sceneego_v2v_op_t& op, void* const* d_buffers, const void* d_blob, int batch, int op_index,
                       cudaStream_t st) {
  if (!(op.ksize == 3 && (op.cin == 16 || op.cin == 32) && op.cout == M2_N0 && op.cout_real == M2_N0)) return SCENEEGO_E_UNSUPPORTED;
  March2Params p;
  memset(&p, 0, sizeof(p));
  p.src = (const __nv_bfloat16*)d_buffers[op.src];
  p.dst = (__nv_bfloat16*)d_buffers[op.dst];
  p.res = op.res >= 0 ? (const __nv_bfloat16*)d_buffers[op.res] : nullptr;
  p.src2 = (op.src2 >= 0 && op.cin2 > 0) ? (const __nv_bfloat16*)d_buffers[op.src2] : nullptr;
  p.w = (const uint8_t*)d_blob + op.w_offset;
  p.bias = (const float*)((const char*)d_blob + op.b_offset);
  p.ls = op.lay_src; p.ld = op.lay_dst; p.batch = batch; p.flags = op.flags;
  const int S = p.ls.side;
  SE_REQUIRE(p.src && p.dst && S >= 2, "v2v_run: op %d has a null buffer", op_index);
  SE_REQUIRE(!(op.flags & SCENEEGO_F_OUT_F32), "v2v_run: op %d: the marching conv writes planar bf16", op_index);
  SE_REQUIRE(p.ls.s2d == 0 && p.ld.s2d == 0 && p.ls.pad >= 1 && p.ld.side == S && p.ls.guard >= p.ls.pitch_y + 1,
             "v2v_run: op %d: layouts incompatible with the marching conv", op_index);
  SE_REQUIRE(!p.src2 || (op.cin2 == 16 && op.cin == 32 && op.res < 0), "v2v_run: op %d: bad fused shortcut", op_index);
  SE_REQUIRE(!(op.flags & (SCENEEGO_F_RESIDUAL | SCENEEGO_F_ADD_AFTER)) || p.res, "v2v_run: op %d needs a residual", op_index);
  const int ksteps = op.cin / 16, ksteps2 = p.src2 ? 1 : 0;
  p.cin_planes = op.cin / 8;
  p.cin2_planes = p.src2 ? 2 : 0;
  p.halo = p.ls.pitch_y + 1;
  p.win_cells = (M2_L + 2 * p.halo + 7) / 8 * 8;
  p.win_bytes = (uint32_t)p.win_cells * 16u;
  p.win2_bytes = (uint32_t)M2_L * 16u;
  p.stage_bytes = (uint32_t)p.cin_planes * p.win_bytes + (uint32_t)p.cin2_planes * p.win2_bytes;
  p.w_bytes = 9u * (uint32_t)p.cin_planes * 3u * M2_N0 * 16u;
  p.w2_bytes = (uint32_t)p.cin2_planes * M2_N0 * 16u;
  SE_REQUIRE(p.win_cells < 16384, "v2v_run: op %d: window too large (side %d)", op_index, S);
  p.cells_per_plane = (S - 1) * p.ls.pitch_y + S;
  p.groups_per_frame = (p.cells_per_plane + M2_L - 1) / M2_L;
  const int base_items = batch * p.groups_per_frame;
  int n_seg = kNumSMs / base_items;
  if (n_seg < 1) n_seg = 1;
  if (n_seg > 16) n_seg = 16;
  { const char* e = getenv("SCENEEGO_MARCH_SEGMENTS"); if (e && atoi(e) >= 1 && atoi(e) <= 32) n_seg = atoi(e); }
  int seg_planes = (S + n_seg - 1) / n_seg;
  if (seg_planes < 4) seg_planes = 4 < S ? 4 : S;
  n_seg = (S + seg_planes - 1) / seg_planes;
  p.n_seg = n_seg; p.seg_planes = seg_planes;
  p.n_items = base_items * n_seg;
  SE_REQUIRE((int64_t)batch * p.ls.frame_pitch + 4096 < (1ll << 31), "v2v_run: op %d: batch * frame_pitch too large for one launch", op_index);
  const uint32_t bar_bytes = 8u * (2 * M2_MAX_PSLOTS + 1 + 2 * M2_TILES * M2_RING) + 64u;
  const uint32_t w_all = (p.w_bytes + p.w2_bytes + 127u) / 128u * 128u;
  const int64_t left = (int64_t)kMaxSmem - w_all - 128 - bar_bytes;
  int p_slots = (int)(left / (int64_t)p.stage_bytes);
  if (p_slots > M2_MAX_PSLOTS) p_slots = M2_MAX_PSLOTS;
  { const char* e = getenv("SCENEEGO_MARCH_STAGES"); if (e && atoi(e) >= 2 && atoi(e) <= p_slots) p_slots = atoi(e); }
  if (p_slots < 2) return SCENEEGO_E_UNSUPPORTED;          // very wide volumes: march.cu's planner handles them
  p.p_slots = p_slots;
  p.off_win = w_all;
  p.off_bias = p.off_win + (uint32_t)p_slots * p.stage_bytes;
  p.off_bar = p.off_bias + 128u;
  p.fd_gpf = make_fastdiv((uint32_t)p.groups_per_frame);
  p.fd_py = make_fastdiv((uint32_t)p.ls.pitch_y);
  p.fd_seg = make_fastdiv((uint32_t)p.n_seg);
  const size_t smem_bytes = (size_t)p.off_bar + bar_bytes;
  march2_fn fn = pick_march2(ksteps, ksteps2);
  if (fn == nullptr || smem_bytes > kMaxSmem) return SCENEEGO_E_UNSUPPORTED;
  if (int rc = ensure_max_dynamic_smem((const void*)fn, (int)kMaxSmem)) return rc;
  const int grid = p.n_items < kNumSMs ? p.n_items : kNumSMs;
  fn<<<grid, M2_THREADS, smem_bytes, st>>>(p);
  SE_CUDA_LAUNCH_CHECK("conv_march2");
  return SCENEEGO_OK;
}

}  // namespace sceneego
