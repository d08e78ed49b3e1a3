// a7, first layer: the 7x7x7 stem Conv3d(33 -> 16) + BN + ReLU (network/v2v.py:8-18,147) on sm_100a.
//
// Why its own kernel.  Cout = 16 is the worst tcgen05 shape: an N = 16 MMA costs as much as N = 64
// (the A operand's shared-memory read, 128 rows x 32 B at 128 B/clk, bounds it), and Cin = 33 wastes a
// third of a 48-channel K loop on the one binary occupancy channel.  This kernel therefore
//   * stacks the 2x2x2 block of output voxels (2X+sx, 2Y+sy, 2Z+sz) into N = 8 x 16 = 128 columns of one
//     GEMM row, so rows enumerate the 32^3 grid of blocks and every tcgen05.mma runs at the tensor
//     pipe's own rate (64 cycles at N = 128, tools/mma_replay.cu);
//   * reads its input in a space-to-depth ("s2d") layout written directly by the unprojection and
//     voxelisation kernels: 8 parity sub-volumes (x&1, y&1, z&1) of side V/2, each in the planar
//     padded layout with pad = 2, so that the input voxel (2X+ox, 2Y+oy, 2Z+oz), ox,oy,oz in [-3,4],
//     is sub-volume (ox&1, oy&1, oz&1) at block shift (ox>>1, oy>>1, oz>>1): a constant position
//     offset, exactly like a tap of the plain kernel.  512 offsets x 2 K-steps (32 feature channels)
//     against Toeplitz-stacked weights W[co][ci][ox-sx+3][oy-sy+3][oz-sz+3] (67 % non-zero);
//   * packs the occupancy channel along K instead of padding it to 16 channels: the scene plane
//     holds, per block, the 8 parity values of the 2x2x2 block as the 8 "channels" of one 16-byte
//     cell, and one MMA (K = 16) covers two neighbouring blocks in z (LBO = 16 bytes): 75 MMAs
//     instead of 512 per tile.
// Per 128-row tile (1024 output voxels): 1024 + 75(+5 padding) MMAs of 128x128x16.
//
// Pipeline: like conv_tc_kernel.  Warp 0 = producer (cp.async.bulk windows + weight chunks),
// warps 1-4 = MMA issuers (one 128-row tile each), warps 5-12 = epilogue.  An item is 4 tiles =
// 512 block positions; its accumulators fill all 512 TMEM columns (single-buffered: the epilogue
// is ~2 % of an item).
#include "tc_common.cuh"
#include <stdlib.h>
#include <math.h>

namespace sceneego {

constexpr int STEM_TILES = 4;
constexpr int STEM_L = STEM_TILES * 128;
constexpr int STEM_THREADS = 32 * (1 + STEM_TILES + 8);
constexpr int STEM_FEAT_STAGES = 32;          // (ox in -3..4) x (py, pz)
constexpr int STEM_SCENE_STAGES = 5;          // block shift bx in -2..2
// Weights stream in 16 KB chunks per CTA.  A feature stage has 16 taps (4 by x 4 bz) of 8 KB, a scene stage
// 15 taps (5 by x 3 bz pairs) + 1 zero tap of 4 KB.  One CTA: 2 feature / 4 scene taps per chunk; a CTA pair
// stages half of every tap's columns in each CTA, so its chunks hold twice as many taps (the pair's issue
// rate is bounded per warp -- tools/mma_pair_rate.cu -- and wants the fewest chunk hand-shakes per MMA).
constexpr int STEM_CHUNK_BYTES = 16384;
constexpr int STEM_FEAT_TAP_BYTES = 8192, STEM_SCENE_TAP_BYTES = 4096;
constexpr size_t STEM_W_BYTES = (size_t)STEM_FEAT_STAGES * 16 * STEM_FEAT_TAP_BYTES + (size_t)STEM_SCENE_STAGES * 16 * STEM_SCENE_TAP_BYTES;   // 4.3 MB
constexpr size_t STEM_SCENE_OFF = (size_t)STEM_FEAT_STAGES * 16 * STEM_FEAT_TAP_BYTES;
constexpr int STEM_MAX_STAGES = 3, STEM_MAX_WSLOTS = 8;

struct StemParams {
  const __nv_bfloat16* src;   // 33 planes: plane = parity * 4 + channel group; plane 32 = occupancy blocks
  __nv_bfloat16* dst;         // 16 channels = 2 planes, layout ld
  const uint8_t* w;           // STEM_W_BYTES, streaming order (half-major for CTA pairs)
  const float* bias;          // 16
  sceneego_vol_layout_t ls, ld;
  int batch, relu, n_items;
  int halo, win_cells;        // halo = 2 * (pitch_y + 1); window = L + 2 * halo + 8 cells
  int win_stages, w_slots;
  int cg;                     // 2: CTA pairs (tcgen05 cta_group::2), weights N-split: blob = [half][chunk][tap][kchunk][64][8]
  uint32_t win_bytes, off_w, off_bias, off_bar;
  FastDiv fd_frame, fd_px, fd_py;
};

template <int CG>
__global__ void __launch_bounds__(STEM_THREADS, 1) stem_s2d_tc_kernel(const __grid_constant__ StemParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  float* s_bias = reinterpret_cast<float*>(smem + p.off_bias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  const uint32_t bar0 = sbase + p.off_bar;
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int B_FULL_WIN = 0, B_EMPTY_WIN = STEM_MAX_STAGES, B_FULL_W = 2 * STEM_MAX_STAGES,
                B_EMPTY_W = B_FULL_W + STEM_MAX_WSLOTS, B_TMEM_FULL = B_EMPTY_W + STEM_MAX_WSLOTS,
                B_TMEM_EMPTY = B_TMEM_FULL + 1, B_COUNT = B_TMEM_EMPTY + 1;
  uint32_t* s_tmem_ptr = reinterpret_cast<uint32_t*>(bars + B_COUNT);

  // CTA pair (CG = 2): rank 0 issues every MMA (M = 256 = one tile of each CTA); see conv_tc_kernel
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  if (threadIdx.x < 16) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  if (threadIdx.x == 0) {
    const int full_cnt = (CG == 2 && is_leader) ? 2 : 1;       // own TMA + the peer's relayed arrival
    for (int i = 0; i < STEM_MAX_STAGES; ++i) { mbar_init(BAR(B_FULL_WIN + i), full_cnt); mbar_init(BAR(B_EMPTY_WIN + i), STEM_TILES); }
    for (int i = 0; i < STEM_MAX_WSLOTS; ++i) { mbar_init(BAR(B_FULL_W + i), full_cnt); mbar_init(BAR(B_EMPTY_W + i), STEM_TILES); }
    mbar_init(BAR(B_TMEM_FULL), STEM_TILES);
    mbar_init(BAR(B_TMEM_EMPTY), 8 * CG);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem_ptr)), "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem_ptr)), "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem_ptr;
  const int n_cl = (int)gridDim.x / CG, cl = (int)blockIdx.x / CG;
  const int n_groups = (p.n_items + CG - 1) / CG;
  const int my_items = (n_groups - cl + n_cl - 1) / n_cl;
  auto item_of = [&](int it) -> int {
    const int item = (cl + it * n_cl) * CG + (int)cta_rank;
    return item < p.n_items ? item : p.n_items - 1;          // odd item count: the peer repeats the last item
  };
  const int pitch_y = p.ls.pitch_y, pitch_x = p.ls.pitch_x;
  const uint32_t stage_bytes = 4u * p.win_bytes;

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      int ws = 0, wph = 0, sl = 0, sph = 0;
      for (int it = 0; it < my_items; ++it) {
        const int64_t q0 = (int64_t)p.ls.guard + (int64_t)item_of(it) * STEM_L;
        int chunk = 0;
        for (int s = 0; s < STEM_FEAT_STAGES + STEM_SCENE_STAGES; ++s) {
          const bool feat = s < STEM_FEAT_STAGES;
          int bx, plane0, n_planes, n_chunks;
          if (feat) {
            const int ox = (s >> 2) - 3;
            bx = ox >> 1;                                   // arithmetic shift: floor(ox / 2)
            plane0 = (((ox & 1) << 2) | (s & 3)) * 4;       // parity (px, py, pz) x 4 channel groups
            n_planes = 4; n_chunks = 16 / (2 * CG);
          } else {
            bx = s - STEM_FEAT_STAGES - 2; plane0 = 32; n_planes = 1; n_chunks = 16 / (4 * CG);
          }
          mbar_wait(BAR(B_EMPTY_WIN + ws), wph ^ 1);
          mbar_expect_tx(BAR(B_FULL_WIN + ws), p.win_bytes * (uint32_t)n_planes);
          const int64_t qs = q0 + (int64_t)bx * pitch_x - p.halo;
          for (int g = 0; g < n_planes; ++g)
            bulk_g2s(sbase + (uint32_t)ws * stage_bytes + (uint32_t)g * p.win_bytes,
                     p.src + ((int64_t)(plane0 + g) * p.ls.plane_stride + qs) * 8, p.win_bytes, BAR(B_FULL_WIN + ws));
          if (++ws == p.win_stages) { ws = 0; wph ^= 1; }
          for (int c = 0; c < n_chunks; ++c, ++chunk) {
            mbar_wait(BAR(B_EMPTY_W + sl), sph ^ 1);
            mbar_expect_tx(BAR(B_FULL_W + sl), STEM_CHUNK_BYTES);
            bulk_g2s(sbase + p.off_w + (uint32_t)sl * STEM_CHUNK_BYTES,
                     p.w + (size_t)cta_rank * (STEM_W_BYTES / CG) + (size_t)chunk * STEM_CHUNK_BYTES, STEM_CHUNK_BYTES,
                     BAR(B_FULL_W + sl));
            if (++sl == p.w_slots) { sl = 0; sph ^= 1; }
          }
        }
      }
    }
  } else if (warp <= STEM_TILES) {
    // ===================== MMA issuers: warp w owns tile w-1 (TMEM columns 128*(w-1)..+127) ===========
    if (CG == 2 && !is_leader) {
      // peer CTA: warp 1 relays "landed" to the leader's full barriers in consumption order
      if (warp == 1 && lane == 0) {
        int ws = 0, wph = 0, sl = 0, sph = 0;
        for (int it = 0; it < my_items; ++it)
          for (int s = 0; s < STEM_FEAT_STAGES + STEM_SCENE_STAGES; ++s) {
            mbar_wait(BAR(B_FULL_WIN + ws), wph);
            mbar_arrive_remote(mapa_shared(BAR(B_FULL_WIN + ws), 0));
            if (++ws == p.win_stages) { ws = 0; wph ^= 1; }
            const int n_chunks = s < STEM_FEAT_STAGES ? 16 / (2 * CG) : 16 / (4 * CG);
            for (int c = 0; c < n_chunks; ++c) {
              mbar_wait(BAR(B_FULL_W + sl), sph);
              mbar_arrive_remote(mapa_shared(BAR(B_FULL_W + sl), 0));
              if (++sl == p.w_slots) { sl = 0; sph ^= 1; }
            }
          }
      }
    } else {
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const bool leader = elect_one();
    auto WAIT = [&](uint32_t bar, uint32_t parity) {
      if constexpr (CG == 2) mbar_wait_warp_cluster(bar, parity); else mbar_wait_warp(bar, parity);
    };
    auto MMA = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t ac) {
      if constexpr (CG == 2) tc_mma_bf16_pair(d, a, b, id, ac); else tc_mma_bf16(d, a, b, id, ac);
    };
    auto COMMIT = [&](uint32_t bar) {
      if constexpr (CG == 2) tc_commit_pair(bar); else tc_commit(bar);
    };
    constexpr uint32_t NH = 128 / CG;                                      // B rows (columns of D) staged per CTA
    const uint32_t idesc = (1u << 4) | kIdescAB | ((uint32_t)(128 >> 3) << 17) | ((CG == 2 ? 16u : 8u) << 24);
    const uint64_t desc_hi = (uint64_t)(8u | (1u << 14)) << 32;            // SBO = 128 B, descriptor version 1
    const uint32_t a_lbo_feat = ((uint32_t)p.win_cells & 0x3FFFu) << 16;   // K chunk 1 = next channel-group plane
    const uint32_t a_lbo_scene = 1u << 16;                                 // K chunk 1 = next block in z (16 B)
    const uint32_t b_lbo = (NH & 0x3FFFu) << 16;                           // [kchunk][NH columns][8]: NH * 16 B apart
    const uint32_t my_tile = (uint32_t)(warp - 1);
    const uint32_t d_mine = tmem_u + my_tile * 128u;
    const uint32_t a_ks = 2u * (uint32_t)p.win_cells;                      // K step = two planes
    int ws = 0, wph = 0, sl = 0, sph = 0;
    for (int it = 0; it < my_items; ++it) {
      WAIT(BAR(B_TMEM_EMPTY), (it & 1) ^ 1);
      tc_fence_after();
      uint32_t acc = 0;
      for (int s = 0; s < STEM_FEAT_STAGES + STEM_SCENE_STAGES; ++s) {
        WAIT(BAR(B_FULL_WIN + ws), wph);
        // first cell of this tile's rows at block shift (.., 0, 0)
        const uint32_t a_org = ((sbase + (uint32_t)ws * stage_bytes) >> 4) + (uint32_t)p.halo + my_tile * 128u;
        if (s < STEM_FEAT_STAGES) {
          const int py = (s >> 1) & 1, pz = s & 1;
          constexpr int TPC = 2 * CG;                     // feature taps per chunk: consecutive in bz, one by
          for (int c = 0; c < 16 / TPC; ++c) {
            WAIT(BAR(B_FULL_W + sl), sph);
            tc_fence_after();
            const uint32_t b_org = (((sbase + p.off_w + (uint32_t)sl * STEM_CHUNK_BYTES) >> 4) & 0x3FFFu) | b_lbo;
            const int tp0 = c * TPC;
            const int by = (tp0 >> 2) - 1 - py;
            const int bz0 = (tp0 & 3) - 1 - pz;
            const uint32_t a_row = a_org + (uint32_t)(by * pitch_y + bz0);
            if (leader) {
#pragma unroll
              for (int j = 0; j < TPC; ++j) {
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
                  MMA(d_mine, desc_hi | (((a_row + (uint32_t)j + (uint32_t)ks * a_ks) & 0x3FFFu) | a_lbo_feat),
                      desc_hi | (b_org + (uint32_t)(j * 4 + ks * 2) * NH), idesc, (j == 0 && ks == 0) ? acc : 1u);
              }
            }
            acc = 1;
            if (leader) COMMIT(BAR(B_EMPTY_W + sl));
            if (++sl == p.w_slots) { sl = 0; sph ^= 1; }
          }
        } else {
          constexpr int TPC = 4 * CG;                     // scene taps per chunk
          for (int c = 0; c < 16 / TPC; ++c) {
            WAIT(BAR(B_FULL_W + sl), sph);
            tc_fence_after();
            const uint32_t b_org = (((sbase + p.off_w + (uint32_t)sl * STEM_CHUNK_BYTES) >> 4) & 0x3FFFu) | b_lbo;
            if (leader) {
#pragma unroll
              for (int j = 0; j < TPC; ++j) {
                const int tp = c * TPC + j;               // 0..14: (by, bz pair); 15: zero-weight filler
                const int by = tp < 15 ? tp / 3 - 2 : 0;
                const int bz0 = tp < 15 ? (tp % 3) * 2 - 2 : 0;
                MMA(d_mine, desc_hi | (((a_org + (uint32_t)(by * pitch_y + bz0)) & 0x3FFFu) | a_lbo_scene),
                    desc_hi | (b_org + (uint32_t)(j * 2) * NH), idesc, 1u);
              }
            }
            if (leader) COMMIT(BAR(B_EMPTY_W + sl));
            if (++sl == p.w_slots) { sl = 0; sph ^= 1; }
          }
        }
        if (leader) COMMIT(BAR(B_EMPTY_WIN + ws));
        if (++ws == p.win_stages) { ws = 0; wph ^= 1; }
      }
      if (leader) COMMIT(BAR(B_TMEM_FULL));
    }
    __syncwarp();
    }
  } else {
    // ===================== epilogue: 8 warps, two per TMEM lane quarter =====================
    const int quarter = warp & 3;
    const int half = (warp - (1 + STEM_TILES)) >> 2;
    const int S2 = p.ls.side;
    for (int it = 0; it < my_items; ++it) {
      const int64_t q0 = (int64_t)p.ls.guard + (int64_t)item_of(it) * STEM_L;
      mbar_wait(BAR(B_TMEM_FULL), it & 1);
      tc_fence_after();
#pragma unroll 1
      for (int t = half; t < STEM_TILES; t += 2) {
        // decode this thread's row: block position -> (frame, X, Y, Z)
        const uint32_t q = (uint32_t)(q0 + t * 128 + quarter * 32 + lane);
        const uint32_t b = fdiv(q, p.fd_frame);
        const int rem = (int)(q - b * (uint32_t)p.ls.frame_pitch) - p.ls.guard;
        bool valid = ((int)b < p.batch) && rem >= 0;
        int X = 0, Y = 0, Z = 0;
        if (valid) {
          X = (int)fdiv((uint32_t)rem, p.fd_px);
          const int r2 = rem - X * pitch_x;
          Y = (int)fdiv((uint32_t)r2, p.fd_py);
          Z = r2 - Y * pitch_y;
          valid = X < S2 && Y < S2 && Z < S2;
        }
        const int64_t d0 = (int64_t)b * p.ld.frame_pitch + p.ld.guard + (int64_t)(2 * X) * p.ld.pitch_x +
                           (int64_t)(2 * Y) * p.ld.pitch_y + 2 * Z;
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * 128);
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {       // column chunk c = stacked voxel (sx, sy, sz), 16 channels
          uint32_t raw[16];
          tc_ld16(taddr + (uint32_t)(c << 4), raw);
          tc_wait_ld();
          if (valid) {
            const int64_t dpos = d0 + (int64_t)(c >> 2) * p.ld.pitch_x + (int64_t)((c >> 1) & 1) * p.ld.pitch_y + (c & 1);
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              float o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float v = __uint_as_float(raw[8 * g + j]) + s_bias[8 * g + j];
                o[j] = p.relu ? fmaxf(v, 0.f) : v;
              }
              *reinterpret_cast<uint4*>(p.dst + ((int64_t)g * p.ld.plane_stride + dpos) * 8) = pack8(o);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_remote(mapa_shared(BAR(B_TMEM_EMPTY), 0));
        else mbar_arrive(BAR(B_TMEM_EMPTY));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();
  if (warp == 1) {
    if constexpr (CG == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---------------------------------------------------------------------------
// CUDA-core checker: same s2d input, same packed blob (walks the streaming order), one thread per
// output voxel.  op.impl = 1 / SCENEEGO_FORCE_SIMT; used by the tests to validate the tensor path.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) stem_s2d_simt_kernel(const __grid_constant__ StemParams p) {
  const int V = p.ld.side;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (n >= V * V * V) return;
  const int z = n % V, y = (n / V) % V, x = n / (V * V);
  const int X = x >> 1, Y = y >> 1, Z = z >> 1;
  const int col0 = (((x & 1) * 2 + (y & 1)) * 2 + (z & 1)) * 16;     // this voxel's 16 columns of the stacked B
  const int64_t q = vol_pos(p.ls, b, X, Y, Z);
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  const uint4* wq = reinterpret_cast<const uint4*>(p.w);
  const int nh = 128 / p.cg, hf = col0 / nh, colh = col0 - hf * nh;   // CTA-pair blobs are half-major
  const size_t half_u4 = STEM_W_BYTES / p.cg / 16;
  for (int s = 0; s < STEM_FEAT_STAGES; ++s) {
    const int ox = (s >> 2) - 3, py = (s >> 1) & 1, pz = s & 1;
    const int bx = ox >> 1, plane0 = (((ox & 1) << 2) | (s & 3)) * 4;
    for (int tp = 0; tp < 16; ++tp) {
      const int by = (tp >> 2) - 1 - py, bz = (tp & 3) - 1 - pz;
      const int64_t qs = q + (int64_t)bx * p.ls.pitch_x + (int64_t)by * p.ls.pitch_y + bz;
      const uint4* wt = wq + (size_t)hf * half_u4 + (size_t)(s * 16 + tp) * (STEM_FEAT_TAP_BYTES / p.cg) / 16;   // [kchunk 4][nh][8]
      for (int g = 0; g < 4; ++g) {
        float a[8];
        unpack8(*reinterpret_cast<const uint4*>(p.src + ((int64_t)(plane0 + g) * p.ls.plane_stride + qs) * 8), a);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float wv[8];
          unpack8(__ldg(wt + g * nh + colh + j), wv);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[j] = fmaf(a[i], wv[i], acc[j]);
        }
      }
    }
  }
  for (int s = 0; s < STEM_SCENE_STAGES; ++s) {
    const int bx = s - 2;
    for (int tp = 0; tp < 15; ++tp) {
      const int by = tp / 3 - 2, bz0 = (tp % 3) * 2 - 2;
      const uint4* wt = wq + (size_t)hf * half_u4 + (STEM_SCENE_OFF + (size_t)(s * 16 + tp) * STEM_SCENE_TAP_BYTES) / p.cg / 16;   // [kchunk 2][nh][8]
      for (int c2 = 0; c2 < 2; ++c2) {
        const int64_t qs = q + (int64_t)bx * p.ls.pitch_x + (int64_t)by * p.ls.pitch_y + bz0 + c2;
        float a[8];
        unpack8(*reinterpret_cast<const uint4*>(p.src + ((int64_t)32 * p.ls.plane_stride + qs) * 8), a);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float wv[8];
          unpack8(__ldg(wt + c2 * nh + colh + j), wv);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[j] = fmaf(a[i], wv[i], acc[j]);
        }
      }
    }
  }
  const int64_t dpos = vol_pos(p.ld, b, x, y, z);
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = acc[8 * g + j] + p.bias[8 * g + j];
      o[j] = p.relu ? fmaxf(v, 0.f) : v;
    }
    *reinterpret_cast<uint4*>(p.dst + ((int64_t)g * p.ld.plane_stride + dpos) * 8) = pack8(o);
  }
}

// Called by sceneego_v2v_run for SCENEEGO_OP_STEM7_S2D.
int launch_stem_s2d(const sceneego_v2v_op_t& op, void* const* d_buffers, const void* d_blob, int batch, int op_index,
                    bool simt, cudaStream_t st) {
  SE_REQUIRE(op.lay_src.s2d == 1 && op.lay_src.pad >= 2 && op.lay_dst.s2d == 0 && op.lay_dst.side == 2 * op.lay_src.side,
             "v2v_run: op %d: stem needs an s2d source (pad >= 2) of half the destination side", op_index);
  SE_REQUIRE(op.cout == 16 && op.cin == 33 && op.ksize == 7, "v2v_run: op %d: stem is 33 -> 16, k = 7", op_index);
  StemParams p;
  memset(&p, 0, sizeof(p));
  p.src = (const __nv_bfloat16*)d_buffers[op.src];
  p.dst = (__nv_bfloat16*)d_buffers[op.dst];
  p.w = (const uint8_t*)d_blob + op.w_offset;
  p.bias = (const float*)((const char*)d_blob + op.b_offset);
  p.ls = op.lay_src; p.ld = op.lay_dst; p.batch = batch; p.relu = (op.flags & SCENEEGO_F_RELU) ? 1 : 0;
  p.cg = op.cta_pair == 2 ? 2 : 1;
  SE_REQUIRE(p.src && p.dst, "v2v_run: op %d has a null buffer", op_index);
  if (simt) {
    const int V = p.ld.side;
    dim3 grid((V * V * V + 127) / 128, batch);
    stem_s2d_simt_kernel<<<grid, 128, 0, st>>>(p);
    SE_CUDA_LAUNCH_CHECK("stem_s2d_simt");
    return SCENEEGO_OK;
  }
  const int64_t n_pos = (int64_t)batch * p.ls.frame_pitch;
  SE_REQUIRE(n_pos + 4096 < (1ll << 31) && (n_pos + 4096) * (int64_t)p.ls.frame_pitch < (1ll << 48),
             "v2v_run: op %d: batch * frame_pitch too large for one launch (use a smaller chunk)", op_index);
  p.fd_frame = make_fastdiv((uint32_t)p.ls.frame_pitch);
  p.fd_px = make_fastdiv((uint32_t)p.ls.pitch_x);
  p.fd_py = make_fastdiv((uint32_t)p.ls.pitch_y);
  p.halo = 2 * (p.ls.pitch_y + 1);
  p.win_cells = (STEM_L + 2 * p.halo + 8 + 7) / 8 * 8;     // +8: the z-pair of the last scene tap reads one block further
  p.win_bytes = (uint32_t)p.win_cells * 16u;
  SE_REQUIRE(p.ls.guard >= 2 * p.ls.pitch_x + p.halo, "v2v_run: op %d: s2d guard too small", op_index);
  const uint32_t fixed = 1024;
  int stages = STEM_MAX_STAGES, slots = 0;
  for (; stages >= 2; --stages) {
    const uint32_t used = 4u * p.win_bytes * stages + fixed;
    if (used + 3u * STEM_CHUNK_BYTES > kMaxSmem) continue;
    slots = (int)((kMaxSmem - used) / STEM_CHUNK_BYTES);
    if (slots > STEM_MAX_WSLOTS) slots = STEM_MAX_WSLOTS;
    break;
  }
  SE_REQUIRE(stages >= 2 && slots >= 3, "v2v_run: op %d: stem windows do not fit shared memory (side %d)", op_index, p.ld.side);
  p.win_stages = stages; p.w_slots = slots;
  p.off_w = 4u * p.win_bytes * stages;
  p.off_bias = p.off_w + (uint32_t)slots * STEM_CHUNK_BYTES;
  p.off_bar = p.off_bias + 128;
  p.n_items = (int)((n_pos + STEM_L - 1) / STEM_L);
  const size_t smem_bytes = (size_t)p.off_bar + 256;
  if (int rc = ensure_max_dynamic_smem((const void*)stem_s2d_tc_kernel<1>, (int)kMaxSmem)) return rc;
  if (int rc = ensure_max_dynamic_smem((const void*)stem_s2d_tc_kernel<2>, (int)kMaxSmem)) return rc;
  if (p.cg == 2) {
    const int groups = (p.n_items + 1) / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * (groups < kNumSMs / 2 ? groups : kNumSMs / 2))); cfg.blockDim = dim3(STEM_THREADS);
    cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, stem_s2d_tc_kernel<2>, p);
    if (e != cudaSuccess) { set_error("stem_s2d_tc (CTA pair): %s", cudaGetErrorString(e)); return SCENEEGO_E_CUDA; }
  } else {
    const int grid = p.n_items < kNumSMs ? p.n_items : kNumSMs;
    stem_s2d_tc_kernel<1><<<grid, STEM_THREADS, smem_bytes, st>>>(p);
  }
  SE_CUDA_LAUNCH_CHECK("stem_s2d_tc");
  return SCENEEGO_OK;
}

}  // namespace sceneego

using namespace sceneego;

extern "C" size_t sceneego_v2v_stem_s2d_weight_bytes(void) { return STEM_W_BYTES; }

extern "C" int sceneego_v2v_pack_stem_s2d(const float* h_weight, const float* h_bias, const float* h_gamma,
                                          const float* h_beta, const float* h_mean, const float* h_var, double eps,
                                          int n_split, uint16_t* h_w_out, float* h_b_out) {
  SE_REQUIRE(h_weight && h_w_out && h_b_out && (n_split == 1 || n_split == 2), "pack_stem_s2d: bad argument");
  const int NHp = 128 / n_split;                                    // columns per half-blob
  const size_t half_elems = STEM_W_BYTES / 2 / n_split;
  constexpr int CO = 16, CI = 33, K = 7;
  memset(h_w_out, 0, STEM_W_BYTES);
  double scale[CO];
  for (int co = 0; co < CO; ++co) {
    double sc = 1.0, sh = 0.0;
    if (h_gamma) {
      sc = (double)h_gamma[co] / sqrt((double)h_var[co] + eps);
      sh = (double)h_beta[co] - (double)h_mean[co] * sc;
    }
    scale[co] = sc;
    h_b_out[co] = (float)((h_bias ? (double)h_bias[co] : 0.0) * sc + sh);
  }
  auto W = [&](int co, int ci, int dx, int dy, int dz) -> uint16_t {
    if (dx < 0 || dx >= K || dy < 0 || dy >= K || dz < 0 || dz >= K) return 0;
    return f2bf((float)((double)h_weight[((((size_t)co * CI + ci) * K + dx) * K + dy) * K + dz] * scale[co]));
  };
  // feature stages: chunk = (stage s, c), tap j: [kchunk 4][n 128][8 channels]
  for (int s = 0; s < STEM_FEAT_STAGES; ++s) {
    const int ox = (s >> 2) - 3, py = (s >> 1) & 1, pz = s & 1;
    for (int tp = 0; tp < 16; ++tp) {
      const int by = (tp >> 2) - 1 - py, bz = (tp & 3) - 1 - pz;
      const int oy = 2 * by + py, oz = 2 * bz + pz;
      // within a half-blob: tap stride 8 KB / n_split, [kchunk 4][NHp][8]
      uint16_t* dst = h_w_out + (size_t)(s * 16 + tp) * STEM_FEAT_TAP_BYTES / 2 / n_split;
      for (int n = 0; n < 128; ++n) {
        const int sx = n >> 6, sy = (n >> 5) & 1, sz = (n >> 4) & 1, co = n & 15;
        const int hf = n / NHp, nn = n % NHp;
        for (int ci = 0; ci < 32; ++ci)
          dst[hf * half_elems + ((size_t)(ci >> 3) * NHp + nn) * 8 + (ci & 7)] = W(co, ci, ox - sx + 3, oy - sy + 3, oz - sz + 3);
      }
    }
  }
  // scene stages: tap tp = (by, bz pair): [kchunk 2 = block bz0, bz0+1][n 128][8 parities]
  for (int s = 0; s < STEM_SCENE_STAGES; ++s) {
    const int bx = s - 2;
    for (int tp = 0; tp < 15; ++tp) {
      const int by = tp / 3 - 2, bz0 = (tp % 3) * 2 - 2;
      uint16_t* dst = h_w_out + (STEM_SCENE_OFF + (size_t)(s * 16 + tp) * STEM_SCENE_TAP_BYTES) / 2 / n_split;
      for (int c2 = 0; c2 < 2; ++c2)
        for (int n = 0; n < 128; ++n) {
          const int sx = n >> 6, sy = (n >> 5) & 1, sz = (n >> 4) & 1, co = n & 15;
          const int hf = n / NHp, nn = n % NHp;
          for (int e = 0; e < 8; ++e) {
            const int ox = 2 * bx + (e >> 2), oy = 2 * by + ((e >> 1) & 1), oz = 2 * (bz0 + c2) + (e & 1);
            dst[hf * half_elems + ((size_t)c2 * NHp + nn) * 8 + e] = W(co, 32, ox - sx + 3, oy - sy + 3, oz - sz + 3);
          }
        }
    }
  }
  return SCENEEGO_OK;
}
