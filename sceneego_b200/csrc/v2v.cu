// a7: V2V encoder-decoder (network/v2v.py) on sm_100a.
//
// Data layout ("planar padded"): a volume with C channels is stored as C/8
// planes; a plane is a linear array of 16-byte cells (8 bf16 channels of one voxel)
// indexed by a POSITION  q = b*frame_pitch + guard + x*pitch_x + y*pitch_y + z,
// where every z-line is followed by `pad` zero cells and every x-plane by `pad` zero
// lines, and every frame is preceded by `guard` zero cells.  With the zeros physically
// present, a conv tap (dx,dy,dz) is the constant position offset dx*pitch_x+dy*pitch_y+dz:
// "same" padding needs no masks, and the A operand of the implicit GEMM for any tap is
// just the staged window shifted by a multiple of 16 bytes.
//
// conv_tc_kernel: persistent, warp-specialised implicit GEMM.
//   rows (M)   = 128 consecutive positions per UMMA tile, `tiles` tiles per work item
//   cols (N)   = xs * Cout (16..256): xs consecutive x-planes of outputs stacked against Toeplitz weights
//   reduction  = taps x Cin, 16 channels (two 8-channel planes) per tcgen05.mma
//   warp 0     = producer: cp.async.bulk (TMA 1-D) of the per-dx activation window
//                (one contiguous run per channel plane) and of the weight chunks
//   warps 1-4  = TMEM allocator + tcgen05.mma issuers, one tile (group) each: a single warp can issue an
//                MMA only every ~41.5 cycles (222 for cta_group::2), less than the pipe needs at N <= 64
//   next 8     = epilogue: tcgen05.ld -> bias / residual / ReLU / pad-zeroing -> bf16 planes
//   smem operands use the no-swizzle K-major canonical layout: 8 rows x 16 B core matrices,
//   SBO = 128 B between 8-row groups, LBO = plane stride between the two K chunks.
//   Accumulators: 2 x (tiles x N) fp32 columns of TMEM, double-buffered across work items.
//   CG = 2: two CTAs of a cluster share one MMA stream (M = 256), weights N-split across the pair.
// The 7^3 stem and the fused 1x1 tail have their own kernels (stem.cu, tail.cu).
#include "tc_common.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <mutex>

namespace sceneego {

// ---------------------------------------------------------------------------
// conv parameters
// ---------------------------------------------------------------------------

struct ConvParams {
  const __nv_bfloat16* src;
  __nv_bfloat16* dst;
  const __nv_bfloat16* res;
  float* dst_f32;
  const __nv_bfloat16* w;   // [tap][cin/8][cout][8]
  const float* bias;        // [cout]
  sceneego_vol_layout_t ls, ld;
  int batch;
  int k, r, cin_planes, ksteps, N, cout_real, flags;
  // x-stacking (XS > 1): one GEMM row produces the outputs of XS consecutive x-planes, N = xs * n0
  // columns, against weights Toeplitz-stacked over n_dx = k + xs - 1 input plane offsets.  Work
  // items are then plane-aligned: (frame, group of xs planes, chunk of L cells inside the plane).
  int xs, n0, n_dx, n_xg, items_per_plane;
  FastDiv fd_frame, fd_px, fd_py;   // by ls.frame_pitch, ls.pitch_x, ls.pitch_y
  // transposed conv k2 s2 on the same kernel: rows = input positions, every output parity is a "tap"
  // with its own weight block (no row shift) accumulating into its own n0 TMEM columns; the epilogue
  // scatters column block `par` to output voxel (2x+i, 2y+j, 2z+l).  One launch covers parities
  // [par0, par0 + taps_dx).
  int deconv, par0;
  int taps_dx;                      // taps per window stage: k*k (conv) or parities per launch (deconv)
  int mma_n;                        // N of one tcgen05.mma: xs*n0 (conv) or n0 (deconv)
  int tiles, L, WL;                 // tiles per item, positions per item, window length (positions)
  int n_items;
  int win_stages, w_slots, wchunk_taps, wchunks_per_dx;
  int wrows;                        // 3^3 only: whole dy-rows per weight chunk (0 = generic tap loop)
  uint32_t win_bytes, wchunk_bytes, tap_bytes;
  uint32_t off_win, off_w, off_bias, off_bar;
  uint32_t tmem_cols, half_cols;
  // CTA pair (cg = 2): two CTAs of one cluster run one tcgen05.mma.cta_group::2 stream (M = 256 = one
  // 128-row tile of each CTA) issued by the leader; each CTA stages its own windows and HALF of the
  // weight columns (B is N-split across the pair), which halves the weight traffic and the B-operand
  // shared-memory reads that bound N <= 128 MMAs.  Weights are packed half-major: [half][tap][cin/8][N/2][8].
  int cg;
  uint32_t w_half_bytes;            // bytes of one half-blob (cg = 2)
  // Fused projection shortcut (Res3DBlock.skip_con, network/v2v.py:32-38): a 1x1 conv of a SECOND source with
  // half the channels, accumulated into the same tile after the stencil taps -- n_dx2 = xs extra stages of one
  // tap each (input plane x0 + e feeds output plane block e).  Weights follow the stencil's inside (each half
  // of) the blob; the two folded biases are summed on the host.
  const __nv_bfloat16* src2;
  int n_dx2;
  uint32_t w_main_bytes;            // bytes of the stencil weights per (half-)blob = offset of the shortcut taps
  int march;                        // checker only: weights in the marching layout [tap(dy,dz)][cin/8][3*n0][8] (march.cu)
  int n0_shift;                     // log2(n0) when n0 is a power of two (the epilogue's column -> block index), else -1
};

constexpr int CONV_EPI_WARPS = 8;
// One warp can issue a tcgen05.mma only every ~41.5 cycles (tools/mma_replay.cu), which is longer than
// the tensor pipe needs for N <= 64 (48 cycles at N = 64).  The tiles of a work item are therefore
// split over up to four issuing warps, one per SM sub-partition.
__host__ __device__ constexpr int conv_mma_warps(int tiles) { return tiles >= 4 ? 4 : tiles; }
__host__ __device__ constexpr int conv_threads(int tiles) { return 32 * (1 + conv_mma_warps(tiles) + CONV_EPI_WARPS); }
constexpr int CONV_MAX_THREADS = 32 * (1 + 4 + CONV_EPI_WARPS);   // producer warp, <= 4 MMA warps, 8 epilogue warps
constexpr int MAX_STAGES = 4, MAX_WSLOTS = 8;

struct RowInfo {
  bool write;    // position belongs to a frame (valid voxel or in-frame pad)
  bool valid;    // a real voxel
  int b, n;      // frame and flat voxel index (valid only)
  int x, y, z;   // voxel coordinates in the source volume
  int64_t dpos;  // position in the destination layout
};

__device__ __forceinline__ RowInfo decode_row(const ConvParams& p, int64_t q64) {
  RowInfo ri;
  ri.write = false; ri.valid = false; ri.b = 0; ri.n = 0; ri.dpos = 0; ri.x = ri.y = ri.z = 0;
  const int S = p.ls.side;
  const uint32_t q = (uint32_t)q64;                       // batch * frame_pitch < 2^29 (checked on the host)
  const uint32_t b = fdiv(q, p.fd_frame);
  const int rem = (int)(q - b * (uint32_t)p.ls.frame_pitch) - p.ls.guard;
  if ((int)b >= p.batch || rem < 0) return ri;
  const int x = (int)fdiv((uint32_t)rem, p.fd_px);
  const int r2 = rem - x * p.ls.pitch_x;
  const int y = (int)fdiv((uint32_t)r2, p.fd_py);
  const int z = r2 - y * p.ls.pitch_y;
  if (x >= S) return ri;
  // cells of the destination layout reachable from this source cell: voxels and the
  // destination's own pad cells (kept zero so the next conv can use them as padding)
  if (y >= S + p.ld.pad || z >= S + p.ld.pad) return ri;
  ri.write = true;
  ri.valid = (y < S) && (z < S);
  ri.b = (int)b;
  ri.x = x; ri.y = y; ri.z = z;
  ri.n = (x * S + y) * S + z;
  // transposed conv: position of output voxel (2x, 2y, 2z); the epilogue adds the parity offset per column block
  const int m = p.deconv ? 2 : 1;
  ri.dpos = (int64_t)b * p.ld.frame_pitch + p.ld.guard + (int64_t)(m * x) * p.ld.pitch_x + (int64_t)(m * y) * p.ld.pitch_y + m * z;
  return ri;
}

// Plane-aligned variant (x-stacked items): `cell` is the index inside x-plane `x` of frame `b`.
__device__ __forceinline__ RowInfo decode_row_plane(const ConvParams& p, int b, int x, int cell) {
  RowInfo ri;
  ri.write = false; ri.valid = false; ri.b = b; ri.n = 0; ri.dpos = 0; ri.x = x; ri.y = ri.z = 0;
  const int S = p.ls.side;
  if (cell >= p.ls.pitch_x) return ri;
  const int y = (int)fdiv((uint32_t)cell, p.fd_py);
  const int z = cell - y * p.ls.pitch_y;
  if (y >= S + p.ld.pad || z >= S + p.ld.pad) return ri;
  ri.write = true;
  ri.valid = (y < S) && (z < S);
  ri.y = y; ri.z = z;
  ri.n = (x * S + y) * S + z;
  ri.dpos = (int64_t)b * p.ld.frame_pitch + p.ld.guard + (int64_t)x * p.ld.pitch_x + (int64_t)y * p.ld.pitch_y + z;
  return ri;
}

// Shared epilogue: v[0..15] are output channels c0..c0+15 of one row; r0/r1 hold the residual
// cells (8 channels each) already loaded by the caller (ignored unless a residual flag is set).
__device__ __forceinline__ void store_row16_pre(const ConvParams& p, const RowInfo& ri, int c0, float (&v)[16],
                                                const float* s_bias, const uint4& r0, const uint4& r1) {
  if (!ri.write) return;
  const int S = p.ls.side;
  if (p.flags & SCENEEGO_F_OUT_F32) {
    if (!ri.valid) return;
    const size_t N3 = (size_t)S * S * S;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int c = c0 + j;
      if (c < p.cout_real) {
        float o = v[j] + s_bias[c];
        if (p.flags & SCENEEGO_F_RELU) o = fmaxf(o, 0.f);
        p.dst_f32[((size_t)ri.b * p.cout_real + c) * N3 + ri.n] = o;
      }
    }
    return;
  }
  float bs[16];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c0 + 4 * j);
    bs[4 * j] = b4.x; bs[4 * j + 1] = b4.y; bs[4 * j + 2] = b4.z; bs[4 * j + 3] = b4.w;
  }
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    float o[8];
    const int64_t cell = (int64_t)((c0 >> 3) + g) * p.ld.plane_stride + ri.dpos;
    if (ri.valid) {
      float rr[8];
      unpack8(g ? r1 : r0, rr);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = v[8 * g + j] + bs[8 * g + j];
        if (p.flags & SCENEEGO_F_RESIDUAL) t += rr[j];
        if (p.flags & SCENEEGO_F_RELU) t = fmaxf(t, 0.f);
        if (p.flags & SCENEEGO_F_ADD_AFTER) t += rr[j];
        o[j] = t;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = 0.f;
    }
    *reinterpret_cast<uint4*>(p.dst + cell * 8) = pack8(o);
  }
}

// Residual cells of channels c0..c0+15 of one row (zeros when the op has no residual).
__device__ __forceinline__ void load_res16(const ConvParams& p, const RowInfo& ri, int c0, uint4& r0, uint4& r1) {
  r0 = make_uint4(0, 0, 0, 0);
  r1 = r0;
  if ((p.flags & (SCENEEGO_F_RESIDUAL | SCENEEGO_F_ADD_AFTER)) && ri.valid && !(p.flags & SCENEEGO_F_OUT_F32)) {
    const int64_t cell = (int64_t)(c0 >> 3) * p.ld.plane_stride + ri.dpos;
    r0 = *reinterpret_cast<const uint4*>(p.res + cell * 8);
    r1 = *reinterpret_cast<const uint4*>(p.res + (cell + p.ld.plane_stride) * 8);
  }
}

__device__ __forceinline__ void store_row16(const ConvParams& p, const RowInfo& ri, int c0, float (&v)[16],
                                            const float* s_bias) {
  uint4 r0, r1;
  load_res16(p, ri, c0, r0, r1);
  store_row16_pre(p, ri, c0, v, s_bias, r0, r1);
}

// ---------------------------------------------------------------------------
// tcgen05 implicit-GEMM conv.  KSTEPS = Cin/16 and TILES = 128-row tiles per work item are
// compile-time so that the per-tap block of TILES*KSTEPS MMAs is straight-line code whose
// descriptors differ by constants: the single issuing lane must sustain one tcgen05.mma per
// ~41 cycles (measured floor at N<=32, tools/mma_rate.cu).
// ---------------------------------------------------------------------------
template <int KSTEPS, int TILES, int XS, int WROWS, int CG>
__global__ void __launch_bounds__(conv_threads(TILES), 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  // warp index through a shuffle: tells ptxas it is warp-uniform, so the role branches below are
  // uniform branches and the MMA issuer's address arithmetic can live in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  float* s_bias = reinterpret_cast<float*>(smem + p.off_bias);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  // barrier map
  const uint32_t bar0 = sbase + p.off_bar;
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const int B_FULL_WIN = 0, B_EMPTY_WIN = MAX_STAGES, B_FULL_W = 2 * MAX_STAGES, B_EMPTY_W = 2 * MAX_STAGES + MAX_WSLOTS,
            B_TMEM_FULL = 2 * MAX_STAGES + 2 * MAX_WSLOTS, B_TMEM_EMPTY = B_TMEM_FULL + 2, B_COUNT = B_TMEM_EMPTY + 2;
  uint32_t* s_tmem_ptr = reinterpret_cast<uint32_t*>(bars + B_COUNT);

  constexpr int NW = conv_mma_warps(TILES);          // MMA-issuing warps: warps 1..NW
  constexpr int FIRST_EPI = 1 + NW;                  // epilogue warps: FIRST_EPI .. FIRST_EPI+7
  for (int i = threadIdx.x; i < p.n0; i += conv_threads(TILES)) s_bias[i] = p.bias[i];
  // CTA pair: rank 0 (leader) issues every MMA for both CTAs; its "full" barriers also count one arrival
  // relayed from the peer (the peer's data has landed), its accumulator-empty barrier counts both epilogues.
  const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  if (threadIdx.x == 0) {
    const int full_cnt = (CG == 2 && is_leader) ? 2 : 1;
    for (int i = 0; i < MAX_STAGES; ++i) { mbar_init(BAR(B_FULL_WIN + i), full_cnt); mbar_init(BAR(B_EMPTY_WIN + i), NW); }
    for (int i = 0; i < MAX_WSLOTS; ++i) { mbar_init(BAR(B_FULL_W + i), full_cnt); mbar_init(BAR(B_EMPTY_W + i), NW); }
    for (int i = 0; i < 2; ++i) { mbar_init(BAR(B_TMEM_FULL + i), NW); mbar_init(BAR(B_TMEM_EMPTY + i), CONV_EPI_WARPS * CG); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem_ptr)),
                   "r"(p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem_ptr)),
                   "r"(p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();       // the peer's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem_ptr;

  // work split: cluster c of n_cl clusters takes item groups c, c + n_cl, ...; CTA r of the pair takes item
  // CG * group + r (the last group of an odd item count repeats the last item: same values written twice)
  const int n_cl = (int)gridDim.x / CG, cl = (int)blockIdx.x / CG;
  const int n_groups = (p.n_items + CG - 1) / CG;
  const int my_items = (n_groups - cl + n_cl - 1) / n_cl;
  auto item_of = [&](int it) -> int {
    const int item = (cl + it * n_cl) * CG + (int)cta_rank;
    return item < p.n_items ? item : p.n_items - 1;
  };
  const int halo = p.r * (p.ls.pitch_y + 1);   // window starts `halo` positions before the item
  constexpr int L = TILES * 128;
  // first position of work item `item`; for x-stacked items also frame / first plane / first cell
  auto item_origin = [&](int item, int& b, int& x0, int& cell0) -> int64_t {
    if (XS == 1) { b = 0; x0 = 0; cell0 = 0; return (int64_t)p.ls.guard + (int64_t)item * L; }
    const int c = item % p.items_per_plane;
    const int t = item / p.items_per_plane;
    const int xg = t % p.n_xg;
    b = t / p.n_xg; x0 = xg * XS; cell0 = c * L;
    return (int64_t)b * p.ls.frame_pitch + p.ls.guard + (int64_t)x0 * p.ls.pitch_x + cell0;
  };

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      int ws = 0, wph = 0, sl = 0, sph = 0;
      for (int it = 0; it < my_items; ++it) {
        int ib, ix0, icell0;
        const int64_t q0 = item_origin(item_of(it), ib, ix0, icell0);
        for (int dx = 0; dx < p.n_dx; ++dx) {
          mbar_wait(BAR(B_EMPTY_WIN + ws), wph ^ 1);
          mbar_expect_tx(BAR(B_FULL_WIN + ws), p.win_bytes * (2 * KSTEPS));
          const int64_t qs = q0 + (int64_t)(dx - p.r) * p.ls.pitch_x - halo;
#pragma unroll
          for (int g = 0; g < 2 * KSTEPS; ++g)
            bulk_g2s(sbase + p.off_win + (uint32_t)(ws * 2 * KSTEPS + g) * p.win_bytes,
                     p.src + ((int64_t)g * p.ls.plane_stride + qs) * 8, p.win_bytes, BAR(B_FULL_WIN + ws));
          if (++ws == p.win_stages) { ws = 0; wph ^= 1; }
          for (int wc = 0; wc < p.wchunks_per_dx; ++wc) {
            mbar_wait(BAR(B_EMPTY_W + sl), sph ^ 1);
            mbar_expect_tx(BAR(B_FULL_W + sl), p.wchunk_bytes);
            const char* wsrc = reinterpret_cast<const char*>(p.w) + (size_t)cta_rank * p.w_half_bytes +
                               (size_t)(dx * p.wchunks_per_dx + wc) * p.wchunk_bytes;
            bulk_g2s(sbase + p.off_w + (uint32_t)sl * p.wchunk_bytes, wsrc, p.wchunk_bytes, BAR(B_FULL_W + sl));
            if (++sl == p.w_slots) { sl = 0; sph ^= 1; }
          }
        }
        if constexpr (KSTEPS % 2 == 0) {
          for (int e = 0; e < p.n_dx2; ++e) {          // fused 1x1 shortcut: KSTEPS planes of src2, one tap
            mbar_wait(BAR(B_EMPTY_WIN + ws), wph ^ 1);
            mbar_expect_tx(BAR(B_FULL_WIN + ws), p.win_bytes * KSTEPS);
            const int64_t qs = q0 + (int64_t)e * p.ls.pitch_x - halo;
#pragma unroll
            for (int g = 0; g < KSTEPS; ++g)
              bulk_g2s(sbase + p.off_win + (uint32_t)(ws * 2 * KSTEPS + g) * p.win_bytes,
                       p.src2 + ((int64_t)g * p.ls.plane_stride + qs) * 8, p.win_bytes, BAR(B_FULL_WIN + ws));
            if (++ws == p.win_stages) { ws = 0; wph ^= 1; }
            mbar_wait(BAR(B_EMPTY_W + sl), sph ^ 1);
            mbar_expect_tx(BAR(B_FULL_W + sl), p.tap_bytes / 2);
            bulk_g2s(sbase + p.off_w + (uint32_t)sl * p.wchunk_bytes,
                     reinterpret_cast<const char*>(p.w) + (size_t)cta_rank * p.w_half_bytes + p.w_main_bytes +
                         (size_t)e * (p.tap_bytes / 2),
                     p.tap_bytes / 2, BAR(B_FULL_W + sl));
            if (++sl == p.w_slots) { sl = 0; sph ^= 1; }
          }
        }
      }
    }
  } else if (warp <= NW) {
    // ===================== MMA issuers =====================
    // Warp w issues the MMAs of tiles (w-1), (w-1)+NW, ...; every issuing warp commits to the
    // window / weight / accumulator barriers, whose arrival counts are NW.
    // The whole warp runs this loop converged so that every descriptor is a warp-uniform value
    // (uniform registers, no per-MMA R2UR/ELECT loop); only the tcgen05 instructions themselves
    // are predicated on one elected lane, which is also the lane that commits.
    if (CG == 2 && !is_leader) {
      // peer CTA of a pair: no MMAs of its own.  Warp 1 relays "my window / weight half has landed" to the
      // leader's full barriers, in consumption order; warps 2..NW have nothing to do.
      if (warp == 1 && lane == 0) {
        int ws = 0, wph = 0, sl = 0, sph = 0;
        for (int it = 0; it < my_items; ++it) {
          for (int dx = 0; dx < p.n_dx; ++dx) {
            mbar_wait(BAR(B_FULL_WIN + ws), wph);
            mbar_arrive_remote(mapa_shared(BAR(B_FULL_WIN + ws), 0));
            if (++ws == p.win_stages) { ws = 0; wph ^= 1; }
            for (int wc = 0; wc < p.wchunks_per_dx; ++wc) {
              mbar_wait(BAR(B_FULL_W + sl), sph);
              mbar_arrive_remote(mapa_shared(BAR(B_FULL_W + sl), 0));
              if (++sl == p.w_slots) { sl = 0; sph ^= 1; }
            }
          }
          for (int e = 0; e < p.n_dx2; ++e) {          // shortcut stages
            mbar_wait(BAR(B_FULL_WIN + ws), wph);
            mbar_arrive_remote(mapa_shared(BAR(B_FULL_WIN + ws), 0));
            if (++ws == p.win_stages) { ws = 0; wph ^= 1; }
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
    // instruction descriptor: D=f32, A=B=bf16, both K-major, M=128 (256 over a CTA pair), N=p.mma_n
    const uint32_t idesc = (1u << 4) | kIdescAB | ((uint32_t)(p.mma_n >> 3) << 17) | ((CG == 2 ? 16u : 8u) << 24);
    // K-major no-swizzle: LBO = byte stride between the two 8-channel K chunks of one MMA,
    // SBO = byte stride between 8-row core matrices (validated on B200 hardware).
    // hi word: SBO = 128 B (>>4 = 8) | descriptor version 1 (bit 46 -> bit 14 of the hi word)
    const uint64_t desc_hi = (uint64_t)(8u | (1u << 14)) << 32;
    const uint32_t a_lo_flags = ((p.win_bytes >> 4) & 0x3FFFu) << 16;          // LBO = window plane stride
    const uint32_t n_mine = (uint32_t)p.mma_n / CG;                            // B rows staged in this CTA (N-split over a pair)
    const uint32_t b_lo_flags = ((n_mine * 16u >> 4) & 0x3FFFu) << 16;         // LBO = rows * 16 B
    const uint32_t a_ks_step = (2u * p.win_bytes) >> 4;                        // two planes per K step
    const uint32_t b_ks_step = (2u * n_mine * 16u) >> 4;
    const uint32_t b_tap_step = p.tap_bytes >> 4;
    const uint32_t n_cols = (uint32_t)p.N;
    const int pitch_y = p.ls.pitch_y, ksz = p.k, wtaps = p.wchunk_taps;
    const uint32_t is_deconv = (uint32_t)p.deconv, par_cols = (uint32_t)p.n0;
    const uint32_t my_tile = (uint32_t)(warp - 1);                             // first tile of this warp
    const uint32_t a_mine = my_tile * 128u;                                    // its row offset, in 16-B cells
    int ws = 0, wph = 0, sl = 0, sph = 0;
    for (int it = 0; it < my_items; ++it) {
      const int buf = it & 1;
      WAIT(BAR(B_TMEM_EMPTY + buf), ((it >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_mine = tmem_u + (uint32_t)buf * p.half_cols + my_tile * n_cols;
      uint32_t acc = 0;
      for (int dx = 0; dx < p.n_dx; ++dx) {
        WAIT(BAR(B_FULL_WIN + ws), wph);
        const uint32_t win_lo = (((sbase + p.off_win + (uint32_t)(ws * 2 * KSTEPS) * p.win_bytes) >> 4) & 0x3FFFu) |
                                a_lo_flags;
        int dy = 0, dz = 0;
        uint32_t dcol = 0;                                                     // deconv: parity column block
        for (int wc = 0; wc < p.wchunks_per_dx; ++wc) {
          WAIT(BAR(B_FULL_W + sl), sph);
          tc_fence_after();
          uint32_t b_lo = (((sbase + p.off_w + (uint32_t)sl * p.wchunk_bytes) >> 4) & 0x3FFFu) | b_lo_flags;
          if constexpr (WROWS == 0) {
            // generic chunk: `wtaps` taps, tap coordinates carried in running counters
            for (int tp = 0; tp < wtaps; ++tp) {
              const uint32_t a_lo = win_lo + (uint32_t)(dy * pitch_y + dz);    // tap shift, in 16-B cells
              if (leader) {
#pragma unroll
                for (int tt = 0; tt < TILES / NW; ++tt) {
#pragma unroll
                  for (int ks = 0; ks < KSTEPS; ++ks)
                    MMA(d_mine + (uint32_t)(tt * NW) * n_cols + dcol,
                                desc_hi | (a_lo + a_mine + (uint32_t)(tt * NW) * 128u + (uint32_t)ks * a_ks_step),
                                desc_hi | (b_lo + (uint32_t)ks * b_ks_step), idesc, ks == 0 ? acc : 1u);
                }
              }
              b_lo += b_tap_step;
              if (is_deconv) { dcol += par_cols; }                             // next parity: fresh columns, acc stays 0
              else { acc = 1; if (++dz == ksz) { dz = 0; ++dy; } }
            }
          } else {
            // 3^3 stencil, chunk = WROWS whole dy-rows of 3 taps: straight-line code, the tap shifts are
            // (dy + r) * pitch_y + j with compile-time r, j -- no per-tap loop control or divergence
            if (leader) {
#pragma unroll
              for (int r = 0; r < WROWS; ++r) {
                const uint32_t a_row = win_lo + a_mine + (uint32_t)((dy + r) * pitch_y);
#pragma unroll
                for (int j = 0; j < 3; ++j) {
#pragma unroll
                  for (int tt = 0; tt < TILES / NW; ++tt) {
#pragma unroll
                    for (int ks = 0; ks < KSTEPS; ++ks)
                      MMA(d_mine + (uint32_t)(tt * NW) * n_cols,
                                  desc_hi | (a_row + (uint32_t)j + (uint32_t)(tt * NW) * 128u + (uint32_t)ks * a_ks_step),
                                  desc_hi | (b_lo + (uint32_t)(r * 3 + j) * b_tap_step + (uint32_t)ks * b_ks_step), idesc,
                                  (ks == 0 && r == 0 && j == 0) ? acc : 1u);
                  }
                }
              }
            }
            acc = 1;
            dy += WROWS;
          }
          if (leader) COMMIT(BAR(B_EMPTY_W + sl));
          if (++sl == p.w_slots) { sl = 0; sph ^= 1; }
        }
        if (leader) COMMIT(BAR(B_EMPTY_WIN + ws));
        if (++ws == p.win_stages) { ws = 0; wph ^= 1; }
      }
      if constexpr (KSTEPS % 2 == 0) {
        for (int e = 0; e < p.n_dx2; ++e) {            // fused 1x1 shortcut: centre tap of the second source
          WAIT(BAR(B_FULL_WIN + ws), wph);
          WAIT(BAR(B_FULL_W + sl), sph);
          tc_fence_after();
          const uint32_t a_c = ((((sbase + p.off_win + (uint32_t)(ws * 2 * KSTEPS) * p.win_bytes) >> 4) & 0x3FFFu) | a_lo_flags) +
                               a_mine + (uint32_t)(p.r * (pitch_y + 1));
          const uint32_t b_c = (((sbase + p.off_w + (uint32_t)sl * p.wchunk_bytes) >> 4) & 0x3FFFu) | b_lo_flags;
          if (leader) {
#pragma unroll
            for (int tt = 0; tt < TILES / NW; ++tt) {
#pragma unroll
              for (int ks = 0; ks < KSTEPS / 2; ++ks)
                MMA(d_mine + (uint32_t)(tt * NW) * n_cols, desc_hi | (a_c + (uint32_t)(tt * NW) * 128u + (uint32_t)ks * a_ks_step),
                    desc_hi | (b_c + (uint32_t)ks * b_ks_step), idesc, 1u);
            }
          }
          if (leader) COMMIT(BAR(B_EMPTY_W + sl));
          if (++sl == p.w_slots) { sl = 0; sph ^= 1; }
          if (leader) COMMIT(BAR(B_EMPTY_WIN + ws));
          if (++ws == p.win_stages) { ws = 0; wph ^= 1; }
        }
      }
      if (leader) COMMIT(BAR(B_TMEM_FULL + buf));
    }
    __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    // Chunks of 16 output channels; the residual cells of chunk i+1 are requested before chunk i
    // is processed so that their global-load latency overlaps the TMEM read and the math.
    // Eight warps: two per TMEM lane quarter (a warp may only read lanes 32*(warp%4)..+31).  The pair
    // splits the item by tile parity (by column-chunk parity when the item is a single tile).
    const int quarter = warp & 3;
    const int half = (warp - FIRST_EPI) >> 2;
    const int nch = p.N >> 4;
    const int S = p.ls.side;
    const int t_first = TILES >= 2 ? half : 0, t_step = TILES >= 2 ? 2 : 1;
    const int c_first = TILES >= 2 ? 0 : half, c_step = TILES >= 2 ? 1 : 2;
    // x-stacked ops: column chunk c belongs to output plane x0 + (16c / n0), channels (16c % n0)..+15
    auto shifted = [&](const RowInfo& ri, int x0, int c, int& ch0) -> RowInfo {
      if (XS == 1 && p.deconv) {
        const int blk = p.n0_shift >= 0 ? ((c << 4) >> p.n0_shift) : (c << 4) / p.n0;
        const int par = p.par0 + blk;
        ch0 = (c << 4) - blk * p.n0;
        RowInfo r = ri;
        r.write = ri.valid;                                  // pads of the 2x volume are never touched
        r.dpos = ri.dpos + ((par >> 2) * p.ld.pitch_x + ((par >> 1) & 1) * p.ld.pitch_y + (par & 1));
        return r;
      }
      if (XS == 1) { ch0 = c << 4; return ri; }
      const int sft = p.n0_shift >= 0 ? ((c << 4) >> p.n0_shift) : (c << 4) / p.n0;
      ch0 = (c << 4) - sft * p.n0;
      RowInfo r = ri;
      const bool in = (x0 + sft) < S;
      r.write = ri.write && in;
      r.valid = ri.valid && in;
      r.dpos = ri.dpos + (int64_t)sft * p.ld.pitch_x;
      r.n = ri.n + sft * S * S;
      return r;
    };
    for (int it = 0; it < my_items; ++it) {
      const int buf = it & 1;
      int ib, ix0, icell0;
      const int64_t q0 = item_origin(item_of(it), ib, ix0, icell0);
      auto row_info = [&](int t) -> RowInfo {
        const int r = t * 128 + quarter * 32 + lane;
        return XS == 1 ? decode_row(p, q0 + r) : decode_row_plane(p, ib, ix0, icell0 + r);
      };
      RowInfo ri = row_info(t_first);
      if constexpr (TILES == 1) {
        // single-tile items (transposed convs: N = 256 against 128 rows; the deep, small levels) are short, so the
        // residual / skip cells of FOUR column chunks are kept in flight -- one chunk ahead left the 64->32
        // transposed conv latency-bound at ~1 TB/s on its skip tensor
        constexpr int D = 4;
        uint4 q0r[D], q1r[D];
#pragma unroll
        for (int d = 0; d < D; ++d) {
          q0r[d] = make_uint4(0, 0, 0, 0); q1r[d] = q0r[d];
          const int c = c_first + d * c_step;
          if (c < nch) { int chn; const RowInfo rp = shifted(ri, ix0, c, chn); load_res16(p, rp, chn, q0r[d], q1r[d]); }
        }
        mbar_wait(BAR(B_TMEM_FULL + buf), (it >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)buf * p.half_cols;
        int cpf = c_first + D * c_step;
#pragma unroll 1
        for (int c = c_first; c < nch; c += c_step) {
          const uint4 r0 = q0r[0], r1 = q1r[0];
#pragma unroll
          for (int d = 0; d + 1 < D; ++d) { q0r[d] = q0r[d + 1]; q1r[d] = q1r[d + 1]; }
          if (cpf < nch) { int chn; const RowInfo rp = shifted(ri, ix0, cpf, chn); load_res16(p, rp, chn, q0r[D - 1], q1r[D - 1]); }
          cpf += c_step;
          int ch0;
          const RowInfo rs = shifted(ri, ix0, c, ch0);
          uint32_t raw[16];
          tc_ld16(taddr + (uint32_t)(c << 4), raw);
          tc_wait_ld();
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
          store_row16_pre(p, rs, ch0, v, s_bias, r0, r1);
        }
      } else {
      uint4 rn0 = make_uint4(0, 0, 0, 0), rn1 = rn0;
      if (c_first < nch) {
        int ch0;
        const RowInfo rs = shifted(ri, ix0, c_first, ch0);
        load_res16(p, rs, ch0, rn0, rn1);
      }
      mbar_wait(BAR(B_TMEM_FULL + buf), (it >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int t = t_first; t < TILES; t += t_step) {
        RowInfo ri_next = ri;
        if (t + t_step < TILES) ri_next = row_info(t + t_step);
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)buf * p.half_cols +
                               (uint32_t)(t * p.N);
        for (int c = c_first; c < nch; c += c_step) {
          const uint4 r0 = rn0, r1 = rn1;
          int ch0, chn;
          const RowInfo rs = shifted(ri, ix0, c, ch0);
          if (c + c_step < nch) { const RowInfo rp = shifted(ri, ix0, c + c_step, chn); load_res16(p, rp, chn, rn0, rn1); }
          else if (t + t_step < TILES) { const RowInfo rp = shifted(ri_next, ix0, c_first, chn); load_res16(p, rp, chn, rn0, rn1); }
          uint32_t raw[16];
          tc_ld16(taddr + (uint32_t)(c << 4), raw);
          tc_wait_ld();
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]);
          store_row16_pre(p, rs, ch0, v, s_bias, r0, r1);
        }
        ri = ri_next;
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_remote(mapa_shared(BAR(B_TMEM_EMPTY + buf), 0));   // the leader issues for both
        else mbar_arrive(BAR(B_TMEM_EMPTY + buf));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();        // both CTAs are done with the pair's tensor memory
  if (warp == 1) {
    if constexpr (CG == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

typedef void (*conv_tc_fn)(const ConvParams);
// CTA-pair instantiations: the full-resolution 3^3 layers with N = 64 (where the B operand's shared-memory
// read is a quarter of the MMA's operand traffic) and the 64-channel layers at half resolution.
static conv_tc_fn pick_conv_tc_pair(int ksteps, int tiles, int xs, int wrows) {
  if (xs == 2 && ksteps == 2 && tiles == 4 && wrows == 3) return conv_tc_kernel<2, 4, 2, 3, 2>;
  if (xs == 2 && ksteps == 2 && tiles == 4 && wrows == 1) return conv_tc_kernel<2, 4, 2, 1, 2>;
  if (xs == 2 && ksteps == 1 && tiles == 4 && wrows == 3) return conv_tc_kernel<1, 4, 2, 3, 2>;
  if (xs == 1 && ksteps == 4 && tiles == 4 && wrows == 3) return conv_tc_kernel<4, 4, 1, 3, 2>;
  if (xs == 1 && ksteps == 4 && tiles == 4 && wrows == 1) return conv_tc_kernel<4, 4, 1, 1, 2>;
  if (xs == 1 && ksteps == 2 && tiles == 4 && wrows == 3) return conv_tc_kernel<2, 4, 1, 3, 2>;
  if (xs == 1 && ksteps == 2 && tiles == 4 && wrows == 0) return conv_tc_kernel<2, 4, 1, 0, 2>;
  return nullptr;
}
template <int WROWS>
static conv_tc_fn pick_conv_tc_w(int ksteps, int tiles, int xs) {
  if (WROWS == 0 && xs == 4 && ksteps == 5 && tiles == 2) return conv_tc_kernel<5, 2, 4, 0, 1>;   // 65-channel stem
  if (WROWS == 0 && xs == 4 && ksteps == 5 && tiles == 1) return conv_tc_kernel<5, 1, 4, 0, 1>;   // (with_intersection)
  if (xs == 4 && ksteps == 3 && tiles == 4) return conv_tc_kernel<3, 4, 4, WROWS, 1>;
  if (xs == 4 && ksteps == 3 && tiles == 2) return conv_tc_kernel<3, 2, 4, WROWS, 1>;
  if (xs == 4 && ksteps == 2 && tiles == 4) return conv_tc_kernel<2, 4, 4, WROWS, 1>;
  if (xs == 4 && ksteps == 2 && tiles == 2) return conv_tc_kernel<2, 2, 4, WROWS, 1>;
  if (xs == 2 && ksteps == 2 && tiles == 4) return conv_tc_kernel<2, 4, 2, WROWS, 1>;
  if (xs == 2 && ksteps == 1 && tiles == 4) return conv_tc_kernel<1, 4, 2, WROWS, 1>;
  if (xs == 2 && ksteps == 4 && tiles == 2) return conv_tc_kernel<4, 2, 2, WROWS, 1>;
  if (xs != 1) return nullptr;
#define SE_CASE(K, T) if (ksteps == K && tiles == T) return conv_tc_kernel<K, T, 1, WROWS, 1>;
  SE_CASE(2, 4) SE_CASE(4, 4) SE_CASE(4, 2) SE_CASE(8, 2) SE_CASE(1, 4) SE_CASE(1, 2) SE_CASE(2, 2)
  SE_CASE(8, 1) SE_CASE(4, 1) SE_CASE(2, 1) SE_CASE(1, 1) SE_CASE(1, 8) SE_CASE(2, 8)
  if (WROWS == 0) {
    SE_CASE(3, 4) SE_CASE(3, 8) SE_CASE(3, 2) SE_CASE(3, 1)
  }
#undef SE_CASE
  return nullptr;
}
static conv_tc_fn pick_conv_tc(int ksteps, int tiles, int xs, int wrows) {
  if (wrows == 3) return pick_conv_tc_w<3>(ksteps, tiles, xs);
  if (wrows == 1) return pick_conv_tc_w<1>(ksteps, tiles, xs);
  return pick_conv_tc_w<0>(ksteps, tiles, xs);
}

// ---------------------------------------------------------------------------
// CUDA-core checker conv: same layouts, same packed weights, same epilogue.  Used for the
// tiny deep levels' validation and to bisect tensor-core descriptor errors (op.impl = 1).
// One thread = one position x 16 output channels.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) conv_simt_kernel(const __grid_constant__ ConvParams p, int64_t n_pos) {
  __shared__ __align__(16) float s_bias[128];
  for (int i = threadIdx.x; i < p.n0; i += blockDim.x) s_bias[i] = p.bias[i];
  __syncthreads();
  const int64_t q = (int64_t)p.ls.guard + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = blockIdx.y * 16;
  if (q - p.ls.guard >= n_pos) return;
  const RowInfo ri = decode_row(p, q);
  if (!ri.write) return;
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  if (ri.valid) {
    for (int dx = 0; dx < p.k; ++dx)
      for (int dy = 0; dy < p.k; ++dy)
        for (int dz = 0; dz < p.k; ++dz) {
          const int tap = (dx * p.k + dy) * p.k + dz;
          const int64_t qs = q + (int64_t)(dx - p.r) * p.ls.pitch_x + (int64_t)(dy - p.r) * p.ls.pitch_y + (dz - p.r);
          for (int g = 0; g < p.cin_planes; ++g) {
            float a[8];
            unpack8(*reinterpret_cast<const uint4*>(p.src + ((int64_t)g * p.ls.plane_stride + qs) * 8), a);
            // p.N = xs*n0 columns; CTA-pair blobs are packed half-major ([half][tap][cin/8][N/2][8])
            const int nh = p.cg == 2 ? p.N / 2 : p.N, hf = c0 / nh;
            const uint4* wrow = p.march
                ? reinterpret_cast<const uint4*>(p.w) + ((size_t)(dy * 3 + dz) * p.cin_planes + g) * (3 * p.n0) + (2 - dx) * p.n0 + c0
                : reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(p.w) + (size_t)hf * p.w_half_bytes) +
                      ((size_t)tap * p.cin_planes + g) * nh + (c0 - hf * nh);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float wv[8];
              unpack8(__ldg(wrow + j), wv);
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[j] = fmaf(a[i], wv[i], acc[j]);
            }
          }
        }
  }
  if (ri.valid && p.n_dx2 > 0) {   // fused 1x1 shortcut of the second source (column block 0 of its first tap)
    const int nh = p.cg == 2 ? p.N / 2 : p.N, hf = c0 / nh;
    const int planes2 = p.cin_planes / 2;
    for (int g = 0; g < planes2; ++g) {
      float a[8];
      unpack8(*reinterpret_cast<const uint4*>(p.src2 + ((int64_t)g * p.ls.plane_stride + q) * 8), a);
      const uint4* wrow = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(p.w) + (size_t)hf * p.w_half_bytes + p.w_main_bytes) +
                          (size_t)g * nh + (c0 - hf * nh);
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float wv[8];
        unpack8(__ldg(wrow + j), wv);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j] = fmaf(a[i], wv[i], acc[j]);
      }
    }
  }
  store_row16(p, ri, c0, acc, s_bias);
}

// ---------------------------------------------------------------------------
// 2x2x2 max-pool (network/v2v.py:46-52).  One thread = one output cell (8 channels).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool2_kernel(const __nv_bfloat16* __restrict__ src,
                                                      __nv_bfloat16* __restrict__ dst, sceneego_vol_layout_t ls,
                                                      sceneego_vol_layout_t ld, int planes) {
  const int So = ld.side;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int g = blockIdx.y, b = blockIdx.z;
  if (n >= So * So * So || g >= planes) return;
  const int z = n % So, y = (n / So) % So, x = n / (So * So);
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const int64_t q = vol_pos(ls, b, 2 * x + (t >> 2), 2 * y + ((t >> 1) & 1), 2 * z + (t & 1));
    float a[8];
    unpack8(*reinterpret_cast<const uint4*>(src + ((int64_t)g * ls.plane_stride + q) * 8), a);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], a[j]);
  }
  *reinterpret_cast<uint4*>(dst + ((int64_t)g * ld.plane_stride + vol_pos(ld, b, x, y, z)) * 8) = pack8(m);
}

// ---------------------------------------------------------------------------
// ConvTranspose3d k2 s2 + folded BN + ReLU, then + skip (network/v2v.py:55-67,125-137).
// out[co, 2x+i, 2y+j, 2z+l] = relu(sum_ci W[parity][ci][co] * in[ci,x,y,z] + b[co]) + skip.
// One thread = one output cell x 8 output channels; weights packed like a conv with 8 taps.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) deconv2_kernel(const __grid_constant__ ConvParams p) {
  const int So = p.ld.side;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int go = blockIdx.y, b = blockIdx.z;
  if (n >= So * So * So) return;
  const int z = n % So, y = (n / So) % So, x = n / (So * So);
  const int tap = ((x & 1) * 2 + (y & 1)) * 2 + (z & 1);
  const int64_t qs = vol_pos(p.ls, b, x >> 1, y >> 1, z >> 1);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int g = 0; g < p.cin_planes; ++g) {
    float a[8];
    unpack8(*reinterpret_cast<const uint4*>(p.src + ((int64_t)g * p.ls.plane_stride + qs) * 8), a);
    // weights: [group of npar parities][cin/8][npar * cout][8]   (p.N = cout here)
    int npar = 256 / p.N;
    if (npar > 8) npar = 8;
    const uint4* wrow = reinterpret_cast<const uint4*>(p.w) +
                        ((size_t)(tap / npar) * p.cin_planes + g) * ((size_t)npar * p.N) + (size_t)(tap % npar) * p.N + go * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float wv[8];
      unpack8(__ldg(wrow + j), wv);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[j] = fmaf(a[i], wv[i], acc[j]);
    }
  }
  const int64_t cell = (int64_t)go * p.ld.plane_stride + vol_pos(p.ld, b, x, y, z);
  float rr[8];
  if (p.flags & SCENEEGO_F_ADD_AFTER) unpack8(*reinterpret_cast<const uint4*>(p.res + cell * 8), rr);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float t = acc[j] + p.bias[go * 8 + j];
    if (p.flags & SCENEEGO_F_RELU) t = fmaxf(t, 0.f);
    if (p.flags & SCENEEGO_F_ADD_AFTER) t += rr[j];
    acc[j] = t;
  }
  *reinterpret_cast<uint4*>(p.dst + cell * 8) = pack8(acc);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";
int ensure_max_dynamic_smem(const void* kernel, int bytes) {
  struct Entry { int dev; const void* fn; int bytes; };
  static Entry seen[512];
  static int n_seen = 0;
  static std::mutex mu;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return SCENEEGO_E_CUDA; }
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < n_seen; ++i)
    if (seen[i].dev == dev && seen[i].fn == kernel && seen[i].bytes >= bytes) return SCENEEGO_OK;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(MaxDynamicSharedMemorySize): %s", cudaGetErrorString(e)); return SCENEEGO_E_CUDA; }
  if (n_seen < 512) seen[n_seen++] = Entry{dev, kernel, bytes};
  return SCENEEGO_OK;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static thread_local int g_launches = 0;

// Choose tiles / stages / weight chunking for a conv and fill the kernel parameters.
// SCENEEGO_TILES / SCENEEGO_STAGES override the choice (tuning only).
static bool try_plan(ConvParams& p, int tiles, int want_stages) {
  const int taps_dx = p.taps_dx;
  const int L = tiles * 128;
  const int WL = L + 2 * p.r * (p.ls.pitch_y + 1);
  const uint32_t win_bytes = (uint32_t)WL * 16u;
  const uint32_t stage = win_bytes * p.cin_planes;
  const uint32_t fixed = 1024;  // bias + barriers + tmem ptr
  if ((uint64_t)stage * want_stages + fixed > kMaxSmem) return false;
  const uint32_t used = stage * want_stages + fixed;
  const uint32_t room = kMaxSmem - used;
  // weight chunk.  3^3 stencils: whole dy-rows (3 taps) per chunk so the issue loop is straight-line
  // code -- all 9 taps of a dx when three slots of that size fit, else one row; anything else: as many
  // taps of one dx as fit ~24 KB (must divide the taps of a dx).
  int wct = taps_dx, wrows = 0;
  if (p.k == 3 && !p.deconv && 3u * 9u * p.tap_bytes <= room) { wct = 9; wrows = 3; }
  else if (p.k == 3 && !p.deconv && 2u * 3u * p.tap_bytes <= room) { wct = 3; wrows = 1; }
  else {
    while (wct > 1 && (uint32_t)wct * p.tap_bytes > 24576u) {
      int nx = wct - 1;
      while (nx > 1 && taps_dx % nx) --nx;
      wct = nx;
    }
  }
  { const char* e = getenv("SCENEEGO_WCHUNK"); if (e && atoi(e) > 0 && taps_dx % atoi(e) == 0) { wct = atoi(e); wrows = (p.k == 3 && !p.deconv && wct % 3 == 0) ? wct / 3 : 0; if (wrows == 2) wrows = 0; } }
  const uint32_t wchunk = (uint32_t)wct * p.tap_bytes;
  if (2ull * wchunk > room) return false;
  int slots = (int)(room / wchunk);
  if (slots > MAX_WSLOTS) slots = MAX_WSLOTS;
  if (slots > 4 && wchunk > 8192) slots = 4;
  { const char* e = getenv("SCENEEGO_WSLOTS"); if (e && atoi(e) >= 2 && atoi(e) <= MAX_WSLOTS && (uint64_t)used + (uint64_t)atoi(e) * wchunk <= kMaxSmem) slots = atoi(e); }
  p.wrows = wrows;
  p.tiles = tiles; p.L = L; p.WL = WL;
  p.win_bytes = win_bytes; p.win_stages = want_stages;
  p.wchunk_taps = wct; p.wchunks_per_dx = taps_dx / wct; p.wchunk_bytes = wchunk; p.w_slots = slots;
  p.off_win = 0;
  p.off_w = stage * want_stages;
  p.off_bias = p.off_w + wchunk * slots;
  p.off_bar = p.off_bias + 512;
  p.half_cols = (uint32_t)(tiles * p.N);
  uint32_t cols = 32;
  while (cols < 2 * p.half_cols) cols <<= 1;
  p.tmem_cols = cols;
  return true;
}

static int plan_conv(ConvParams& p, int64_t n_pos) {
  p.tap_bytes = (uint32_t)p.cin_planes * (p.mma_n / p.cg) * 16u;   // per CTA; mma_n includes the x-stacking factor
  int max_tiles = 256 / p.N;
  if (max_tiles > 8) max_tiles = 8;
  // the deep, small levels (S <= 8 at 64 frames) have fewer work items than SMs: smaller items put more CTAs to
  // work on them (each streams the whole weight array anyway; what counts there is latency, not reuse)
  // and while there are only a few rounds of items, the tile count with the fewest (rounds x tiles) wins: 175 two-tile
  // items are two rounds on 148 SMs, 350 one-tile items three rounds of half the length
  if (p.xs == 1 && p.cg == 1) {
    auto items_of = [&](int t) { return (n_pos + (int64_t)t * 128 - 1) / ((int64_t)t * 128); };
    if (items_of(max_tiles) < 4 * kNumSMs) {
      int best = max_tiles;
      int64_t best_cost = ((items_of(max_tiles) + kNumSMs - 1) / kNumSMs) * max_tiles;
      for (int t = max_tiles >> 1; t >= 1; t >>= 1) {
        const int64_t cost = ((items_of(t) + kNumSMs - 1) / kNumSMs) * t;
        if (cost < best_cost) { best_cost = cost; best = t; }
      }
      max_tiles = best;
    }
  }
  const char* et = getenv("SCENEEGO_TILES");
  const char* es = getenv("SCENEEGO_STAGES");
  if (et || es) {
    const int tiles = et ? atoi(et) : max_tiles;
    const int stages = es ? atoi(es) : 2;
    if (tiles >= 1 && tiles <= max_tiles && stages >= 1 && stages <= MAX_STAGES && try_plan(p, tiles, stages))
      return (int)(p.off_bar + 512);
  }
  // measured on B200 (tools/tune_conv.py, tools/pair_sweep.py): the largest item wins; a 2-stage window ring is
  // enough for the single-CTA 3^3 / 7^3 layers, but CTA pairs (the peer's "landed" signal is relayed through
  // the leader: one more hop of latency) and the 16-channel input layer (short stages) want 3; 1x1 convs have
  // no halo, so their small windows get a deeper ring
  const int first_stages = p.k == 1 ? 4 : (p.cg == 2 || p.cin_planes <= 2) ? 3 : 2;
  for (int tiles = max_tiles; tiles >= 1; tiles >>= 1)
    for (int stages = first_stages; stages >= 2; --stages)
      if (try_plan(p, tiles, stages)) return (int)(p.off_bar + 512);
  return -1;
}

}  // namespace sceneego

using namespace sceneego;

extern "C" int sceneego_abi_version(void) { return SCENEEGO_ABI_VERSION; }
extern "C" int sceneego_act_dtype(void) { return kActDtype; }
extern "C" const char* sceneego_last_error(void) { return g_err; }
extern "C" int sceneego_v2v_last_launch_count(void) { return g_launches; }

extern "C" int sceneego_v2v_pack_conv(const float* h_weight, const float* h_bias, const float* h_gamma,
                                      const float* h_beta, const float* h_mean, const float* h_var, double eps,
                                      int cout, int cin, int ksize, int transposed, int cout_pad, int cin_pad,
                                      int xstack, int n_split, uint16_t* h_w_out, float* h_b_out) {
  SE_REQUIRE(h_weight && h_w_out && h_b_out, "pack_conv: null argument");
  SE_REQUIRE(cout_pad >= cout && cin_pad >= cin && cout_pad % 8 == 0 && cin_pad % 8 == 0, "pack_conv: bad padding");
  SE_REQUIRE(xstack >= 1 && (xstack == 1 || !transposed), "pack_conv: bad xstack");
  SE_REQUIRE(n_split == 1 || (n_split == 2 && !transposed && (xstack * cout_pad) % 32 == 0), "pack_conv: bad n_split");
  const int taps = ksize * ksize * ksize;
  const int kk = ksize * ksize;
  const int n_stk = xstack * cout_pad;                 // columns of the stacked B operand
  const int n_dx = ksize + xstack - 1;                 // input plane offsets
  memset(h_w_out, 0, (size_t)n_dx * kk * cin_pad * n_stk * sizeof(uint16_t));
  for (int co = 0; co < cout_pad; ++co) {
    double scale = 1.0, shift = 0.0;
    if (co < cout) {
      if (h_gamma) {
        scale = (double)h_gamma[co] / sqrt((double)h_var[co] + eps);
        shift = (double)h_beta[co] - (double)h_mean[co] * scale;
      }
      h_b_out[co] = (float)((h_bias ? (double)h_bias[co] : 0.0) * scale + shift);
    } else {
      h_b_out[co] = 0.f;
    }
    if (co >= cout) continue;
    for (int ci = 0; ci < cin; ++ci)
      for (int t = 0; t < taps; ++t) {
        // Conv3d (cout,cin,kx,ky,kz); ConvTranspose3d (cin,cout,kx,ky,kz): tap = output parity
        const size_t src = transposed ? (((size_t)ci * cout + co) * taps + t) : (((size_t)co * cin + ci) * taps + t);
        const uint16_t wv = f2bf((float)((double)h_weight[src] * scale));
        // tap t = (dx, dy, dz); output-plane shift s reads input plane offset dxp = dx + s
        const int dx = t / kk, rest = t % kk;
        if (transposed) {
          // output parity t belongs to launch group t / npar; inside a group the parities are stacked along N
          int npar = 256 / cout_pad;
          if (npar > 8) npar = 8;
          const size_t grp = (size_t)t / npar, nrow = (size_t)(t % npar) * cout_pad + co;
          h_w_out[((grp * (cin_pad / 8) + ci / 8) * ((size_t)npar * cout_pad) + nrow) * 8 + (ci & 7)] = wv;
          continue;
        }
        for (int sft = 0; sft < xstack; ++sft) {
          const size_t tp = (size_t)(dx + sft) * kk + rest;
          const size_t col = (size_t)sft * cout_pad + co;
          const size_t nh = (size_t)n_stk / n_split, hf = col / nh;               // n_split = 2: half-major blob
          const size_t half_elems = (size_t)n_dx * kk * cin_pad * nh;
          const size_t dst = hf * half_elems + ((tp * (cin_pad / 8) + ci / 8) * nh + (col - hf * nh)) * 8 + (ci & 7);
          h_w_out[dst] = wv;
        }
      }
  }
  return SCENEEGO_OK;
}

static int launch_conv_tc(ConvParams& p, int batch, int op_index, cudaStream_t st) {
  p.n0_shift = -1;
  for (int k = 3; k < 10; ++k) if ((1 << k) == p.n0) p.n0_shift = k;
  const int64_t n_pos = (int64_t)batch * p.ls.frame_pitch;
  // fdiv() is exact while n * d < 2^48 and n fits 32 bits
  SE_REQUIRE(n_pos + 4096 < (1ll << 31) && (n_pos + 4096) * (int64_t)p.ls.frame_pitch < (1ll << 48),
             "v2v_run: op %d: batch * frame_pitch too large for one launch (use a smaller chunk)", op_index);
  p.fd_frame = make_fastdiv((uint32_t)p.ls.frame_pitch);
  p.fd_px = make_fastdiv((uint32_t)p.ls.pitch_x);
  p.fd_py = make_fastdiv((uint32_t)p.ls.pitch_y);
  const int smem = plan_conv(p, n_pos);
  SE_REQUIRE(smem > 0, "v2v_run: op %d does not fit shared memory", op_index);
  if (p.xs == 1) {
    p.n_items = (int)((n_pos + p.L - 1) / p.L);
  } else {
    p.items_per_plane = (p.ls.pitch_x + p.L - 1) / p.L;
    p.n_xg = (p.ls.side + p.xs - 1) / p.xs;
    p.n_items = batch * p.n_xg * p.items_per_plane;
  }
  int grid = p.n_items < kNumSMs ? p.n_items : kNumSMs;
  conv_tc_fn fn = nullptr;
  p.w_main_bytes = (uint32_t)(p.n_dx * p.taps_dx) * p.tap_bytes;
  p.w_half_bytes = p.w_main_bytes + (uint32_t)p.n_dx2 * (p.tap_bytes / 2);
  if (p.cg == 2) {
    const int groups = (p.n_items + 1) / 2;
    grid = 2 * (groups < kNumSMs / 2 ? groups : kNumSMs / 2);
    fn = pick_conv_tc_pair(p.ksteps, p.tiles, p.xs, p.wrows);
  } else {
    fn = pick_conv_tc(p.ksteps, p.tiles, p.xs, p.wrows);
    if (fn == nullptr && p.wrows != 0) { p.wrows = 0; fn = pick_conv_tc(p.ksteps, p.tiles, p.xs, 0); }   // same chunking, generic tap loop
  }
  SE_REQUIRE(fn != nullptr, "v2v_run: op %d: no conv_tc instantiation for ksteps=%d tiles=%d xs=%d wrows=%d cta_pair=%d", op_index,
             p.ksteps, p.tiles, p.xs, p.wrows, p.cg);
  if (int rc = ensure_max_dynamic_smem((const void*)fn, (int)kMaxSmem)) return rc;
  if (p.cg == 2) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)conv_threads(p.tiles));
    cfg.dynamicSmemBytes = (size_t)smem; cfg.stream = st;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, fn, p);
    if (e != cudaSuccess) { set_error("conv_tc (CTA pair): %s", cudaGetErrorString(e)); return SCENEEGO_E_CUDA; }
  } else {
    // only the bytes the plan uses: leaves room for a small concurrent kernel (the output-#2 materialisation on a
    // side stream) to share the SM
    fn<<<grid, conv_threads(p.tiles), (size_t)smem, st>>>(p);
  }
  SE_CUDA_LAUNCH_CHECK("conv_tc");
  ++g_launches;
  return SCENEEGO_OK;
}

static int v2v_run_impl(const sceneego_v2v_op_t* ops, int n_ops, void* const* d_buffers, const void* d_blob,
                        int batch, void* stream, cudaEvent_t* ev) {
  SE_REQUIRE(ops && d_buffers && d_blob && n_ops > 0 && batch > 0, "v2v_run: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const char* force_simt = getenv("SCENEEGO_FORCE_SIMT");
  g_launches = 0;
  for (int i = 0; i < n_ops; ++i) {
    if (ev) cudaEventRecord(ev[i], st);
    const sceneego_v2v_op_t& op = ops[i];
    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.src = (const __nv_bfloat16*)d_buffers[op.src];
    p.dst = (__nv_bfloat16*)d_buffers[op.dst];
    p.dst_f32 = (float*)d_buffers[op.dst];
    p.res = op.res >= 0 ? (const __nv_bfloat16*)d_buffers[op.res] : nullptr;
    p.ls = op.lay_src; p.ld = op.lay_dst; p.batch = batch; p.flags = op.flags;
    SE_REQUIRE(p.src && p.dst, "v2v_run: op %d has a null buffer", i);
    SE_REQUIRE(!(op.flags & (SCENEEGO_F_RESIDUAL | SCENEEGO_F_ADD_AFTER)) || p.res, "v2v_run: op %d needs a residual", i);
    if (op.type == SCENEEGO_OP_MAXPOOL2) {
      SE_REQUIRE(op.lay_dst.side * 2 == op.lay_src.side && op.cin % 8 == 0, "v2v_run: op %d bad pool shape", i);
      const int So = op.lay_dst.side;
      dim3 grid((So * So * So + 255) / 256, op.cin / 8, batch);
      maxpool2_kernel<<<grid, 256, 0, st>>>(p.src, p.dst, p.ls, p.ld, op.cin / 8);
      SE_CUDA_LAUNCH_CHECK("maxpool2");
      ++g_launches;
      continue;
    }
    if (op.type == SCENEEGO_OP_STEM7_S2D || op.type == SCENEEGO_OP_TAIL_MLP || op.type == SCENEEGO_OP_STEM7_MARCH) {
      const bool simt = op.impl == 1 || force_simt;
      const int rc = op.type == SCENEEGO_OP_STEM7_S2D ? launch_stem_s2d(op, d_buffers, d_blob, batch, i, simt, st)
                     : op.type == SCENEEGO_OP_STEM7_MARCH ? launch_stem_march(op, d_buffers, d_blob, batch, i, simt, st)
                                                          : launch_tail_mlp(op, d_buffers, d_blob, batch, i, simt, st);
      if (rc != SCENEEGO_OK) return rc;
      ++g_launches;
      continue;
    }
    p.w = (const __nv_bfloat16*)((const char*)d_blob + op.w_offset);
    p.bias = (const float*)((const char*)d_blob + op.b_offset);
    p.cin_planes = op.cin / 8; p.ksteps = op.cin / 16; p.N = op.cout; p.cout_real = op.cout_real;
    p.xs = 1; p.n0 = op.cout; p.n_dx = op.ksize; p.n_xg = 0; p.items_per_plane = 0; p.cg = 1;
    if (op.type == SCENEEGO_OP_DECONV2) {
      SE_REQUIRE(op.lay_dst.side == 2 * op.lay_src.side && op.cin % 8 == 0 && op.cout % 8 == 0, "v2v_run: op %d bad deconv shape", i);
      if (op.impl == 1 || force_simt || op.cin % 16 || op.cout % 16 || op.cout > 256) {
        const int So = op.lay_dst.side;
        dim3 grid((So * So * So + 127) / 128, op.cout / 8, batch);
        deconv2_kernel<<<grid, 128, 0, st>>>(p);
        SE_CUDA_LAUNCH_CHECK("deconv2");
        ++g_launches;
        continue;
      }
      // tensor path: the parities do not shift the rows, so `npar` of them (256 / cout columns of TMEM) are
      // ONE wide GEMM: a single N = npar*cout MMA per K step against weights packed [group][cin/8][npar*cout][8]
      int npar = 256 / op.cout;
      if (npar > 8) npar = 8;
      const __nv_bfloat16* w_all = p.w;
      for (int par0 = 0; par0 < 8; par0 += npar) {
        ConvParams q = p;
        q.k = 1; q.r = 0; q.xs = 1; q.n0 = op.cout; q.mma_n = npar * op.cout; q.N = npar * op.cout; q.n_dx = 1;
        q.deconv = 1; q.par0 = par0; q.taps_dx = 1;
        q.w = w_all + (size_t)par0 * q.cin_planes * op.cout * 8;
        const int rc = launch_conv_tc(q, batch, i, st);
        if (rc != SCENEEGO_OK) return rc;
      }
      continue;
    }
    if (op.type == SCENEEGO_OP_CONV3_MARCH) {
      if (op.impl == 1 || force_simt) {
        // checker: the generic CUDA-core conv reading the marching weight layout
        SE_REQUIRE(op.ksize == 3 && op.cin % 16 == 0 && op.cout % 16 == 0, "v2v_run: op %d bad marching conv", i);
        p.k = 3; p.r = 1; p.xs = 1; p.n0 = op.cout; p.N = op.cout; p.cg = 1; p.march = 1;
        p.fd_frame = make_fastdiv((uint32_t)p.ls.frame_pitch);
        p.fd_px = make_fastdiv((uint32_t)p.ls.pitch_x);
        p.fd_py = make_fastdiv((uint32_t)p.ls.pitch_y);
        p.src2 = nullptr; p.n_dx2 = 0;
        if (op.src2 >= 0 && op.cin2 > 0) { p.src2 = (const __nv_bfloat16*)d_buffers[op.src2]; p.n_dx2 = 1; }
        p.w_main_bytes = 27u * (uint32_t)p.cin_planes * (uint32_t)op.cout * 16u;
        p.w_half_bytes = 0;
        const int64_t n_pos = (int64_t)batch * p.ls.frame_pitch;
        dim3 grid((unsigned)((n_pos + 127) / 128), op.cout / 16);
        conv_simt_kernel<<<grid, 128, 0, st>>>(p, n_pos);
        SE_CUDA_LAUNCH_CHECK("conv_simt (march layout)");
      } else {
        const int rc = launch_conv_march(op, d_buffers, d_blob, batch, i, st);
        if (rc != SCENEEGO_OK) return rc;
      }
      ++g_launches;
      continue;
    }
    SE_REQUIRE(op.type == SCENEEGO_OP_CONV, "v2v_run: op %d unknown type", i);
    SE_REQUIRE(op.ksize == 1 || op.ksize == 3 || op.ksize == 7, "v2v_run: op %d unsupported kernel size", i);
    SE_REQUIRE(op.cin % 16 == 0 && op.cout % 16 == 0 && op.cout >= 16 && op.cout <= 128, "v2v_run: op %d channels must be multiples of 16", i);
    SE_REQUIRE(op.lay_src.side == op.lay_dst.side && op.lay_src.pad >= op.ksize / 2 && op.lay_dst.pad <= op.lay_src.pad,
               "v2v_run: op %d layouts incompatible with the stencil", i);
    p.k = op.ksize; p.r = op.ksize / 2;
    const int64_t n_pos = (int64_t)batch * p.ls.frame_pitch;
    p.fd_frame = make_fastdiv((uint32_t)p.ls.frame_pitch);      // the checker kernel decodes rows too
    p.fd_px = make_fastdiv((uint32_t)p.ls.pitch_x);
    p.fd_py = make_fastdiv((uint32_t)p.ls.pitch_y);
    p.cg = op.cta_pair == 2 ? 2 : 1;
    p.src2 = nullptr; p.n_dx2 = 0;
    if (op.src2 >= 0 && op.cin2 > 0) {
      SE_REQUIRE(op.cin2 * 2 == op.cin && (op.cin / 16) % 2 == 0 && op.ksize == 3 && d_buffers[op.src2],
                 "v2v_run: op %d: a fused shortcut needs a 3^3 conv and a second source with half its channels", i);
      p.src2 = (const __nv_bfloat16*)d_buffers[op.src2];
      p.n_dx2 = op.xstack > 1 ? op.xstack : 1;
    }
    if (op.impl == 1 || force_simt) {
      const int xs_ = op.xstack > 1 ? op.xstack : 1;
      p.n0 = op.cout; p.N = xs_ * op.cout;   // weight row stride of the stacked blob
      const uint32_t tapb = (uint32_t)p.cin_planes * (p.N / p.cg) * 16u;
      p.w_main_bytes = (uint32_t)((op.ksize + xs_ - 1) * op.ksize * op.ksize) * tapb;
      p.w_half_bytes = p.w_main_bytes + (uint32_t)p.n_dx2 * (tapb / 2);
      dim3 grid((unsigned)((n_pos + 127) / 128), op.cout / 16);
      conv_simt_kernel<<<grid, 128, 0, st>>>(p, n_pos);
      SE_CUDA_LAUNCH_CHECK("conv_simt");
      ++g_launches;
      continue;
    }
    const int xs = op.xstack > 1 ? op.xstack : 1;
    SE_REQUIRE(xs * op.cout <= 256, "v2v_run: op %d: xstack * cout exceeds 256 columns", i);
    SE_REQUIRE(p.cg == 1 || (xs * op.cout) % 32 == 0, "v2v_run: op %d: a CTA pair splits N in halves of whole 16-column groups", i);
    p.xs = xs; p.n0 = op.cout; p.N = xs * op.cout; p.mma_n = p.N; p.n_dx = op.ksize + xs - 1;
    p.deconv = 0; p.par0 = 0; p.taps_dx = op.ksize * op.ksize;
    const int rc = launch_conv_tc(p, batch, i, st);
    if (rc != SCENEEGO_OK) return rc;
  }
  if (ev) cudaEventRecord(ev[n_ops], st);
  return SCENEEGO_OK;
}

extern "C" int sceneego_v2v_run(const sceneego_v2v_op_t* ops, int n_ops, void* const* d_buffers, const void* d_blob,
                                int batch, void* stream) {
  return v2v_run_impl(ops, n_ops, d_buffers, d_blob, batch, stream, nullptr);
}

// Measurement variant: brackets every op with CUDA events on `stream`, waits for the stream and
// returns per-op milliseconds in h_ms[n_ops].  (The only entry point that synchronises.)
extern "C" int sceneego_v2v_run_profile(const sceneego_v2v_op_t* ops, int n_ops, void* const* d_buffers,
                                        const void* d_blob, int batch, void* stream, float* h_ms) {
  SE_REQUIRE(h_ms && n_ops > 0 && n_ops < 4096, "v2v_run_profile: bad argument");
  cudaEvent_t* ev = (cudaEvent_t*)malloc(sizeof(cudaEvent_t) * (n_ops + 1));
  for (int i = 0; i <= n_ops; ++i) cudaEventCreate(&ev[i]);
  int rc = v2v_run_impl(ops, n_ops, d_buffers, d_blob, batch, stream, ev);
  if (rc == SCENEEGO_OK) {
    cudaError_t e = cudaEventSynchronize(ev[n_ops]);
    if (e != cudaSuccess) { set_error("v2v_run_profile: %s", cudaGetErrorString(e)); rc = SCENEEGO_E_CUDA; }
    for (int i = 0; i < n_ops && rc == SCENEEGO_OK; ++i) cudaEventElapsedTime(&h_ms[i], ev[i], ev[i + 1]);
  }
  for (int i = 0; i <= n_ops; ++i) cudaEventDestroy(ev[i]);
  free(ev);
  return rc;
}
