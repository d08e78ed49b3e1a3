// a7, first layer: the 7x7x7 stem Conv3d(33 -> 16) + BN + ReLU (network/v2v.py:8-18,147) as an x-MARCHING banded
// implicit GEMM on tcgen05 -- the dense successor of the 2x2x2-stacked kernel in stem.cu.
//
// Why.  stem.cu stacks the 2x2x2 block of outputs into N = 128 and pays for it with Toeplitz zeros: 512 input
// offsets for 343 taps, i.e. 1.9x the algorithmic MACs once K padding is counted; its tensor pipe is 87 % busy doing
// redundant math (profiles/r01_ncu_stem_s2d_final.txt).  Here GEMM rows are 128 cells of the (y,z) plane that march
// along x (csrc/march.cu): input plane x feeds the seven output planes x-3 .. x+3 through W[dx = 6 .. 0], stacked
// along N they are ONE dense B operand per (dy, dz, k-step) -- every multiply is algorithmic.
//   * Accumulators: a ring of 8 tensor-memory slots (16 columns = one output plane each) per tile.  A band of 7 in
//     a ring of 8 would straddle the ring end on 6 planes of 8 (two MMAs instead of one), so the band is issued as
//     N = 128 over the WHOLE ring with the weight rows ROTATED to the ring position of the plane: row block s holds
//     W[dx = 6 - j], j = (s - r) mod 8, r = the plane's ring position, and zeros for the one block (j = 7) that is
//     not in the band.  8 rotations of the 408 KB weight set = 3.3 MB, streamed from L2; the issuer picks the
//     rotation, every MMA has the same shape (64 cycles for 7 useful blocks of 8: 87.5 % of the pipe's own rate,
//     against 52 % for the stacked form).  The idle block accumulates zeros into the slot that the NEXT plane
//     opens, so that slot must already be drained and cleared: the issuer waits for the drain of output x-4
//     before plane x -- one plane earlier than a ring strictly needs, hidden behind the other tiles' MMAs.
//   * Cin = 33: the 32 feature channels are two K = 16 steps per tap.  The binary occupancy channel is stored as a
//     z-WINDOW plane: cell (x,y,z) holds occ[x][y][z-3 .. z+4] as its 8 entries (written 8-fold by the voxelisation
//     scatter, layout flag `zwin`), so ONE cell covers the seven dz taps of a row and one K = 16 MMA (two cells,
//     LBO = one z-line) covers two dy rows: 4 MMAs per plane instead of 49 K-padded ones.
//   * Per tile and input plane: 49 x 2 + 4 = 102 MMAs (128 x 128 x 16).  An item is four neighbouring tiles (512
//     cells) that share every staged window and weight chunk; one MMA-issuing warp per tile (a single warp cannot
//     feed the pipe through its barrier bookkeeping), eight epilogue warps.
//   * Weights stream in (k-step, dy) chunks of 7 taps (28 KB) through a ring of shared-memory slots, windows of two
//     channel-group planes (the K = 16 pair of a k-step) through a second ring; both by cp.async.bulk (1-D TMA).
// Edges: outputs -3..-1 and S..S+2 of a march are accumulated like any other (from the planes that exist) and
// drained without being stored, which keeps every plane's MMA sequence identical.
#include "tc_common.cuh"
#include <stdlib.h>
#include <math.h>

namespace sceneego {

constexpr int SMR_TILES = 4;
constexpr int SMR_L = 128 * SMR_TILES;
constexpr int SMR_RING = 8;
constexpr int SMR_EPI_WARPS = 8;
constexpr int SMR_THREADS = 32 * (2 + SMR_TILES + SMR_EPI_WARPS);      // 448
constexpr int SMR_MMA_B_BYTES = 2 * 128 * 16;                          // [2 k-chunks][128 rows][8] bf16
constexpr int SMR_FEAT_CHUNK = 7 * SMR_MMA_B_BYTES;                    // one (k-step, dy): seven dz taps
constexpr int SMR_OCC_CHUNK = 4 * SMR_MMA_B_BYTES;                     // four dy pairs
constexpr int SMR_FEAT_CHUNKS = 14;
constexpr int SMR_ROT_BYTES = SMR_FEAT_CHUNKS * SMR_FEAT_CHUNK + SMR_OCC_CHUNK;   // 417,792
constexpr size_t SMR_W_BYTES = (size_t)SMR_RING * SMR_ROT_BYTES;       // 3,342,336
constexpr int SMR_MAX_PSLOTS = 4, SMR_MAX_WSLOTS = 4;

struct StemMarchParams {
  const __nv_bfloat16* src;   // 5 planes: 4 feature channel groups + the z-window occupancy plane; layout ls (pad 3)
  __nv_bfloat16* dst;         // 16 channels = 2 planes, layout ld
  const uint8_t* w;           // SMR_W_BYTES: [rotation 8][chunk 15][tap][k-chunk 2][128 rows][8]
  const float* bias;          // 16
  sceneego_vol_layout_t ls, ld;
  int batch, relu;
  int groups_per_frame, n_items, cells_per_plane;
  int n_seg, seg_planes;        // small batches: a march is cut into n_seg x-ranges (3 planes of overlap each side)
  int halo, win_cells;
  int p_slots, w_slots;
  uint32_t win_bytes, pair_bytes;
  uint32_t off_w, off_bias, off_bar;
  FastDiv fd_gpf, fd_py, fd_seg;
};

__host__ __device__ __forceinline__ int smr_occ_dy(int pair, int c) {   // dy rows of the two K chunks of occupancy MMA `pair`
  const int d0 = pair < 3 ? 2 * pair : 5;                      // (0,1) (2,3) (4,5) (5,6): the last pair's first chunk is zero
  return d0 + c;
}

__global__ void __launch_bounds__(SMR_THREADS, 1) stem_march_tc_kernel(const __grid_constant__ StemMarchParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  float* s_bias = reinterpret_cast<float*>(smem + p.off_bias);
  const uint32_t bar0 = sbase + p.off_bar;
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int B_PFULL = 0, B_PEMPTY = B_PFULL + SMR_MAX_PSLOTS, B_WFULL = B_PEMPTY + SMR_MAX_PSLOTS,
                B_WEMPTY = B_WFULL + SMR_MAX_WSLOTS, B_ACC_FULL = B_WEMPTY + SMR_MAX_WSLOTS,
                B_ACC_EMPTY = B_ACC_FULL + SMR_TILES * SMR_RING, B_COUNT = B_ACC_EMPTY + SMR_TILES * SMR_RING;
  uint32_t* s_tmem_ptr = reinterpret_cast<uint32_t*>(smem + p.off_bar + 8 * B_COUNT);

  if (threadIdx.x < 16) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  if (threadIdx.x == 0) {
    for (int i = 0; i < SMR_MAX_PSLOTS; ++i) { mbar_init(BAR(B_PFULL + i), 1); mbar_init(BAR(B_PEMPTY + i), SMR_TILES); }
    for (int i = 0; i < SMR_MAX_WSLOTS; ++i) { mbar_init(BAR(B_WFULL + i), 1); mbar_init(BAR(B_WEMPTY + i), SMR_TILES); }
    for (int i = 0; i < SMR_TILES * SMR_RING; ++i) { mbar_init(BAR(B_ACC_FULL + i), 1); mbar_init(BAR(B_ACC_EMPTY + i), 4); }
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
  // every accumulator slot starts cleared: all MMAs accumulate, the epilogue re-clears a slot after draining it
  if (warp >= 2 + SMR_TILES && warp < 2 + SMR_TILES + 4) {
    const uint32_t t0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (uint32_t c = 0; c < 512; c += 16) tc_st16_zero(t0 + c);
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const int S = p.ls.side;
  const int my_items = ((int)p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  // item = (frame, group of four tiles, x-segment).  A segment stores outputs [x0, x1); it marches over the input
  // planes [xa, xb] = [x0 - 3, x1 + 2] clipped to the volume and drains the outputs xa - 3 .. xb + 3 in order.
  auto item_of = [&](int it, int& b, int& cell0, int& n_act, int& x0, int& x1, int& xa, int& xb) {
    const uint32_t item = blockIdx.x + (uint32_t)it * gridDim.x;
    const uint32_t bg = fdiv(item, p.fd_seg);
    const int seg = (int)(item - bg * (uint32_t)p.n_seg);
    b = (int)fdiv(bg, p.fd_gpf);
    cell0 = (int)(bg - (uint32_t)b * (uint32_t)p.groups_per_frame) * SMR_L;
    const int left = p.cells_per_plane - cell0;
    n_act = left >= SMR_L ? SMR_TILES : (left + 127) / 128;
    x0 = seg * p.seg_planes;
    x1 = x0 + p.seg_planes < S ? x0 + p.seg_planes : S;
    xa = x0 - 3 > 0 ? x0 - 3 : 0;
    xb = x1 + 2 < S - 1 ? x1 + 2 : S - 1;
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
          const uint8_t* wr = p.w + (size_t)((G + (uint32_t)(x - xa)) & 7u) * SMR_ROT_BYTES;
          for (int c = 0; c <= SMR_FEAT_CHUNKS; ++c) {
            const uint32_t bytes = c < SMR_FEAT_CHUNKS ? SMR_FEAT_CHUNK : SMR_OCC_CHUNK;
            mbar_wait(BAR(B_WEMPTY + sl), sph ^ 1);
            mbar_expect_tx(BAR(B_WFULL + sl), bytes);
            const uint32_t dst = sbase + p.off_w + (uint32_t)sl * SMR_FEAT_CHUNK;
            const uint8_t* srcw = wr + (size_t)c * SMR_FEAT_CHUNK;
            for (uint32_t o = 0; o < bytes; o += 14336u)
              bulk_g2s(dst + o, srcw + o, bytes - o < 14336u ? bytes - o : 14336u, BAR(B_WFULL + sl));
            if (++sl == p.w_slots) { sl = 0; sph ^= 1; }
          }
        }
        G += (uint32_t)(xb - xa + 1 + 6);
      }
    }
  } else if (warp == 1) {
    // ===================== window producer =====================
    if (lane == 0) {
      int ps = 0, pph = 0;
      for (int it = 0; it < my_items; ++it) {
        int b, cell0, n_act, x0, x1, xa, xb;
        item_of(it, b, cell0, n_act, x0, x1, xa, xb);
        const int64_t q0 = (int64_t)b * p.ls.frame_pitch + p.ls.guard + cell0 - p.halo;
        for (int x = xa; x <= xb; ++x) {
          const int64_t qx = q0 + (int64_t)x * p.ls.pitch_x;
          for (int ph = 0; ph < 3; ++ph) {             // planes (0,1), (2,3), (4 = occupancy)
            const int n_planes = ph < 2 ? 2 : 1;
            mbar_wait(BAR(B_PEMPTY + ps), pph ^ 1);
            mbar_expect_tx(BAR(B_PFULL + ps), p.win_bytes * (uint32_t)n_planes);
            for (int g = 0; g < n_planes; ++g)
              bulk_g2s(sbase + (uint32_t)ps * p.pair_bytes + (uint32_t)g * p.win_bytes,
                       p.src + ((int64_t)(2 * ph + g) * p.ls.plane_stride + qx) * 8, p.win_bytes, BAR(B_PFULL + ps));
            if (++ps == p.p_slots) { ps = 0; pph ^= 1; }
          }
        }
      }
    }
  } else if (warp < 2 + SMR_TILES) {
    // ===================== MMA issuers: warp 2 + t owns tile t (tensor-memory columns 128 t .. 128 t + 127) ====
    const int t = warp - 2;
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const bool leader = elect_one();
    constexpr uint32_t idesc = (1u << 4) | kIdescAB | ((uint32_t)(128 >> 3) << 17) | (8u << 24);
    const uint64_t desc_hi = (uint64_t)(8u | (1u << 14)) << 32;            // SBO = 128 B, descriptor version 1
    const uint32_t a_lbo_feat = ((uint32_t)p.win_cells & 0x3FFFu) << 16;   // K chunk 1 = the pair's second plane
    const uint32_t a_lbo_occ = ((uint32_t)pitch_y & 0x3FFFu) << 16;        // K chunk 1 = the next dy row
    constexpr uint32_t b_lbo = 128u << 16;                                 // [k-chunk][128 rows][8]: 2 KB apart
    const uint32_t d_mine = tmem_u + (uint32_t)t * 128u;
    int ps = 0, sl = 0;
    uint32_t pph = 0, sph = 0;
    uint32_t G = 0;
    for (int it = 0; it < my_items; ++it) {
      int b_, cell0_, n_act, x0_, x1_, xa, xb;
      item_of(it, b_, cell0_, n_act, x0_, x1_, xa, xb);
      // A tile beyond the end of the plane (the last group of a frame) runs the whole barrier protocol -- its commits
      // arrive at once, the epilogue drains zeros -- and skips only the MMA instructions, so every tile keeps the same
      // ring position G (the running count of drained outputs) that the weight producer derives the rotation from.
      const bool active = t < n_act;
      if (it > 0) {
        // all eight slots of the previous march have been drained and cleared
        for (uint32_t k = 1; k <= SMR_RING; ++k)
          mbar_wait_warp(BAR(B_ACC_EMPTY + t * SMR_RING + (int)((G - k) & 7u)), ((G - k) >> 3) & 1u);
      }
      for (int x = xa; x <= xb; ++x) {
        const uint32_t gx = G + (uint32_t)(x - xa);       // ring index of the first band block (output x - 3)
        if (x > xa)     // the slot this plane's idle block touches (and the next plane opens): output x-4 is gone
          mbar_wait_warp(BAR(B_ACC_EMPTY + t * SMR_RING + (int)((gx - 1u) & 7u)), ((gx - 1u) >> 3) & 1u);
        for (int ks = 0; ks < 2; ++ks) {
          mbar_wait_warp(BAR(B_PFULL + ps), pph);
          const uint32_t a_org = ((sbase + (uint32_t)ps * p.pair_bytes) >> 4) + (uint32_t)p.halo + (uint32_t)t * 128u;
          for (int dy = 0; dy < 7; ++dy) {
            mbar_wait_warp(BAR(B_WFULL + sl), sph);
            tc_fence_after();
            const uint32_t b_org = (((sbase + p.off_w + (uint32_t)sl * SMR_FEAT_CHUNK) >> 4) & 0x3FFFu) | b_lbo;
            const uint32_t a_row = a_org + (uint32_t)((dy - 3) * pitch_y - 3);
            if (leader) {
              if (active) {
#pragma unroll
                for (int dz = 0; dz < 7; ++dz)
                  tc_mma_bf16(d_mine, desc_hi | (uint64_t)(((a_row + (uint32_t)dz) & 0x3FFFu) | a_lbo_feat),
                              desc_hi | (uint64_t)(b_org + (uint32_t)dz * (SMR_MMA_B_BYTES / 16)), idesc, 1u);
              }
              tc_commit(BAR(B_WEMPTY + sl));
            }
            if (++sl == p.w_slots) { sl = 0; sph ^= 1u; }
          }
          if (leader) tc_commit(BAR(B_PEMPTY + ps));
          if (++ps == p.p_slots) { ps = 0; pph ^= 1u; }
        }
        {   // occupancy: one window plane, four MMAs over (dy pair) x (all dz inside the z-window cell)
          mbar_wait_warp(BAR(B_PFULL + ps), pph);
          mbar_wait_warp(BAR(B_WFULL + sl), sph);
          tc_fence_after();
          const uint32_t a_org = ((sbase + (uint32_t)ps * p.pair_bytes) >> 4) + (uint32_t)p.halo + (uint32_t)t * 128u;
          const uint32_t b_org = (((sbase + p.off_w + (uint32_t)sl * SMR_FEAT_CHUNK) >> 4) & 0x3FFFu) | b_lbo;
          if (leader) {
            if (active) {
#pragma unroll
              for (int pr = 0; pr < 4; ++pr)
                tc_mma_bf16(d_mine, desc_hi | (uint64_t)(((a_org + (uint32_t)((smr_occ_dy(pr, 0) - 3) * pitch_y)) & 0x3FFFu) | a_lbo_occ),
                            desc_hi | (uint64_t)(b_org + (uint32_t)pr * (SMR_MMA_B_BYTES / 16)), idesc, 1u);
            }
            tc_commit(BAR(B_WEMPTY + sl));
            tc_commit(BAR(B_PEMPTY + ps));
            tc_commit(BAR(B_ACC_FULL + t * SMR_RING + (int)(gx & 7u)));                      // output x-3 is complete
            if (x == xb)
              for (uint32_t k = 1; k <= 6; ++k)                                               // and so are xb-2 .. xb+3
                tc_commit(BAR(B_ACC_FULL + t * SMR_RING + (int)((gx + k) & 7u)));
          }
          if (++sl == p.w_slots) { sl = 0; sph ^= 1u; }
          if (++ps == p.p_slots) { ps = 0; pph ^= 1u; }
        }
      }
      G += (uint32_t)(xb - xa + 1 + 6);
    }
    __syncwarp();
  } else {
    // ===================== epilogue: 8 warps, two per tensor-memory lane quarter; warp pair h drains tiles 2h, 2h+1 ====
    const int quarter = warp & 3;
    const int half = (warp - (2 + SMR_TILES)) >> 2;
    float bs[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) bs[j] = s_bias[j];
    uint32_t G = 0;
    for (int it = 0; it < my_items; ++it) {
      int b, cell0, n_act, x0, x1, xa, xb;
      item_of(it, b, cell0, n_act, x0, x1, xa, xb);
      const int OUTS = xb - xa + 1 + 6;
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
        const int o = xa - 3 + oi;
        const uint32_t gi = G + (uint32_t)oi;
        const int slot = (int)(gi & 7u);
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
          const int t = 2 * half + tt;
          mbar_wait(BAR(B_ACC_FULL + t * SMR_RING + slot), (gi >> 3) & 1u);
          tc_fence_after();
          uint32_t raw[16];
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * 128 + slot * 16);
          tc_ld16(taddr, raw);
          tc_wait_ld();
          tc_st16_zero(taddr);                              // hand the slot back cleared
          tc_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(BAR(B_ACC_EMPTY + t * SMR_RING + slot));
          if (valid[tt] && o >= x0 && o < x1) {
            const int64_t dpos = dpos0[tt] + (int64_t)o * p.ld.pitch_x;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              float ov[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float v = __uint_as_float(raw[8 * g + j]) + bs[8 * g + j];
                ov[j] = p.relu ? fmaxf(v, 0.f) : v;
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
// CUDA-core checker: same input (z-window occupancy plane), same packed blob walked in the kernel's order, one
// thread per output voxel; rotation r = input plane & 7 so that all eight rotations are exercised.
// op.impl = 1 / SCENEEGO_FORCE_SIMT; used by the tests to validate the tensor path and the packer.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) stem_march_simt_kernel(const __grid_constant__ StemMarchParams p) {
  const int V = p.ld.side;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (n >= V * V * V) return;
  const int z = n % V, y = (n / V) % V, x = n / (V * V);
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  for (int xi = x - 3; xi <= x + 3; ++xi) {
    if (xi < 0 || xi >= V) continue;
    const int j = x - xi + 3;                       // band block of input plane xi that feeds output x
    const int r = xi & 7, s = (r + j) & 7;
    const uint4* wr = reinterpret_cast<const uint4*>(p.w + (size_t)r * SMR_ROT_BYTES);
    const int64_t q = vol_pos(p.ls, b, xi, y, z);
    for (int ks = 0; ks < 2; ++ks)
      for (int dy = 0; dy < 7; ++dy)
        for (int dz = 0; dz < 7; ++dz) {
          const uint4* wt = wr + ((size_t)(ks * 7 + dy) * SMR_FEAT_CHUNK + (size_t)dz * SMR_MMA_B_BYTES) / 16;
          const int64_t qs = q + (int64_t)(dy - 3) * p.ls.pitch_y + (dz - 3);
          for (int c = 0; c < 2; ++c) {
            float a[8];
            unpack8(*reinterpret_cast<const uint4*>(p.src + ((int64_t)(2 * ks + c) * p.ls.plane_stride + qs) * 8), a);
#pragma unroll
            for (int co = 0; co < 16; ++co) {
              float wv[8];
              unpack8(__ldg(wt + c * 128 + s * 16 + co), wv);
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[co] = fmaf(a[i], wv[i], acc[co]);
            }
          }
        }
    for (int pr = 0; pr < 4; ++pr) {
      const uint4* wt = wr + ((size_t)SMR_FEAT_CHUNKS * SMR_FEAT_CHUNK + (size_t)pr * SMR_MMA_B_BYTES) / 16;
      for (int c = 0; c < 2; ++c) {
        const int64_t qs = q + (int64_t)(smr_occ_dy(pr, c) - 3) * p.ls.pitch_y;
        float a[8];
        unpack8(*reinterpret_cast<const uint4*>(p.src + ((int64_t)4 * p.ls.plane_stride + qs) * 8), a);
#pragma unroll
        for (int co = 0; co < 16; ++co) {
          float wv[8];
          unpack8(__ldg(wt + c * 128 + s * 16 + co), wv);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[co] = fmaf(a[i], wv[i], acc[co]);
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

// Called by sceneego_v2v_run for SCENEEGO_OP_STEM7_MARCH.
int launch_stem_march(const sceneego_v2v_op_t& op, void* const* d_buffers, const void* d_blob, int batch, int op_index,
                      bool simt, cudaStream_t st) {
  SE_REQUIRE(op.lay_src.s2d == 0 && op.lay_src.zwin == 1 && op.lay_src.pad >= 3 && op.lay_dst.s2d == 0 &&
                 op.lay_dst.side == op.lay_src.side && op.lay_dst.pad >= 1,
             "v2v_run: op %d: the marching stem needs a z-window source layout (pad >= 3) of the destination's side", op_index);
  SE_REQUIRE(op.cout == 16 && op.cin == 33 && op.ksize == 7, "v2v_run: op %d: stem is 33 -> 16, k = 7", op_index);
  StemMarchParams p;
  memset(&p, 0, sizeof(p));
  p.src = (const __nv_bfloat16*)d_buffers[op.src];
  p.dst = (__nv_bfloat16*)d_buffers[op.dst];
  p.w = (const uint8_t*)d_blob + op.w_offset;
  p.bias = (const float*)((const char*)d_blob + op.b_offset);
  p.ls = op.lay_src; p.ld = op.lay_dst; p.batch = batch; p.relu = (op.flags & SCENEEGO_F_RELU) ? 1 : 0;
  SE_REQUIRE(p.src && p.dst, "v2v_run: op %d has a null buffer", op_index);
  const int S = p.ls.side;
  if (simt) {
    dim3 grid((S * S * S + 127) / 128, batch);
    stem_march_simt_kernel<<<grid, 128, 0, st>>>(p);
    SE_CUDA_LAUNCH_CHECK("stem_march_simt");
    return SCENEEGO_OK;
  }
  p.halo = 3 * (p.ls.pitch_y + 1);
  p.win_cells = SMR_L + 2 * p.halo;
  p.win_bytes = (uint32_t)p.win_cells * 16u;
  p.pair_bytes = 2u * p.win_bytes;
  SE_REQUIRE(p.win_cells < 16384 && p.ls.guard >= p.halo, "v2v_run: op %d: stem window too large / guard too small (side %d)", op_index, S);
  p.cells_per_plane = (S - 1) * p.ls.pitch_y + S;
  p.groups_per_frame = (p.cells_per_plane + SMR_L - 1) / SMR_L;
  // small batches leave most SMs without an item (B = 1: nine groups): cut every march into x-segments, each paying
  // six extra input planes, until the launch fills the machine once
  const int base_items = batch * p.groups_per_frame;
  int n_seg = kNumSMs / base_items;
  if (n_seg < 1) n_seg = 1;
  if (n_seg > 16) n_seg = 16;
  { const char* e = getenv("SCENEEGO_STEM_SEGMENTS"); if (e && atoi(e) >= 1 && atoi(e) <= 32) n_seg = atoi(e); }
  int seg_planes = (S + n_seg - 1) / n_seg;
  if (seg_planes < 4) seg_planes = 4 < S ? 4 : S;
  n_seg = (S + seg_planes - 1) / seg_planes;
  p.n_seg = n_seg; p.seg_planes = seg_planes;
  p.n_items = base_items * n_seg;
  SE_REQUIRE((int64_t)batch * p.ls.frame_pitch + 4096 < (1ll << 31), "v2v_run: op %d: batch * frame_pitch too large for one launch", op_index);
  const uint32_t fixed = 64 + 8 * (2 * SMR_MAX_PSLOTS + 2 * SMR_MAX_WSLOTS + 2 * SMR_TILES * SMR_RING) + 64;
  int p_slots = 3, w_slots = 0;
  for (; p_slots >= 2; --p_slots) {
    const int64_t left = (int64_t)kMaxSmem - fixed - (int64_t)p_slots * p.pair_bytes;
    w_slots = (int)(left / SMR_FEAT_CHUNK);
    if (w_slots >= 3) break;
  }
  if (w_slots > SMR_MAX_WSLOTS) w_slots = SMR_MAX_WSLOTS;
  { const char* e = getenv("SCENEEGO_STEM_WSLOTS"); if (e && atoi(e) >= 2 && atoi(e) <= w_slots) w_slots = atoi(e); }
  SE_REQUIRE(p_slots >= 2 && w_slots >= 2, "v2v_run: op %d: stem windows do not fit shared memory (side %d)", op_index, S);
  p.p_slots = p_slots; p.w_slots = w_slots;
  p.off_w = (uint32_t)p_slots * p.pair_bytes;
  p.off_bias = p.off_w + (uint32_t)w_slots * SMR_FEAT_CHUNK;
  p.off_bar = p.off_bias + 64;
  p.fd_gpf = make_fastdiv((uint32_t)p.groups_per_frame);
  p.fd_py = make_fastdiv((uint32_t)p.ls.pitch_y);
  p.fd_seg = make_fastdiv((uint32_t)p.n_seg);
  const size_t smem_bytes = (size_t)p.off_bar + 8 * (2 * SMR_MAX_PSLOTS + 2 * SMR_MAX_WSLOTS + 2 * SMR_TILES * SMR_RING) + 64;
  SE_REQUIRE(smem_bytes <= kMaxSmem, "v2v_run: op %d: stem shared memory plan exceeds 227 KB", op_index);
  if (int rc = ensure_max_dynamic_smem((const void*)stem_march_tc_kernel, (int)kMaxSmem)) return rc;
  const int grid = p.n_items < kNumSMs ? p.n_items : kNumSMs;
  stem_march_tc_kernel<<<grid, SMR_THREADS, smem_bytes, st>>>(p);
  SE_CUDA_LAUNCH_CHECK("stem_march_tc");
  return SCENEEGO_OK;
}

}  // namespace sceneego

using namespace sceneego;

extern "C" size_t sceneego_v2v_stem_march_weight_bytes(void) { return SMR_W_BYTES; }

extern "C" int sceneego_v2v_pack_stem_march(const float* h_weight, const float* h_bias, const float* h_gamma,
                                            const float* h_beta, const float* h_mean, const float* h_var, double eps,
                                            uint16_t* h_w_out, float* h_b_out) {
  SE_REQUIRE(h_weight && h_w_out && h_b_out, "pack_stem_march: bad argument");
  constexpr int CO = 16, CI = 33, K = 7;
  memset(h_w_out, 0, SMR_W_BYTES);
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
    return f2bf((float)((double)h_weight[((((size_t)co * CI + ci) * K + dx) * K + dy) * K + dz] * scale[co]));
  };
  for (int r = 0; r < SMR_RING; ++r) {
    uint16_t* rot = h_w_out + (size_t)r * SMR_ROT_BYTES / 2;
    for (int s = 0; s < SMR_RING; ++s) {
      const int j = (s - r) & 7;                    // band block at ring position s: feeds output (input plane - 3 + j)
      if (j == 7) continue;                         // the one slot outside the band: zeros
      const int dx = 6 - j;
      for (int co = 0; co < CO; ++co) {
        const int row = s * 16 + co;
        for (int ks = 0; ks < 2; ++ks)
          for (int dy = 0; dy < K; ++dy)
            for (int dz = 0; dz < K; ++dz) {
              uint16_t* tap = rot + ((size_t)(ks * 7 + dy) * SMR_FEAT_CHUNK + (size_t)dz * SMR_MMA_B_BYTES) / 2;
              for (int c = 0; c < 2; ++c)
                for (int e = 0; e < 8; ++e)
                  tap[((size_t)c * 128 + row) * 8 + e] = W(co, ks * 16 + c * 8 + e, dx, dy, dz);
            }
        for (int pr = 0; pr < 4; ++pr) {
          uint16_t* tap = rot + ((size_t)SMR_FEAT_CHUNKS * SMR_FEAT_CHUNK + (size_t)pr * SMR_MMA_B_BYTES) / 2;
          for (int c = 0; c < 2; ++c) {
            if (pr == 3 && c == 0) continue;        // dy = 5 is already served by pair 2
            const int dy = smr_occ_dy(pr, c);
            for (int e = 0; e < 7; ++e)             // entry e of a z-window cell is the dz = e tap; entry 7 is unused
              tap[((size_t)c * 128 + row) * 8 + e] = W(co, 32, dx, dy, e);
          }
        }
      }
    }
  }
  return SCENEEGO_OK;
}
