// a7: the 3^3 convolutions with few output channels (Cout = 32 at full resolution: ten of V2V's layers,
// network/v2v.py:21-43,147-156) as an x-MARCHING, BANDED implicit GEMM on tcgen05.
//
// Why another kernel.  With both operands in shared memory a tcgen05.mma (M = 128, K = 16) costs
// max(N/2, 32 + N/4) cycles (tools/mma_rate.cu): the 128 x 32 B A tile is re-read for every MMA, so N = 32
// runs at 40 % and the x-stacked N = 64 Toeplitz form of conv_tc_kernel at 2/3 of the pipe -- with a quarter of
// its columns multiplying structural zeros.  Here one work item is a tile of 128 cells of the (y,z) plane that
// MARCHES along x.  Input plane p contributes to output planes p-1, p, p+1 through W[dx=2], W[dx=1], W[dx=0]:
// stacked along N in that order ([tap(dy,dz)][cin/8][3*Cout][8]) the three blocks are ONE dense B operand, so
// each (dy,dz,k-step) is one N = 3*Cout = 96 MMA (56 cycles for three useful blocks, no zeros; N = 2*Cout
// sub-bands of the same array at the volume faces).  Consequences:
//   * every input plane is staged once per item and feeds three outputs (conv_tc x-stacking: 4 planes per 2);
//   * the whole weight array (27 taps: 55 KB for 32 -> 32) stays RESIDENT in shared memory -- no weight stream;
//   * accumulators are a RING of 512/Cout tensor-memory slots, one output plane each: output p is complete
//     when plane p+1 has been issued, its slot is committed to the epilogue and recycled 16 planes later.
//     Where the three slots of a band straddle the end of the ring the MMA is split in two (2 planes per ring
//     revolution); every MMA accumulates -- the epilogue hands a drained slot back cleared (tcgen05.st of zeros),
//     so the slot a plane opens needs no overwriting MMA of its own.
//   warp 0 = producer (cp.async.bulk of one (cin/8)-plane window per input plane, 2..8-stage ring)
//   warp 1 = TMEM allocator + the one MMA issuer (an N = 96 MMA takes longer than the 41.5-cycle issue floor)
//   then 4 or 8 epilogue warps: one or two per TMEM lane quarter (two alternate output planes); row decode
//   (y, z, validity) is done once per item because it is the same for every plane of the march.
// Measured (tools/mma_band.cu, tools/tune_march.py, profiles/r01_mma_band.txt, r01_tune_march.txt): the band MMAs
// cost what the model says (N = 96: 56.1 cycles at any column / row offset; 58.5 with the split first MMA and the
// commits), but the issuing warp's per-plane bookkeeping (~300-600 cycles of waits, commits and ring arithmetic
// executed by ONE warp) is NOT hidden behind the tensor pipe's short queue.  Hence two CTAs per SM, each with
// half of tensor memory (ring of 8 slots) -- one CTA's bookkeeping overlaps the other's MMAs: 15.7 -> 12.4 us per
// 32->32 layer and frame (conv_tc with CTA pairs: 15.3), tensor-only floor 9.9 us.
// The fused 1x1 projection shortcut (Res3DBlock.skip_con) is one extra N = Cout MMA per plane on the plane's
// own slot, from a halo-free window of the second source staged with the plane.
#include "tc_common.cuh"
#include <stdlib.h>

namespace sceneego {

constexpr int MARCH_L = 128;             // cells per item (one UMMA M tile)
constexpr int MARCH_MAX_STAGES = 8;
constexpr int MARCH_MAX_SLOTS = 32;      // 512 columns / 16
// TWO = 1: two CTAs per SM, each with half of tensor memory (a ring of 256/Cout slots), four epilogue warps and
// at most half of shared memory: the per-plane bookkeeping of one CTA's single MMA-issuing warp (barrier waits,
// commits, ring arithmetic -- ~300-600 cycles that the tensor pipe would otherwise idle through, measured with
// tools/tune_march.py) overlaps with the other CTA's MMAs.
__host__ __device__ constexpr int march_epi_warps(int two) { return two ? 4 : 8; }
__host__ __device__ constexpr int march_threads(int two) { return 32 * (2 + march_epi_warps(two)); }
__host__ __device__ constexpr int march_tmem_cols(int two) { return two ? 256 : 512; }

struct MarchParams {
  const __nv_bfloat16* src;
  const __nv_bfloat16* src2;   // fused shortcut source (cin2_planes planes) or nullptr
  const __nv_bfloat16* res;
  __nv_bfloat16* dst;
  const __nv_bfloat16* w;      // [9][cin/8][3*n0][8], then the shortcut's [cin2/8][n0][8]
  const float* bias;
  sceneego_vol_layout_t ls, ld;
  int batch, flags;
  int cin_planes, cin2_planes, n0, n_slots;
  int tiles_per_plane, n_items;
  int halo;
  int stages;
  uint32_t win_bytes, win2_bytes, stage_bytes;
  uint32_t w_bytes, w2_bytes;
  uint32_t off_win, off_bias, off_bar;   // weights live at offset 0
  FastDiv fd_py, fd_tpp;
};

// KSTEPS = Cin/16 of the stencil, KSTEPS2 = Cin2/16 of the fused shortcut (0 = none), NCH = Cout/16.
template <int KSTEPS, int KSTEPS2, int NCH, int TWO>
__global__ void __launch_bounds__(march_threads(TWO), TWO ? 2 : 1) conv_march_kernel(const __grid_constant__ MarchParams p) {
  constexpr int MARCH_THREADS = march_threads(TWO), MARCH_EPI_WARPS = march_epi_warps(TWO);
  constexpr uint32_t TMEM_COLS = march_tmem_cols(TWO);
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  float* s_bias = reinterpret_cast<float*>(smem + p.off_bias);
  const uint32_t bar0 = sbase + p.off_bar;
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int B_WIN_FULL = 0, B_WIN_EMPTY = MARCH_MAX_STAGES, B_ACC_FULL = 2 * MARCH_MAX_STAGES,
                B_ACC_EMPTY = B_ACC_FULL + MARCH_MAX_SLOTS, B_W_FULL = B_ACC_EMPTY + MARCH_MAX_SLOTS, B_COUNT = B_W_FULL + 1;
  uint32_t* s_tmem_ptr = reinterpret_cast<uint32_t*>(smem + p.off_bar + 8 * B_COUNT);
  constexpr int N0 = 16 * NCH;
  const int S = p.ls.side;
  constexpr int NS = ((int)TMEM_COLS / N0) > MARCH_MAX_SLOTS ? MARCH_MAX_SLOTS : ((int)TMEM_COLS / N0);   // accumulator ring (== p.n_slots)

  for (int i = threadIdx.x; i < N0; i += MARCH_THREADS) s_bias[i] = p.bias[i];
  if (threadIdx.x == 0) {
    for (int i = 0; i < MARCH_MAX_STAGES; ++i) { mbar_init(BAR(B_WIN_FULL + i), 1); mbar_init(BAR(B_WIN_EMPTY + i), 1); }
    for (int i = 0; i < MARCH_MAX_SLOTS; ++i) { mbar_init(BAR(B_ACC_FULL + i), 1); mbar_init(BAR(B_ACC_EMPTY + i), 4); }   // one epilogue warp per TMEM lane quarter
    mbar_init(BAR(B_W_FULL), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem_ptr)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem_ptr;
  // every accumulator slot starts cleared: all MMAs accumulate, and the epilogue re-clears a slot after draining it
  if (warp >= 2 && warp < 6) {                       // four warps = the four TMEM lane quarters
    const uint32_t t0 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    for (uint32_t c = 0; c < TMEM_COLS; c += 16) tc_st16_zero(t0 + c);
    tc_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // Work split.  Whole rounds: CTA k marches items k, k + grid, ... -- neighbouring CTAs work on neighbouring
  // tiles of the same frame at the same plane at the same time, so the halo cells two tiles share are read from
  // DRAM once and hit in L2 for the other (with contiguous item ranges per CTA ncu showed a 2 % L2 hit rate and
  // 2.05x the compulsory DRAM reads).  The items of the last, incomplete round (2112 items on 296 CTAs would
  // leave 12 % of the machine idle) are cut by PLANES into gridDim.x equal ranges: a march over [x0, x1) reads
  // input planes max(x0-1, 0) .. min(x1, S-1).
  const uint32_t n_whole = (uint32_t)p.n_items / gridDim.x;                       // rounds of whole items
  const uint32_t left_items = (uint32_t)p.n_items - n_whole * gridDim.x;
  const uint32_t left_total = left_items * (uint32_t)S;                           // < gridDim.x * S
  const uint32_t lbase = left_total / gridDim.x, lrem = left_total % gridDim.x;
  const uint32_t Q0 = blockIdx.x * lbase + (blockIdx.x < lrem ? blockIdx.x : lrem);
  const uint32_t Q1 = Q0 + lbase + (blockIdx.x < lrem ? 1u : 0u);
  const uint32_t lfirst = Q1 > Q0 ? Q0 / (uint32_t)S : 0u, llast = Q1 > Q0 ? (Q1 - 1) / (uint32_t)S : 0u;
  const int my_items = (int)n_whole + (Q1 > Q0 ? (int)(llast - lfirst + 1) : 0);
  auto item_of = [&](int it, int& b, int& cell0, int& x0, int& x1) {
    uint32_t item;
    if ((uint32_t)it < n_whole) {
      item = blockIdx.x + (uint32_t)it * gridDim.x;
      x0 = 0; x1 = S;
    } else {
      const uint32_t li = lfirst + ((uint32_t)it - n_whole);
      item = n_whole * gridDim.x + li;
      x0 = li == lfirst ? (int)(Q0 - lfirst * (uint32_t)S) : 0;
      x1 = li == llast ? (int)(Q1 - llast * (uint32_t)S) : S;
    }
    b = (int)fdiv(item, p.fd_tpp);
    cell0 = (int)(item - (uint32_t)b * (uint32_t)p.tiles_per_plane) * MARCH_L;
  };
  constexpr uint32_t G_START = 4u * NS;   // running output counter; the offset keeps (G - 2) non-negative, slot 0 / parity 0

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      const uint32_t wtot = p.w_bytes + p.w2_bytes;
      mbar_expect_tx(BAR(B_W_FULL), wtot);
      for (uint32_t o = 0; o < wtot; o += 16384u) {
        const uint32_t n = wtot - o < 16384u ? wtot - o : 16384u;
        bulk_g2s(sbase + o, reinterpret_cast<const char*>(p.w) + o, n, BAR(B_W_FULL));
      }
      int ws = 0, wph = 0;
      for (int it = 0; it < my_items; ++it) {
        int b, cell0, x0, x1;
        item_of(it, b, cell0, x0, x1);
        const int64_t q0 = (int64_t)b * p.ls.frame_pitch + p.ls.guard + cell0;
        const int xa = x0 > 0 ? x0 - 1 : 0, xb = x1 < S ? x1 : S - 1;
        for (int x = xa; x <= xb; ++x) {
          mbar_wait(BAR(B_WIN_EMPTY + ws), wph ^ 1);
          mbar_expect_tx(BAR(B_WIN_FULL + ws), p.stage_bytes);
          const int64_t qc = q0 + (int64_t)x * p.ls.pitch_x;
          const uint32_t dst0 = sbase + p.off_win + (uint32_t)ws * p.stage_bytes;
#pragma unroll
          for (int g = 0; g < 2 * KSTEPS; ++g)
            bulk_g2s(dst0 + (uint32_t)g * p.win_bytes, p.src + ((int64_t)g * p.ls.plane_stride + qc - p.halo) * 8, p.win_bytes,
                     BAR(B_WIN_FULL + ws));
#pragma unroll
          for (int g = 0; g < 2 * KSTEPS2; ++g)
            bulk_g2s(dst0 + (uint32_t)(2 * KSTEPS) * p.win_bytes + (uint32_t)g * p.win2_bytes,
                     p.src2 + ((int64_t)g * p.ls.plane_stride + qc) * 8, p.win2_bytes, BAR(B_WIN_FULL + ws));
          if (++ws == p.stages) { ws = 0; wph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The warp runs converged (waits are warp votes) so descriptors stay in uniform registers; only the
    // tcgen05.mma / commit instructions are predicated on the elected lane.  Everything per MMA is an add of a
    // compile-time constant to a per-plane base: one warp must issue an MMA every ~56 cycles.
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const bool leader = elect_one();
    constexpr uint32_t DHI = 8u | (1u << 14);                                   // SBO = 128 B, descriptor version 1
    auto DESC = [](uint32_t lo) { return ((uint64_t)DHI << 32) | (uint64_t)lo; };
    const uint32_t a_lbo = ((p.win_bytes >> 4) & 0x3FFFu) << 16;                // K chunk stride = one window plane
    const uint32_t a2_lbo = ((p.win2_bytes >> 4) & 0x3FFFu) << 16;
    constexpr uint32_t b_lbo = (uint32_t)(3 * N0) << 16;                        // rows * 16 B, in 16-B units
    constexpr uint32_t b2_lbo = (uint32_t)N0 << 16;
    const uint32_t a_ks_step = (2u * p.win_bytes) >> 4, a2_ks_step = (2u * p.win2_bytes) >> 4;
    constexpr uint32_t b_ks_step = 2u * 3u * N0, b2_ks_step = 2u * N0;
    constexpr uint32_t tap_step = (uint32_t)(2 * KSTEPS) * 3u * N0;             // one (dy,dz) block, in 16-B units
    const uint32_t w_b = ((sbase >> 4) & 0x3FFFu) | b_lbo;
    const uint32_t w2_b = (((sbase + p.w_bytes) >> 4) & 0x3FFFu) | b2_lbo;
    constexpr uint32_t idesc0 = (1u << 4) | kIdescAB | (8u << 24);   // D f32, A/B bf16 K-major, M = 128
    constexpr uint32_t ID1 = idesc0 | ((uint32_t)(N0 >> 3) << 17), ID2 = idesc0 | ((uint32_t)(2 * N0 >> 3) << 17),
                       ID3 = idesc0 | ((uint32_t)(3 * N0 >> 3) << 17);
    const uint32_t pitch_y = (uint32_t)p.ls.pitch_y;
    const uint32_t win0 = ((sbase + p.off_win) >> 4) & 0x3FFFu;                  // stage 0, in 16-B units
    const uint32_t stage16 = p.stage_bytes >> 4;
    // all (dy,dz,k-step) MMAs of one run of the band: B rows from block ja on, N = idesc's, into column d
    auto RUN = [&](uint32_t a0, uint32_t bb, uint32_t d, uint32_t idesc, uint32_t acc_first, bool skip_first) {
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const uint32_t a_row = a0 + (uint32_t)dy * pitch_y;
#pragma unroll
        for (int dz = 0; dz < 3; ++dz) {
#pragma unroll
          for (int ks = 0; ks < KSTEPS; ++ks) {
            const bool first = dy == 0 && dz == 0 && ks == 0;
            if (first && skip_first) continue;
            tc_mma_bf16(d, DESC(a_row + (uint32_t)dz + (uint32_t)ks * a_ks_step),
                        DESC(bb + (uint32_t)(dy * 3 + dz) * tap_step + (uint32_t)ks * b_ks_step), idesc, first ? acc_first : 1u);
          }
        }
      }
    };
    mbar_wait_warp(BAR(B_W_FULL), 0);
    int ws = 0;
    uint32_t wph = 0;
    uint32_t G = G_START;                        // outputs opened before this item
    auto SLOT = [&](uint32_t g) { return g % (uint32_t)NS; };
    auto PAR = [&](uint32_t g) { return (g / (uint32_t)NS) & 1u; };
    for (int it = 0; it < my_items; ++it) {
      int b_, cell0_, x0, x1;
      item_of(it, b_, cell0_, x0, x1);
      const int xa = x0 > 0 ? x0 - 1 : 0, xb = x1 < S ? x1 : S - 1;
      for (int xi = xa; xi <= xb; ++xi) {
        // band: row block j of the weight array feeds output xi - 1 + j, kept to the outputs of this march
        const int j_lo = xi - 1 >= x0 ? 0 : (xi >= x0 ? 1 : 2);
        const int j_hi = xi + 1 < x1 ? 2 : (xi < x1 ? 1 : 0);
        const int jf = (xi == xa && xa == x0) ? 1 : 2;          // first block this plane OPENS (the epilogue cleared it)
        const uint32_t gj0 = G + (uint32_t)(xi - x0) - 1u;      // running index of block 0's output
        for (int j = jf > j_lo ? jf : j_lo; j <= j_hi; ++j)      // the epilogue has drained the slots this plane opens
          mbar_wait_warp(BAR(B_ACC_EMPTY + (int)SLOT(gj0 + (uint32_t)j)), PAR(gj0 + (uint32_t)j) ^ 1u);
        mbar_wait_warp(BAR(B_WIN_FULL + ws), wph);
        tc_fence_after();
        const uint32_t stage = win0 + (uint32_t)ws * stage16;
        const uint32_t a0 = stage | a_lbo;
        if (leader) {
          // consecutive ring slots: run A up to the end of the ring, run B from slot 0
          const int nb = j_hi - j_lo + 1;
          const uint32_t sA = SLOT(gj0 + (uint32_t)j_lo);
          const int nA = nb < NS - (int)sA ? nb : NS - (int)sA, nB = nb - nA;
          const uint32_t dA = tmem_u + sA * N0, dB = tmem_u;
          const uint32_t bA = w_b + (uint32_t)(j_lo * N0), bB = w_b + (uint32_t)((j_lo + nA) * N0);
          auto IDN = [&](int n) { return ID1 + (uint32_t)(n - 1) * ((uint32_t)(N0 >> 3) << 17); };
          // every MMA accumulates: the epilogue leaves a drained slot cleared (tcgen05.st of zeros), so the slot a
          // plane opens needs no overwriting first MMA of its own (the N = 64 + N = 32 split cost 33 cycles per plane)
          RUN(a0, bA, dA, IDN(nA), 1u, false);
          if (nB > 0) RUN(a0, bB, dB, IDN(nB), 1u, false);
          if constexpr (KSTEPS2 > 0) {
            // fused 1x1 shortcut: the plane's own output, from the halo-free window of the second source
            if (xi >= x0 && xi < x1) {
              const uint32_t a2 = (stage + (uint32_t)(2 * KSTEPS) * (p.win_bytes >> 4)) | a2_lbo;
              const uint32_t dc = tmem_u + SLOT(gj0 + 1u) * N0;
#pragma unroll
              for (int ks = 0; ks < KSTEPS2; ++ks)
                tc_mma_bf16(dc, DESC(a2 + (uint32_t)ks * a2_ks_step), DESC(w2_b + (uint32_t)ks * b2_ks_step), ID1, 1u);
            }
          }
        }
        if (leader) {
          tc_commit(BAR(B_WIN_EMPTY + ws));
          if (j_lo == 0) tc_commit(BAR(B_ACC_FULL + (int)SLOT(gj0)));                             // output xi-1 is complete
          if (xi == xb && j_lo <= 1 && j_hi >= 1) tc_commit(BAR(B_ACC_FULL + (int)SLOT(gj0 + 1u)));  // so is the last one
        }
        if (++ws == p.stages) { ws = 0; wph ^= 1u; }
      }
      G += (uint32_t)(x1 - x0);
    }
    __syncwarp();
  } else {
    // ===================== epilogue =====================
    const int quarter = warp & 3;                   // TMEM lanes 32*quarter .. +31
    constexpr int HALVES = MARCH_EPI_WARPS / 4;   // warps per lane quarter: they alternate output planes
    const uint32_t half = (uint32_t)(warp - 2) >> 2;
    const bool has_res = (p.flags & (SCENEEGO_F_RESIDUAL | SCENEEGO_F_ADD_AFTER)) != 0;
    float bs[N0];
#pragma unroll
    for (int j = 0; j < N0; ++j) bs[j] = s_bias[j];
    uint32_t G = G_START;
    for (int it = 0; it < my_items; ++it) {
      int b, cell0, x0, x1;
      item_of(it, b, cell0, x0, x1);
      const int cell = cell0 + quarter * 32 + lane;
      const int y = (int)fdiv((uint32_t)cell, p.fd_py);
      const int z = cell - y * p.ls.pitch_y;
      const bool valid = y < S && z < S;            // pads keep their zeros: nothing is written there
      const int64_t dpos0 = (int64_t)b * p.ld.frame_pitch + p.ld.guard + (int64_t)y * p.ld.pitch_y + z;
      uint4 rn[2 * NCH];
      auto load_res = [&](int x) {
#pragma unroll
        for (int g = 0; g < 2 * NCH; ++g) {
          rn[g] = make_uint4(0, 0, 0, 0);
          if (has_res && valid && x < x1)
            rn[g] = *reinterpret_cast<const uint4*>(p.res + ((int64_t)g * p.ld.plane_stride + dpos0 + (int64_t)x * p.ld.pitch_x) * 8);
        }
      };
      int x = x0 + (HALVES == 2 ? (int)((G & 1u) ^ half) : 0);   // this warp's planes: running output counter parity == half
      load_res(x);
      for (; x < x1; x += HALVES) {
        const uint32_t gp = G + (uint32_t)(x - x0);
        const int slot = (int)(gp % (uint32_t)NS);
        uint4 rc[2 * NCH];
#pragma unroll
        for (int g = 0; g < 2 * NCH; ++g) rc[g] = rn[g];
        load_res(x + HALVES);                       // in flight while this plane is drained
        mbar_wait(BAR(B_ACC_FULL + slot), (gp / (uint32_t)NS) & 1u);
        tc_fence_after();
        uint32_t raw[NCH][16];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)slot * N0;
#pragma unroll
        for (int c = 0; c < NCH; ++c) tc_ld16(taddr + (uint32_t)(16 * c), raw[c]);
        tc_wait_ld();
#pragma unroll
        for (int c = 0; c < NCH; ++c) tc_st16_zero(taddr + (uint32_t)(16 * c));    // hand the slot back cleared
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_ACC_EMPTY + slot));
        if (valid) {
          const int64_t dpos = dpos0 + (int64_t)x * p.ld.pitch_x;
#pragma unroll
          for (int g = 0; g < 2 * NCH; ++g) {
            float rr[8], o[8];
            unpack8(rc[g], rr);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float t = __uint_as_float(raw[g >> 1][(g & 1) * 8 + j]) + bs[8 * g + j];
              if (p.flags & SCENEEGO_F_RESIDUAL) t += rr[j];
              if (p.flags & SCENEEGO_F_RELU) t = fmaxf(t, 0.f);
              if (p.flags & SCENEEGO_F_ADD_AFTER) t += rr[j];
              o[j] = t;
            }
            *reinterpret_cast<uint4*>(p.dst + ((int64_t)g * p.ld.plane_stride + dpos) * 8) = pack8(o);
          }
        }
      }
      G += (uint32_t)(x1 - x0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
}

typedef void (*march_fn)(const MarchParams);
template <int TWO>
static march_fn pick_march_t(int ksteps, int ksteps2, int nch) {
  if (nch == 2 && ksteps == 2 && ksteps2 == 0) return conv_march_kernel<2, 0, 2, TWO>;
  if (nch == 2 && ksteps == 2 && ksteps2 == 1) return conv_march_kernel<2, 1, 2, TWO>;
  if (nch == 2 && ksteps == 1 && ksteps2 == 0) return conv_march_kernel<1, 0, 2, TWO>;
  if (nch == 1 && ksteps == 2 && ksteps2 == 0) return conv_march_kernel<2, 0, 1, TWO>;
  if (nch == 1 && ksteps == 1 && ksteps2 == 0) return conv_march_kernel<1, 0, 1, TWO>;
  return nullptr;
}
static march_fn pick_march(int ksteps, int ksteps2, int nch, int two) {
  return two ? pick_march_t<1>(ksteps, ksteps2, nch) : pick_march_t<0>(ksteps, ksteps2, nch);
}

// Called by sceneego_v2v_run for SCENEEGO_OP_CONV3_MARCH (tensor path; the CUDA-core checker lives in v2v.cu).
int launch_conv_march(const sceneego_v2v_op_t& op, void* const* d_buffers, const void* d_blob, int batch, int op_index,
                      cudaStream_t st) {
  MarchParams p;
  memset(&p, 0, sizeof(p));
  p.src = (const __nv_bfloat16*)d_buffers[op.src];
  p.dst = (__nv_bfloat16*)d_buffers[op.dst];
  p.res = op.res >= 0 ? (const __nv_bfloat16*)d_buffers[op.res] : nullptr;
  p.src2 = (op.src2 >= 0 && op.cin2 > 0) ? (const __nv_bfloat16*)d_buffers[op.src2] : nullptr;
  p.w = (const __nv_bfloat16*)((const char*)d_blob + op.w_offset);
  p.bias = (const float*)((const char*)d_blob + op.b_offset);
  p.ls = op.lay_src; p.ld = op.lay_dst; p.batch = batch; p.flags = op.flags;
  const int S = p.ls.side;
  SE_REQUIRE(op.ksize == 3 && op.cin % 16 == 0 && op.cout % 16 == 0 && 3 * op.cout <= 256 && S >= 2,
             "v2v_run: op %d: the marching conv needs a 3^3 stencil, channels in multiples of 16 and 3*cout <= 256", op_index);
  SE_REQUIRE(!(op.flags & SCENEEGO_F_OUT_F32) && op.cout_real == op.cout, "v2v_run: op %d: the marching conv writes planar bf16", op_index);
  SE_REQUIRE(p.ls.s2d == 0 && p.ld.s2d == 0 && p.ls.pad >= 1 && p.ld.side == S && p.ls.guard >= p.ls.pitch_y + 1,
             "v2v_run: op %d: layouts incompatible with the marching conv", op_index);
  SE_REQUIRE(!p.src2 || (op.cin2 % 16 == 0 && op.res < 0), "v2v_run: op %d: bad fused shortcut", op_index);
  p.cin_planes = op.cin / 8;
  p.cin2_planes = p.src2 ? op.cin2 / 8 : 0;
  p.n0 = op.cout;
  int two = 1;                         // two CTAs per SM (see march_epi_warps) unless the weights need more than half an SM
  { const char* e = getenv("SCENEEGO_MARCH_CTAS"); if (e && atoi(e) == 1) two = 0; }
  p.n_slots = march_tmem_cols(two) / p.n0;
  if (p.n_slots > MARCH_MAX_SLOTS) p.n_slots = MARCH_MAX_SLOTS;
  p.halo = p.ls.pitch_y + 1;
  p.tiles_per_plane = ((S - 1) * p.ls.pitch_y + S + MARCH_L - 1) / MARCH_L;
  p.n_items = batch * p.tiles_per_plane;
  p.win_bytes = (uint32_t)(MARCH_L + 2 * p.halo) * 16u;
  p.win2_bytes = (uint32_t)MARCH_L * 16u;
  p.stage_bytes = (uint32_t)p.cin_planes * p.win_bytes + (uint32_t)p.cin2_planes * p.win2_bytes;
  p.w_bytes = 9u * (uint32_t)p.cin_planes * 3u * (uint32_t)p.n0 * 16u;
  p.w2_bytes = (uint32_t)p.cin2_planes * (uint32_t)p.n0 * 16u;
  const uint32_t fixed = 512 + 1024;   // bias + barriers
  const uint32_t kHalfSmem = (kMaxSmem - 2048) / 2;           // two CTAs per SM, 1 KB reserved by the system per CTA
  if (two && (uint64_t)p.w_bytes + p.w2_bytes + 2ull * p.stage_bytes + fixed > kHalfSmem) {
    two = 0;
    p.n_slots = march_tmem_cols(0) / p.n0;
    if (p.n_slots > MARCH_MAX_SLOTS) p.n_slots = MARCH_MAX_SLOTS;
  }
  const uint32_t smem_cap = two ? kHalfSmem : kMaxSmem;
  SE_REQUIRE((uint64_t)p.w_bytes + p.w2_bytes + 2ull * p.stage_bytes + fixed <= smem_cap,
             "v2v_run: op %d: resident weights + two window stages exceed shared memory", op_index);
  int stages = (int)((smem_cap - p.w_bytes - p.w2_bytes - fixed) / p.stage_bytes);
  if (two) {
    // two CTAs at ~107 KB each leave the SM ~12 KB of L1 for the epilogue's residual loads and stores: three window
    // stages measured 3 % SLOWER than two on the 32 -> 32 layers (12.5 vs 12.1 us).  Keep each CTA under 92 KB
    // (>= 44 KB of L1) unless that would leave fewer than two stages.
    int capped = (int)(((int64_t)92 * 1024 - p.w_bytes - p.w2_bytes - fixed) / (int64_t)p.stage_bytes);
    if (capped < 2) capped = 2;
    if (stages > capped) stages = capped;
  }
  if (stages > MARCH_MAX_STAGES) stages = MARCH_MAX_STAGES;
  { const char* e = getenv("SCENEEGO_MARCH_STAGES"); if (e && atoi(e) >= 2 && atoi(e) <= stages) stages = atoi(e); }
  p.stages = stages;
  p.off_win = p.w_bytes + p.w2_bytes;
  p.off_bias = p.off_win + (uint32_t)stages * p.stage_bytes;
  p.off_bar = p.off_bias + 512;
  p.fd_py = make_fastdiv((uint32_t)p.ls.pitch_y);
  p.fd_tpp = make_fastdiv((uint32_t)p.tiles_per_plane);
  SE_REQUIRE((int64_t)batch * p.ls.frame_pitch + 4096 < (1ll << 31), "v2v_run: op %d: batch * frame_pitch too large for one launch", op_index);
  march_fn fn = pick_march(op.cin / 16, p.cin2_planes / 2, op.cout / 16, two);
  SE_REQUIRE(fn != nullptr, "v2v_run: op %d: no conv_march instantiation for cin=%d cin2=%d cout=%d", op_index, op.cin,
             p.src2 ? op.cin2 : 0, op.cout);
  if (int rc = ensure_max_dynamic_smem((const void*)fn, (int)kMaxSmem)) return rc;
  // CTAs take equal contiguous ranges of (item, plane) pairs; at least 16 planes each so that the one or two
  // extra input planes at the ends of a partial march stay a small fraction
  const int slots = kNumSMs * (two ? 2 : 1);
  int grid = (int)(((int64_t)p.n_items * S + 15) / 16);
  if (grid > slots) grid = slots;
  if (grid < 1) grid = 1;
  fn<<<grid, march_threads(two), (size_t)p.off_bar + 1024, st>>>(p);
  SE_CUDA_LAUNCH_CHECK("conv_march");
  return SCENEEGO_OK;
}

}  // namespace sceneego

using namespace sceneego;

extern "C" int sceneego_v2v_pack_conv_march(const float* h_weight, const float* h_bias, const float* h_gamma,
                                            const float* h_beta, const float* h_mean, const float* h_var, double eps,
                                            int cout, int cin, int cout_pad, int cin_pad, uint16_t* h_w_out, float* h_b_out) {
  SE_REQUIRE(h_weight && h_w_out && h_b_out, "pack_conv_march: null argument");
  SE_REQUIRE(cout_pad >= cout && cin_pad >= cin && cout_pad % 16 == 0 && cin_pad % 16 == 0, "pack_conv_march: bad padding");
  // fold exactly like sceneego_v2v_pack_conv ([tap][cin/8][cout][8], taps ordered (dx,dy,dz)), then regroup
  const size_t tap_elems = (size_t)cin_pad * cout_pad;
  uint16_t* tmp = (uint16_t*)malloc(27 * tap_elems * sizeof(uint16_t));
  SE_REQUIRE(tmp != nullptr, "pack_conv_march: out of memory");
  const int rc = sceneego_v2v_pack_conv(h_weight, h_bias, h_gamma, h_beta, h_mean, h_var, eps, cout, cin, 3, 0, cout_pad,
                                        cin_pad, 1, 1, tmp, h_b_out);
  if (rc == SCENEEGO_OK) {
    const int g8 = cin_pad / 8;
    for (int dx = 0; dx < 3; ++dx)
      for (int t = 0; t < 9; ++t)
        for (int g = 0; g < g8; ++g)
          memcpy(h_w_out + ((((size_t)t * g8 + g) * 3 + (2 - dx)) * cout_pad) * 8,
                 tmp + (((size_t)(dx * 9 + t) * g8 + g) * cout_pad) * 8, (size_t)cout_pad * 8 * sizeof(uint16_t));
  }
  free(tmp);
  return rc;
}
