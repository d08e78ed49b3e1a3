// a7: ConvTranspose3d(k2, s2) + folded BN + ReLU, then "+ skip" (network/v2v.py:55-67,125-137) for the two transposed
// convs that carry data (128 -> 64 at 16^3 -> 32^3 and 64 -> 32 at 32^3 -> 64^3: 16 of the 18.7 us per frame the five
// take).  HBM-bound: 4.2 MB read + 16.8 MB skip read + 16.8 MB written for 64 -> 32, 5.8 us at the copy peak.
//
// On conv_tc_kernel (round 1) the GEMM is fine -- rows = input cells, N = (parities) x Cout -- but its epilogue lets
// every thread (= one input cell) store its eight output cells itself: for a fixed parity a warp writes 32 cells at a
// stride of 32 bytes, every store instruction half-fills 32 sectors, and the kernel ran at 0.41 of the HBM peak.  Here
// the epilogue is two-phase: (1) tensor memory -> bias, ReLU, + skip, 16-bit -> a SHARED-MEMORY image of the output
// z-lines (cell 2 r + pz of line (plane, px, py) for input row r); (2) the image goes out with consecutive lanes
// writing consecutive 16-byte cells -- 512 contiguous bytes per warp instruction whatever the alignment of the line.
// The skip cells are still fetched by the row's thread (strided 16-byte loads, prefetched per column chunk).
// warp 0 = producer (weights once -- they stay resident -- and one halo-free window per item), warp 1 = MMA issuer,
// warps 2..5 = epilogue.  TWO CTAs per SM (112 KB of shared memory and 256 tensor-memory columns each): within a CTA the
// item is a strict chain (window -> 4 MMAs -> loads + math -> barrier -> stores), so one CTA keeps either loads or
// stores on the wire; two interleave them.  One CTA per SM measured 13.5 us (4 epilogue warps: 22 us) for 64 -> 32
// against the 12.2 us of the generic path; all sixteen column chunks' skip cells are requested before the accumulator
// is awaited (64 KB in flight per SM).
// Weight blob and CUDA-core checker are the ones of the generic path (sceneego_v2v_pack_conv transposed, deconv2_kernel).
#include "tc_common.cuh"
#include <stdlib.h>

namespace sceneego {

constexpr int DC_EPI_WARPS = 4;                          // the four tensor-memory lane quarters
constexpr int DC_THREADS = 32 * (2 + DC_EPI_WARPS);
constexpr int DC_STAGE_BYTES = 128 * 256 * 2;            // one parity group of a tile as 16-bit cells: 64 KB

struct DeconvParams {
  const __nv_bfloat16* src;
  const __nv_bfloat16* res;    // skip tensor (ADD_AFTER) or nullptr
  __nv_bfloat16* dst;
  const uint8_t* w;            // [group][cin/8][256 rows][8]
  const float* bias;           // [cout]
  sceneego_vol_layout_t ls, ld;
  int batch, flags, cin_planes, cout, n_groups, n_items, win_stages;
  uint32_t win_bytes, w_bytes, group_bytes;
  uint32_t off_win, off_stage, off_rows, off_bias, off_bar;
  FastDiv fd_frame, fd_px, fd_py;
};

template <int KSTEPS, int COUT>
__global__ void __launch_bounds__(DC_THREADS, 2) deconv_tc_kernel(const __grid_constant__ DeconvParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  float* s_bias = reinterpret_cast<float*>(smem + p.off_bias);
  int* s_rowpos = reinterpret_cast<int*>(smem + p.off_rows);        // per input row: output position of (2X, 2Y, 2Z), or -1
  const uint32_t bar0 = sbase + p.off_bar;
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  constexpr int B_W_FULL = 0, B_WIN_FULL = 1, B_WIN_EMPTY = 3, B_TM_FULL = 5, B_TM_EMPTY = 7, B_COUNT = 9;
  uint32_t* s_tmem_ptr = reinterpret_cast<uint32_t*>(smem + p.off_bar + 8 * B_COUNT);
  constexpr int NPAR = 256 / COUT;                  // parities per group (8 or 4)
  constexpr int NGROUPS = 8 / NPAR;
  constexpr int PLANES = COUT / 8;

  for (int i = threadIdx.x; i < COUT; i += DC_THREADS) s_bias[i] = p.bias[i];
  if (threadIdx.x == 0) {
    mbar_init(BAR(B_W_FULL), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(BAR(B_WIN_FULL + i), 1); mbar_init(BAR(B_WIN_EMPTY + i), 1);
      mbar_init(BAR(B_TM_FULL + i), 1); mbar_init(BAR(B_TM_EMPTY + i), DC_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem_ptr)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem_ptr;
  const int my_items = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    // ===================== producer =====================
    if (lane == 0) {
      mbar_expect_tx(BAR(B_W_FULL), p.w_bytes);
      for (uint32_t o = 0; o < p.w_bytes; o += 16384u)
        bulk_g2s(sbase + o, p.w + o, p.w_bytes - o < 16384u ? p.w_bytes - o : 16384u, BAR(B_W_FULL));
      for (int it = 0; it < my_items; ++it) {
        const int ws = p.win_stages == 2 ? (it & 1) : 0;
        const int wph = p.win_stages == 2 ? ((it >> 1) & 1) : (it & 1);
        const int64_t q0 = (int64_t)p.ls.guard + ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 128;
        mbar_wait(BAR(B_WIN_EMPTY + ws), wph ^ 1);
        mbar_expect_tx(BAR(B_WIN_FULL + ws), p.win_bytes);
#pragma unroll
        for (int g = 0; g < 2 * KSTEPS; ++g)
          bulk_g2s(sbase + p.off_win + (uint32_t)ws * p.win_bytes + (uint32_t)g * 2048u, p.src + ((int64_t)g * p.ls.plane_stride + q0) * 8, 2048u,
                   BAR(B_WIN_FULL + ws));
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const bool leader = elect_one();
    constexpr uint32_t idesc = (1u << 4) | kIdescAB | ((256u >> 3) << 17) | (8u << 24);
    const uint64_t desc_hi = (uint64_t)(8u | (1u << 14)) << 32;
    constexpr uint32_t a_lbo = 128u << 16;               // window planes are 128 cells (2 KB) apart
    constexpr uint32_t b_lbo = 256u << 16;               // [k-chunk][256 rows][8]
    mbar_wait_warp(BAR(B_W_FULL), 0);
    uint32_t k = 0;                                      // running (item, group) counter: accumulator buffer k & 1
    for (int it = 0; it < my_items; ++it) {
      const int ws = p.win_stages == 2 ? (it & 1) : 0;
      mbar_wait_warp(BAR(B_WIN_FULL + ws), (uint32_t)(p.win_stages == 2 ? ((it >> 1) & 1) : (it & 1)));
      const uint32_t a16 = ((sbase + p.off_win + (uint32_t)ws * p.win_bytes) >> 4) & 0x3FFFu;
      for (int gi = 0; gi < NGROUPS; ++gi, ++k) {
        const uint32_t buf = 0u;                           // one accumulator buffer per CTA (the other CTA of the SM has its own)
        mbar_wait_warp(BAR(B_TM_EMPTY + (int)buf), (k & 1u) ^ 1u);
        tc_fence_after();
        if (leader) {
          const uint32_t b16 = (((sbase + (uint32_t)gi * p.group_bytes) >> 4) & 0x3FFFu) | b_lbo;
#pragma unroll
          for (int ks = 0; ks < KSTEPS; ++ks)
            tc_mma_bf16(tmem_u + buf * 256u, desc_hi | (uint64_t)((a16 + (uint32_t)ks * 256u) | a_lbo), desc_hi | (uint64_t)(b16 + (uint32_t)ks * 512u),
                        idesc, ks == 0 ? 0u : 1u);
          if (gi == NGROUPS - 1) tc_commit(BAR(B_WIN_EMPTY + ws));
          tc_commit(BAR(B_TM_FULL + (int)buf));
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue: warps 2..5 = the four tensor-memory lane quarters =====================
    const int quarter = warp & 3;
    constexpr int half = 0;
    const int row = quarter * 32 + lane;
    const int et = threadIdx.x - 64;                     // 0..127 among the epilogue threads
    const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const bool add_after = (p.flags & SCENEEGO_F_ADD_AFTER) != 0;
    const bool relu = (p.flags & SCENEEGO_F_RELU) != 0;
    auto epi_sync = [&]() { asm volatile("bar.sync 1, 128;" ::: "memory"); };
    uint8_t* stage = smem + p.off_stage;
    uint32_t k = 0;
    for (int it = 0; it < my_items; ++it) {
      // this thread's input row -> output position of voxel (2X, 2Y, 2Z)
      const uint32_t q = (uint32_t)((int64_t)p.ls.guard + ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * 128 + row);
      const uint32_t b = fdiv(q, p.fd_frame);
      const int rem = (int)(q - b * (uint32_t)p.ls.frame_pitch) - p.ls.guard;
      int dpos = -1;
      if ((int)b < p.batch && rem >= 0) {
        const int X = (int)fdiv((uint32_t)rem, p.fd_px);
        const int r2 = rem - X * p.ls.pitch_x;
        const int Y = (int)fdiv((uint32_t)r2, p.fd_py);
        const int Z = r2 - Y * p.ls.pitch_y;
        if (X < p.ls.side && Y < p.ls.side && Z < p.ls.side)
          dpos = (int)((int64_t)b * p.ld.frame_pitch + p.ld.guard + (int64_t)(2 * X) * p.ld.pitch_x + (int64_t)(2 * Y) * p.ld.pitch_y + 2 * Z);
      }
      epi_sync();                                        // phase 2 of the previous item has read s_rowpos
      s_rowpos[row] = dpos;
      for (int gi = 0; gi < NGROUPS; ++gi, ++k) {
        const uint32_t buf = 0u;
        // ---- phase 1: accumulators -> bias, ReLU, + skip -> 16-bit cells in the shared-memory image of the output lines
        auto skip_cells = [&](int c, uint4& s0, uint4& s1) {    // the two planes of column chunk c (16 channels of one parity)
          s0 = make_uint4(0, 0, 0, 0); s1 = s0;
          if (add_after && dpos >= 0) {
            const int par = gi * NPAR + (16 * c) / COUT, ch0 = (16 * c) % COUT;
            const int64_t cell = (int64_t)dpos + (par >> 2) * p.ld.pitch_x + ((par >> 1) & 1) * p.ld.pitch_y + (par & 1);
            s0 = *reinterpret_cast<const uint4*>(p.res + ((int64_t)(ch0 >> 3) * p.ld.plane_stride + cell) * 8);
            s1 = *reinterpret_cast<const uint4*>(p.res + ((int64_t)((ch0 >> 3) + 1) * p.ld.plane_stride + cell) * 8);
          }
        };
        // the skip cells of eight column chunks ahead are kept in flight
        constexpr int D = 8;
        uint4 q0[D], q1[D];
#pragma unroll
        for (int d = 0; d < D; ++d) skip_cells(d, q0[d], q1[d]);
        mbar_wait(BAR(B_TM_FULL + (int)buf), k & 1u);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 16; ++c) {
          const uint4 s0 = q0[0], s1 = q1[0];
#pragma unroll
          for (int d = 0; d + 1 < D; ++d) { q0[d] = q0[d + 1]; q1[d] = q1[d + 1]; }
          if (c + D < 16) skip_cells(c + D, q0[D - 1], q1[D - 1]);
          uint32_t raw[16];
          tc_ld16(taddr0 + buf * 256u + (uint32_t)(16 * c), raw);
          tc_wait_ld();
          const int par_l = (16 * c) / COUT, ch0 = (16 * c) % COUT;       // parity inside the group, first channel
          const int par = gi * NPAR + par_l;
          float sk[16];
          {
            float t8[8];
            unpack8(s0, t8);
#pragma unroll
            for (int j = 0; j < 8; ++j) sk[j] = t8[j];
            unpack8(s1, t8);
#pragma unroll
            for (int j = 0; j < 8; ++j) sk[8 + j] = t8[j];
          }
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float t = __uint_as_float(raw[8 * g + j]) + s_bias[ch0 + 8 * g + j];
              if (relu) t = fmaxf(t, 0.f);
              if (add_after) t += sk[8 * g + j];
              o[j] = t;
            }
            // image: [plane][px, py of this group's parities][cell 2 r + pz]
            const int line = ((ch0 >> 3) + g) * (NPAR / 2) + (par_l >> 1);
            *reinterpret_cast<uint4*>(stage + ((size_t)line * 256 + 2 * row + (par & 1)) * 16) = pack8(o);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(BAR(B_TM_EMPTY + (int)buf));
        epi_sync();                                      // the image (and s_rowpos) is complete
        // ---- phase 2: consecutive lanes write consecutive output cells
        constexpr int LINES = PLANES * (NPAR / 2);
#pragma unroll 8
        for (int i = et; i < LINES * 256; i += 128) {
          const int line = i >> 8, cell = i & 255;
          const int rp = s_rowpos[cell >> 1];
          if (rp >= 0) {
            const int plane = line / (NPAR / 2), pxy = line % (NPAR / 2);
            const int par_hi = gi * (NPAR / 2) + pxy;    // (px, py) = bits of par >> 1
            const int64_t dst = (int64_t)rp + (par_hi >> 1) * p.ld.pitch_x + (par_hi & 1) * p.ld.pitch_y + (cell & 1);
            *reinterpret_cast<uint4*>(p.dst + ((int64_t)plane * p.ld.plane_stride + dst) * 8) =
                *reinterpret_cast<const uint4*>(stage + (size_t)i * 16);
          }
        }
        epi_sync();                                      // the image may be overwritten
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
}

// Tensor path of SCENEEGO_OP_DECONV2 for 64 -> 32 and 128 -> 64 channels.  Returns SCENEEGO_E_UNSUPPORTED for every other
// shape: the caller then runs the transposed conv on conv_tc_kernel as before.
int launch_deconv_tc(const sceneego_v2v_op_t& op, void* const* d_buffers, const void* d_blob, int batch, int op_index,
                     cudaStream_t st) {
  // 64 -> 32 only: the 128 -> 64 layer's 128 KB of weights leave no room for a second CTA, and with one CTA per SM this
  // kernel is slower than the generic path (4.1 - 4.5 vs 3.8 us)
  const bool s6432 = op.cin == 64 && op.cout == 32;
  if (!s6432 || op.cout_real != op.cout) return SCENEEGO_E_UNSUPPORTED;
  { static const bool off = getenv("SCENEEGO_DECONV_IMPL") && atoi(getenv("SCENEEGO_DECONV_IMPL")) == 0; if (off) return SCENEEGO_E_UNSUPPORTED; }
  DeconvParams p;
  memset(&p, 0, sizeof(p));
  p.src = (const __nv_bfloat16*)d_buffers[op.src];
  p.dst = (__nv_bfloat16*)d_buffers[op.dst];
  p.res = op.res >= 0 ? (const __nv_bfloat16*)d_buffers[op.res] : nullptr;
  p.w = (const uint8_t*)d_blob + op.w_offset;
  p.bias = (const float*)((const char*)d_blob + op.b_offset);
  p.ls = op.lay_src; p.ld = op.lay_dst; p.batch = batch; p.flags = op.flags;
  SE_REQUIRE(p.src && p.dst && p.ls.s2d == 0 && p.ld.s2d == 0 && p.ld.side == 2 * p.ls.side, "v2v_run: op %d bad deconv", op_index);
  SE_REQUIRE(!(op.flags & SCENEEGO_F_ADD_AFTER) || p.res, "v2v_run: op %d needs a skip tensor", op_index);
  SE_REQUIRE((int64_t)batch * p.ld.frame_pitch + p.ld.guard + 4096 < (1ll << 31), "v2v_run: op %d: batch too large for one launch", op_index);
  p.cin_planes = op.cin / 8;
  p.cout = op.cout;
  p.n_groups = 8 / (256 / op.cout);
  p.group_bytes = (uint32_t)p.cin_planes * 256u * 16u;
  p.w_bytes = (uint32_t)p.n_groups * p.group_bytes;
  p.win_bytes = (uint32_t)p.cin_planes * 2048u;
  const int64_t n_pos = (int64_t)batch * p.ls.frame_pitch;
  p.n_items = (int)((n_pos + 127) / 128);
  p.off_win = (p.w_bytes + 127u) / 128u * 128u;
  p.win_stages = 1;                                      // two CTAs per SM: 32 + 16 + 64 KB each
  p.off_stage = p.off_win + (uint32_t)p.win_stages * p.win_bytes;
  p.off_rows = p.off_stage + DC_STAGE_BYTES;
  p.off_bias = p.off_rows + 512u;
  p.off_bar = p.off_bias + 256u;
  const size_t smem_bytes = (size_t)p.off_bar + 8 * 16 + 64;
  if (smem_bytes > (kMaxSmem - 2048) / 2) return SCENEEGO_E_UNSUPPORTED;
  p.fd_frame = make_fastdiv((uint32_t)p.ls.frame_pitch);
  p.fd_px = make_fastdiv((uint32_t)p.ls.pitch_x);
  p.fd_py = make_fastdiv((uint32_t)p.ls.pitch_y);
  const int grid = p.n_items < 2 * kNumSMs ? p.n_items : 2 * kNumSMs;
  if (int rc = ensure_max_dynamic_smem((const void*)deconv_tc_kernel<4, 32>, (int)kMaxSmem)) return rc;
  deconv_tc_kernel<4, 32><<<grid, DC_THREADS, smem_bytes, st>>>(p);
  SE_CUDA_LAUNCH_CHECK("deconv_tc");
  return SCENEEGO_OK;
}

}  // namespace sceneego
