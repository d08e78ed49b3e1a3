// a7, last three layers: Basic3DBlock(32,32,1) x 2 + Conv3d(32,15,1) (network/v2v.py:150-161,168-169)
// fused into one pass.  A per-voxel 32 -> 32 -> 32 -> 15 MLP is HBM-bound (read 16.8 MB of bf16
// activations, write 15.7 MB of f32 logits per 64^3 frame; 1.3 GFLOP), so the point is to touch
// memory once: three separate 1x1 launches cost 29.6 us/frame on B200, the compulsory traffic is 5 us.
// The intermediates never leave registers: the fp32 accumulator fragment of mma.sync.m16n8k16 has
// exactly the thread layout of the next layer's bf16 A fragment (two n8 tiles = one k16 tile), so
// bias + ReLU + bf16 rounding happen in place -- the same roundings as the unfused chain.
// (Warp-level mma.sync rather than tcgen05: the chain of three dependent GEMMs per 16 rows is
// latency-, not throughput-bound, and needs no shared-memory round trips this way.)
#include "tc_common.cuh"

namespace sceneego {

struct TailParams {
  const __nv_bfloat16* src;   // 32 channels = 4 planes, layout ls
  float* dst;                 // (B, cout_real, S, S, S) f32
  const __nv_bfloat16* w1;    // [4][32][8]   (sceneego_v2v_pack_conv layout, one tap)
  const __nv_bfloat16* w2;    // [4][32][8]
  const __nv_bfloat16* w3;    // [4][16][8]
  const float *b1, *b2, *b3;  // 32, 32, 16
  sceneego_vol_layout_t ls;
  int batch, cout_real;
  int64_t n_pos;              // positions to scan, starting at ls.guard
  FastDiv fd_frame, fd_px, fd_py;
};

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32" SE_MMA_SYNC_AB ".f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  return act_pack2(lo, hi);
}

// B fragments of a [cin/8][cout][8] weight block: thread (g = lane/4, q = lane%4) of n-tile nt, k-tile kt holds
// (k = 16kt + 2q, +1; n = 8nt + g) and (k = 16kt + 8 + 2q, +1; n = 8nt + g).
template <int NT>
__device__ __forceinline__ void load_b(const __nv_bfloat16* w, int cout, int g, int q, uint32_t (&b)[NT][2][2]) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {
      b[nt][kt][0] = *reinterpret_cast<const uint32_t*>(w + ((size_t)(2 * kt) * cout + nt * 8 + g) * 8 + 2 * q);
      b[nt][kt][1] = *reinterpret_cast<const uint32_t*>(w + ((size_t)(2 * kt + 1) * cout + nt * 8 + g) * 8 + 2 * q);
    }
}

// One hidden layer on a 16-row tile: a (2 k-tiles) -> relu(a W^T + bias) as the next layer's A fragments.
__device__ __forceinline__ void hidden_layer(const uint32_t (&a)[2][4], const uint32_t (&b)[4][2][2],
                                             const float* bias /* shared: this thread's column pair of n-tile 0 */,
                                             uint32_t (&out)[2][4]) {
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const float2 bv = *reinterpret_cast<const float2*>(bias + nt * 8);
    float c[4] = {bv.x, bv.y, bv.x, bv.y};
    mma_bf16_16816(c, a[0], b[nt][0][0], b[nt][0][1]);
    mma_bf16_16816(c, a[1], b[nt][1][0], b[nt][1][1]);
    // accumulator (rows g / g+8, cols 8nt + 2q, +1) == A fragment slots of k-tile nt/2: even nt -> a0,a1; odd -> a2,a3
    out[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(fmaxf(c[0], 0.f), fmaxf(c[1], 0.f));
    out[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(fmaxf(c[2], 0.f), fmaxf(c[3], 0.f));
  }
}

__global__ void __launch_bounds__(256, 2) tail_mlp_kernel(const __grid_constant__ TailParams p) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  uint32_t B1[4][2][2], B2[4][2][2], B3[2][2][2];
  load_b<4>(p.w1, 32, g, q, B1);
  load_b<4>(p.w2, 32, g, q, B2);
  load_b<2>(p.w3, 16, g, q, B3);
  __shared__ __align__(8) float s_bias[80];      // b1 | b2 | b3 (registers go to the A tiles in flight instead)
  if (threadIdx.x < 80) s_bias[threadIdx.x] = threadIdx.x < 32 ? p.b1[threadIdx.x] : threadIdx.x < 64 ? p.b2[threadIdx.x - 32] : p.b3[threadIdx.x - 64];
  __syncthreads();
  const float* bias1 = s_bias + 2 * q;
  const float* bias2 = s_bias + 32 + 2 * q;
  const float* bias3 = s_bias + 64 + 2 * q;

  const int S = p.ls.side;
  const size_t N3 = (size_t)S * S * S;
  const int64_t warp_global = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  constexpr int TPI = 4;                                  // 16-row tiles per warp iteration: 64 positions, 4 KB in flight
  const int64_t n_chunks = (p.n_pos + 16 * TPI - 1) / (16 * TPI);
  const uint32_t* src32 = reinterpret_cast<const uint32_t*>(p.src);
  const int64_t plane32 = p.ls.plane_stride * 4;          // plane stride in 4-byte words
  for (int64_t ch = warp_global; ch < n_chunks; ch += n_warps) {
    const int64_t q0 = (int64_t)p.ls.guard + ch * (16 * TPI);
    // A fragments of the 16-row tiles: rows (g, g+8) of tile t, 4-byte word q of the row's cell in plane 2kt / 2kt+1
    uint32_t a[TPI][2][4];
#pragma unroll
    for (int t = 0; t < TPI; ++t)
#pragma unroll
      for (int kt = 0; kt < 2; ++kt) {
        const int64_t w0 = (q0 + t * 16 + g) * 4 + q + (int64_t)(2 * kt) * plane32;
        a[t][kt][0] = __ldcs(src32 + w0);
        a[t][kt][1] = __ldcs(src32 + w0 + 8 * 4);
        a[t][kt][2] = __ldcs(src32 + w0 + plane32);
        a[t][kt][3] = __ldcs(src32 + w0 + plane32 + 8 * 4);
      }
#pragma unroll
    for (int t = 0; t < TPI; ++t) {
      uint32_t h1[2][4], h2[2][4];
      hidden_layer(a[t], B1, bias1, h1);
      hidden_layer(h1, B2, bias2, h2);
      float c[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const float2 bv = *reinterpret_cast<const float2*>(bias3 + nt * 8);
        c[nt][0] = bv.x; c[nt][1] = bv.y; c[nt][2] = bv.x; c[nt][3] = bv.y;
        mma_bf16_16816(c[nt], h2[0], B3[nt][0][0], B3[nt][0][1]);
        mma_bf16_16816(c[nt], h2[1], B3[nt][1][0], B3[nt][1][1]);
      }
      // rows g and g+8 of this tile -> (frame, voxel); pads and guards are skipped
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const uint32_t pos = (uint32_t)(q0 + t * 16 + g + 8 * r);
        const uint32_t b = fdiv(pos, p.fd_frame);
        const int rem = (int)(pos - b * (uint32_t)p.ls.frame_pitch) - p.ls.guard;
        if ((int)b >= p.batch || rem < 0) continue;
        const int x = (int)fdiv((uint32_t)rem, p.fd_px);
        const int r2 = rem - x * p.ls.pitch_x;
        const int y = (int)fdiv((uint32_t)r2, p.fd_py);
        const int z = r2 - y * p.ls.pitch_y;
        if (x >= S || y >= S || z >= S) continue;
        float* o = p.dst + (size_t)b * p.cout_real * N3 + ((size_t)x * S + y) * S + z;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int col = nt * 8 + 2 * q + j;
            if (col < p.cout_real) __stcs(o + (size_t)col * N3, c[nt][2 * r + j]);
          }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// tcgen05 variant (default).  The mma.sync kernel above is latency-bound (ncu: issue active 50 %, 16 warps per
// SM at 128 registers; rolling prefetch, more occupancy and cross-tile interleaving all made it slower), so the
// chain of three GEMMs moves to the 5th-generation tensor cores with the intermediates staying on the SM:
//   one CTA per SM = 8 independent groups of 4 warps; a group walks its own 128-position tiles:
//     TMA (4 planes x 128 cells x 16 B, double-buffered)              -> A tile in shared memory
//     tcgen05.mma M=128 N=32 (W1 resident in smem)                    -> 32 TMEM columns
//     tcgen05.ld, ReLU + bf16 in one F2FP (the unfused chain's rounding), tcgen05.st -> 16 TMEM columns = layer 2's A operand
//     tcgen05.mma N=32 (W2, A from tensor memory) -> ld / pack / st -> tcgen05.mma N=16 (W3, A from tensor memory) -> ld -> f32 logits
//   The biases ride in the MMAs: every layer starts with a K=16 MMA of a constant "ones" tile against a B block whose
//   first two K rows hold the bias split in two 16-bit parts (hi + lo: exact to 2^-17 of the bias), so the epilogues
//   have no adds and no shared-memory bias reads.  Round-2 form: with the hidden activations going through shared
//   memory (four 16-byte stores per row and layer + a proxy fence) and the bias added in the epilogue the kernel was
//   bound by instruction issue (~330 instructions per voxel, ncu: issue active 69 %, DRAM 47 %).
//   The 4 warps of a group are the four TMEM lane quarters; its first lane issues the TMA and the MMAs, a named
//   barrier per group orders "hidden tile written" before "MMA issued".  Eight groups keep eight such dependent chains
//   in flight per SM, which is what hides their latency.
// ---------------------------------------------------------------------------
constexpr int TT_GROUPS = 8;
constexpr int TT_THREADS = TT_GROUPS * 128;
constexpr uint32_t TT_TILE_BYTES = 4u * 128u * 16u;          // 4 channel-group planes x 128 rows x 16 B
// shared-memory layout: [W1 2048][W2 2048][W3 1024][bias blocks 1024 + 1024 + 512][ones tile 4096][barriers 256][tmem ptr 128]
//                       then 8 groups x 2 buffers x 8 KB tiles (128-byte aligned)
constexpr uint32_t TT_OFF_W1 = 0, TT_OFF_W2 = 2048, TT_OFF_W3 = 4096, TT_OFF_BB = 5120, TT_OFF_ONES = 7680, TT_OFF_BAR = 11776,
                   TT_OFF_TMEM = 12032, TT_OFF_TILES = 12160;
constexpr uint32_t TT_GROUP_COLS = 64;                        // per group: 32 accumulator columns + 16 hidden (packed) columns

__global__ void __launch_bounds__(TT_THREADS, 1) tail_tc_kernel(const __grid_constant__ TailParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int grp = warp >> 2, quarter = warp & 3;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + TT_OFF_TMEM);
  for (int i = threadIdx.x; i < 5120 / 16; i += TT_THREADS)
    reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(p.w1)[i];      // w1 | w2 | w3 are contiguous in the blob
  // bias blocks, one per layer, in the weights' [k-chunk 2][n rows][8] form: K rows 0 and 1 = bias hi / lo, the rest 0
  for (int i = threadIdx.x; i < (1024 + 1024 + 512) / 16; i += TT_THREADS) {
    const int layer = i < 64 ? 0 : i < 128 ? 1 : 2;
    const int n = layer == 2 ? 16 : 32;
    const int cell = i - (layer == 0 ? 0 : layer == 1 ? 64 : 128);                    // [k-chunk][row]
    uint4 v = make_uint4(0, 0, 0, 0);
    if (cell < n) {
      const float bv = layer == 0 ? p.b1[cell] : layer == 1 ? p.b2[cell] : p.b3[cell];
      const float hi = act_to_float(act_from_float(bv));
      v.x = act_pack2(hi, bv - hi);
    }
    reinterpret_cast<uint4*>(smem + TT_OFF_BB)[i] = v;
  }
  // the constant A tile of the bias MMAs: K elements 0 and 1 of every row are 1, the other 14 are 0
  for (int i = threadIdx.x; i < 4096 / 16; i += TT_THREADS)
    reinterpret_cast<uint4*>(smem + TT_OFF_ONES)[i] = i < 128 ? make_uint4(act_pack2(1.f, 1.f), 0, 0, 0) : make_uint4(0, 0, 0, 0);
  auto BAR = [&](int g, int i) { return sbase + TT_OFF_BAR + 8u * (uint32_t)(g * 3 + i); };   // 0,1: TMA full[buf]; 2: MMA done
  if (threadIdx.x == 0) {
    for (int g = 0; g < TT_GROUPS; ++g) { mbar_init(BAR(g, 0), 1); mbar_init(BAR(g, 1), 1); mbar_init(BAR(g, 2), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // weights, bias blocks, ones tile: generic stores
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem + (uint32_t)grp * TT_GROUP_COLS;     // this group's accumulator columns; hidden tile at +32
  const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16);
  const uint32_t tile0 = sbase + TT_OFF_TILES + (uint32_t)grp * 2u * TT_TILE_BYTES;
  const bool issuer = (warp & 3) == 0 && lane == 0;
  const int S = p.ls.side;
  const size_t N3 = (size_t)S * S * S;
  const int64_t n_tiles = (p.n_pos + 127) / 128;
  const int64_t stride = (int64_t)gridDim.x * TT_GROUPS;
  int64_t tile = (int64_t)blockIdx.x * TT_GROUPS + grp;
  // descriptors: K-major no-swizzle, SBO = 128 B between 8-row groups, LBO = distance between the two 8-channel chunks
  constexpr uint32_t DHI = 8u | (1u << 14);
  auto DESC = [](uint32_t lo) { return ((uint64_t)DHI << 32) | (uint64_t)lo; };
  constexpr uint32_t idesc0 = (1u << 4) | kIdescAB | (8u << 24);
  constexpr uint32_t ID32 = idesc0 | ((32u >> 3) << 17), ID16 = idesc0 | ((16u >> 3) << 17);
  const uint32_t a_lbo = (2048u >> 4) << 16;
  const uint32_t ones = (((sbase + TT_OFF_ONES) >> 4) & 0x3FFFu) | a_lbo;
  auto load_tile = [&](int64_t t, int buf) {                       // issuer only
    const uint32_t dst = tile0 + (uint32_t)buf * TT_TILE_BYTES;
    mbar_expect_tx(BAR(grp, buf), TT_TILE_BYTES);
    const int64_t q0 = (int64_t)p.ls.guard + t * 128;
#pragma unroll
    for (int g = 0; g < 4; ++g)
      bulk_g2s(dst + (uint32_t)g * 2048u, p.src + ((int64_t)g * p.ls.plane_stride + q0) * 8, 2048u, BAR(grp, buf));
  };
  auto bdesc = [&](uint32_t off, uint32_t n) { return (((sbase + off) >> 4) & 0x3FFFu) | (n << 16); };   // LBO = n rows x 16 B
  // one layer (issuer only): bias, then the two K-steps of the data; A from the shared-memory tile (layer 1) or from
  // the group's packed hidden columns in tensor memory (layers 2, 3)
  auto gemm = [&](bool a_in_smem, uint32_t a_tile, uint32_t w_off, uint32_t bb_off, uint32_t n, uint32_t idesc) {
    tc_mma_bf16(tmem, DESC(ones), DESC(bdesc(bb_off, n)), idesc, 0u);
    const uint32_t b = bdesc(w_off, n);
    if (a_in_smem) {
      const uint32_t a = ((a_tile >> 4) & 0x3FFFu) | a_lbo;
      tc_mma_bf16(tmem, DESC(a), DESC(b), idesc, 1u);
      tc_mma_bf16(tmem, DESC(a + (4096u >> 4)), DESC(b + 2u * n), idesc, 1u);        // channels 16..31
    } else {
      tc_mma_bf16_ts(tmem, tmem + 32u, DESC(b), idesc, 1u);
      tc_mma_bf16_ts(tmem, tmem + 40u, DESC(b + 2u * n), idesc, 1u);
    }
    tc_commit(BAR(grp, 2));
  };
  const uint32_t bar_id = 1u + (uint32_t)grp;                        // named barrier of this group (0 = __syncthreads)
  auto group_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory"); };
  // ReLU + 16-bit rounding of this row's 32 accumulators -> the row's 16 packed columns of the hidden tile
  auto hidden_to_tmem = [&]() {
    uint32_t raw[2][16];
    tc_ld16(taddr, raw[0]);
    tc_ld16(taddr + 16u, raw[1]);
    tc_wait_ld();
    uint32_t pk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j)
      pk[j] = act_pack2_relu(__uint_as_float(raw[j >> 3][2 * (j & 7)]), __uint_as_float(raw[j >> 3][2 * (j & 7) + 1]));
    tc_st16(taddr + 32u, pk);
    tc_wait_st();
    tc_fence_before();
  };
  if (issuer && tile < n_tiles) load_tile(tile, 0);
  uint32_t ph_full[2] = {0u, 0u}, ph_mma = 0u;
  int buf = 0;
  for (; tile < n_tiles; tile += stride, buf ^= 1) {
    const uint32_t a_tile = tile0 + (uint32_t)buf * TT_TILE_BYTES;
    // ---- layer 1
    if (issuer) {
      mbar_wait(BAR(grp, buf), ph_full[buf]);
      tc_fence_after();
      gemm(true, a_tile, TT_OFF_W1, TT_OFF_BB, 32u, ID32);
      if (tile + stride < n_tiles) load_tile(tile + stride, buf ^ 1);   // that buffer's last reader (an MMA) was awaited
    }
    ph_full[buf] ^= 1u;
    mbar_wait(BAR(grp, 2), ph_mma); ph_mma ^= 1u;
    tc_fence_after();
    hidden_to_tmem();
    group_sync();
    // ---- layer 2
    if (issuer) { tc_fence_after(); gemm(false, 0u, TT_OFF_W2, TT_OFF_BB + 1024u, 32u, ID32); }
    mbar_wait(BAR(grp, 2), ph_mma); ph_mma ^= 1u;
    tc_fence_after();
    hidden_to_tmem();
    group_sync();
    // ---- layer 3 (16 columns, cout_real of them real)
    if (issuer) { tc_fence_after(); gemm(false, 0u, TT_OFF_W3, TT_OFF_BB + 2048u, 16u, ID16); }
    mbar_wait(BAR(grp, 2), ph_mma); ph_mma ^= 1u;
    tc_fence_after();
    uint32_t raw[16];
    tc_ld16(taddr, raw);
    tc_wait_ld();
    tc_fence_before();
    group_sync();                                                     // every quarter has read: the next layer 1 may overwrite
    const uint32_t pos = (uint32_t)((int64_t)p.ls.guard + tile * 128 + quarter * 32 + lane);
    const uint32_t b = fdiv(pos, p.fd_frame);
    const int rem = (int)(pos - b * (uint32_t)p.ls.frame_pitch) - p.ls.guard;
    if ((int)b < p.batch && rem >= 0) {
      const int x = (int)fdiv((uint32_t)rem, p.fd_px);
      const int r2 = rem - x * p.ls.pitch_x;
      const int y = (int)fdiv((uint32_t)r2, p.fd_py);
      const int z = r2 - y * p.ls.pitch_y;
      if (x < S && y < S && z < S) {
        float* o = p.dst + (size_t)b * p.cout_real * N3 + ((size_t)x * S + y) * S + z;
#pragma unroll
        for (int c = 0; c < 16; ++c)
          if (c < p.cout_real) __stcs(o + (size_t)c * N3, __uint_as_float(raw[c]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*s_tmem), "r"(512u) : "memory");
}

// CUDA-core checker (op.impl = 1): one thread per voxel, same packed weights, same bf16 roundings.
__global__ void __launch_bounds__(128) tail_mlp_simt_kernel(const __grid_constant__ TailParams p) {
  const int S = p.ls.side;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (n >= S * S * S) return;
  const int z = n % S, y = (n / S) % S, x = n / (S * S);
  const int64_t pos = vol_pos(p.ls, b, x, y, z);
  float h[32], t[32];
  for (int gq = 0; gq < 4; ++gq) {
    float a[8];
    unpack8(*reinterpret_cast<const uint4*>(p.src + ((int64_t)gq * p.ls.plane_stride + pos) * 8), a);
    for (int i = 0; i < 8; ++i) h[gq * 8 + i] = a[i];
  }
  for (int layer = 0; layer < 2; ++layer) {
    const __nv_bfloat16* w = layer ? p.w2 : p.w1;
    const float* bias = layer ? p.b2 : p.b1;
    for (int co = 0; co < 32; ++co) {
      float acc = 0.f;
      for (int ci = 0; ci < 32; ++ci) acc = fmaf(h[ci], act_to_float(w[((size_t)(ci >> 3) * 32 + co) * 8 + (ci & 7)]), acc);
      t[co] = act_to_float(act_from_float(fmaxf(acc + bias[co], 0.f)));
    }
    for (int co = 0; co < 32; ++co) h[co] = t[co];
  }
  const size_t N3 = (size_t)S * S * S;
  for (int co = 0; co < p.cout_real; ++co) {
    float acc = 0.f;
    for (int ci = 0; ci < 32; ++ci) acc = fmaf(h[ci], act_to_float(p.w3[((size_t)(ci >> 3) * 16 + co) * 8 + (ci & 7)]), acc);
    p.dst[((size_t)b * p.cout_real + co) * N3 + n] = acc + p.b3[co];
  }
}

// Called by sceneego_v2v_run for SCENEEGO_OP_TAIL_MLP.  Blob segment at op.w_offset:
// [w1 2048 B][w2 2048 B][w3 1024 B][b1 32 f32][b2 32 f32][b3 16 f32].
int launch_tail_mlp(const sceneego_v2v_op_t& op, void* const* d_buffers, const void* d_blob, int batch, int op_index,
                    bool simt, cudaStream_t st) {
  SE_REQUIRE(op.cin == 32 && op.cout == 16 && op.cout_real >= 1 && op.cout_real <= 16 && (op.flags & SCENEEGO_F_OUT_F32),
             "v2v_run: op %d: the fused tail is 32 -> 32 -> 32 -> (<=16) with f32 output", op_index);
  SE_REQUIRE(op.lay_src.s2d == 0, "v2v_run: op %d: the fused tail reads a plain layout", op_index);
  TailParams p;
  memset(&p, 0, sizeof(p));
  p.src = (const __nv_bfloat16*)d_buffers[op.src];
  p.dst = (float*)d_buffers[op.dst];
  SE_REQUIRE(p.src && p.dst, "v2v_run: op %d has a null buffer", op_index);
  const char* seg = (const char*)d_blob + op.w_offset;
  p.w1 = (const __nv_bfloat16*)seg;
  p.w2 = (const __nv_bfloat16*)(seg + 2048);
  p.w3 = (const __nv_bfloat16*)(seg + 4096);
  p.b1 = (const float*)(seg + 5120);
  p.b2 = p.b1 + 32;
  p.b3 = p.b2 + 32;
  p.ls = op.lay_src; p.batch = batch; p.cout_real = op.cout_real;
  p.n_pos = (int64_t)batch * p.ls.frame_pitch;
  SE_REQUIRE(p.n_pos + 4096 < (1ll << 31) && (p.n_pos + 4096) * (int64_t)p.ls.frame_pitch < (1ll << 48),
             "v2v_run: op %d: batch * frame_pitch too large for one launch", op_index);
  p.fd_frame = make_fastdiv((uint32_t)p.ls.frame_pitch);
  p.fd_px = make_fastdiv((uint32_t)p.ls.pitch_x);
  p.fd_py = make_fastdiv((uint32_t)p.ls.pitch_y);
  if (simt) {
    const int S = p.ls.side;
    dim3 grid((S * S * S + 127) / 128, batch);
    tail_mlp_simt_kernel<<<grid, 128, 0, st>>>(p);
    SE_CUDA_LAUNCH_CHECK("tail_mlp_simt");
    return SCENEEGO_OK;
  }
  if (op.impl == 2) {            // the register-resident mma.sync variant (kept for A/B measurements and tests)
    const int64_t n_chunks = (p.n_pos + 63) / 64;
    int64_t blocks = (n_chunks + 7) / 8;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    tail_mlp_kernel<<<(int)blocks, 256, 0, st>>>(p);
    SE_CUDA_LAUNCH_CHECK("tail_mlp");
    return SCENEEGO_OK;
  }
  SE_REQUIRE((const char*)p.w2 == (const char*)p.w1 + 2048 && (const char*)p.w3 == (const char*)p.w1 + 4096 &&
                 ((uintptr_t)p.w1 & 15) == 0, "v2v_run: op %d: tail weights must be one aligned 5 KB segment", op_index);
  const size_t smem_bytes = TT_OFF_TILES + (size_t)TT_GROUPS * 2 * TT_TILE_BYTES;
  if (int rc = ensure_max_dynamic_smem((const void*)tail_tc_kernel, (int)smem_bytes)) return rc;
  const int64_t n_tiles = (p.n_pos + 127) / 128;
  int64_t blocks = (n_tiles + TT_GROUPS - 1) / TT_GROUPS;
  if (blocks > kNumSMs) blocks = kNumSMs;
  tail_tc_kernel<<<(int)blocks, TT_THREADS, smem_bytes, st>>>(p);
  SE_CUDA_LAUNCH_CHECK("tail_tc");
  return SCENEEGO_OK;
}

}  // namespace sceneego
