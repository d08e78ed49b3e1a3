// Shared helpers for the sceneego_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/sceneego_b200.h"

namespace sceneego {

// ---------------------------------------------------------------------------
// Storage type of V2V activations and weights.  Default: bf16 (BASELINE configs[1]: "bf16 V2V").  Compiled with
// -DSCENEEGO_ACT_F16 (libsceneego_b200_f16.so) every 16-bit cell holds IEEE fp16 instead: same tcgen05 rate
// (kind::f16 takes either), 3 more mantissa bits -- the storage-rounding error of the 50-layer chain drops 8x
// (profiles/r02_bf16_attribution.txt) -- at the price of fp16's range: stores saturate at +-65504 instead of
// overflowing to infinity.  The element type in signatures stays `__nv_bfloat16` (an opaque 16-bit cell); only the
// conversions below and the MMA instruction descriptor's A/B format bits differ.
// ---------------------------------------------------------------------------
#ifdef SCENEEGO_ACT_F16
constexpr int kActDtype = 1;                               // sceneego_act_dtype()
constexpr uint32_t kIdescAB = 0u;                          // instruction descriptor: A and B are f16
#define SE_MMA_SYNC_AB ".f16.f16"
__device__ __forceinline__ uint32_t act_pack2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t act_pack2_relu(float lo, float hi) {      // max(x, 0) and the rounding in one F2FP
  uint32_t r;
  asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 act_unpack2(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2*>(&u)); }
__device__ __forceinline__ __nv_bfloat16 act_from_float(float v) {
  const uint32_t r = act_pack2(v, 0.f);
  const uint16_t lo = (uint16_t)(r & 0xFFFFu);
  return *reinterpret_cast<const __nv_bfloat16*>(&lo);
}
__device__ __forceinline__ float act_to_float(__nv_bfloat16 v) { return __half2float(*reinterpret_cast<const __half*>(&v)); }
#else
constexpr int kActDtype = 0;
constexpr uint32_t kIdescAB = (1u << 7) | (1u << 10);      // instruction descriptor: A and B are bf16
#define SE_MMA_SYNC_AB ".bf16.bf16"
__device__ __forceinline__ uint32_t act_pack2(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t act_pack2_relu(float lo, float hi) {      // max(x, 0) and the rounding in one F2FP
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float2 act_unpack2(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u)); }
__device__ __forceinline__ __nv_bfloat16 act_from_float(float v) { return __float2bfloat16(v); }
__device__ __forceinline__ float act_to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
#endif

void set_error(const char* fmt, ...);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel): the attribute is per device, so a
// process that drives several GPUs needs it on each of them.  Returns SCENEEGO_OK or SCENEEGO_E_CUDA.
int ensure_max_dynamic_smem(const void* kernel, int bytes);

#define SE_REQUIRE(cond, ...)                      \
  do {                                             \
    if (!(cond)) {                                 \
      ::sceneego::set_error(__VA_ARGS__);          \
      return SCENEEGO_E_INVALID;                   \
    }                                              \
  } while (0)

#define SE_CUDA_LAUNCH_CHECK(what)                                              \
  do {                                                                          \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      ::sceneego::set_error("%s: %s", what, cudaGetErrorString(e__));           \
      return SCENEEGO_E_CUDA;                                                   \
    }                                                                           \
  } while (0)

// Device-side copy of the camera model (fp64 coefficients; fp32 copies are made where needed).
struct CalibDev {
  double cx, cy;
  double c2w[7];
  double w2c[11];
  int width, height;
};

inline CalibDev to_dev(const sceneego_calib_t* c) {
  CalibDev d;
  d.cx = c->cx; d.cy = c->cy;
  for (int i = 0; i < 7; ++i) d.c2w[i] = c->c2w[i];
  for (int i = 0; i < 11; ++i) d.w2c[i] = c->w2c[i];
  d.width = c->width; d.height = c->height;
  return d;
}

// Planar padded layout: position of voxel (x,y,z) of frame b inside one channel-group plane.
__host__ __device__ inline int64_t vol_pos(const sceneego_vol_layout_t& L, int b, int x, int y, int z) {
  return (int64_t)b * L.frame_pitch + L.guard + (int64_t)x * L.pitch_x + (int64_t)y * L.pitch_y + z;
}

// Voxel-centre coordinate exactly as the reference builds it in fp32
// (network/voxel_net_depth.py:127-130): fl(fl(step)*i) + lo, two roundings, no FMA.
__device__ __forceinline__ float voxel_coord(float lo, float step, int i) {
  return __fadd_rn(lo, __fmul_rn(step, (float)i));
}

// Cell (16-byte group of 8 bf16) of voxel (x,y,z), channel group g of a C8-group feature volume; handles
// the space-to-depth stem input (lay.s2d, include/sceneego_b200.h).
__host__ __device__ inline int64_t vol_cell(const sceneego_vol_layout_t& L, int b, int x, int y, int z, int g, int C8) {
  if (L.s2d) {
    const int par = ((x & 1) << 2) | ((y & 1) << 1) | (z & 1);
    return (int64_t)(par * C8 + g) * L.plane_stride + vol_pos(L, b, x >> 1, y >> 1, z >> 1);
  }
  return (int64_t)g * L.plane_stride + vol_pos(L, b, x, y, z);
}
// Element index (in bf16 units) of the occupancy channel that follows C8 feature groups.
__host__ __device__ inline int64_t vol_scene_elem(const sceneego_vol_layout_t& L, int b, int x, int y, int z, int C8) {
  if (L.s2d) {
    const int par = ((x & 1) << 2) | ((y & 1) << 1) | (z & 1);
    return ((int64_t)(8 * C8) * L.plane_stride + vol_pos(L, b, x >> 1, y >> 1, z >> 1)) * 8 + par;
  }
  return ((int64_t)C8 * L.plane_stride + vol_pos(L, b, x, y, z)) * 8;
}

constexpr int kNumSMs = 148;

int launch_stem_s2d(const sceneego_v2v_op_t& op, void* const* d_buffers, const void* d_blob, int batch, int op_index,
                    bool simt, cudaStream_t st);
int launch_feature_conv1x1_simt(const float* d_feat, const float* d_weight, const float* d_bias, float* d_out, int batch,
                                int cin, int cout, int h, int w, cudaStream_t st);
int launch_stem_march(const sceneego_v2v_op_t& op, void* const* d_buffers, const void* d_blob, int batch, int op_index,
                      bool simt, cudaStream_t st);
int launch_tail_mlp(const sceneego_v2v_op_t& op, void* const* d_buffers, const void* d_blob, int batch, int op_index,
                    bool simt, cudaStream_t st);
int launch_conv_march(const sceneego_v2v_op_t& op, void* const* d_buffers, const void* d_blob, int batch, int op_index,
                      cudaStream_t st);

}  // namespace sceneego
