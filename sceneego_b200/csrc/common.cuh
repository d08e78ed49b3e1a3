// Shared helpers for the sceneego_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/sceneego_b200.h"

namespace sceneego {

void set_error(const char* fmt, ...);

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

constexpr int kNumSMs = 148;

}  // namespace sceneego
