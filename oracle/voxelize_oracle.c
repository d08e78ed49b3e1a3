/*
 * TEST INFRASTRUCTURE ONLY -- plain-C restatement of the integer/fp64 part of the hot path,
 * used by tests/ to cross-check the NumPy oracle and the CUDA kernels.  Never linked into
 * or called from the product (sceneego_b200/).
 *
 * Restates, for one frame:
 *   camera2world_ray            /root/reference/utils/fisheye/FishEyeCalibrated.py:36-51
 *   calculated_ray_direction    /root/reference/network/voxel_net_depth.py:147-155
 *   depth_map_to_voxel_numpy    /root/reference/network/voxel_net_depth.py:194-205
 *   point_cloud_to_voxel_numpy  /root/reference/network/voxel_net_depth.py:207-222
 *     (scatter with the torch-1.13.1 per-point (x,y,z) meaning)
 * Build: see oracle/Makefile (-ffp-contract=off: NumPy never fuses multiply-add).
 * Parity pin: compared against tests/golden/voxel.npz (reference-generated) in
 * tests/test_oracle_c.py.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* ray[(x*height + y)*3 + c], fp64 -- the reference's x-major order */
void oracle_ray_table(double cx, double cy, const double* c2w /*7, ascending*/, int width, int height,
                      double* ray) {
  for (int x = 0; x < width; ++x)
    for (int y = 0; y < height; ++y) {
      const double xc = (double)x - cx, yc = (double)y - cy;
      const double r2 = xc * xc + yc * yc;
      const double d = sqrt(r2);
      double z = 0.0;
      for (int i = 6; i >= 0; --i) z = z * d + c2w[i]; /* np.polyval, highest power first */
      const double nz = -z;
      const double norm = sqrt(r2 + nz * nz);
      double* o = ray + ((size_t)x * height + y) * 3;
      o[0] = xc / norm;
      o[1] = yc / norm;
      o[2] = nz / norm;
    }
}

/* depth (h,w) f32 -> occ (V,V,V) f32 {0,1}; returns the number of in-bounds points */
long oracle_voxelize(const float* depth, int h, int w, const double* ray, int img_h, int pad, int V,
                     double side, float* occ) {
  const int img_w = img_h + 2 * pad;
  long inb = 0;
  memset(occ, 0, sizeof(float) * (size_t)V * V * V);
  for (int x = 0; x < img_w; ++x)
    for (int y = 0; y < img_h; ++y) {
      float dv = 0.0f;
      const int xs = x - pad;
      if (xs >= 0 && xs < img_h) { /* cv2.resize INTER_NEAREST then np.pad */
        /* OpenCV resizeNN: cvFloor(dst * ifx), ifx = 1. / ((double)n_dst / n_src) */
        int sy = (int)floor((double)y * (1.0 / ((double)img_h / (double)h)));
        int sx = (int)floor((double)xs * (1.0 / ((double)img_h / (double)w)));
        if (sy > h - 1) sy = h - 1;
        if (sx > w - 1) sx = w - 1;
        dv = depth[(size_t)sy * w + sx];
      }
      const double* r = ray + ((size_t)x * img_h + y) * 3;
      const double d = (double)dv;
      const double qx = rint(((r[0] * d + side / 2) * V) / side);
      const double qy = rint(((r[1] * d + side / 2) * V) / side);
      const double qz = rint(((r[2] * d) * V) / side);
      if (qx >= 0 && qx <= V - 1 && qy >= 0 && qy <= V - 1 && qz >= 0 && qz <= V - 1) {
        occ[((size_t)qx * V + (size_t)qy) * V + (size_t)qz] = 1.0f;
        ++inb;
      }
    }
  return inb;
}
