"""Scaramuzza omnidirectional camera model, device-side.

Mirrors the two methods of the reference class the hot path uses
(utils/fisheye/FishEyeCalibrated.py:8-16, :36-51, :137-187); both run as CUDA
kernels through the C-ABI -- there is no NumPy/torch CPU implementation here.
"""
import json

import numpy as np
import torch

from ... import _lib


class FishEyeCameraCalibrated:
    def __init__(self, calibration_file_path, use_gpu=True):
        with open(calibration_file_path) as f:
            data = json.load(f)
        self.intrinsic = np.array(data["intrinsic"])
        self.img_size = np.array(data["size"])  # w, h
        self.fisheye_polynomial = np.array(data["polynomialC2W"])
        self.fisheye_inverse_polynomial = np.array(data["polynomialW2C"])
        self.img_center = np.array([self.intrinsic[0][2], self.intrinsic[1][2]])
        self.use_gpu = use_gpu

    def calib_struct(self, width=None, height=None) -> "_lib.Calib":
        w = int(self.img_size[0]) if width is None else int(width)
        h = int(self.img_size[1]) if height is None else int(height)
        return _lib.make_calib(self.img_center[0], self.img_center[1], self.fisheye_polynomial,
                               self.fisheye_inverse_polynomial, w, h)

    def ray_table_device(self, width, height, device="cuda") -> torch.Tensor:
        """(height, width, 3) fp64 unit rays on the device, row-major."""
        return _lib.ray_table(self.calib_struct(width, height), device)

    def camera2world_ray(self, point: np.ndarray) -> np.ndarray:
        """Rays for integer pixel coordinates (n,2) -> (n,3) fp64, like
        FishEyeCalibrated.py:36-51; evaluated on the device table."""
        w, h = int(self.img_size[0]), int(self.img_size[1])
        p = np.asarray(point)
        xs, ys = p[:, 0].astype(np.int64), p[:, 1].astype(np.int64)
        if not (np.array_equal(xs, p[:, 0]) and np.array_equal(ys, p[:, 1])):
            raise _lib.SceneEgoError("camera2world_ray: integer pixel coordinates expected")
        if xs.min() < 0 or ys.min() < 0 or xs.max() >= w or ys.max() >= h:
            raise _lib.SceneEgoError("camera2world_ray: pixel outside the calibrated image")
        table = self.ray_table_device(w, h).cpu().numpy()
        return table[ys, xs]

    def world2camera_pytorch(self, point3d_original: torch.Tensor, normalize=False) -> torch.Tensor:
        if normalize:
            raise _lib.SceneEgoError("world2camera_pytorch(normalize=True) is not on the accelerated path")
        pts = point3d_original.detach().to(dtype=torch.float32).contiguous()
        if not pts.is_cuda:
            pts = pts.cuda()
        return _lib.world2camera(self.calib_struct(), pts)
