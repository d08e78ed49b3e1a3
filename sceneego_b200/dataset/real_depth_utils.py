"""Dataset-side depth handling on the GPU (reference: dataset/real_depth_utils.py:29-60,
dataset/demo_dataset.py:86-96, dataset/test_dataset.py:138-148).

The reference's datasets decode an EXR, nearest-resize it to 1280x1024 when needed, clamp depth > 10 m to 10 m and
-- with `voxel_output=True` -- turn it into a (V,V,V) occupancy grid on the host.  Here the resize, the clamp and
the voxelisation are ONE kernel launch over a batch of raw maps (`sceneego_voxelize_depth_raw_f64`); EXR decoding
stays host I/O.  Same function names and argument meaning as the reference; tensors live on the CUDA device.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib

PRE_H, PRE_W, CLAMP_MAX = 1024, 1280, 10.0     # demo_dataset.py:87-91

_ray_cache = {}


def _device_ray(ray, device, image_width=PRE_W, image_height=PRE_H) -> torch.Tensor:
    """The reference's ray table is a NumPy (W*H, 3) array in x-major order (network/voxel_net_depth.py:147-155);
    the kernel reads a row-major (H, W, 3) fp64 device table.  Converted once per table."""
    if isinstance(ray, torch.Tensor) and ray.is_cuda and ray.dim() == 3:
        return ray
    key = (id(ray), str(device))
    hit = _ray_cache.get(key)
    if hit is None:
        arr = np.asarray(ray, dtype=np.float64).reshape(image_width, image_height, 3)
        hit = torch.from_numpy(np.ascontiguousarray(arr.transpose(1, 0, 2))).to(device)
        _ray_cache[key] = hit
    return hit


def depth_maps_to_voxels(ray, depth_raw: torch.Tensor, cuboid_side: float, volume_size: int,
                         preprocess: bool = True) -> torch.Tensor:
    """Batch form: depth_raw (B,h,w) f32 CUDA as decoded -> (B,V,V,V) f32 {0,1}.  preprocess=False skips the
    dataset's resize/clamp (the maps are already 1024x1280 and clamped)."""
    if not depth_raw.is_cuda:
        raise _lib.SceneEgoError("depth_maps_to_voxels needs CUDA tensors (no CPU fallback)")
    d = depth_raw.contiguous().float()
    b = d.shape[0]
    occ = torch.zeros(b, volume_size, volume_size, volume_size, dtype=torch.float32, device=d.device)
    r = _device_ray(ray, d.device)
    if preprocess:
        _lib.voxelize_depth_raw(d, (PRE_H, PRE_W), CLAMP_MAX, r, PRE_H, PRE_W, volume_size, float(cuboid_side), occ,
                                None, None)
    else:
        _lib.voxelize_depth(d, r, PRE_H, PRE_W, volume_size, float(cuboid_side), occ, None, None)
    return occ


def depth_map_to_voxel(ray, depth, cuboid_side, volume_size):
    """dataset/real_depth_utils.py:29-43, one PREPROCESSED (1024,1280) map (what the reference's datasets pass)."""
    d = torch.as_tensor(depth)
    if not d.is_cuda:
        d = d.cuda()
    return depth_maps_to_voxels(ray, d.reshape(1, d.shape[-2], d.shape[-1]), cuboid_side, volume_size,
                                preprocess=False)[0]
