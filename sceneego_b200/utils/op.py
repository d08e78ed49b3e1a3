"""Lifting / integration ops with the reference's signatures (utils/op.py).

Every function runs CUDA kernels through the C-ABI (sceneego_b200._lib); CPU
tensors are rejected, there is no PyTorch fallback.
"""
import torch

from .. import _lib
from . import multiview


def get_projected_2d_points_with_coord_volumes(fisheye_model, coord_volume):
    """utils/op.py:98-116 -- (V,V,V,3) voxel centres -> (N,2) pixel coordinates."""
    grid_coord = coord_volume.reshape((-1, 3))
    return multiview.project_3d_points_to_image_fisheye_camera(fisheye_model, grid_coord)


def get_grid_coord_proj_batch(grid_coord_proj, batch_size, heatmap_shape):
    """utils/op.py:177-184 -- normalise to [-1,1], shape (B,N,1,2), batch dim stride 0.
    Table preparation done once at construction.  The arithmetic runs on the host: torch's CUDA
    division by a scalar multiplies by the reciprocal, which differs by one ulp from the IEEE
    division the reference performs on the CPU."""
    p = grid_coord_proj.detach().cpu()
    g = torch.zeros_like(p)
    g[:, 0] = 2 * (p[:, 0] / heatmap_shape[1] - 0.5)
    g[:, 1] = 2 * (p[:, 1] / heatmap_shape[0] - 0.5)
    g = g.to(grid_coord_proj.device)
    return g.unsqueeze(1).unsqueeze(0).expand(batch_size, -1, -1, -1)


def _grid_and_stride(grid_batch, batch):
    """(B,N,1,2) grid -> (tensor whose storage holds the rows, float stride between frames)."""
    if grid_batch.dim() != 4 or grid_batch.shape[-1] != 2 or grid_batch.shape[-2] != 1:
        raise _lib.SceneEgoError("grid must have shape (B, N, 1, 2)")
    if grid_batch.shape[0] < batch:
        raise _lib.SceneEgoError("grid batch smaller than the heatmap batch")
    if grid_batch.stride(0) == 0 or grid_batch.shape[0] == 1:
        return grid_batch[0].contiguous(), 0
    g = grid_batch[:batch].contiguous()
    return g, g.stride(0)


def unproject_heatmaps_one_view_batch(heatmaps, grid_coord_proj_transformed_batch, volume_size):
    """utils/op.py:194-214 -- bilinear gather of (B,C,H,W) at the projected voxel centres,
    returned as (B,C,V,V,V)."""
    b, c = heatmaps.shape[0], heatmaps.shape[1]
    grid, stride = _grid_and_stride(grid_coord_proj_transformed_batch, b)
    out = _lib.grid_sample(heatmaps.contiguous().float(), grid, stride)
    return out.view(b, c, volume_size, volume_size, volume_size)


def unproject_heatmaps_one_view(heatmaps, grid_coord_proj, volume_size):
    """utils/op.py:135-175 (per-frame loop variant): same result as the batched call."""
    shape = tuple(heatmaps.shape[2:])
    grid = get_grid_coord_proj_batch(grid_coord_proj, 1, shape)
    return unproject_heatmaps_one_view_batch(heatmaps, grid, volume_size)


def integrate_tensor_3d_with_coordinates(volumes, coord_volumes, softmax=True):
    """utils/op.py:83-96 -- (B,J,V,V,V) logits -> ((B,J,3) expected coordinates, softmaxed volumes)."""
    b = volumes.shape[0]
    v = volumes.shape[2]
    if coord_volumes.dim() != 5 or coord_volumes.shape[0] < b:
        raise _lib.SceneEgoError("coord_volumes must have shape (>=B, V, V, V, 3)")
    if coord_volumes.stride(0) != 0 and coord_volumes.shape[0] > 1:
        if not bool((coord_volumes[:b] == coord_volumes[:1]).all()):
            raise _lib.SceneEgoError("per-frame coordinate volumes are not supported (the reference "
                                     "always expands one table, network/voxel_net_depth.py:81-83)")
    coords = coord_volumes[0].reshape(-1, 3).contiguous().float()
    kp, vol = _lib.softargmax3d(volumes.contiguous().float(), 1.0, softmax, None, coords, True)
    return kp, vol.view(volumes.shape)
