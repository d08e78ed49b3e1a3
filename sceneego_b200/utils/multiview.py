"""Dispatch helper with the reference's name (utils/multiview.py:114-132)."""
import torch


def project_3d_points_to_image_fisheye_camera(fisheye_camera_model, points_3d):
    if torch.is_tensor(points_3d):
        return fisheye_camera_model.world2camera_pytorch(points_3d)
    raise TypeError("Works only with PyTorch tensors (CUDA path).")
