"""Evaluation math on the device (reference: utils/calculate_errors.py:22-28,60-91,
utils/rigid_transform_with_scale.py:18-43, dataset/test_dataset.py:102-112).

Same function names and argument meaning as the reference; poses are CUDA tensors (B,15,3) and everything runs
in ONE kernel launch per call (`sceneego_pose_errors_f64`, fp64 like the reference's NumPy) instead of a
per-frame Python loop around np.linalg.svd.  `skeleton_model` resizing (bone-length normalisation from a .mat
file, unused by test.py) is not part of this path.
"""
from __future__ import annotations

import torch

from .. import _lib


def _as_cuda(x, dtype):
    """dtype=None keeps the caller's floating dtype (float32 network output stays float32 and is promoted element by
    element inside the kernel, float64 stays float64 -- what NumPy does in the reference); integers become float64."""
    t = torch.as_tensor(x)
    if not t.is_cuda:
        raise _lib.SceneEgoError("evaluation ops need CUDA tensors (no CPU fallback)")
    if dtype is None:
        return t if t.dtype in (torch.float32, torch.float64) else t.to(torch.float64)
    return t.to(dtype)


def calculate_error(estimated_seq, gt_seq):
    """utils/calculate_errors.py:22-28 -- mean over frames and joints of the Euclidean distance."""
    est, gt = _as_cuda(estimated_seq, None), _as_cuda(gt_seq, torch.float64)
    mp, _, _, _, _ = _lib.pose_errors(est, gt)
    return float(mp.mean().item())


def align_skeleton(estimated_seq, gt_seq, skeleton_model=None, scale=True):
    """utils/calculate_errors.py:60-91 -- per-frame Umeyama alignment; returns (aligned (B,J,3) f64, gt (B,J,3) f64)."""
    if skeleton_model is not None:
        raise _lib.SceneEgoError("skeleton_model resizing is outside the device path (test.py passes None)")
    est, gt = _as_cuda(estimated_seq, None), _as_cuda(gt_seq, torch.float64)
    _, _, aligned, gt_out, _ = _lib.pose_errors(est, gt, scale=scale, want_aligned=True)
    return aligned, gt_out


def umeyama(P, Q):
    """utils/rigid_transform_with_scale.py:18-43 for one pose pair (n,3): returns c (float), R (3,3), t (3,)."""
    est, gt = _as_cuda(P, None)[None], _as_cuda(Q, torch.float64)[None]
    _, _, _, _, tr = _lib.pose_errors(est, gt, want_aligned=True)
    return float(tr[0, 0].item()), tr[0, 1:10].reshape(3, 3), tr[0, 10:13]


def evaluate_mpjpe(predicted_pose_list, gt_pose_list):
    """dataset/test_dataset.py:102-112 -- (mpjpe, pa-mpjpe) as test.py prints them."""
    est, gt = _as_cuda(predicted_pose_list, None), _as_cuda(gt_pose_list, torch.float64)
    mp, pa, _, _, _ = _lib.pose_errors(est, gt)
    return float(mp.mean().item()), float(pa.mean().item())
