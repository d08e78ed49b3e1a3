"""Evaluation math on the device vs the reference's umeyama outputs and the oracle."""
import numpy as np
import pytest
import torch

from oracle import sceneego_oracle as orc
from tests import util

pytestmark = pytest.mark.gpu


def test_pose_errors_vs_reference_golden():
    """The kernel computes in fp64 from the float32 predictions; the reference keeps them float32 through the
    alignment (np.zeros_like(estimated_seq)), so the gate vs the reference's numbers is float32 resolution of
    metre-scale coordinates (2e-6 m, 250x below the 0.5 mm north-star tolerance); vs the fp64 evaluation 1e-11."""
    from sceneego_b200.utils import calculate_errors as ce
    g = util.golden("eval_poses.npz")
    pred, gt = torch.from_numpy(g["pred"]).cuda(), torch.from_numpy(g["gt"]).cuda()
    mp, pa = ce.evaluate_mpjpe(pred, gt)
    assert abs(mp - float(g["mpjpe"])) <= 1e-12
    assert abs(pa - float(g["pampjpe"])) <= 2e-6 and abs(pa - float(g["pampjpe_f64"])) <= 1e-11
    assert abs(ce.calculate_error(pred, gt) - float(g["mpjpe"])) <= 1e-12
    aligned, gt_out = ce.align_skeleton(pred, gt)
    assert np.abs(aligned.cpu().numpy() - g["aligned_f64"]).max() <= 1e-10
    assert np.abs(aligned.cpu().numpy() - g["aligned"]).max() <= 2e-6
    assert torch.equal(gt_out, gt)
    for b in (0, 3, 5):                                                  # generic, reflected (det < 0), planar gt
        c, R, t = ce.umeyama(pred[b], gt[b])
        T = g["transform"][b]
        assert abs(c - T[0]) <= 1e-6 and np.abs(R.cpu().numpy().reshape(-1) - T[1:10]).max() <= 1e-6
        assert np.abs(t.cpu().numpy() - T[10:]).max() <= 1e-6
        assert abs(np.linalg.det(R.cpu().numpy()) - 1.0) <= 1e-9


def test_pose_errors_scale_false_and_large_batch():
    from sceneego_b200.utils import calculate_errors as ce
    rng = np.random.default_rng(4)
    B = 1000
    gt = rng.normal(0, 0.5, (B, 15, 3))
    pred = (gt + rng.normal(0, 0.05, (B, 15, 3))).astype(np.float32)
    a_ref, g_ref = orc.align_skeleton(pred.astype(np.float64), gt, scale=False)
    a, g0 = ce.align_skeleton(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda(), scale=False)
    assert np.abs(a.cpu().numpy() - a_ref).max() <= 1e-10 and np.abs(g0.cpu().numpy() - g_ref).max() <= 1e-12
    mp, pa = ce.evaluate_mpjpe(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda())
    mp_ref, pa_ref = orc.evaluate_mpjpe(pred.astype(np.float64), gt)
    assert abs(mp - mp_ref) <= 1e-12 and abs(pa - pa_ref) <= 1e-11
    # identical poses: zero error, identity transform
    same = torch.from_numpy(gt.astype(np.float32)).cuda()
    mp0, pa0 = ce.evaluate_mpjpe(same, same.double())
    assert mp0 == 0.0 and pa0 <= 1e-12
    with pytest.raises(Exception):
        ce.calculate_error(torch.zeros(2, 15, 3), torch.zeros(2, 15, 3))   # CPU tensors: no fallback


def test_pose_errors_keep_float64_predictions():
    """The reference's NumPy code keeps the caller's dtype: float64 predictions (e.g. poses loaded from a pickle) are
    used as they are, not rounded to float32 first."""
    from sceneego_b200.utils import calculate_errors as ce
    rng = np.random.default_rng(8)
    gt = rng.normal(0, 0.5, (7, 15, 3))
    pred = gt + rng.normal(0, 0.05, (7, 15, 3))                          # float64, not representable in float32
    mp, pa = ce.evaluate_mpjpe(torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda())
    mp_ref, pa_ref = orc.evaluate_mpjpe(pred, gt)
    assert abs(mp - mp_ref) <= 1e-14 and abs(pa - pa_ref) <= 1e-12
    mp32, _ = ce.evaluate_mpjpe(torch.from_numpy(pred.astype(np.float32)).cuda(), torch.from_numpy(gt).cuda())
    assert abs(mp32 - mp_ref) > 1e-12                                     # the float32 rounding is visible at this gate
    c, R, t = ce.umeyama(torch.from_numpy(pred[0]).cuda(), torch.from_numpy(gt[0]).cuda())
    c_ref, R_ref, t_ref = orc.umeyama(pred[0], gt[0])
    assert abs(c - c_ref) <= 1e-10 and np.abs(R.cpu().numpy() - R_ref).max() <= 1e-10
