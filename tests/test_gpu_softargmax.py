"""Soft-argmax (utils/op.py:83-96) on the GPU vs the oracle.  fp32: 1e-5 m on coordinates."""
import pytest
import torch

from oracle import sceneego_oracle as orc
from tests import util

pytestmark = pytest.mark.gpu


def _axis(coord):
    return torch.stack([coord[:, 0, 0, 0], coord[0, :, 0, 1], coord[0, 0, :, 2]]).contiguous()


@pytest.mark.parametrize("V,B,J,scale", [(64, 2, 15, 1.0), (64, 1, 15, 40.0), (32, 5, 3, 8.0), (128, 1, 2, 3.0)])
def test_softargmax_vs_oracle(V, B, J, scale):
    from sceneego_b200 import _lib
    coord = orc.build_coord_volume(V, 2.0)
    x = torch.randn(B, J, V, V, V, generator=torch.Generator().manual_seed(V + B)) * scale
    kp_ref, vol_ref = orc.soft_argmax(x, coord.unsqueeze(0).expand(B, -1, -1, -1, -1))
    # The fp32 CPU reference itself carries ~1e-4 relative accumulation error on the all-positive
    # z sums (2M terms at V=128), so it is checked at 3e-4 m and an fp64 evaluation at 5e-6 m.
    kp_f64, _ = orc.soft_argmax(x.double(), coord.double().unsqueeze(0).expand(B, -1, -1, -1, -1))
    for mode in ("axis", "coords"):
        kp, vol = _lib.softargmax3d(x.cuda(), 1.0, True, _axis(coord).cuda() if mode == "axis" else None,
                                    coord.reshape(-1, 3).cuda().contiguous() if mode == "coords" else None, True)
        assert (kp.cpu() - kp_ref).abs().max().item() <= 3e-4, mode
        assert (kp.cpu().double() - kp_f64).abs().max().item() <= 5e-6, mode
        assert torch.allclose(vol.cpu(), vol_ref, rtol=2e-4, atol=1e-9), mode
        assert torch.allclose(vol.sum(dim=(2, 3, 4)).cpu(), torch.ones(B, J), atol=1e-4)


def test_one_hot_known_answer_and_multiplier():
    """The reference's own smoke idea (voxel_net_depth.py:302-320): a sharply peaked volume
    returns that voxel's coordinate."""
    from sceneego_b200 import _lib
    V = 64
    coord = orc.build_coord_volume(V, 2.0)
    x = torch.zeros(1, 2, V, V, V)
    x[0, 0, 10, 20, 30] = 50.0
    x[0, 1, 63, 0, 5] = 50.0
    kp, _ = _lib.softargmax3d(x.cuda(), 2.0, True, _axis(coord).cuda(), None, False)
    assert torch.allclose(kp[0, 0].cpu(), coord[10, 20, 30], atol=1e-6)
    assert torch.allclose(kp[0, 1].cpu(), coord[63, 0, 5], atol=1e-6)
    kp_ref, _ = orc.soft_argmax(x.double() * 0.05, coord.double().unsqueeze(0))
    kp2, _ = _lib.softargmax3d(x.cuda(), 0.05, True, _axis(coord).cuda(), None, False)
    assert (kp2.cpu().double() - kp_ref).abs().max().item() <= 5e-6


def test_relu_mode_and_op_signature():
    from sceneego_b200.utils import op
    V, B, J = 32, 2, 4
    coord = orc.build_coord_volume(V, 2.0)
    cv = coord.unsqueeze(0).expand(8, -1, -1, -1, -1)
    x = torch.randn(B, J, V, V, V, generator=torch.Generator().manual_seed(1))
    for softmax in (True, False):
        kp_ref, vol_ref = orc.soft_argmax(x, cv, softmax=softmax)
        kp, vol = op.integrate_tensor_3d_with_coordinates(x.cuda(), cv.cuda(), softmax=softmax)
        assert vol.shape == x.shape
        assert torch.allclose(kp.cpu(), kp_ref, rtol=1e-4, atol=1e-2 if not softmax else 3e-4)
        assert torch.allclose(vol.cpu(), vol_ref, rtol=2e-4, atol=1e-9)
