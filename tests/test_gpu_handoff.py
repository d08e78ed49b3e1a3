"""Backbone hand-off: pose_resnet's last deconvolution stage fused with process_features[0]."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import sceneego_oracle as orc
from sceneego_b200 import _lib
from sceneego_b200.utils import synth
from tests import util

pytestmark = pytest.mark.gpu


def _modules(seed):
    g = torch.Generator().manual_seed(seed)
    dc = nn.ConvTranspose2d(256, 256, 4, stride=2, padding=1, output_padding=0, bias=False)
    bn = nn.BatchNorm2d(256)
    pf = nn.Conv2d(256, 32, 1)
    with torch.no_grad():
        dc.weight.copy_(torch.randn(dc.weight.shape, generator=g) * (2.0 / (256 * 4)) ** 0.5)
        bn.weight.copy_(torch.rand(256, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(256, generator=g) * 0.1)
        bn.running_mean.copy_(torch.randn(256, generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(256, generator=g) * 1.5 + 0.5)
        pf.weight.copy_(torch.randn(pf.weight.shape, generator=g) * (1.0 / 256) ** 0.5)
        pf.bias.copy_(torch.randn(32, generator=g) * 0.1)
    return dc.cuda().eval(), bn.cuda().eval(), pf.cuda().eval()


@pytest.mark.parametrize("B,h,w", [(1, 32, 32), (3, 32, 32), (5, 16, 24), (64, 32, 32)])
def test_handoff_kernel_vs_torch_fp32(B, h, w):
    """deconv(k4,s2,p1) + BN + ReLU + 1x1 conv as one tcgen05 kernel vs the same ops in torch fp32 (TF32 off).
    bf16 operands with fp32 accumulation: 5e-3 relative Frobenius; against the same computation with the operands
    rounded like the kernel rounds them (bf16, or fp16 on the fp16-storage build) (input, folded weights, hidden activations, 1x1 weights) 1e-3 of the range."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dc, bn, pf = _modules(7)
    x = torch.relu(torch.randn(B, 256, h, w, generator=torch.Generator().manual_seed(B + h))).cuda()
    wts, bias = _lib.handoff_pack(dc.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, pf.weight, pf.bias, "cuda")
    got = _lib.backbone_handoff(x, wts, bias)
    assert got.shape == (B, 2 * h, 2 * w, 32)
    with torch.no_grad():
        ref = pf(F.relu(bn(dc(x)))).permute(0, 2, 3, 1).contiguous()
        rel = ((got - ref).norm() / ref.norm()).item()
        # the kernel's own roundings, emulated
        sc = (bn.weight.double() / torch.sqrt(bn.running_var.double() + bn.eps))
        w1 = util.act_round((dc.weight.double() * sc.view(1, -1, 1, 1)).float())
        sh = (bn.bias.double() - bn.running_mean.double() * sc).float()
        hid = util.act_round(F.relu(F.conv_transpose2d(util.act_round(x), w1, None, stride=2, padding=1) + sh.view(1, -1, 1, 1)))
        emu = (F.conv2d(hid, util.act_round(pf.weight), pf.bias)).permute(0, 2, 3, 1).contiguous()
    err = ((got - emu).abs().max() / (emu.max() - emu.min())).item()
    print(f"handoff B={B} {h}x{w}: rel-Frobenius vs fp32 {rel:.3e}, max-abs/range vs 16-bit emulation {err:.3e}")
    assert rel <= 5e-3 and err <= 1e-3
    again = _lib.backbone_handoff(x, wts, bias)
    assert torch.equal(again, got)


def test_forward_with_backbone_handoff_vs_reference():
    """The WHOLE forward from images (network/voxel_net_depth.py:224-275) against the unmodified reference's outputs
    (tests/golden/forward_v64.npz: seeded weights for all 699 state-dict entries, B = 2): the stock-backbone path and
    the hand-off path both within the 0.5 mm north-star gate; the hand-off's feat32 against the 1x1 conv of the torch
    features."""
    from sceneego_b200.network.voxel_net_depth import VoxelNetwork_depth
    g = util.golden("forward_v64.npz")
    tabs = orc.StageTables(util.CALIB, 64, 2.0)
    torch.manual_seed(0)
    nets = {h: VoxelNetwork_depth(util.load_config(batch_size=2), device="cuda", v2v_chunk=2, backbone_handoff=h).eval()
            for h in (False, True)}
    sd = synth.synthetic_state_dict(util.manifest(), seed=0, mode="random_bn")
    img = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(3)).cuda()
    depth = synth.synthetic_depth_room(2, tabs.ray, seed=4).cuda()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    out = {}
    for h, net in nets.items():
        net.load_state_dict(sd, strict=True)
        with torch.no_grad():
            out[h] = net(img, net.grid_coord_proj_batch, net.coord_volumes, depth_map_batch=depth)
        err_mm = orc.mpjpe(out[h][0].cpu().numpy(), g["kp"]) * 1000.0
        print(f"forward from images, backbone_handoff={h}: MPJPE vs reference {err_mm:.4f} mm, {net.last_launches} launches")
        assert err_mm <= 0.5
        assert out[h][1].shape == (2, 32, 1024, 1280) and out[h][2].shape == (2, 15, 64, 64, 64)
    net = nets[True]
    with torch.no_grad():
        _, feats = net.backbone(img)
        sub = feats[:, ::16, ::4, ::4].cpu().numpy()
        assert np.abs(sub - g["backbone_features"]).max() <= 2e-3 * float(g["backbone_features_absmax"])   # cuDNN vs CPU fp32
        conv = net.process_features[0]
        want = _lib.feature_conv1x1(feats.contiguous(), conv.weight, conv.bias)
        x2 = net.backbone.forward_before_last_deconv(img)
        w, bias = net._handoff_weights(x2.device)
        got = _lib.backbone_handoff(x2, w, bias)
    assert ((got - want).norm() / want.norm()).item() <= 5e-3
    # output #2 of both paths is the same map up to the hand-off's bf16 operands
    assert ((out[True][1] - out[False][1]).norm() / out[False][1].norm()).item() <= 5e-3
    # in-place edit of a source weight repacks
    with torch.no_grad():
        net.process_features[0].bias.add_(1.0)
        w2, bias2 = net._handoff_weights(x2.device)
        assert not torch.equal(bias2, bias)
        net.process_features[0].bias.sub_(1.0)
