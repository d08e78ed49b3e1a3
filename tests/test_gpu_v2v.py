"""V2V kernels vs torch fp32 (floating point -> tolerance stated per test)."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from tests import util

pytestmark = pytest.mark.gpu


def _mk_conv(cin, cout, k, seed, transposed=False):
    g = torch.Generator().manual_seed(seed)
    conv = (nn.ConvTranspose3d(cin, cout, 2, stride=2) if transposed
            else nn.Conv3d(cin, cout, k, padding=(k - 1) // 2))
    bn = nn.BatchNorm3d(cout)
    with torch.no_grad():
        conv.weight.copy_(torch.randn(conv.weight.shape, generator=g) * (2.0 / (cin * k ** 3)) ** 0.5)
        conv.bias.copy_(torch.randn(cout, generator=g) * 0.1)
        bn.weight.copy_(torch.rand(cout, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(cout, generator=g) * 0.1)
        bn.running_mean.copy_(torch.randn(cout, generator=g) * 0.1)
        bn.running_var.copy_(torch.rand(cout, generator=g) * 1.5 + 0.5)
    return conv.cuda().eval(), bn.cuda().eval()


def _ref(x, conv, bn, relu, res=None, add_after=None):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    with torch.no_grad():
        y = bn(conv(x))
        if res is not None:
            y = y + res
        if relu:
            y = F.relu(y)
        if add_after is not None:
            y = y + add_after
    return y


def _close(got, ref, what):
    # bf16 weights + bf16 output: 2^-8 relative per rounding; K up to 16k terms accumulate in fp32
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    assert err <= 2e-2 * scale + 1e-3, f"{what}: max abs err {err} vs scale {scale}"
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel <= 6e-3, f"{what}: relative Frobenius error {rel}"


CASES = [
    # (cin, cout, k, S, B, pad_src)
    (32, 32, 3, 16, 2, 1),
    (16, 32, 3, 16, 1, 1),
    (32, 64, 3, 8, 3, 1),
    (64, 64, 3, 16, 2, 1),
    (64, 128, 3, 8, 2, 1),
    (128, 128, 3, 8, 2, 1),
    (128, 128, 3, 2, 2, 1),
    (16, 32, 1, 16, 2, 1),
    (32, 32, 1, 16, 2, 1),
    (32, 15, 1, 16, 2, 1),
    (33, 16, 7, 16, 2, 3),
    (32, 32, 3, 64, 1, 1),
]


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("cin,cout,k,S,B,pad_src", CASES)
def test_single_conv(cin, cout, k, S, B, pad_src, impl):
    conv, bn = _mk_conv(cin, cout, k, seed=cin * 1000 + cout * 10 + k)
    g = torch.Generator().manual_seed(S + B)
    x = util.act_round(torch.randn(B, cin, S, S, S, generator=g)).cuda()
    res = util.act_round(torch.randn(B, cout, S, S, S, generator=g)).cuda()
    got, dst, lay = util.run_single_op(x, conv, bn, relu=True, res=res, impl=impl, pad_src=pad_src)
    _close(got, _ref(x, conv, bn, True, res=res), f"conv {cin}->{cout} k{k} S{S} impl{impl}")
    # pad cells of the destination must be exactly zero (they are the next layer's padding)
    full = dst.float().sum().item()
    assert np.isfinite(full)
    plane = dst[0].float()  # (plane_stride, 8)
    mask = torch.ones(lay.plane_stride, dtype=torch.bool, device=dst.device)
    idx = torch.arange(S, device=dst.device)
    for b in range(B):
        pos = (b * lay.frame_pitch + lay.guard + idx[:, None, None] * lay.pitch_x + idx[None, :, None] * lay.pitch_y
               + idx[None, None, :]).reshape(-1)
        mask[pos] = False
    assert plane[mask].abs().max().item() == 0.0, "pad / guard cells were written"


@pytest.mark.parametrize("cin,cout,k,S,B,pad_src,xs", [(33, 16, 7, 16, 2, 3, 4), (33, 16, 7, 64, 1, 3, 4),
                                                        (32, 32, 3, 16, 3, 1, 2), (32, 16, 3, 8, 2, 1, 4)])
def test_x_stacked_conv(cin, cout, k, S, B, pad_src, xs):
    """x-stacking: GEMM rows produce `xs` consecutive x-planes against Toeplitz-stacked weights."""
    conv, bn = _mk_conv(cin, cout, k, seed=99 + xs)
    g = torch.Generator().manual_seed(S * 7 + B)
    x = util.act_round(torch.randn(B, cin, S, S, S, generator=g)).cuda()
    res = util.act_round(torch.randn(B, cout, S, S, S, generator=g)).cuda()
    got, dst, lay = util.run_single_op(x, conv, bn, relu=True, res=res, impl=0, pad_src=pad_src, xstack=xs)
    _close(got, _ref(x, conv, bn, True, res=res), f"xstack{xs} conv {cin}->{cout} k{k} S{S}")
    plain, _, _ = util.run_single_op(x, conv, bn, relu=True, res=res, impl=0, pad_src=pad_src, xstack=1)
    assert ((got - plain).abs() <= 0.0079 * plain.abs() + 1e-4).all()      # same math, other summation order
    ref_simt, _, _ = util.run_single_op(x, conv, bn, relu=True, res=res, impl=1, pad_src=pad_src, xstack=xs)
    assert ((got - ref_simt).abs() <= 0.0079 * ref_simt.abs() + 1e-4).all()  # checker reads the stacked blob
    plane = dst[0].float()
    mask = torch.ones(lay.plane_stride, dtype=torch.bool, device=dst.device)
    idx = torch.arange(S, device=dst.device)
    for b in range(B):
        pos = (b * lay.frame_pitch + lay.guard + idx[:, None, None] * lay.pitch_x + idx[None, :, None] * lay.pitch_y
               + idx[None, None, :]).reshape(-1)
        mask[pos] = False
    assert plane[mask].abs().max().item() == 0.0, "pad / guard cells were written"


@pytest.mark.parametrize("cin,cout,k,S,B,xs", [(32, 32, 3, 16, 3, 2), (32, 32, 3, 64, 1, 2), (16, 32, 3, 16, 2, 2),
                                                (64, 64, 3, 16, 5, 1), (32, 64, 3, 8, 3, 1)])
def test_cta_pair_conv(cin, cout, k, S, B, xs):
    """tcgen05 cta_group::2: M = 256 over two CTAs of a cluster, weights N-split across the pair.
    Same math as the single-CTA kernel on the same inputs (odd item counts repeat the last item)."""
    conv, bn = _mk_conv(cin, cout, k, seed=17 + xs)
    g = torch.Generator().manual_seed(S * 3 + B)
    x = util.act_round(torch.randn(B, cin, S, S, S, generator=g)).cuda()
    res = util.act_round(torch.randn(B, cout, S, S, S, generator=g)).cuda()
    got, dst, lay = util.run_single_op(x, conv, bn, relu=True, res=res, impl=0, xstack=xs, cta_pair=2)
    _close(got, _ref(x, conv, bn, True, res=res), f"cta pair conv {cin}->{cout} k{k} S{S}")
    one, _, _ = util.run_single_op(x, conv, bn, relu=True, res=res, impl=0, xstack=xs, cta_pair=1)
    assert ((got - one).abs() <= 0.0079 * one.abs() + 1e-4).all()
    simt, _, _ = util.run_single_op(x, conv, bn, relu=True, res=res, impl=1, xstack=xs, cta_pair=2)
    assert ((got - simt).abs() <= 0.0079 * simt.abs() + 1e-4).all()        # checker reads the half-major blob
    plane = dst[0].float()
    mask = torch.ones(lay.plane_stride, dtype=torch.bool, device=dst.device)
    idx = torch.arange(S, device=dst.device)
    for b in range(B):
        pos = (b * lay.frame_pitch + lay.guard + idx[:, None, None] * lay.pitch_x + idx[None, :, None] * lay.pitch_y
               + idx[None, None, :]).reshape(-1)
        mask[pos] = False
    assert plane[mask].abs().max().item() == 0.0, "pad / guard cells were written"


@pytest.mark.parametrize("c,S,B,xs,pair", [(32, 16, 3, 2, 2), (32, 16, 2, 2, 1), (64, 16, 2, 1, 2), (128, 8, 2, 1, 1), (32, 64, 1, 2, 2)])
def test_fused_projection_shortcut(c, S, B, xs, pair):
    """Res3DBlock with a 1x1 projection shortcut (network/v2v.py:32-43): relu(bn(conv3(t)) + bn(conv1(x))) as ONE op --
    the shortcut's taps join the stencil's accumulation (second source with half the channels)."""
    conv, bn = _mk_conv(c, c, 3, seed=3 * c)
    sc_conv, sc_bn = _mk_conv(c // 2, c, 1, seed=5 * c)
    g = torch.Generator().manual_seed(S + B + c)
    t_in = util.act_round(torch.randn(B, c, S, S, S, generator=g)).cuda()
    x_in = util.act_round(torch.randn(B, c // 2, S, S, S, generator=g)).cuda()
    got, dst, lay = util.run_single_op(t_in, conv, bn, relu=True, impl=0, xstack=xs, cta_pair=pair,
                                       shortcut=(sc_conv, sc_bn, x_in))
    with torch.no_grad():
        ref = F.relu(bn(conv(t_in)) + sc_bn(sc_conv(x_in)))
    _close(got, ref, f"fused shortcut c{c} S{S} xs{xs} pair{pair}")
    simt, _, _ = util.run_single_op(t_in, conv, bn, relu=True, impl=1, xstack=xs, cta_pair=pair,
                                    shortcut=(sc_conv, sc_bn, x_in))
    assert ((got - simt).abs() <= 0.0079 * simt.abs() + 1e-4).all()
    plane = dst[0].float()
    mask = torch.ones(lay.plane_stride, dtype=torch.bool, device=dst.device)
    idx = torch.arange(S, device=dst.device)
    for b in range(B):
        pos = (b * lay.frame_pitch + lay.guard + idx[:, None, None] * lay.pitch_x + idx[None, :, None] * lay.pitch_y
               + idx[None, None, :]).reshape(-1)
        mask[pos] = False
    assert plane[mask].abs().max().item() == 0.0, "pad / guard cells were written"


def _pads_clean(dst, lay, S, B):
    plane = dst[0].float()
    mask = torch.ones(lay.plane_stride, dtype=torch.bool, device=dst.device)
    idx = torch.arange(S, device=dst.device)
    for b in range(B):
        pos = (b * lay.frame_pitch + lay.guard + idx[:, None, None] * lay.pitch_x + idx[None, :, None] * lay.pitch_y
               + idx[None, None, :]).reshape(-1)
        mask[pos] = False
    return plane[mask].abs().max().item() == 0.0


@pytest.mark.parametrize("cin,cout,S,B,res", [(32, 32, 16, 3, True), (32, 32, 16, 2, False), (16, 32, 16, 2, False),
                                              (32, 32, 32, 5, True), (32, 32, 64, 1, True), (16, 32, 64, 1, False),
                                              (32, 16, 8, 2, True), (32, 32, 2, 3, True), (32, 32, 40, 2, True),
                                              (32, 32, 32, 40, True), (16, 32, 32, 70, False)])
def test_marching_conv(cin, cout, S, B, res):
    """SCENEEGO_OP_CONV3_MARCH (csrc/march.cu): x-marching banded GEMM, resident weights, ring of accumulator
    slots (B > 1 and S = 40 make the ring wrap at every phase; S = 2 has only face planes; B = 40 / 70 at S = 32 are
    360 / 630 items on 296 CTAs: whole rounds of interleaved items plus a last round cut by planes)."""
    conv, bn = _mk_conv(cin, cout, 3, seed=31 + cin + S)
    g = torch.Generator().manual_seed(S * 5 + B)
    x = util.act_round(torch.randn(B, cin, S, S, S, generator=g)).cuda()
    r = util.act_round(torch.randn(B, cout, S, S, S, generator=g)).cuda() if res else None
    got, dst, lay = util.run_single_op(x, conv, bn, relu=True, res=r, impl=0, march=True)
    _close(got, _ref(x, conv, bn, True, res=r), f"marching conv {cin}->{cout} S{S}")
    simt, _, _ = util.run_single_op(x, conv, bn, relu=True, res=r, impl=1, march=True)
    assert ((got - simt).abs() <= 0.0079 * simt.abs() + 1e-4).all()      # checker reads the marching blob
    plain, _, _ = util.run_single_op(x, conv, bn, relu=True, res=r, impl=0)
    assert ((got - plain).abs() <= 0.0079 * plain.abs() + 1e-4).all()    # conv_tc: same math, other summation order
    assert _pads_clean(dst, lay, S, B), "pad / guard cells were written"


@pytest.mark.parametrize("S,B", [(16, 3), (64, 1)])
def test_marching_conv_fused_shortcut(S, B):
    """relu(bn(conv3(t)) + bn(conv1(x))) on the marching kernel: the shortcut is one N = Cout MMA per plane."""
    conv, bn = _mk_conv(32, 32, 3, seed=77)
    sc_conv, sc_bn = _mk_conv(16, 32, 1, seed=78)
    g = torch.Generator().manual_seed(S + B)
    t_in = util.act_round(torch.randn(B, 32, S, S, S, generator=g)).cuda()
    x_in = util.act_round(torch.randn(B, 16, S, S, S, generator=g)).cuda()
    got, dst, lay = util.run_single_op(t_in, conv, bn, relu=True, impl=0, march=True, shortcut=(sc_conv, sc_bn, x_in))
    with torch.no_grad():
        ref = F.relu(bn(conv(t_in)) + sc_bn(sc_conv(x_in)))
    _close(got, ref, f"marching fused shortcut S{S}")
    simt, _, _ = util.run_single_op(t_in, conv, bn, relu=True, impl=1, march=True, shortcut=(sc_conv, sc_bn, x_in))
    assert ((got - simt).abs() <= 0.0079 * simt.abs() + 1e-4).all()
    assert _pads_clean(dst, lay, S, B), "pad / guard cells were written"


def test_marching_conv_single_cta_per_sm_variant():
    """SCENEEGO_MARCH_CTAS=1: one CTA per SM with all 512 TMEM columns (ring of 16 slots, eight epilogue warps) --
    the configuration the kernel falls back to when resident weights + two stages exceed half an SM."""
    import os
    conv, bn = _mk_conv(32, 32, 3, seed=9)
    g = torch.Generator().manual_seed(4)
    x = util.act_round(torch.randn(3, 32, 24, 24, 24, generator=g)).cuda()
    r = util.act_round(torch.randn(3, 32, 24, 24, 24, generator=g)).cuda()
    two, _, _ = util.run_single_op(x, conv, bn, relu=True, res=r, impl=0, march=True)
    os.environ["SCENEEGO_MARCH_CTAS"] = "1"
    try:
        one, dst, lay = util.run_single_op(x, conv, bn, relu=True, res=r, impl=0, march=True)
    finally:
        del os.environ["SCENEEGO_MARCH_CTAS"]
    _close(one, _ref(x, conv, bn, True, res=r), "marching conv, one CTA per SM")
    assert torch.equal(one, two)            # same accumulation order: the ring size does not change the arithmetic
    assert _pads_clean(dst, lay, 24, 3)


def test_marching_conv_is_deterministic_and_reuses_buffers():
    """Two runs over the same buffers give bit-identical results (one issuer, fixed accumulation order)."""
    conv, bn = _mk_conv(32, 32, 3, seed=5)
    x = util.act_round(torch.randn(4, 32, 32, 32, 32, generator=torch.Generator().manual_seed(2))).cuda()
    a, _, _ = util.run_single_op(x, conv, bn, relu=True, impl=0, march=True)
    b, _, _ = util.run_single_op(x, conv, bn, relu=True, impl=0, march=True)
    assert torch.equal(a, b)


@pytest.mark.parametrize("V,B", [(16, 2), (32, 3), (64, 1), (32, 7), (64, 5)])
def test_stem_march(V, B):
    """7^3 stem as an x-marching banded GEMM (csrc/stem_march.cu): N = 128 over a ring of 8 accumulator slots with
    rotated weight rows, occupancy as a z-window plane.  vs torch fp32 Conv3d + BN + ReLU, vs the CUDA-core checker
    walking the same blob, pads untouched, deterministic; (32, 7) and (64, 5) give a CTA several items (ring positions
    carried from one march to the next) and the last group of a frame with tiles beyond the plane."""
    from sceneego_b200 import _lib
    conv, bn = _mk_conv(33, 16, 7, seed=5)
    g = torch.Generator().manual_seed(V + B)
    x = util.act_round(torch.randn(B, 33, V, V, V, generator=g))
    x[:, 32] = (x[:, 32] > 0.8).float()                       # occupancy channel is {0,1}
    x = x.cuda()
    got, dst, lay, src, lay_s = util.run_stem_s2d(x, conv, bn, impl=0, kind="march")
    assert lay_s.zwin == 1 and lay_s.pad == 3
    assert torch.equal(_lib.unpack_volume(src, lay_s, B, 33), x)          # z-window pack/unpack round trip
    # the z-window plane really holds occ[z-3 .. z+4] in every cell
    cell = src[4].view(-1, 8)[lay_s.guard + 5 * lay_s.pitch_x + 6 * lay_s.pitch_y + 7].float().cpu()
    want = torch.stack([x[0, 32, 5, 6, 7 - 3 + e] if 0 <= 7 - 3 + e < V else torch.zeros(()) .cuda() for e in range(8)]).cpu()
    assert torch.equal(cell, want)
    _close(got, _ref(x, conv, bn, True), f"stem march V{V}")
    simt, _, _, _, _ = util.run_stem_s2d(x, conv, bn, impl=1, kind="march")
    assert ((got - simt).abs() <= 0.0079 * simt.abs() + 1e-4).all()      # same blob, other summation order
    again, _, _, _, _ = util.run_stem_s2d(x, conv, bn, impl=0, kind="march")
    assert torch.equal(again, got)
    plane = dst[0].float()
    mask = torch.ones(lay.plane_stride, dtype=torch.bool, device=dst.device)
    idx = torch.arange(V, device=dst.device)
    for b in range(B):
        pos = (b * lay.frame_pitch + lay.guard + idx[:, None, None] * lay.pitch_x + idx[None, :, None] * lay.pitch_y
               + idx[None, None, :]).reshape(-1)
        mask[pos] = False
    assert plane[mask].abs().max().item() == 0.0, "pad / guard cells were written"


@pytest.mark.parametrize("pair", [1, 2])
@pytest.mark.parametrize("V,B", [(16, 2), (32, 3), (64, 1)])
def test_stem_s2d(V, B, pair):
    """7^3 stem from the space-to-depth input: 2x2x2 output stacking, occupancy packed along K."""
    from sceneego_b200 import _lib
    conv, bn = _mk_conv(33, 16, 7, seed=5)
    g = torch.Generator().manual_seed(V + B)
    x = util.act_round(torch.randn(B, 33, V, V, V, generator=g))
    x[:, 32] = (x[:, 32] > 0.8).float()                       # occupancy channel is {0,1}
    x = x.cuda()
    got, dst, lay, src, lay_s = util.run_stem_s2d(x, conv, bn, impl=0, cta_pair=pair)
    assert torch.equal(_lib.unpack_volume(src, lay_s, B, 33), x)          # s2d pack/unpack round trip (bf16-exact input)
    _close(got, _ref(x, conv, bn, True), f"stem s2d V{V}")
    simt, _, _, _, _ = util.run_stem_s2d(x, conv, bn, impl=1, cta_pair=pair)
    assert ((got - simt).abs() <= 0.0079 * simt.abs() + 1e-4).all()      # same blob, other summation order
    plane = dst[0].float()
    mask = torch.ones(lay.plane_stride, dtype=torch.bool, device=dst.device)
    idx = torch.arange(V, device=dst.device)
    for b in range(B):
        pos = (b * lay.frame_pitch + lay.guard + idx[:, None, None] * lay.pitch_x + idx[None, :, None] * lay.pitch_y
               + idx[None, None, :]).reshape(-1)
        mask[pos] = False
    assert plane[mask].abs().max().item() == 0.0, "pad / guard cells were written"


@pytest.mark.parametrize("cin,cout,S,B", [(32, 32, 16, 2), (128, 128, 4, 2), (16, 32, 32, 1)])
def test_tc_matches_simt_bitwise_close(cin, cout, S, B):
    """Same packed weights, same bf16 inputs: tensor-core and CUDA-core paths may differ only
    by fp32 accumulation order, i.e. at most one bf16 ulp after the output rounding."""
    conv, bn = _mk_conv(cin, cout, 3, seed=7)
    x = util.act_round(torch.randn(B, cin, S, S, S, generator=torch.Generator().manual_seed(1))).cuda()
    a, _, _ = util.run_single_op(x, conv, bn, relu=False, impl=0)
    b, _, _ = util.run_single_op(x, conv, bn, relu=False, impl=1)
    assert ((a - b).abs() <= 0.0079 * b.abs() + 1e-4).all()


def test_maxpool_and_deconv():
    from sceneego_b200 import _lib
    import ctypes as C
    S, B, c = 16, 2, 64
    x = util.act_round(torch.randn(B, c, S, S, S, generator=torch.Generator().manual_seed(3))).cuda()
    lay_s, lay_d = _lib.vol_layout(S, 1, B), _lib.vol_layout(S // 2, 1, B)
    src, dst = _lib.alloc_volume(lay_s, c, x.device), _lib.alloc_volume(lay_d, c, x.device)
    _lib.pack_volume(x, src, lay_s)
    op = _lib.V2VOp()
    op.type, op.cin, op.cout, op.cout_real, op.src, op.dst, op.res = _lib.OP_MAXPOOL2, c, c, c, 0, 1, -1
    op.lay_src, op.lay_dst = lay_s, lay_d
    ops = (_lib.V2VOp * 1)(op)
    bufs = (C.c_void_p * 2)(src.data_ptr(), dst.data_ptr())
    dummy = torch.zeros(16, device=x.device)
    _lib._check(_lib.load_library().sceneego_v2v_run(ops, 1, bufs, C.c_void_p(dummy.data_ptr()), B, _lib._stream()),
                "pool")
    got = _lib.unpack_volume(dst, lay_d, B, c)
    assert torch.equal(got, F.max_pool3d(x, 2, 2))          # exact: max of bf16 values

    for cin, cout, s in ((64, 32, 8), (128, 128, 2), (128, 64, 4)):
        conv, bn = _mk_conv(cin, cout, 2, seed=11, transposed=True)
        x = util.act_round(torch.randn(B, cin, s, s, s, generator=torch.Generator().manual_seed(4))).cuda()
        skip = util.act_round(torch.randn(B, cout, 2 * s, 2 * s, 2 * s, generator=torch.Generator().manual_seed(5))).cuda()
        got, _, _ = util.run_single_op(x, conv, bn, relu=True, deconv=True, add_after=skip)
        _close(got, _ref(x, conv, bn, True, add_after=skip), f"deconv {cin}->{cout}")


@pytest.mark.parametrize("mode", ["default", "random_bn"])
def test_v2v_v32_vs_reference_golden(mode):
    """Whole V2V at V=32 against outputs of the UNMODIFIED reference (tests/golden/v2v_v32.npz).
    Tolerance: bf16 activations through 52 layers -> 3% of the logit range, 1.5% Frobenius."""
    from sceneego_b200.network.v2v import V2VModel
    from sceneego_b200.utils import synth
    m = V2VModel(33, 15).eval()
    sd = synth.synthetic_state_dict([(k, tuple(v.shape)) for k, v in m.state_dict().items()], seed=1, mode=mode)
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 33, 32, 32, 32, generator=g).abs()
    x[:, 32] = (x[:, 32] > 1.0).float()
    with torch.no_grad():
        out = m(x.cuda())
    ref = torch.from_numpy(util.golden("v2v_v32.npz")[mode])
    got = out.reshape(15, -1)[:, ::13].cpu()
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() <= 3e-2 * scale
    assert ((got - ref).norm() / ref.norm()).item() <= 1.5e-2


def test_v2v_simt_and_tc_whole_network_agree():
    from sceneego_b200.network.v2v import V2VModel
    torch.manual_seed(0)
    m = V2VModel(33, 15).eval().cuda()
    x = torch.randn(2, 33, 32, 32, 32, device="cuda").abs()
    pg = m.program(32, 2, x.device)
    from sceneego_b200 import _lib
    _lib.pack_volume(x, pg.buffers[pg.in_buf], pg.lay_in)
    a = torch.empty(2, 15, 32, 32, 32, device="cuda")
    b = torch.empty_like(a)
    m.run_chunk(pg, 2, a, impl=0)
    m.run_chunk(pg, 2, b, impl=1)
    m.run_chunk(pg, 2, a, impl=0)   # again: buffers reused, pads must still be clean
    torch.cuda.synchronize()
    assert ((a - b).norm() / b.norm()).item() <= 2e-2   # accumulation order x 52 bf16-rounded layers


def test_fused_tail_matches_unfused_chain():
    """SCENEEGO_OP_TAIL_MLP (three 1x1 convs in registers, csrc/tail.cu) vs the same layers run one by one:
    identical bf16 roundings, only the fp32 summation order differs."""
    from sceneego_b200.network.v2v import V2VModel
    from sceneego_b200.utils import synth
    m = V2VModel(33, 15).eval()
    sd = synth.synthetic_state_dict([(k, tuple(v.shape)) for k, v in m.state_dict().items()], seed=4, mode="random_bn")
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    x = torch.randn(2, 33, 32, 32, 32, generator=torch.Generator().manual_seed(9)).abs().cuda()
    with torch.no_grad():
        fused = m(x)
        m.fuse_tail = False
        m.invalidate()
        plain = m(x)
    assert [op.type for op in m.program(32, 2, x.device).ops].count(4) == 0
    scale = plain.abs().max().item()
    assert (fused - plain).abs().max().item() <= 8e-3 * scale      # one bf16 ulp of the hidden activations, propagated
    assert ((fused - plain).norm() / plain.norm()).item() <= 2e-3


@pytest.mark.parametrize("mode", ["default", "random_bn"])
def test_v2v_simple_vs_reference_golden(mode):
    """V2VModelSimple at V=32 against outputs of the UNMODIFIED reference class (tests/golden/v2v_simple_v32.npz).
    Tolerance as for V2VModel: bf16 activations -> 3 % of the logit range, 1.5 % Frobenius."""
    import json
    from sceneego_b200.network.v2v import V2VModelSimple
    from sceneego_b200.utils import synth
    g = util.golden("v2v_simple_v32.npz")
    shapes = [(k, tuple(s)) for k, s in json.loads(str(g["state_dict_shapes"]))]
    m = V2VModelSimple(33, 15).eval()
    m.load_state_dict(synth.synthetic_state_dict(shapes, seed=2, mode=mode), strict=True)
    m = m.cuda()
    x = torch.randn(1, 33, 32, 32, 32, generator=torch.Generator().manual_seed(6)).abs()
    x[:, 32] = (x[:, 32] > 1.0).float()
    with torch.no_grad():
        out = m(x.cuda())
    ref = torch.from_numpy(g[mode])
    got = out.reshape(15, -1)[:, ::13].cpu()
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() <= 3e-2 * scale
    assert ((got - ref).norm() / ref.norm()).item() <= 1.5e-2


def test_tail_tcgen05_matches_mma_sync_variant():
    """The fused 1x1 tail on tcgen05 (default, csrc/tail.cu: tail_tc_kernel) vs the register-resident mma.sync
    variant (op.impl = 2) and the CUDA-core checker: same packed weights, same bf16 roundings of the hidden
    activations, only the fp32 summation order differs."""
    from sceneego_b200.network.v2v import V2VModel
    from sceneego_b200.utils import synth
    from sceneego_b200 import _lib
    m = V2VModel(33, 15).eval()
    sd = synth.synthetic_state_dict([(k, tuple(v.shape)) for k, v in m.state_dict().items()], seed=8, mode="random_bn")
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    B = 3
    x = torch.randn(B, 33, 32, 32, 32, generator=torch.Generator().manual_seed(2)).abs().cuda()
    pg = m.program(32, B, x.device)
    _lib.pack_volume(x, pg.buffers[pg.in_buf], pg.lay_in)
    outs = []
    for impl in (0, 2, 0):
        o = torch.full((B, 15, 32, 32, 32), float("nan"), device="cuda")
        m.run_chunk(pg, B, o, impl=impl)
        torch.cuda.synchronize()
        assert torch.isfinite(o).all()
        outs.append(o)
    assert torch.equal(outs[0], outs[2])                                   # deterministic
    scale = outs[1].abs().max().item()
    assert (outs[0] - outs[1]).abs().max().item() <= 8e-3 * scale          # one bf16 ulp of a hidden activation, propagated
    assert ((outs[0] - outs[1]).norm() / outs[1].norm()).item() <= 2e-3
