"""CPU: host-side logic -- state-dict contract, op-program construction, BN folding / weight
packing (host C code), sharding arithmetic, config loading."""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from sceneego_b200 import _lib
from sceneego_b200.network import pose_resnet
from sceneego_b200.network.v2v import V2VModel
from sceneego_b200.parallel import shard_range
from sceneego_b200.utils import synth
from tests import util


def test_state_dict_names_and_shapes_match_reference():
    man = util.manifest()
    v2v = [(k, tuple(v.shape)) for k, v in V2VModel(33, 15).state_dict().items()]
    assert v2v == [(k[len("volume_net."):], s) for k, s in man if k.startswith("volume_net.")]
    bb = [(k, tuple(v.shape)) for k, v in pose_resnet.get_pose_net(None).state_dict().items()]
    assert bb == [(k[len("backbone."):], s) for k, s in man if k.startswith("backbone.")]
    rest = [k for k, _ in man if not k.startswith(("volume_net.", "backbone."))]
    assert rest == ["process_features.0.weight", "process_features.0.bias"]
    assert len(man) == 699


def test_program_structure_and_flops():
    m = V2VModel(33, 15)
    assert abs(m.flops_per_frame(64) / 1e9 - 299.11) < 0.01
    assert abs(V2VModel(32, 15).flops_per_frame(64) / 1e9 - 296.2) < 0.1
    pg = m.program(32, 2, torch.device("cpu"))
    kinds = [op.type for op in pg.ops]
    assert kinds.count(_lib.OP_CONV) == 47 and kinds.count(_lib.OP_MAXPOOL2) == 5 and kinds.count(_lib.OP_DECONV2) == 5
    assert pg.flops * 8 == m.flops_per_frame(64)
    assert pg.ops[0].ksize == 7 and pg.ops[0].cin == 48 and pg.ops[0].cout == 16 and pg.ops[0].lay_src.pad == 3
    last = pg.ops[-1]
    assert last.flags & _lib.F_OUT_F32 and last.cout_real == 15 and last.cout == 16
    # every op reads a buffer some earlier op (or the input staging) wrote, and never its own output
    written = {pg.in_buf}
    for op in pg.ops:
        assert op.src in written and op.src != op.dst
        if op.res >= 0:
            assert op.res in written and op.res != op.dst
        written.add(op.dst)
    # loading new weights drops the packed program
    m.load_state_dict(m.state_dict())
    assert m._programs == {}


def _unpack(w_packed, taps, cin_pad, cout_pad):
    a = w_packed.reshape(taps, cin_pad // 8, cout_pad, 8)
    f = (a.astype(np.uint32) << 16).view(np.float32)
    return f.transpose(2, 1, 3, 0).reshape(cout_pad, cin_pad, taps)     # (co, ci, tap)


def test_pack_conv_folds_batchnorm():
    rng = np.random.default_rng(0)
    lib = _lib.load_library()
    for (cout, cin, k, tr) in ((16, 33, 7, 0), (32, 16, 3, 0), (15, 32, 1, 0), (32, 64, 2, 1)):
        taps = k ** 3
        shape = (cin, cout, k, k, k) if tr else (cout, cin, k, k, k)
        w = rng.standard_normal(shape).astype(np.float32)
        b = rng.standard_normal(cout).astype(np.float32)
        gamma, beta = rng.random(cout).astype(np.float32) + 0.5, rng.standard_normal(cout).astype(np.float32)
        mean, var = rng.standard_normal(cout).astype(np.float32), rng.random(cout).astype(np.float32) + 0.5
        cin_p, cout_p = (cin + 15) // 16 * 16, (cout + 15) // 16 * 16
        wp = np.zeros(taps * cin_p * cout_p, np.uint16)
        bp = np.zeros(cout_p, np.float32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        assert lib.sceneego_v2v_pack_conv(p(w), p(b), p(gamma), p(beta), p(mean), p(var), C.c_double(1e-5), cout, cin,
                                          k, tr, cout_p, cin_p, 1, p(wp), p(bp)) == 0
        scale = gamma.astype(np.float64) / np.sqrt(var.astype(np.float64) + 1e-5)
        wf = (w.transpose(1, 0, 2, 3, 4) if tr else w).reshape(cout, cin, taps) * scale[:, None, None]
        got = _unpack(wp, taps, cin_p, cout_p)
        ref = torch.from_numpy(wf.astype(np.float32)).to(torch.bfloat16).float().numpy()
        assert np.array_equal(got[:cout, :cin], ref)                 # bf16 round-to-nearest-even, bit-exact
        assert np.all(got[cout:] == 0) and np.all(got[:, cin:] == 0)  # channel padding is zero
        assert np.allclose(bp[:cout], b * scale + (beta - mean * scale), rtol=1e-6, atol=1e-6)
        assert np.all(bp[cout:] == 0)
    # no-BN variant (output layer)
    w = rng.standard_normal((15, 32, 1, 1, 1)).astype(np.float32)
    wp, bp = np.zeros(32 * 16, np.uint16), np.zeros(16, np.float32)
    assert lib.sceneego_v2v_pack_conv(w.ctypes.data_as(C.c_void_p), None, None, None, None, None, C.c_double(0), 15, 32,
                                      1, 0, 16, 32, 1, wp.ctypes.data_as(C.c_void_p), bp.ctypes.data_as(C.c_void_p)) == 0
    assert np.all(bp == 0)


def test_shard_range_partitions_exactly():
    for total in (1, 7, 8, 64, 1000, 1024):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_config_and_synth():
    c = util.load_config(batch_size=3)
    assert c.model.volume_size == 64 and c.model.cuboid_side == 2 and c.opt.batch_size == 3
    assert c.model.backbone.num_joints == 15 and list(c.heatmap_shape) == [1024, 1280]
    a = synth.synthetic_state_dict(util.stage_shapes(), seed=0)
    b = synth.synthetic_state_dict(list(reversed(util.stage_shapes())), seed=0)
    assert all(torch.equal(a[k], b[k]) for k in a)                    # order independent
    assert torch.equal(synth.synthetic_features(1), synth.synthetic_features(1))
