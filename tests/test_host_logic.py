"""CPU: host-side logic -- state-dict contract, op-program construction, BN folding / weight
packing (host C code), sharding arithmetic, config loading."""
import ctypes as C

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

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
    # ten 3^3 convs with 32 output channels march along x (csrc/march.cu), the other 30 stay on conv_tc
    assert kinds.count(_lib.OP_CONV3_MARCH) == 10 and all(op.cout == 32 and op.ksize == 3 for op in pg.ops if op.type == _lib.OP_CONV3_MARCH)
    assert kinds.count(_lib.OP_CONV) == 30 and kinds.count(_lib.OP_TAIL_MLP) == 1 and kinds.count(_lib.OP_MAXPOOL2) == 5 and kinds.count(_lib.OP_DECONV2) == 5
    assert kinds[0] == _lib.OP_STEM7_MARCH
    assert pg.flops * 8 == m.flops_per_frame(64)
    assert pg.ops[0].ksize == 7 and pg.ops[0].cin == 33 and pg.ops[0].cout == 16
    assert pg.ops[0].lay_src.zwin == 1 and pg.ops[0].lay_src.s2d == 0 and pg.ops[0].lay_src.side == 32 and pg.ops[0].lay_src.pad == 3
    # 4 feature planes + the z-window occupancy plane, which is written whole from the program's plain f32 grid
    assert pg.extra_zero_planes == 0 and pg.buffers[pg.in_buf].shape[0] == 5
    m2 = V2VModel(33, 15)
    m2.stem = "s2d"                                                                    # the round-1 stem stays selectable
    pg2 = m2.program(32, 2, torch.device("cpu"))
    assert pg2.ops[0].type == _lib.OP_STEM7_S2D and pg2.ops[0].lay_src.s2d == 1 and pg2.ops[0].lay_src.side == 16 and pg2.ops[0].lay_src.pad == 2
    pg32 = V2VModel(32, 15).program(32, 1, torch.device("cpu"))          # no occupancy channel: plain x-stacked stem
    assert pg32.ops[0].type == _lib.OP_CONV and pg32.ops[0].cin == 32 and pg32.ops[0].lay_src.pad == 3
    last = pg.ops[-1]
    assert last.type == _lib.OP_TAIL_MLP and last.flags & _lib.F_OUT_F32 and last.cout_real == 15 and last.cout == 16
    # every op reads a buffer some earlier op (or the input staging) wrote, and never its own output
    written = {pg.in_buf}
    assert sum(1 for op in pg.ops if op.src2 >= 0) == 3           # 16->32, 32->64, 64->128 projection shortcuts fused
    for op in pg.ops:
        assert op.src in written and op.src != op.dst
        if op.src2 >= 0:
            assert op.src2 in written and op.src2 != op.dst and op.cin2 * 2 == op.cin
        if op.res >= 0:
            assert op.res in written and op.res != op.dst
        written.add(op.dst)
    # loading new weights drops the packed program
    m.load_state_dict(m.state_dict())
    assert m._programs == {}


def _unpack(w_packed, taps, cin_pad, cout_pad):
    a = w_packed.reshape(taps, cin_pad // 8, cout_pad, 8)
    f = util.decode_act(a)
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
                                          k, tr, cout_p, cin_p, 1, 1, p(wp), p(bp)) == 0
        scale = gamma.astype(np.float64) / np.sqrt(var.astype(np.float64) + 1e-5)
        wf = (w.transpose(1, 0, 2, 3, 4) if tr else w).reshape(cout, cin, taps) * scale[:, None, None]
        if tr:   # transposed: parities stacked along N in groups of npar = min(8, 256 / cout_pad)
            npar = min(8, 256 // cout_p)
            a = wp.reshape(taps // npar, cin_p // 8, npar, cout_p, 8)
            f = util.decode_act(a)
            got = f.transpose(3, 1, 4, 0, 2).reshape(cout_p, cin_p, taps)
        else:
            got = _unpack(wp, taps, cin_p, cout_p)
        ref = util.act_round(torch.from_numpy(wf.astype(np.float32))).numpy()
        assert np.array_equal(got[:cout, :cin], ref)                 # bf16 round-to-nearest-even, bit-exact
        assert np.all(got[cout:] == 0) and np.all(got[:, cin:] == 0)  # channel padding is zero
        assert np.allclose(bp[:cout], b * scale + (beta - mean * scale), rtol=1e-6, atol=1e-6)
        assert np.all(bp[cout:] == 0)
    # no-BN variant (output layer)
    w = rng.standard_normal((15, 32, 1, 1, 1)).astype(np.float32)
    wp, bp = np.zeros(32 * 16, np.uint16), np.zeros(16, np.float32)
    assert lib.sceneego_v2v_pack_conv(w.ctypes.data_as(C.c_void_p), None, None, None, None, None, C.c_double(0), 15, 32,
                                      1, 0, 16, 32, 1, 1, wp.ctypes.data_as(C.c_void_p), bp.ctypes.data_as(C.c_void_p)) == 0
    assert np.all(bp == 0)
    # CTA-pair blob (n_split = 2): x-stacked columns split in two half-major blobs, same values
    w = rng.standard_normal((32, 32, 3, 3, 3)).astype(np.float32)
    taps_s = 4 * 9                                               # xstack 2: (3 + 1) input plane offsets x 9
    one, two = np.zeros(taps_s * 32 * 64, np.uint16), np.zeros(taps_s * 32 * 64, np.uint16)
    bp = np.zeros(32, np.float32)
    for ns, out in ((1, one), (2, two)):
        assert lib.sceneego_v2v_pack_conv(w.ctypes.data_as(C.c_void_p), None, None, None, None, None, C.c_double(0), 32, 32,
                                          3, 0, 32, 32, 2, ns, out.ctypes.data_as(C.c_void_p), bp.ctypes.data_as(C.c_void_p)) == 0
    a = one.reshape(taps_s, 4, 64, 8)
    b2 = two.reshape(2, taps_s, 4, 32, 8)
    assert np.array_equal(a[:, :, :32], b2[0]) and np.array_equal(a[:, :, 32:], b2[1])


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


def test_stem_s2d_packing_reproduces_conv3d():
    """The 2x2x2-stacked, occupancy-along-K weight blob of the s2d stem (csrc/stem.cu), walked on the
    CPU in the kernel's streaming order over an s2d-arranged input, equals Conv3d(33,16,7,pad 3) + BN."""
    lib = _lib.load_library()
    torch.manual_seed(3)
    conv = nn.Conv3d(33, 16, 7, padding=3)
    bn = nn.BatchNorm3d(16).eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.1); bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 2)
    w_out = np.zeros(lib.sceneego_v2v_stem_s2d_weight_bytes() // 2, dtype=np.uint16)
    b_out = np.zeros(16, dtype=np.float32)
    fp = lambda a: a.detach().float().contiguous().numpy().ctypes.data_as(C.c_void_p)
    arrs = [conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]
    keep = [a.detach().float().contiguous().numpy() for a in arrs]
    ptrs = [k.ctypes.data_as(C.c_void_p) for k in keep]
    assert lib.sceneego_v2v_pack_stem_s2d(*ptrs, C.c_double(bn.eps), 1, w_out.ctypes.data_as(C.c_void_p),
                                          b_out.ctypes.data_as(C.c_void_p)) == 0
    wf = torch.from_numpy(util.decode_act(w_out).copy()).double()   # storage type -> f64
    V, S2, P = 8, 4, 2
    x = torch.randn(33, V, V, V).double()
    x[32] = (x[32] > 0.5).double()
    # s2d arrangement with pad 2: sub[par][c][X+P][Y+P][Z+P]
    sub = torch.zeros(8, 33, S2 + 2 * P + 1, S2 + 2 * P + 1, S2 + 2 * P + 1, dtype=torch.float64)
    for par in range(8):
        px, py, pz = par >> 2, (par >> 1) & 1, par & 1
        sub[par, :, P:P + S2, P:P + S2, P:P + S2] = x[:, px::2, py::2, pz::2]
    acc = torch.zeros(S2, S2, S2, 128, dtype=torch.float64)
    win = lambda par, c0, c1, bx, by, bz: sub[par, c0:c1, P + bx:P + bx + S2, P + by:P + by + S2, P + bz:P + bz + S2]
    chunk = 16384 // 2
    for s in range(32):
        ox, py, pz = (s >> 2) - 3, (s >> 1) & 1, s & 1
        par, bx = ((ox & 1) << 2) | (s & 3), ox >> 1
        for tp in range(16):
            by, bz = (tp >> 2) - 1 - py, (tp & 3) - 1 - pz
            B = wf[s * 8 * chunk + tp * 4096: s * 8 * chunk + (tp + 1) * 4096].reshape(4, 128, 8)     # [kchunk][n][8]
            A = win(par, 0, 32, bx, by, bz).reshape(4, 8, S2, S2, S2)
            acc += torch.einsum("gcxyz,gnc->xyzn", A, B)
    for s in range(5):
        bx = s - 2
        for tp in range(15):
            by, bz0 = tp // 3 - 2, (tp % 3) * 2 - 2
            base = (256 + s * 4) * chunk + tp * 2048
            B = wf[base: base + 2048].reshape(2, 128, 8)
            for c2 in range(2):
                A = torch.stack([win(par, 32, 33, bx, by, bz0 + c2)[0] for par in range(8)])           # (8 parities, X, Y, Z)
                acc += torch.einsum("exyz,ne->xyzn", A, B[c2])
    out = torch.zeros(16, V, V, V, dtype=torch.float64)
    for n0 in range(8):
        sx, sy, sz = n0 >> 2, (n0 >> 1) & 1, n0 & 1
        out[:, sx::2, sy::2, sz::2] = acc[..., n0 * 16:(n0 + 1) * 16].permute(3, 0, 1, 2)
    out += torch.from_numpy(b_out).double()[:, None, None, None]
    with torch.no_grad():
        ref = bn.double()(conv.double()(x[None]))[0]
    assert (out - ref).abs().max().item() <= 4e-3 * ref.abs().max().item()      # bf16 weights (2^-9 relative each)
    # and the blob's non-zero count is exactly the 343 x 33 x 16 taps, each stored once per stacked voxel it serves
    assert int((w_out != 0).sum()) <= 343 * 33 * 16 * 8


def test_stem_march_packing_walked_like_the_kernel_reproduces_conv3d():
    """sceneego_v2v_pack_stem_march (csrc/stem_march.cu): 8 rotations x 15 chunks; walking the blob the way the kernel
    does -- input plane x adds one N = 128 product into a ring of 8 accumulator slots with the rotation of its ring
    position, outputs -3..S+2 are drained in order and the out-of-range ones dropped, the occupancy channel read from
    z-window cells two dy rows per MMA -- equals Conv3d(33,16,7,pad 3) + BN on bf16 inputs."""
    lib = _lib.load_library()
    torch.manual_seed(4)
    conv = nn.Conv3d(33, 16, 7, padding=3)
    bn = nn.BatchNorm3d(16).eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.1); bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 2)
    w_out = np.zeros(lib.sceneego_v2v_stem_march_weight_bytes() // 2, dtype=np.uint16)
    b_out = np.zeros(16, dtype=np.float32)
    keep = [a.detach().float().contiguous().numpy() for a in
            (conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var)]
    ptrs = [k.ctypes.data_as(C.c_void_p) for k in keep]
    assert lib.sceneego_v2v_pack_stem_march(*ptrs, C.c_double(bn.eps), w_out.ctypes.data_as(C.c_void_p),
                                            b_out.ctypes.data_as(C.c_void_p)) == 0
    wf = torch.from_numpy(util.decode_act(w_out).copy()).double()
    ROT, FEAT, MMA = 417792 // 2, 28672 // 2, 4096 // 2                    # elements
    assert wf.numel() == 8 * ROT
    S, P = 6, 3
    x = torch.randn(33, S, S, S).bfloat16().double()
    x[32] = (x[32] > 0.3).double()
    xp = torch.zeros(33, S, S + 2 * P + 1, S + 2 * P + 1, dtype=torch.float64)       # y/z zero pads (+1 for the pair's 2nd row)
    xp[:, :, P:P + S, P:P + S] = x
    # z-window occupancy cells: zw[x][y'][z][e] = occ[x][y'][z-3+e]
    zw = torch.zeros(S, S + 2 * P + 1, S, 8, dtype=torch.float64)
    for e in range(8):
        for z in range(S):
            if 0 <= z - 3 + e < S:
                zw[:, :, z, e] = xp[32, :, :, P + z - 3 + e]
    ring = torch.zeros(8, S, S, 16, dtype=torch.float64)                  # slot, y, z, cout
    out = torch.zeros(16, S, S, S, dtype=torch.float64)
    G = 70 * 3                                                            # as if this were the CTA's fourth march of a 64^3 run
    occ_dy = lambda pr, c: (2 * pr if pr < 3 else 5) + c

    def drain(gi, o):
        if 0 <= o < S:
            out[:, o] = ring[gi & 7].permute(2, 0, 1)
        ring[gi & 7] = 0

    for xi in range(S):
        r = (G + xi) & 7
        rot = wf[r * ROT:(r + 1) * ROT]
        acc = torch.zeros(S, S, 128, dtype=torch.float64)
        for ks in range(2):
            for dy in range(7):
                for dz in range(7):
                    B = rot[(ks * 7 + dy) * FEAT + dz * MMA:(ks * 7 + dy) * FEAT + (dz + 1) * MMA].reshape(2, 128, 8)
                    A = xp[ks * 16:(ks + 1) * 16, xi, dy:dy + S, dz:dz + S].reshape(2, 8, S, S)
                    acc += torch.einsum("geyz,gne->yzn", A, B)
        for pr in range(4):
            B = rot[14 * FEAT + pr * MMA:14 * FEAT + (pr + 1) * MMA].reshape(2, 128, 8)
            A = torch.stack([zw[xi, occ_dy(pr, c):occ_dy(pr, c) + S] for c in range(2)])          # (2, y, z, 8)
            acc += torch.einsum("gyze,gne->yzn", A, B)
        idle = (r + 7) & 7
        assert acc[..., idle * 16:(idle + 1) * 16].abs().max().item() == 0.0                       # the block outside the band is zero
        assert ring[idle].abs().max().item() == 0.0                                                # ... and lands on a cleared slot
        ring += acc.reshape(S, S, 8, 16).permute(2, 0, 1, 3)
        drain(G + xi, xi - 3)                                                                      # output xi-3 is complete
    for k in range(1, 7):
        drain(G + S - 1 + k, S - 1 + k - 3)
    assert ring.abs().max().item() == 0.0
    out += torch.from_numpy(b_out).double()[:, None, None, None]
    with torch.no_grad():
        ref = bn.double()(conv.double()(x[None]))[0]
    assert (out - ref).abs().max().item() <= 4e-3 * ref.abs().max().item()      # bf16 weights (2^-9 relative each)
    # dense: every tap stored once per rotation, the eighth row block of every MMA is zero
    assert int((w_out != 0).sum()) <= 8 * 343 * 33 * 16


def test_march_packing_walked_like_the_kernel_reproduces_conv3d():
    """sceneego_v2v_pack_conv_march: [tap(dy,dz)][cin/8][3*cout][8] with row block j = W[dx = 2 - j].  Walking it the
    way conv_march_kernel does (input plane x feeds outputs x-1, x, x+1 through row blocks 0, 1, 2; a ring of
    accumulator slots; N = 2*cout sub-bands at the faces) over bf16 inputs equals Conv3d(k3, pad 1) + BN."""
    lib = _lib.load_library()
    torch.manual_seed(5)
    cin, cout, S = 16, 32, 5
    conv = nn.Conv3d(cin, cout, 3, padding=1)
    bn = nn.BatchNorm3d(cout).eval()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.1); bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 2)
    keep = [a.detach().float().contiguous().numpy() for a in
            (conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var)]
    w_out = np.zeros(27 * cin * cout, dtype=np.uint16)
    b_out = np.zeros(cout, dtype=np.float32)
    args = [a.ctypes.data_as(C.c_void_p) for a in keep]
    assert lib.sceneego_v2v_pack_conv_march(*args, C.c_double(bn.eps), cout, cin, cout, cin,
                                            w_out.ctypes.data_as(C.c_void_p), b_out.ctypes.data_as(C.c_void_p)) == 0
    wf = util.decode_act(w_out).reshape(9, cin // 8, 3 * cout, 8)     # [t][g][row][c8]
    wm = wf.transpose(0, 2, 1, 3).reshape(9, 3 * cout, cin).astype(np.float64)                  # [t][row][ci]
    x = torch.randn(1, cin, S, S, S).to(torch.bfloat16).float()
    xp = np.zeros((cin, S, S + 2, S + 2))
    xp[:, :, 1:-1, 1:-1] = x[0].numpy()
    NS = 4                                                   # a short ring so that it wraps inside the march
    ring = np.zeros((NS, cout, S, S))
    out = np.zeros((cout, S, S, S))
    for xin in range(S):
        j_lo, j_hi = (0 if xin > 0 else 1), (2 if xin + 1 < S else 1)
        if xin == 0:
            ring[0] = 0
        if xin + 1 < S:
            ring[(xin + 1) % NS] = 0                         # the slot this plane opens
        for t in range(9):
            dy, dz = t // 3, t % 3
            win = xp[:, xin, dy:dy + S, dz:dz + S]           # (ci, y, z)
            band = np.einsum("rc,cyz->ryz", wm[t, j_lo * cout:(j_hi + 1) * cout], win)
            for j in range(j_lo, j_hi + 1):
                ring[(xin - 1 + j) % NS] += band[(j - j_lo) * cout:(j - j_lo + 1) * cout]
        if xin > 0:
            out[:, xin - 1] = ring[(xin - 1) % NS]           # complete: committed to the epilogue
        if xin + 1 == S:
            out[:, xin] = ring[xin % NS]
    out += b_out.astype(np.float64)[:, None, None, None]
    with torch.no_grad():
        wq = conv.weight * (bn.weight / torch.sqrt(bn.running_var + bn.eps))[:, None, None, None, None]
        ref = F.conv3d(x.double(), util.act_round(wq).double(), padding=1)[0].numpy()
        ref += ((conv.bias - bn.running_mean) * bn.weight / torch.sqrt(bn.running_var + bn.eps) + bn.bias).double().numpy()[:, None, None, None]
    assert np.abs(out - ref).max() <= 1e-5


def test_v2v_simple_structure_matches_reference():
    """V2VModelSimple (network/v2v.py:224-257): the reference's state-dict names/shapes (recorded by make_golden from
    the reference class) and an op program that reads only what earlier ops wrote."""
    import json
    from sceneego_b200.network.v2v import V2VModelSimple
    m = V2VModelSimple(33, 15)
    ref_shapes = [(k, tuple(s)) for k, s in json.loads(str(util.golden("v2v_simple_v32.npz")["state_dict_shapes"]))]
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == ref_shapes
    pg = m.program(32, 1, torch.device("cpu"))
    kinds = [op.type for op in pg.ops]
    assert kinds.count(_lib.OP_MAXPOOL2) == 2 and kinds.count(_lib.OP_DECONV2) == 2
    assert kinds.count(_lib.OP_CONV3_MARCH) == 2                      # skip_res1: the only 32-channel Res block
    # stem + 6 Res blocks on conv_tc (two 3^3 convs each; the two projection shortcuts are fused) + back 1x1 + output
    assert kinds.count(_lib.OP_CONV) == 1 + 6 * 2 + 2
    assert pg.ops[0].ksize == 7 and pg.ops[0].cout == 32 and pg.ops[0].lay_src.s2d == 0 and pg.ops[0].lay_src.pad == 3
    assert pg.ops[-1].flags & _lib.F_OUT_F32 and pg.ops[-1].cout_real == 15
    assert pg.flops == m.flops_per_frame(32)
    written = {pg.in_buf}
    for op in pg.ops:
        assert op.src in written and op.src != op.dst
        if op.res >= 0:
            assert op.res in written
        written.add(op.dst)
