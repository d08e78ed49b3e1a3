"""Shared helpers for the test-suite (checker side: may import oracle/)."""
import ctypes as C
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CALIB = os.path.join(ROOT, "sceneego_b200", "data", "fisheye.calibration_05_08.json")


LAST_PROGRAM = None


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def manifest():
    return [(k, tuple(s)) for k, s in json.load(open(os.path.join(GOLDEN, "state_dict_manifest.json")))]


def stage_shapes():
    return [(k, s) for k, s in manifest() if not k.startswith("backbone.")]


def load_config(batch_size=4, volume_size=64):
    from sceneego_b200.utils import cfg
    return cfg.default_config(batch_size, volume_size)


def unpack_bits(packed, V):
    return np.unpackbits(packed)[: V ** 3].reshape(V, V, V).astype(np.float32)


def run_single_op(x, conv, bn, relu, res=None, impl=0, pad_src=1, deconv=False, add_after=None, xstack=1, cta_pair=1,
                  shortcut=None, march=False):
    """Run one V2V op (conv or deconv) through sceneego_v2v_run; x (B,Cin,S,S,S) f32 cuda."""
    from sceneego_b200 import _lib
    from sceneego_b200.network.v2v import _Program, _pad16
    B, cin, S = x.shape[0], x.shape[1], x.shape[2]
    pg = _Program.__new__(_Program)
    pg.side, pg.chunk, pg.device = S, B, x.device
    pg.ops, pg.buffers, pg.buf_level, pg.free, pg.blob_parts, pg.blob_bytes = [], [], [], {}, [], 0
    pg.flops = 0
    pg.meta = []
    So = 2 * S if deconv else S
    lay_s = _lib.vol_layout(S, pad_src, B)
    lay_d = _lib.vol_layout(So, 1, B)
    cin_p, cout_p = _pad16(cin), _pad16(conv.out_channels)
    src = _lib.alloc_volume(lay_s, cin_p, x.device)
    dst = _lib.alloc_volume(lay_d, cout_p, x.device)
    _lib.pack_volume(x.contiguous(), src, lay_s)
    bufs = [src, dst]
    lays = [lay_s, lay_d]
    r_idx = -1
    extra = res if res is not None else add_after
    if extra is not None:
        rb = _lib.alloc_volume(lay_d, cout_p, x.device)
        _lib.pack_volume(extra.contiguous(), rb, lay_d)
        bufs.append(rb)
        lays.append(lay_d)
        r_idx = 2
    sc = None
    if shortcut is not None:      # (conv1x1, bn, x2): second source with half the channels, same layout as src
        sc_conv, sc_bn, x2 = shortcut
        sb = _lib.alloc_volume(lay_s, _pad16(x2.shape[1]), x.device)
        _lib.pack_volume(x2.contiguous(), sb, lay_s)
        bufs.append(sb)
        lays.append(lay_s)
        sc = (sc_conv, sc_bn, len(bufs) - 1)
    pg.buffers = bufs
    pg.buf_level = [0] * len(bufs)
    pg.lay_of = lambda i: lays[i]
    if deconv:
        pg.deconv(conv, bn, 0, 1, add=r_idx)
    else:
        pg.conv(conv, bn, 0, 1, relu=relu, res=r_idx, xstack=xstack, cta_pair=cta_pair, shortcut=sc, march=march)
    pg.ops[0].impl = impl
    pg.finalize()
    global LAST_PROGRAM
    LAST_PROGRAM = pg
    lib = _lib.load_library()
    rc = lib.sceneego_v2v_run(pg.op_array, 1, pg.buf_ptrs, C.c_void_p(pg.blob.data_ptr()), B, _lib._stream())
    _lib._check(rc, "v2v_run")
    torch.cuda.synchronize()
    return _lib.unpack_volume(dst, lay_d, B, conv.out_channels), dst, lay_d


def run_stem_s2d(x, conv, bn, impl=0, cta_pair=1, kind="s2d"):
    """Run the 7^3 stem op on x (B,33,V,V,V) f32 cuda; returns (B,16,V,V,V) f32.
    kind "s2d": SCENEEGO_OP_STEM7_S2D from the space-to-depth input; "march": SCENEEGO_OP_STEM7_MARCH from the
    z-window input."""
    from sceneego_b200 import _lib
    from sceneego_b200.network.v2v import _Program
    B, V = x.shape[0], x.shape[2]
    pg = _Program.__new__(_Program)
    pg.side, pg.chunk, pg.device = V, B, x.device
    pg.ops, pg.buffers, pg.buf_level, pg.free, pg.blob_parts, pg.blob_bytes = [], [], [], {}, [], 0
    pg.flops, pg.meta = 0, []
    lay_s = _lib.vol_layout_s2d(V, B) if kind == "s2d" else _lib.vol_layout_zwin(V, B)
    lay_d = _lib.vol_layout(V, 1, B)
    src = _lib.alloc_volume(lay_s, 33 * 8 if kind == "s2d" else 40, x.device)
    dst = _lib.alloc_volume(lay_d, 16, x.device)
    _lib.pack_volume(x.contiguous(), src, lay_s)
    lays = [lay_s, lay_d]
    pg.buffers = [src, dst]
    pg.buf_level = [0, 0]
    pg.lay_of = lambda i: lays[i]
    if kind == "s2d":
        pg.stem_s2d(conv, bn, 0, 1, cta_pair=cta_pair)
    else:
        pg.stem_march(conv, bn, 0, 1)
    pg.ops[0].impl = impl
    pg.finalize()
    global LAST_PROGRAM
    LAST_PROGRAM = pg
    lib = _lib.load_library()
    rc = lib.sceneego_v2v_run(pg.op_array, 1, pg.buf_ptrs, C.c_void_p(pg.blob.data_ptr()), B, _lib._stream())
    _lib._check(rc, "v2v_run")
    torch.cuda.synchronize()
    return _lib.unpack_volume(dst, lay_d, B, 16), dst, lay_d, src, lay_s


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def decode_act(a):
    """uint16 array of packed weights / activations -> float32, in the library's 16-bit storage type
    (bf16 by default, IEEE fp16 under SCENEEGO_ACT_DTYPE=f16)."""
    from sceneego_b200 import _lib
    a = np.ascontiguousarray(a)
    if _lib.act_dtype_name() == "f16":
        return a.view(np.float16).astype(np.float32)
    return (a.astype(np.uint32) << 16).view(np.float32)


def act_round(t):
    """Round a float tensor to the library's 16-bit storage type and back."""
    from sceneego_b200 import _lib
    return t.to(_lib.act_torch_dtype()).to(torch.float32)
