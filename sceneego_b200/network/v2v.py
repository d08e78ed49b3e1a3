"""V2V encoder-decoder (reference: network/v2v.py:8-181) for the B200 engine.

The nn.Modules below exist to own parameters and BatchNorm buffers under exactly
the reference's state-dict names ("front_layers.0.block.0.weight", ...), so
checkpoints load strictly.  Their forward is NOT torch: `V2VModel` compiles its
own structure into a flat op program (conv / max-pool / deconv steps over planar
padded bf16 buffers) that `sceneego_v2v_run` executes with hand-written sm_100a
kernels.  BatchNorm (eval) is folded into the packed bf16 weights when the
program is built; folded weights are derived data and never enter the state dict.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from .. import _lib


def _conv_bn(cin, cout, k, relu):
    layers = [nn.Conv3d(cin, cout, kernel_size=k, stride=1, padding=(k - 1) // 2), nn.BatchNorm3d(cout)]
    if relu:
        layers.append(nn.ReLU(True))
    return layers


class Basic3DBlock(nn.Module):
    """conv(k, same) - BN - ReLU (v2v.py:8-18); children: block.0, block.1."""

    def __init__(self, in_planes, out_planes, kernel_size):
        super().__init__()
        self.block = nn.Sequential(*_conv_bn(in_planes, out_planes, kernel_size, True))


class Res3DBlock(nn.Module):
    """v2v.py:21-43; children: res_branch.{0,1,3,4}, skip_con.{0,1} when channels change."""

    def __init__(self, in_planes, out_planes):
        super().__init__()
        self.res_branch = nn.Sequential(*(_conv_bn(in_planes, out_planes, 3, True)
                                          + _conv_bn(out_planes, out_planes, 3, False)))
        self.skip_con = nn.Sequential() if in_planes == out_planes else nn.Sequential(
            *_conv_bn(in_planes, out_planes, 1, False))


class Pool3DBlock(nn.Module):
    def __init__(self, pool_size):
        super().__init__()
        assert pool_size == 2
        self.pool_size = pool_size


class Upsample3DBlock(nn.Module):
    """ConvTranspose3d(k2,s2) - BN - ReLU (v2v.py:55-67)."""

    def __init__(self, in_planes, out_planes, kernel_size, stride):
        super().__init__()
        assert kernel_size == 2 and stride == 2
        self.block = nn.Sequential(
            nn.ConvTranspose3d(in_planes, out_planes, kernel_size=2, stride=2, padding=0, output_padding=0),
            nn.BatchNorm3d(out_planes), nn.ReLU(True))


class EncoderDecorder(nn.Module):  # (sic) -- the reference's spelling is part of no key, kept for familiarity
    """v2v.py:70-139: five pool/res levels, mid res, five res/upsample levels with skips."""

    CHANNELS = [32, 64, 128, 128, 128, 128]   # channels at level 0..5

    def __init__(self):
        super().__init__()
        ch = self.CHANNELS
        for lvl in range(1, 6):
            setattr(self, f"encoder_pool{lvl}", Pool3DBlock(2))
            setattr(self, f"encoder_res{lvl}", Res3DBlock(ch[lvl - 1], ch[lvl]))
        self.mid_res = Res3DBlock(128, 128)
        for lvl in range(5, 0, -1):
            setattr(self, f"decoder_res{lvl}", Res3DBlock(ch[lvl], ch[lvl]))
            setattr(self, f"decoder_upsample{lvl}", Upsample3DBlock(ch[lvl], ch[lvl - 1], 2, 2))
        for lvl in range(1, 6):
            setattr(self, f"skip_res{lvl}", Res3DBlock(ch[lvl - 1], ch[lvl - 1]))


def _pad16(c: int) -> int:
    return (c + 15) // 16 * 16


class _Program:
    """Op list + buffer pool + packed weight blob for one (volume_size, chunk) pair."""

    def __init__(self, side: int, chunk: int, in_channels: int, device, allow_s2d: bool = True, stem: str = "march"):
        self.side, self.chunk, self.device = side, chunk, device
        self.ops: List[_lib.V2VOp] = []
        self.buffers: List[torch.Tensor] = []
        self.buf_level: List[int] = []
        self.free: Dict[int, List[int]] = {}
        self.blob_parts: List[np.ndarray] = []
        self.blob_bytes = 0
        self.in_channels = in_channels
        self.in_pad = _pad16(in_channels)
        # stem input.  33 channels (32 lifted features + occupancy): space-to-depth storage read by the
        # 2x2x2-stacked stem kernel (csrc/stem.cu); anything else: plain layout, 7^3 stencil needs 3 zero cells
        # "march" (default): plain layout with pad 3 whose fifth plane is the z-window occupancy plane, read by the
        # x-marching stem (csrc/stem_march.cu); "s2d": the round-1 2x2x2-stacked stem (csrc/stem.cu)
        special = in_channels == 33 and allow_s2d
        self.s2d = special and stem == "s2d"
        self.zwin = special and stem == "march"
        self.lay_in = (_lib.vol_layout_s2d(side, chunk) if self.s2d else _lib.vol_layout_zwin(side, chunk) if self.zwin
                       else _lib.vol_layout(side, 3, chunk))
        # zero planes the unprojection kernel clears behind the 32 feature channels (the occupancy plane(s)); none for
        # the z-window layout: its occupancy plane is written whole, from `occ_scratch` (a plain f32 grid the
        # voxelisation scatters into with one store per pixel; _lib.occ_expand_zwin leaves it all-zero again) or by
        # pack_volume (the scene_volumes= path)
        self.extra_zero_planes = 0 if self.zwin else (self.in_pad - 32) // 8
        self.occ_scratch = (torch.zeros(chunk, side, side, side, dtype=torch.float32, device=device)
                            if self.zwin and torch.device(device).type == "cuda" else None)
        self.lays = [_lib.vol_layout(side >> l, 1, chunk) for l in range(6)]
        self.level_channels = EncoderDecorder.CHANNELS
        self.in_buf = self._new_buffer(-1)
        self.logits_buf: Optional[int] = None
        self.flops = 0
        self.meta: List[dict] = []      # per op: kind, real channels, kernel size, side, algorithmic flops/frame

    # -- buffers -------------------------------------------------------------
    def _new_buffer(self, level: int) -> int:
        if level < 0:
            # s2d: 8 parity sub-volumes x 4 channel groups + 1 plane of occupancy blocks
            t = _lib.alloc_volume(self.lay_in, 33 * 8 if self.s2d else 40 if self.zwin else self.in_pad, self.device)
        else:
            t = _lib.alloc_volume(self.lays[level], self.level_channels[level], self.device)
        self.buffers.append(t)
        self.buf_level.append(level)
        return len(self.buffers) - 1

    def acquire(self, level: int) -> int:
        pool = self.free.setdefault(level, [])
        return pool.pop() if pool else self._new_buffer(level)

    def release(self, idx: int) -> None:
        self.free.setdefault(self.buf_level[idx], []).append(idx)

    def lay_of(self, idx: int):
        lvl = self.buf_level[idx]
        return self.lay_in if lvl < 0 else self.lays[lvl]

    # -- weights -------------------------------------------------------------
    def _append_blob(self, arr: np.ndarray) -> int:
        off = self.blob_bytes
        raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
        padn = (-raw.size) % 256
        if padn:
            raw = np.concatenate([raw, np.zeros(padn, np.uint8)])
        self.blob_parts.append(raw)
        self.blob_bytes += raw.size
        return off

    def pack(self, conv: nn.Module, bn: Optional[nn.BatchNorm3d], cin_pad: int, cout_pad: int,
             xstack: int = 1, n_split: int = 1) -> Tuple[int, int]:
        w_out, b_out = self.pack_arrays(conv, bn, cin_pad, cout_pad, xstack, n_split)
        return self._append_blob(w_out), self._append_blob(b_out)

    def pack_arrays(self, conv: nn.Module, bn: Optional[nn.BatchNorm3d], cin_pad: int, cout_pad: int,
                    xstack: int = 1, n_split: int = 1, march=False) -> Tuple[np.ndarray, np.ndarray]:
        transposed = isinstance(conv, nn.ConvTranspose3d)
        w = conv.weight.detach().float().cpu().contiguous().numpy()
        k = w.shape[-1]
        cin, cout = (w.shape[0], w.shape[1]) if transposed else (w.shape[1], w.shape[0])
        bias = conv.bias.detach().float().cpu().contiguous().numpy() if conv.bias is not None else None
        taps = (k + xstack - 1) * k * k
        w_out = np.zeros(taps * cin_pad * cout_pad * xstack, dtype=np.uint16)
        b_out = np.zeros(cout_pad, dtype=np.float32)

        def fp(a):
            return a.ctypes.data_as(C.c_void_p) if a is not None else None
        if bn is not None:
            g = bn.weight.detach().float().cpu().contiguous().numpy()
            bt = bn.bias.detach().float().cpu().contiguous().numpy()
            mu = bn.running_mean.detach().float().cpu().contiguous().numpy()
            var = bn.running_var.detach().float().cpu().contiguous().numpy()
            eps = float(bn.eps)
        else:
            g = bt = mu = var = None
            eps = 0.0
        if march:
            # marching layout (csrc/march.cu): [tap(dy,dz)][cin/8][3*cout][8], no Toeplitz zeros
            assert k == 3 and not transposed and xstack == 1 and n_split == 1
            rc = _lib.load_library().sceneego_v2v_pack_conv_march(fp(w), fp(bias), fp(g), fp(bt), fp(mu), fp(var),
                                                                  C.c_double(eps), cout, cin, cout_pad, cin_pad,
                                                                  fp(w_out), fp(b_out))
            _lib._check(rc, "v2v_pack_conv_march")
            return w_out, b_out
        rc = _lib.load_library().sceneego_v2v_pack_conv(fp(w), fp(bias), fp(g), fp(bt), fp(mu), fp(var),
                                                        C.c_double(eps), cout, cin, k, int(transposed), cout_pad,
                                                        cin_pad, int(xstack), int(n_split), fp(w_out), fp(b_out))
        _lib._check(rc, "v2v_pack_conv")
        return w_out, b_out

    # -- ops -----------------------------------------------------------------
    def conv(self, conv: nn.Conv3d, bn, src: int, dst: int, relu: bool, res: int = -1, out_f32: bool = False,
             xstack: int = 1, cta_pair: int = 1, shortcut=None, march=False):
        """shortcut = (conv1x1, bn, src2): the projection shortcut of a Res3DBlock (v2v.py:32-43) accumulated into
        the same GEMM tile instead of being written out and re-read as a residual.
        march: run as SCENEEGO_OP_CONV3_MARCH (x-marching banded GEMM, csrc/march.cu); xstack / cta_pair unused."""
        k = conv.kernel_size[0]
        cin_pad, cout_pad = _pad16(conv.in_channels), _pad16(conv.out_channels)
        if march:
            assert k == 3 and not out_f32 and cout_pad == conv.out_channels
            xstack = cta_pair = 1
            w_main, b_main = self.pack_arrays(conv, bn, cin_pad, cout_pad, march=march)
            if shortcut is None:
                w_off, b_off = self._append_blob(w_main), self._append_blob(b_main)
            else:
                sc_conv, sc_bn, _ = shortcut
                assert res < 0 and sc_conv.kernel_size[0] == 1 and _pad16(sc_conv.in_channels) * 2 == cin_pad
                w_sc, b_sc = self.pack_arrays(sc_conv, sc_bn, cin_pad // 2, cout_pad)
                w_off = self._append_blob(np.concatenate([w_main, w_sc]))
                b_off = self._append_blob((b_main.astype(np.float64) + b_sc.astype(np.float64)).astype(np.float32))
        elif shortcut is None:
            w_off, b_off = self.pack(conv, bn, cin_pad, cout_pad, xstack, cta_pair)
        else:
            sc_conv, sc_bn, _ = shortcut
            assert k == 3 and res < 0 and sc_conv.kernel_size[0] == 1 and _pad16(sc_conv.in_channels) * 2 == cin_pad
            w_main, b_main = self.pack_arrays(conv, bn, cin_pad, cout_pad, xstack, cta_pair)
            w_sc, b_sc = self.pack_arrays(sc_conv, sc_bn, cin_pad // 2, cout_pad, xstack, cta_pair)
            wm, ws = w_main.reshape(cta_pair, -1), w_sc.reshape(cta_pair, -1)     # half-major blobs
            w_off = self._append_blob(np.concatenate([np.concatenate([wm[h], ws[h]]) for h in range(cta_pair)]))
            b_off = self._append_blob((b_main.astype(np.float64) + b_sc.astype(np.float64)).astype(np.float32))
        op = _lib.V2VOp()
        op.type = _lib.OP_CONV3_MARCH if march else _lib.OP_CONV
        op.flags = (_lib.F_RELU if relu else 0) | (_lib.F_RESIDUAL if res >= 0 else 0) | (_lib.F_OUT_F32 if out_f32 else 0)
        op.ksize, op.cin, op.cout, op.cout_real = k, cin_pad, cout_pad, conv.out_channels
        op.src, op.dst, op.res, op.impl, op.xstack, op.cta_pair = src, dst, res, 0, xstack, cta_pair
        op.w_offset, op.b_offset = w_off, b_off
        op.src2, op.cin2 = (shortcut[2], cin_pad // 2) if shortcut is not None else (-1, 0)
        op.lay_src = self.lay_of(src)
        op.lay_dst = self.lay_of(dst) if not out_f32 else self.lay_of(src)
        self.ops.append(op)
        fl = 2 * conv.in_channels * conv.out_channels * k ** 3 * op.lay_src.side ** 3
        if shortcut is not None:
            fl += 2 * shortcut[0].in_channels * shortcut[0].out_channels * op.lay_src.side ** 3
        self.flops += fl
        self.meta.append(dict(kind="conv", cin=conv.in_channels, cout=conv.out_channels, k=k,
                              side=op.lay_src.side, flops=fl))

    def stem_s2d(self, conv: nn.Conv3d, bn, src: int, dst: int, cta_pair: int = 1):
        """7^3 stem from the space-to-depth input (SCENEEGO_OP_STEM7_S2D)."""
        lib = _lib.load_library()
        assert conv.in_channels == 33 and conv.out_channels == 16 and conv.kernel_size[0] == 7
        w = conv.weight.detach().float().cpu().contiguous().numpy()
        bias = conv.bias.detach().float().cpu().contiguous().numpy() if conv.bias is not None else None
        w_out = np.zeros(lib.sceneego_v2v_stem_s2d_weight_bytes() // 2, dtype=np.uint16)
        b_out = np.zeros(16, dtype=np.float32)

        def fp(a):
            return a.ctypes.data_as(C.c_void_p) if a is not None else None
        g = bn.weight.detach().float().cpu().contiguous().numpy()
        bt = bn.bias.detach().float().cpu().contiguous().numpy()
        mu = bn.running_mean.detach().float().cpu().contiguous().numpy()
        var = bn.running_var.detach().float().cpu().contiguous().numpy()
        _lib._check(lib.sceneego_v2v_pack_stem_s2d(fp(w), fp(bias), fp(g), fp(bt), fp(mu), fp(var), C.c_double(float(bn.eps)),
                                                   int(cta_pair), fp(w_out), fp(b_out)), "v2v_pack_stem_s2d")
        op = _lib.V2VOp()
        op.type, op.flags = _lib.OP_STEM7_S2D, _lib.F_RELU
        op.src2, op.cin2 = -1, 0
        op.ksize, op.cin, op.cout, op.cout_real = 7, 33, 16, 16
        op.src, op.dst, op.res, op.impl, op.xstack, op.cta_pair = src, dst, -1, 0, 1, cta_pair
        op.w_offset, op.b_offset = self._append_blob(w_out), self._append_blob(b_out)
        op.lay_src, op.lay_dst = self.lay_of(src), self.lay_of(dst)
        self.ops.append(op)
        fl = 2 * 33 * 16 * 343 * op.lay_dst.side ** 3
        self.flops += fl
        self.meta.append(dict(kind="conv", cin=33, cout=16, k=7, side=op.lay_dst.side, flops=fl))

    def stem_march(self, conv: nn.Conv3d, bn, src: int, dst: int):
        """7^3 stem as an x-marching banded GEMM from the z-window input (SCENEEGO_OP_STEM7_MARCH)."""
        lib = _lib.load_library()
        assert conv.in_channels == 33 and conv.out_channels == 16 and conv.kernel_size[0] == 7
        w = conv.weight.detach().float().cpu().contiguous().numpy()
        bias = conv.bias.detach().float().cpu().contiguous().numpy() if conv.bias is not None else None
        w_out = np.zeros(lib.sceneego_v2v_stem_march_weight_bytes() // 2, dtype=np.uint16)
        b_out = np.zeros(16, dtype=np.float32)

        def fp(a):
            return a.ctypes.data_as(C.c_void_p) if a is not None else None
        g = bn.weight.detach().float().cpu().contiguous().numpy()
        bt = bn.bias.detach().float().cpu().contiguous().numpy()
        mu = bn.running_mean.detach().float().cpu().contiguous().numpy()
        var = bn.running_var.detach().float().cpu().contiguous().numpy()
        _lib._check(lib.sceneego_v2v_pack_stem_march(fp(w), fp(bias), fp(g), fp(bt), fp(mu), fp(var), C.c_double(float(bn.eps)),
                                                     fp(w_out), fp(b_out)), "v2v_pack_stem_march")
        op = _lib.V2VOp()
        op.type, op.flags = _lib.OP_STEM7_MARCH, _lib.F_RELU
        op.src2, op.cin2 = -1, 0
        op.ksize, op.cin, op.cout, op.cout_real = 7, 33, 16, 16
        op.src, op.dst, op.res, op.impl, op.xstack, op.cta_pair = src, dst, -1, 0, 1, 1
        op.w_offset, op.b_offset = self._append_blob(w_out), self._append_blob(b_out)
        op.lay_src, op.lay_dst = self.lay_of(src), self.lay_of(dst)
        self.ops.append(op)
        fl = 2 * 33 * 16 * 343 * op.lay_dst.side ** 3
        self.flops += fl
        self.meta.append(dict(kind="conv", cin=33, cout=16, k=7, side=op.lay_dst.side, flops=fl))

    def tail_mlp(self, blocks, out_conv: nn.Conv3d, src: int, dst: int):
        """Two 1x1 conv+BN+ReLU blocks and the 1x1 output conv as one op (SCENEEGO_OP_TAIL_MLP)."""
        parts_w, parts_b = [], []
        for conv, bn in [(b.block[0], b.block[1]) for b in blocks] + [(out_conv, None)]:
            assert conv.kernel_size[0] == 1 and conv.in_channels == 32
            w, b = self.pack_arrays(conv, bn, 32, _pad16(conv.out_channels))
            parts_w.append(w.view(np.uint8))
            parts_b.append(b.view(np.uint8))
        seg = np.concatenate(parts_w + parts_b)
        assert seg.size == 2048 + 2048 + 1024 + 4 * (32 + 32 + 16)
        op = _lib.V2VOp()
        op.type, op.flags = _lib.OP_TAIL_MLP, _lib.F_OUT_F32
        op.src2, op.cin2 = -1, 0
        op.ksize, op.cin, op.cout, op.cout_real = 1, 32, 16, out_conv.out_channels
        op.src, op.dst, op.res, op.impl, op.xstack = src, dst, -1, 0, 1
        op.w_offset = self._append_blob(seg)
        op.b_offset = op.w_offset + 5120
        op.lay_src = self.lay_of(src)
        op.lay_dst = self.lay_of(src)
        self.ops.append(op)
        side = op.lay_src.side
        fl = 2 * (32 * 32 * 2 + 32 * out_conv.out_channels) * side ** 3
        self.flops += fl
        self.meta.append(dict(kind="tail", cin=32, cout=out_conv.out_channels, k=1, side=side, flops=fl))

    def pool(self, src: int, dst: int, channels: int):
        op = _lib.V2VOp()
        op.type, op.cin, op.cout, op.cout_real = _lib.OP_MAXPOOL2, channels, channels, channels
        op.src2, op.cin2 = -1, 0
        op.src, op.dst, op.res = src, dst, -1
        op.lay_src, op.lay_dst = self.lay_of(src), self.lay_of(dst)
        self.ops.append(op)
        self.meta.append(dict(kind="pool", cin=channels, cout=channels, k=2, side=op.lay_src.side, flops=0))

    def deconv(self, conv: nn.ConvTranspose3d, bn, src: int, dst: int, add: int):
        cin_pad, cout_pad = _pad16(conv.in_channels), _pad16(conv.out_channels)
        w_off, b_off = self.pack(conv, bn, cin_pad, cout_pad)
        op = _lib.V2VOp()
        op.type = _lib.OP_DECONV2
        op.src2, op.cin2 = -1, 0
        op.flags = _lib.F_RELU | (_lib.F_ADD_AFTER if add >= 0 else 0)
        op.ksize, op.cin, op.cout, op.cout_real = 2, cin_pad, cout_pad, conv.out_channels
        op.src, op.dst, op.res = src, dst, add
        op.w_offset, op.b_offset = w_off, b_off
        op.lay_src, op.lay_dst = self.lay_of(src), self.lay_of(dst)
        self.ops.append(op)
        fl = 2 * conv.in_channels * conv.out_channels * 8 * op.lay_src.side ** 3
        self.flops += fl
        self.meta.append(dict(kind="deconv", cin=conv.in_channels, cout=conv.out_channels, k=2,
                              side=op.lay_src.side, flops=fl))

    def finalize(self):
        self.blob = torch.from_numpy(np.concatenate(self.blob_parts)).to(self.device)
        self.op_array = (_lib.V2VOp * len(self.ops))(*self.ops)
        self.buf_ptrs = (C.c_void_p * len(self.buffers))(*[C.c_void_p(t.data_ptr()) for t in self.buffers])
        self.blob_parts = []


class V2VModel(nn.Module):
    """network/v2v.py:142-181.  forward(x) takes (B,C,V,V,V) f32 like the reference and
    returns (B,out,V,V,V) f32 logits; the fused path used by VoxelNetwork_depth writes the
    stem input directly (see `input_buffer`) and calls `run_chunk`."""

    def __init__(self, input_channels, output_channels, max_chunk: int = 32):
        super().__init__()
        self.input_channels, self.output_channels = input_channels, output_channels
        self.max_chunk = max_chunk
        self.stem_xstack = 4
        self.c32_xstack = 2      # 3^3 convs with Cout = 32: stack two x-planes (N = 64)
        self.fuse_tail = True
        self.fuse_shortcut = True  # 1x1 projection shortcuts accumulate into the second conv of their Res3DBlock
        self.cta_pair = 2        # 1 = every conv on single CTAs
        self.march = True        # 3^3 convs with Cout = 32: x-marching banded GEMM (csrc/march.cu) instead of x-stacking
        self.stem = "march"      # 33 -> 16 stem: "march" (csrc/stem_march.cu) or "s2d" (csrc/stem.cu, 2x2x2-stacked)
        self.front_layers = nn.Sequential(Basic3DBlock(input_channels, 16, 7), Res3DBlock(16, 32),
                                          Res3DBlock(32, 32), Res3DBlock(32, 32))
        self.encoder_decoder = EncoderDecorder()
        self.back_layers = nn.Sequential(Res3DBlock(32, 32), Basic3DBlock(32, 32, 1), Basic3DBlock(32, 32, 1))
        self.output_layer = nn.Conv3d(32, output_channels, kernel_size=1, stride=1, padding=0)
        self._initialize_weights()
        self._programs: Dict[int, _Program] = {}
        self._weights_version = None

    def _initialize_weights(self):
        # v2v.py:172-181
        for m in self.modules():
            if isinstance(m, (nn.Conv3d, nn.ConvTranspose3d)):
                nn.init.xavier_normal_(m.weight)
                nn.init.constant_(m.bias, 0)

    # -- program construction --------------------------------------------------
    def invalidate(self):
        """Drop packed weights and buffer pools (also happens by itself when a parameter or BatchNorm buffer
        changes, see `_current_version`)."""
        self._programs.clear()
        self._weights_version = None

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        self.invalidate()

    def _current_version(self):
        """Fingerprint of every tensor the packed blob is derived from: in-place edits bump `_version`,
        `.to()` / `.half()` / `load_state_dict(assign=True)` change the storage address."""
        acc, n = 0, 0
        for t in self._tracked_tensors():
            acc = (acc * 1000003 + t._version * 31 + t.data_ptr()) & 0xFFFFFFFFFFFFFFFF
            n += 1
        return acc, n

    def _tracked_tensors(self):
        for prm in self.parameters():
            yield prm
        for name, buf in self.named_buffers():
            if not name.endswith("num_batches_tracked"):
                yield buf

    def _refuse_training(self):
        # BatchNorm is folded from the running statistics: there is no training-mode forward (the reference is
        # inference-only too: demo.py:32 / test.py:29 call .eval()); batch statistics would silently differ
        if self.training:
            raise _lib.SceneEgoError("V2VModel is inference-only (BatchNorm is folded from the running statistics): "
                                     "call .eval() first")

    def _res(self, pg: _Program, blk: Res3DBlock, x: int, level: int) -> int:
        xs = self.c32_xstack if blk.res_branch[0].out_channels == 32 else 1
        # CTA pairs (tcgen05 cta_group::2) where the B operand's shared-memory read bounds the MMA: N = 64
        wide = xs * blk.res_branch[0].out_channels == 64 and level <= 1
        cg = self.cta_pair if wide else 1
        cg0 = cg if blk.res_branch[0].in_channels >= 32 else 1
        mr = self.march and blk.res_branch[0].out_channels == 32
        t = pg.acquire(level)
        pg.conv(blk.res_branch[0], blk.res_branch[1], x, t, relu=True, xstack=xs, cta_pair=cg0, march=mr)
        y = pg.acquire(level)
        if len(blk.skip_con) > 0 and self.fuse_shortcut and blk.skip_con[0].in_channels * 2 == blk.skip_con[0].out_channels:
            # relu(conv3(t) + bn(conv1(x))): the 1x1 projection of x joins the second conv's accumulation
            pg.conv(blk.res_branch[3], blk.res_branch[4], t, y, relu=True, xstack=xs, cta_pair=cg,
                    shortcut=(blk.skip_con[0], blk.skip_con[1], x), march=mr)
            pg.release(t)
            return y
        if len(blk.skip_con) > 0:
            s = pg.acquire(level)
            pg.conv(blk.skip_con[0], blk.skip_con[1], x, s, relu=False)
        else:
            s = x
        pg.conv(blk.res_branch[3], blk.res_branch[4], t, y, relu=True, res=s, xstack=xs, cta_pair=cg, march=mr)
        pg.release(t)
        if s != x:
            pg.release(s)
        return y

    def _build(self, side: int, chunk: int, device) -> _Program:
        if side % 32 != 0:
            raise _lib.SceneEgoError("V2V needs a volume side divisible by 32 (five 2x poolings)")
        pg = _Program(side, chunk, self.input_channels, device, stem=self.stem)
        ed = self.encoder_decoder
        x = pg.acquire(0)
        # 7^3 stem, Cout = 16: the worst tcgen05 shape (N = 16 costs as much as N = 32 per MMA), so four
        # adjacent x-planes of outputs are stacked into N = 64 (tools/mma_rate.cu for the cost model)
        if pg.zwin:
            pg.stem_march(self.front_layers[0].block[0], self.front_layers[0].block[1], pg.in_buf, x)
        elif pg.s2d:
            pg.stem_s2d(self.front_layers[0].block[0], self.front_layers[0].block[1], pg.in_buf, x, cta_pair=self.cta_pair)
        else:
            pg.conv(self.front_layers[0].block[0], self.front_layers[0].block[1], pg.in_buf, x, relu=True,
                    xstack=self.stem_xstack)
        for i in (1, 2, 3):
            y = self._res(pg, self.front_layers[i], x, 0)
            pg.release(x)
            x = y
        skips = []
        for lvl in range(1, 6):
            skips.append(self._res(pg, getattr(ed, f"skip_res{lvl}"), x, lvl - 1))
            p = pg.acquire(lvl)
            pg.pool(x, p, EncoderDecorder.CHANNELS[lvl - 1])
            pg.release(x)
            x = self._res(pg, getattr(ed, f"encoder_res{lvl}"), p, lvl)
            pg.release(p)
        y = self._res(pg, ed.mid_res, x, 5)
        pg.release(x)
        x = y
        for lvl in range(5, 0, -1):
            y = self._res(pg, getattr(ed, f"decoder_res{lvl}"), x, lvl)
            pg.release(x)
            up = getattr(ed, f"decoder_upsample{lvl}")
            u = pg.acquire(lvl - 1)
            pg.deconv(up.block[0], up.block[1], y, u, add=skips[lvl - 1])
            pg.release(y)
            pg.release(skips[lvl - 1])
            x = u
        y = self._res(pg, self.back_layers[0], x, 0)
        pg.release(x)
        x = y
        # f32 logits (B,J,V,V,V): the dst buffer slot is patched per call
        pg.buffers.append(torch.empty(0, device=device))
        pg.buf_level.append(0)
        pg.logits_buf = len(pg.buffers) - 1
        if self.fuse_tail and self.output_channels <= 16:
            # the three trailing 1x1 convs are HBM-bound: one fused pass (csrc/tail.cu)
            pg.tail_mlp([self.back_layers[1], self.back_layers[2]], self.output_layer, x, pg.logits_buf)
        else:
            for i in (1, 2):
                y = pg.acquire(0)
                pg.conv(self.back_layers[i].block[0], self.back_layers[i].block[1], x, y, relu=True)
                pg.release(x)
                x = y
            pg.conv(self.output_layer, None, x, pg.logits_buf, relu=False, out_f32=True)
        pg.finalize()
        return pg

    def program(self, side: int, chunk: int, device) -> _Program:
        """The op program for volumes of `side` holding at least `chunk` frames.  ONE program per side is kept: a
        smaller batch runs on the existing (larger) buffer pool, a larger one (up to max_chunk, rounded up to a
        power of two) replaces it -- a ragged last DataLoader batch or the pipeline's ramp never builds a second
        multi-GB pool.  Rebuilt when a weight or BatchNorm buffer changed since it was packed."""
        device = torch.device(device)
        if device.type == "cuda":            # (a CPU device only builds the op list / packed blob: host-logic tests)
            device = _lib._as_device(device)
        ver = self._current_version()
        if ver != self._weights_version:
            self._programs.clear()
            self._weights_version = ver
        pg = self._programs.get(side)
        if pg is None or pg.device != device or pg.chunk < chunk:
            want = 1
            while want < chunk:
                want *= 2
            want = max(chunk, min(want, self.max_chunk))
            self._programs.pop(side, None)
            pg = None
            if device.type == "cuda":
                with _lib.on_device(device):
                    pg = self._build(side, want, device)
            else:
                pg = self._build(side, want, device)
            self._programs[side] = pg
        return pg

    # -- execution -------------------------------------------------------------
    def run_chunk(self, pg: _Program, batch: int, logits_out: torch.Tensor, impl: Optional[int] = None) -> int:
        """Run the program on the first `batch` frames staged in pg.in_buf; writes
        logits_out (batch, out, V, V, V) f32.  Returns the number of kernels launched."""
        assert batch <= pg.chunk and logits_out.is_contiguous() and logits_out.dtype == torch.float32
        self._refuse_training()
        if logits_out.device != pg.device:
            raise _lib.SceneEgoError(f"v2v_run: logits on {logits_out.device}, program on {pg.device}")
        pg.buf_ptrs[pg.logits_buf] = C.c_void_p(logits_out.data_ptr())
        if impl is not None:
            for op in pg.op_array:
                op.impl = impl
        lib = _lib.load_library()
        with _lib.on_device(pg.device):
            rc = lib.sceneego_v2v_run(pg.op_array, len(pg.ops), pg.buf_ptrs, C.c_void_p(pg.blob.data_ptr()), int(batch),
                                      _lib._stream(pg.device))
        _lib._check(rc, "v2v_run")
        return lib.sceneego_v2v_last_launch_count()

    def profile_chunk(self, pg: _Program, batch: int, logits_out: torch.Tensor):
        """Like run_chunk but returns [(op, milliseconds)] measured with CUDA events per op."""
        self._refuse_training()
        pg.buf_ptrs[pg.logits_buf] = C.c_void_p(logits_out.data_ptr())
        ms = (C.c_float * len(pg.ops))()
        with _lib.on_device(pg.device):
            rc = _lib.load_library().sceneego_v2v_run_profile(pg.op_array, len(pg.ops), pg.buf_ptrs,
                                                              C.c_void_p(pg.blob.data_ptr()), int(batch),
                                                              _lib._stream(pg.device), ms)
        _lib._check(rc, "v2v_run_profile")
        return [(pg.meta[i], float(ms[i])) for i in range(len(pg.ops))]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise _lib.SceneEgoError("V2VModel runs on CUDA only (no CPU fallback)")
        b, c, side = x.shape[0], x.shape[1], x.shape[2]
        assert c == self.input_channels
        chunk = min(self.max_chunk, b)
        pg = self.program(side, chunk, x.device)
        out = torch.empty(b, self.output_channels, side, side, side, dtype=torch.float32, device=x.device)
        x = x.contiguous().float()
        for s in range(0, b, chunk):
            n = min(chunk, b - s)
            _lib.pack_volume(x[s:s + n], pg.buffers[pg.in_buf], pg.lay_in)
            self.run_chunk(pg, n, out[s:s + n])
        return out

    def flops_per_frame(self, side: int) -> int:
        """Algorithmic FLOPs of one frame (2*Cin*Cout*k^3 per output voxel)."""
        total = 0
        res = {"front_layers": side, "back_layers": side, "output_layer": side}
        for name, m in self.named_modules():
            if isinstance(m, (nn.Conv3d, nn.ConvTranspose3d)):
                s = _side_of(name, side)
                k = m.kernel_size[0]
                if isinstance(m, nn.ConvTranspose3d):
                    total += 2 * m.in_channels * m.out_channels * 8 * (s // 2) ** 3
                else:
                    total += 2 * m.in_channels * m.out_channels * k ** 3 * s ** 3
        return total


class EncoderDecoderSimple(nn.Module):
    """v2v.py:184-221: two pool/res levels (32 -> 64 -> 128), mid res, two res/upsample levels with skips."""

    def __init__(self):
        super().__init__()
        self.encoder_pool1 = Pool3DBlock(2)
        self.encoder_res1 = Res3DBlock(32, 64)
        self.encoder_pool2 = Pool3DBlock(2)
        self.encoder_res2 = Res3DBlock(64, 128)
        self.mid_res = Res3DBlock(128, 128)
        self.decoder_res2 = Res3DBlock(128, 128)
        self.decoder_upsample2 = Upsample3DBlock(128, 64, 2, 2)
        self.decoder_res1 = Res3DBlock(64, 64)
        self.decoder_upsample1 = Upsample3DBlock(64, 32, 2, 2)
        self.skip_res1 = Res3DBlock(32, 32)
        self.skip_res2 = Res3DBlock(64, 64)


class V2VModelSimple(V2VModel):
    """network/v2v.py:224-257 -- the shallow V2V variant (not reachable from VoxelNetwork_depth, which builds
    V2VModel; provided for users of the reference class).  Same state-dict names as the reference; the same op
    program machinery: 7^3 stem with 32 outputs on conv_tc_kernel (plain padded input), the 32-channel Res block on
    the marching kernel, 64/128-channel blocks, pools and transposed convs as in V2VModel, the two trailing 1x1
    convs unfused."""

    def __init__(self, input_channels, output_channels, max_chunk: int = 32):
        nn.Module.__init__(self)
        self.input_channels, self.output_channels = input_channels, output_channels
        self.max_chunk = max_chunk
        self.c32_xstack, self.fuse_shortcut, self.cta_pair, self.march = 2, True, 2, True
        self.stem = "march"
        self.front_layers = nn.Sequential(Basic3DBlock(input_channels, 32, 7))
        self.encoder_decoder = EncoderDecoderSimple()
        self.back_layers = nn.Sequential(Basic3DBlock(32, 32, 1))
        self.output_layer = nn.Conv3d(32, output_channels, kernel_size=1, stride=1, padding=0)
        self._initialize_weights()
        self._programs = {}
        self._weights_version = None

    def _build(self, side: int, chunk: int, device) -> _Program:
        if side % 4 != 0:
            raise _lib.SceneEgoError("V2VModelSimple needs a volume side divisible by 4 (two 2x poolings)")
        pg = _Program(side, chunk, self.input_channels, device, allow_s2d=False)
        ed = self.encoder_decoder
        x = pg.acquire(0)
        pg.conv(self.front_layers[0].block[0], self.front_layers[0].block[1], pg.in_buf, x, relu=True)
        skip1 = self._res(pg, ed.skip_res1, x, 0)
        p = pg.acquire(1)
        pg.pool(x, p, 32)
        pg.release(x)
        x = self._res(pg, ed.encoder_res1, p, 1)
        pg.release(p)
        skip2 = self._res(pg, ed.skip_res2, x, 1)
        p = pg.acquire(2)
        pg.pool(x, p, 64)
        pg.release(x)
        x = self._res(pg, ed.encoder_res2, p, 2)
        pg.release(p)
        for blk in (ed.mid_res, ed.decoder_res2):
            y = self._res(pg, blk, x, 2)
            pg.release(x)
            x = y
        u = pg.acquire(1)
        pg.deconv(ed.decoder_upsample2.block[0], ed.decoder_upsample2.block[1], x, u, add=skip2)
        pg.release(x)
        pg.release(skip2)
        x = self._res(pg, ed.decoder_res1, u, 1)
        pg.release(u)
        u = pg.acquire(0)
        pg.deconv(ed.decoder_upsample1.block[0], ed.decoder_upsample1.block[1], x, u, add=skip1)
        pg.release(x)
        pg.release(skip1)
        y = pg.acquire(0)
        pg.conv(self.back_layers[0].block[0], self.back_layers[0].block[1], u, y, relu=True)
        pg.release(u)
        pg.buffers.append(torch.empty(0, device=device))
        pg.buf_level.append(0)
        pg.logits_buf = len(pg.buffers) - 1
        pg.conv(self.output_layer, None, y, pg.logits_buf, relu=False, out_f32=True)
        pg.finalize()
        return pg

    def flops_per_frame(self, side: int) -> int:
        total = 0
        for name, m in self.named_modules():
            if isinstance(m, nn.ConvTranspose3d):
                lvl = int(name.split(".")[1][-1])
                total += 2 * m.in_channels * m.out_channels * 8 * (side >> lvl) ** 3
            elif isinstance(m, nn.Conv3d):
                part = name.split(".")[1] if name.startswith("encoder_decoder.") else ""
                lvl = 0 if not part else 2 if part == "mid_res" else (int(part[-1]) - 1 if part.startswith("skip_res") else int(part[-1]))
                total += 2 * m.in_channels * m.out_channels * m.kernel_size[0] ** 3 * (side >> lvl) ** 3
        return total


def _side_of(name: str, side: int) -> int:
    """Spatial side of the OUTPUT of the conv called `name` inside V2VModel."""
    if not name.startswith("encoder_decoder."):
        return side
    part = name.split(".")[1]
    if part == "mid_res":
        return side >> 5
    lvl = int(part[-1])
    if part.startswith("skip_res"):
        return side >> (lvl - 1)
    if part.startswith("encoder_res") or part.startswith("decoder_res"):
        return side >> lvl
    if part.startswith("decoder_upsample"):
        return side >> (lvl - 1)
    raise KeyError(name)
