"""Timing of the backbone hand-off on one B200.

    python tools/bench_handoff.py [--batch 64] [--iters 10] [--forward] [--out profiles/...json]

(a) the fused kernel pair (handoff_pack_kernel + handoff_tc_kernel; 20 calls replayed as one CUDA graph, because one
    eager call costs the CPU more than the kernels cost the GPU) against the torch ops it replaces --
    ConvTranspose2d(256,256,4,2,1) + BatchNorm2d + ReLU of pose_resnet's head (fp32, TF32 off: what the reference runs)
    followed by this repo's feature_conv1x1_kernel -- CUDA events on the current stream, after warm-up;
(b) with --forward: VoxelNetwork_depth.forward() from images with and without the hand-off.
A diagnostic, not the bench: bench.py measures lift(), whose input is the backbone's feature map by definition.
"""
import argparse
import json
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from sceneego_b200 import _lib  # noqa: E402


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ms[len(ms) // 2]


def timed_graph(fn, reps=20, iters=5):
    """GPU time of one call with the CPU launch overhead taken out: `reps` calls captured in one CUDA graph."""
    fn()
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps):
                fn()
    torch.cuda.current_stream().wait_stream(st)
    return timed(g.replay, iters, 2) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--forward", action="store_true")
    ap.add_argument("--kernel-only", action="store_true", help="run the fused kernel a few times and exit (for ncu)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B, h, w = a.batch, 32, 32
    g = torch.Generator().manual_seed(1)
    dc = nn.ConvTranspose2d(256, 256, 4, stride=2, padding=1, bias=False).cuda().eval()
    bn = nn.BatchNorm2d(256).cuda().eval()
    pf = nn.Conv2d(256, 32, 1).cuda().eval()
    x = torch.relu(torch.randn(B, 256, h, w, generator=g)).cuda()
    wts, bias = _lib.handoff_pack(dc.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, pf.weight, pf.bias, "cuda")
    out = torch.empty(B, 2 * h, 2 * w, 32, device="cuda")
    if a.kernel_only:
        for _ in range(3):
            _lib.backbone_handoff(x, wts, bias, out=out)
        torch.cuda.synchronize()
        return
    res = {"batch": B, "input": [B, 256, h, w]}
    flops = 2.0 * B * (2 * h) * (2 * w) * 256 * (256 * 4 + 32)
    res["algorithmic_GFLOP"] = flops / 1e9
    with torch.no_grad():
        res["fused_ms_eager_launch"] = timed(lambda: _lib.backbone_handoff(x, wts, bias, out=out), a.iters)
        res["fused_ms"] = timed_graph(lambda: _lib.backbone_handoff(x, wts, bias, out=out))
        res["fused_TFLOP_per_s"] = flops / res["fused_ms"] / 1e9
        w1 = pf.weight.reshape(32, 256).contiguous()
        f256 = F.relu(bn(dc(x)))
        res["torch_deconv_bn_relu_fp32_ms"] = timed(lambda: F.relu(bn(dc(x))), a.iters)
        res["feature_conv1x1_kernel_ms"] = timed(lambda: _lib.feature_conv1x1(f256, w1, pf.bias), a.iters)
        res["replaced_ms"] = res["torch_deconv_bn_relu_fp32_ms"] + res["feature_conv1x1_kernel_ms"]
        res["speedup_vs_replaced"] = res["replaced_ms"] / res["fused_ms"]
        torch.backends.cudnn.allow_tf32 = True
        res["torch_deconv_bn_relu_tf32_ms"] = timed(lambda: F.relu(bn(dc(x))), a.iters)
        torch.backends.cudnn.allow_tf32 = False
    if a.forward:
        from sceneego_b200.network.voxel_net_depth import VoxelNetwork_depth
        from sceneego_b200.utils.cfg import default_config
        cfg = default_config(batch_size=B)
        img = torch.randn(B, 3, 256, 256, generator=g).cuda()
        depth = (torch.rand(B, 1024, 1280, generator=g) * 3.0).cuda()
        for hand in (False, True):
            torch.manual_seed(0)
            net = VoxelNetwork_depth(cfg, device="cuda", backbone_handoff=hand, materialize_features=False,
                                     materialize_volumes=False).eval()
            with torch.no_grad():
                ms = timed(lambda: net(img, net.grid_coord_proj_batch, net.coord_volumes, depth_map_batch=depth), max(3, a.iters // 2), 2)
                bb = timed(lambda: net.backbone(img), max(3, a.iters // 2), 2)
            res[f"forward_ms_handoff_{hand}"] = ms
            res[f"backbone_only_ms_handoff_{hand}"] = bb
            del net
            torch.cuda.empty_cache()
    line = json.dumps(res)
    print(line)
    if a.out:
        with open(a.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
