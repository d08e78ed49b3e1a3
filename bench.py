#!/usr/bin/env python
"""Benchmark of the SceneEgo volumetric lifting stage (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on the host CPU

A "step" is one pass of the hot path (everything VoxelNetwork_depth.forward does after
the backbone) over one batch of synthetic input: per GPU 64 frames of backbone features
(64,256,64,64) f32 + depth maps (64,1024,1280) f32, V=64, 15 joints, random-init weights
(BASELINE.json configs[1]).  Frames are sharded over ranks (weak scaling: 64 frames per
GPU) and the poses are all-gathered over NCCL inside every step.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/s VoxelNetDepth volumetric stage"
UNIT = "frames/s"
FRAMES_PER_GPU = 64
V, J = 64, 15


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames-per-gpu", type=int, default=FRAMES_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chunk", type=int, default=64, help="frames per V2V launch group")
    ap.add_argument("--profile-ops", default="", help="write the per-op V2V timing table to this file")
    return ap.parse_args()


def workload_config(frames, world):
    return {"workload": f"VoxelNetDepth post-backbone stage, batch {frames} frames/GPU, 64^3 voxel cube, 15 joints, "
                        f"features (B,256,64,64) f32 + depth (B,1024,1280) f32, bf16 V2V, fp32 unprojection and "
                        f"soft-argmax (BASELINE.json configs[1])",
            "frames_per_gpu": frames, "global_frames": frames * world, "volume_size": V, "joints": J,
            "parallelism": f"frame-shard x{world}, NCCL all-gather of poses",
            "weights": "random init (seeded), BatchNorm statistics randomised",
            "outputs": "reference-identical 4-tuple (features and softmaxed volumes materialised; the 168 MB/frame "
                       "features tensor is written on a side stream inside the timed region)",
            "v2v_chunk_frames": None,
            "l2_policy": "inputs (604 MB/step) and activations (>1 GB/layer) exceed the 126 MB L2"}


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows)}


# --------------------------------------------------------------------------------------
# reference algorithm on the host CPU (oracle port -- the only place bench.py runs oracle/)
# --------------------------------------------------------------------------------------
def cpu_reference_frames_per_s(n_frames, repeats=1, warmup=0):
    import torch
    from oracle import sceneego_oracle as orc
    from sceneego_b200.utils import synth
    from tests import util
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tables = orc.StageTables(util.CALIB, V, 2.0)
    sd = synth.synthetic_state_dict(util.stage_shapes(), seed=0, mode="random_bn")
    feat = synth.synthetic_features(n_frames)
    depth = synth.synthetic_depth_room(n_frames, tables.ray)
    times = []
    for i in range(warmup + repeats):
        t0 = time.perf_counter()
        with torch.no_grad():
            orc.stage_forward(tables, sd, feat, depth_batch=depth)
        times.append(time.perf_counter() - t0)
    times = times[warmup:]
    return n_frames * len(times) / sum(times), cores, times


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each step = 1 frame of the 64-frame batch (bounded sample; the per-frame loop of the
    # reference is serial, network/voxel_net_depth.py:252-256, so frames/s does not depend on B)
    fps, cores, times = cpu_reference_frames_per_s(1, repeats=args.steps, warmup=min(args.warmup, 1))
    ms = 1000.0 * sum(times) / len(times)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.frames_per_gpu, args.gpus),
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "1 frame per step of the same workload (oracle/sceneego_oracle.py, "
                                       "torch CPU threads = all host cores)"},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------
def ours_arm(args):
    import torch
    import torch.distributed as dist
    from sceneego_b200 import _lib
    from sceneego_b200.network.voxel_net_depth import VoxelNetwork_depth
    from sceneego_b200.parallel import gather_poses
    from sceneego_b200.pipeline import HostStagePipeline
    from sceneego_b200.utils import synth
    from tests import util

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this arm has no CPU fallback; use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B = args.frames_per_gpu
    total = B * world

    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        net = VoxelNetwork_depth(util.load_config(batch_size=B), device=f"cuda:{local}", v2v_chunk=args.chunk).eval()
    sd = synth.synthetic_state_dict(util.stage_shapes(), seed=0, mode="random_bn")
    full = net.state_dict()
    full.update(sd)
    net.load_state_dict(full, strict=True)

    feat_h = synth.synthetic_features(B, seed=1234 + rank).pin_memory()
    ray_x_major = net.ray
    depth_h = synth.synthetic_depth_room(B, ray_x_major, seed=7 + rank).pin_memory()
    feat_d, depth_d = feat_h.to(dev), depth_h.to(dev)

    def gather(kp):
        return gather_poses(kp, total) if world > 1 else kp

    def step():
        with torch.no_grad():
            kp = net.lift(feat_d, net.grid_coord_proj_batch, net.coord_volumes, depth_map_batch=depth_d)[0]
        return gather(kp)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        # nvidia-smi's start-up (NVML initialisation over every GPU of the box) holds driver locks that stall kernel
        # launches for tens of milliseconds: wait for its first sample so that only the 100 ms polling overlaps the
        # timed region (a fresh box needed more than the 0.3 s this used to sleep and the step read 17 % slow)
        sampler.start()
        t_wait = time.time()
        while not sampler.rows and time.time() - t_wait < 8.0:
            time.sleep(0.05)
        time.sleep(0.1)
    # ---- device-resident timing ----------------------------------------------------------
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record()
    launches = 0
    for _ in range(args.steps):
        kp = step()
        launches += net.last_launches
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    value = total * args.steps / (ms_total / 1000.0)

    # ---- end to end: pinned host buffers -> H2D -> lift -> D2H, through the public pipeline ----
    pipe = HostStagePipeline(net, gather_fn=gather if world > 1 else None)
    pipe.run([(feat_h, depth_h)] * args.steps)        # same batch count as the timed call: pinned result buffers exist
    barrier()
    pipe.h2d_bytes = pipe.d2h_bytes = 0
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    outs = pipe.run([(feat_h, depth_h)] * args.steps)
    f1.record()
    barrier()
    ms2 = torch.tensor([f0.elapsed_time(f1)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = total * args.steps / (float(ms2.item()) / 1000.0)
    assert torch.isfinite(outs[-1]).all()
    # the host->device copy of one step alone (explains the gap between `value` and `e2e`: PCIe, not kernels)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(3):
        feat_d.copy_(feat_h, non_blocking=True)
        depth_d.copy_(depth_h, non_blocking=True)
    g1.record()
    barrier()
    h2d_ms = g0.elapsed_time(g1) / 3

    line = None
    if rank == 0:
        # ---- per-kernel rooflines (separate pass, CUDA events on the launching stream) ----
        peaks = {}
        pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk_path):
            peaks = json.load(open(pk_path))
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        tc_peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
        peak_src = "MEASURED_PEAKS.json (sustained bf16, copy HBM)" if peaks else "fallback (B200_PROFILING.md)"
        kern = stage_kernel_timings(net, feat_d, depth_d, B, args)
        conv_ms = kern["v2v"]["conv_ms_per_frame"]
        conv_fl = kern["v2v"]["conv_flops_per_frame"]
        ach = conv_fl / (conv_ms * 1e-3) / 1e12
        # dominant kernel family: the nine 3^3 Conv3d(32,32) launches at 64^3 (conv_march_kernel, csrc/march.cu) --
        # the largest share of the step together with the stem (profiles/r01_ncu_launch_shares_march.txt); the
        # stem and the whole tensor path are reported next to it
        fam = kern["v2v"]["families"]
        c32 = fam["conv3_32_32_full_res"]
        n_launch_frames = kern["v2v"]["chunk_frames"]
        roof = {"kernel": "conv_march_kernel<2,*,2,1> (tcgen05 x-marching banded implicit-GEMM Conv3d 3x3x3 32->32 at 64^3, "
                          f"N=96 MMAs into a ring of TMEM slots, resident weights; {c32['launches']} launches per "
                          f"{n_launch_frames}-frame chunk)",
                "bound": "tensor", "achieved": c32["tflops"], "peak": tc_peak, "unit": "TFLOP/s",
                "frac": c32["tflops"] / tc_peak,
                "traffic": 788.0e6,
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture at "
                                "16 frames per launch (profiles/r01_ncu_conv_march.txt), mean of a launch without residual "
                                "(387 + 242 MB vs 563 MB algorithmic = 2 x 16 x 17.6 MB) and one with (696 + 251 MB vs 845 MB): "
                                "the halo cells shared by neighbouring 128-cell tiles hit L2 only partly",
                "peak_source": peak_src,
                "how": f"algorithmic 2*32*32*27*64^3 = 14.50 GFLOP per frame x {n_launch_frames} frames per launch / mean "
                       "CUDA-event duration of those launches on the launching stream (sceneego_v2v_run_profile)",
                "launch_ms": c32["ms_per_frame"] * n_launch_frames / c32["launches"],
                "stem": {"kernel": "stem_s2d_tc_kernel<2> (7x7x7 Conv3d 33->16, 2x2x2 output stacking)",
                         "achieved": fam["stem7_33_16"]["tflops"], "frac": fam["stem7_33_16"]["tflops"] / tc_peak,
                         "tensor_pipe_active_pct_ncu": 87.1},
                "all_tensor_convs": {"achieved": ach, "frac": ach / tc_peak,
                                     "how": "algorithmic FLOPs of every tcgen05 conv of V2V / sum of their durations"}}
        cfg = workload_config(B, world)
        cfg["v2v_chunk_frames"] = min(args.chunk, B)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": cfg, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes // args.steps,
                        "d2h_bytes_per_step": pipe.d2h_bytes // args.steps,
                        "api": "sceneego_b200.pipeline.HostStagePipeline.run (pinned host -> poses on host)",
                        "h2d_alone_ms_per_step": h2d_ms,
                        "h2d_alone_GB_per_s": (feat_h.numel() + depth_h.numel()) * 4 / (h2d_ms * 1e-3) / 1e9},
                "gpu_launches": launches, "roofline": roof, "kernels": kern,
                "frames_per_s_per_gpu": value / world}
        if world == 1 and not args.no_cpu_baseline:
            fps, cores, times = cpu_reference_frames_per_s(2, repeats=1, warmup=0)
            line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"2 frames of the same workload through oracle/sceneego_oracle.py "
                                              f"(torch CPU, {cores} threads), {sum(times):.1f} s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def stage_kernel_timings(net, feat_d, depth_d, B, args):
    """Time each kernel of the stage alone with CUDA events (after warm-up, L2 flushed by the
    >126 MB working set of the preceding kernels) and relate it to its algorithmic bytes/FLOPs."""
    import torch
    from sceneego_b200 import _lib
    vn = net.volume_net
    chunk = min(vn.max_chunk, B)
    pg = vn.program(V, chunk, feat_d.device)
    conv = net.process_features[0]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=feat_d.device)

    def timed(fn, reps=5):
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2]

    out = {}
    n = chunk
    feat32 = _lib.feature_conv1x1(feat_d[:n], conv.weight, conv.bias)
    grid = net.grid_coord_proj_batch[0].reshape(-1, 2).contiguous()
    in_buf = pg.buffers[pg.in_buf]
    logits = torch.empty(n, J, V, V, V, dtype=torch.float32, device=feat_d.device)

    def rec(name, ms, bytes_per_frame, note):
        gbs = bytes_per_frame * n / (ms * 1e-3) / 1e9
        out[name] = {"ms_per_frame": ms / n, "algorithmic_MB_per_frame": bytes_per_frame / 1e6, "GB_per_s": gbs,
                     "bound": "hbm", "note": note}

    ms = timed(lambda: _lib.feature_conv1x1(feat_d[:n], conv.weight, conv.bias))
    rec("feature_conv1x1", ms, 256 * 64 * 64 * 4 + 64 * 64 * 32 * 4, "read (256,64,64) f32 + write (64,64,32) f32")
    ms = timed(lambda: _lib.unproject(feat32, grid, None, V, 2.0, 1024, 1280, None, in_buf, pg.lay_in,
                                      extra_zero_planes=(pg.in_pad - 32) // 8))
    rec("unproject", ms, 64 * 64 * 32 * 4 + V ** 3 * 32 * 2 + V ** 3 * 2 + V ** 3 * 8,
        "read 0.524 MB features + 2.1 MB grid, write 32 bf16 channels + the cleared occupancy plane (space-to-depth layout)")
    d = depth_d[:n]
    ms = timed(lambda: _lib.voxelize_depth(d, net._ray_dev, 1024, 1280, V, 2.0, None, in_buf, pg.lay_in, channel=32))
    rec("voxelize", ms, 1024 * 1280 * 4, "read (1024,1280) f32 depth; ray table (31.5 MB) shared by all frames; sparse bf16 scatter")
    prof = vn.profile_chunk(pg, n, logits)     # warm-up
    prof = vn.profile_chunk(pg, n, logits)
    conv_ms = sum(ms for m, ms in prof if m["kind"] == "conv")
    conv_fl = sum(m["flops"] for m, ms in prof if m["kind"] == "conv")
    other_ms = sum(ms for m, ms in prof if m["kind"] != "conv")
    table = [{"op": i, "kind": m["kind"], "cin": m["cin"], "cout": m["cout"], "k": m["k"], "side": m["side"],
              "ms_per_frame": ms / n, "tflops": (m["flops"] * n / (ms * 1e-3) / 1e12) if m["flops"] else None}
             for i, (m, ms) in enumerate(prof)]
    def family(pred):
        sel = [(m, ms) for m, ms in prof if pred(m)]
        t = sum(ms for _, ms in sel)
        f = sum(m["flops"] for m, _ in sel)
        return {"launches": len(sel), "ms_per_frame": t / n, "flops_per_frame": f, "tflops": f * n / (t * 1e-3) / 1e12 if t else 0.0}
    fams = {"conv3_32_32_full_res": family(lambda m: m["kind"] == "conv" and m["k"] == 3 and m["cin"] == 32 and m["cout"] == 32 and m["side"] == V),
            "stem7_33_16": family(lambda m: m["kind"] == "conv" and m["k"] == 7),
            "other_tensor_convs": family(lambda m: m["kind"] == "conv" and m["k"] != 7 and not (m["k"] == 3 and m["cin"] == 32 and m["cout"] == 32 and m["side"] == V)),
            "tail_1x1_fused": family(lambda m: m["kind"] == "tail"),
            "deconv": family(lambda m: m["kind"] == "deconv"),
            "maxpool": family(lambda m: m["kind"] == "pool")}
    out["v2v"] = {"conv_ms_per_frame": conv_ms / n, "conv_flops_per_frame": conv_fl, "pool_deconv_ms_per_frame": other_ms / n,
                  "bound": "tensor", "launches_per_chunk": len(prof), "chunk_frames": n, "families": fams}
    if args.profile_ops:
        os.makedirs(os.path.dirname(os.path.abspath(args.profile_ops)), exist_ok=True)
        json.dump(table, open(args.profile_ops, "w"), indent=1)
    ms = timed(lambda: _lib.softargmax3d(logits, 1.0, True, net._axis, None, False))
    rec("softargmax", ms, J * V ** 3 * 4, "read (15,64,64,64) f32 logits once (online softmax)")
    ms_full = timed(lambda: _lib.softargmax3d(logits, 1.0, True, net._axis, None, True))
    rec("softmax_volume_write", max(ms_full - ms, 1e-6), 2 * J * V ** 3 * 4,
        "output #3 of the reference forward: read logits again + write (15,64,64,64) f32")
    ms = timed(lambda: _lib.features_upsample_pad(feat32, 1024, 128))
    rec("features_upsample_pad", ms, 32 * 1024 * 1280 * 4 + 64 * 64 * 32 * 4,
        "output #2 of the reference forward: write (32,1024,1280) f32")
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        ours_arm(a)
