"""BASELINE.json configs[4]: isolated unprojection + scene voxelisation, 1024x1024 depth maps (~1M points),
64^3 / 96^3 / 128^3 grids.  Prints one JSON line per case with algorithmic bytes and achieved GB/s
(CUDA events, inputs larger than L2 or L2 flushed between repetitions).  Run on the GPU box:
    python tools/microbench_geometry.py [frames]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sceneego_b200 import _lib
from sceneego_b200.utils.fisheye.FishEyeCalibrated import FishEyeCameraCalibrated
from sceneego_b200.utils import synth
from tests import util

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cam = FishEyeCameraCalibrated(util.CALIB)
calib = cam.calib_struct(1280, 1024)
ray = cam.ray_table_device(1280, 1024, "cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
peaks = json.load(open(os.path.join(util.ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(util.ROOT, "MEASURED_PEAKS.json")) else {}
hbm = float(peaks.get("hbm_gbs", 6650.0))


def timed(fn, reps=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


feat = torch.relu(torch.randn(B, 64, 64, 32, device="cuda"))
depth = (torch.rand(B, 1024, 1024, device="cuda") * 4.0)
for V in (64, 96, 128):
    n = V ** 3
    px, grid = _lib.project_voxels(calib, V, 2.0, (1024, 1280), "cuda")
    grid = grid.reshape(-1, 2).contiguous()
    # f32 NCDHW output (the reference's tensor) and bf16 planar output (what V2V consumes)
    out32 = torch.empty(B, 32, V, V, V, device="cuda")
    ms = timed(lambda: _lib.unproject(feat, grid, None, V, 2.0, 1024, 1280, out32, None, None))
    bytes_f32 = B * (64 * 64 * 32 * 4 + 32 * n * 4)
    print(json.dumps({"case": "unproject f32 NCDHW", "V": V, "frames": B, "ms": ms, "algorithmic_MB": bytes_f32 / 1e6,
                      "GB_per_s": bytes_f32 / ms / 1e6, "frac_of_hbm_peak": bytes_f32 / ms / 1e6 / hbm}))
    lay = _lib.vol_layout(V, 1, B)
    outb = _lib.alloc_volume(lay, 32, "cuda")
    ms = timed(lambda: _lib.unproject(feat, grid, None, V, 2.0, 1024, 1280, None, outb, lay))
    bytes_bf = B * (64 * 64 * 32 * 4 + 32 * n * 2)
    print(json.dumps({"case": "unproject bf16 planar", "V": V, "frames": B, "ms": ms, "algorithmic_MB": bytes_bf / 1e6,
                      "GB_per_s": bytes_bf / ms / 1e6, "frac_of_hbm_peak": bytes_bf / ms / 1e6 / hbm}))
    ms = timed(lambda: _lib.unproject(feat, None, calib, V, 2.0, 1024, 1280, None, outb, lay))
    print(json.dumps({"case": "unproject bf16 planar, projection fused (no grid table)", "V": V, "frames": B, "ms": ms,
                      "algorithmic_MB": bytes_bf / 1e6, "GB_per_s": bytes_bf / ms / 1e6}))
    occ = torch.zeros(B, V, V, V, device="cuda")
    ms = timed(lambda: _lib.voxelize_depth(depth, ray, 1024, 1280, V, 2.0, occ, None, None))
    bytes_v = B * 1024 * 1024 * 4
    print(json.dumps({"case": "voxelize 1024x1024 depth (1M points/frame)", "V": V, "frames": B, "ms": ms,
                      "algorithmic_MB": bytes_v / 1e6, "GB_per_s": bytes_v / ms / 1e6, "Mpoints_per_s": B * 1024 * 1024 / ms / 1e3,
                      "occupied_voxels_frame0": int(occ[0].sum().item())}))
