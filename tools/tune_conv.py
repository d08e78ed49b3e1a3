"""Time single V2V conv layers at real size under planner overrides (tuning aid, not a test).
usage: python tools/tune_conv.py   (run on the GPU box; prints a table)"""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from tests import util

def time_layer(cin, cout, k, S, B, pad, xs, tiles, stages, reps=5):
    os.environ.pop("SCENEEGO_TILES", None); os.environ.pop("SCENEEGO_STAGES", None)
    if tiles: os.environ["SCENEEGO_TILES"] = str(tiles)
    if stages: os.environ["SCENEEGO_STAGES"] = str(stages)
    torch.manual_seed(0)
    conv = nn.Conv3d(cin, cout, k, padding=k // 2).cuda().eval()
    bn = nn.BatchNorm3d(cout).cuda().eval()
    x = torch.randn(B, cin, S, S, S, device="cuda")
    import ctypes as C
    from sceneego_b200 import _lib
    from sceneego_b200.network.v2v import _Program, _pad16
    # build once, launch repeatedly
    got, dst, lay = util.run_single_op(x, conv, bn, relu=True, pad_src=pad, xstack=xs, cta_pair=int(os.environ.get('SCENEEGO_TEST_PAIR', '1')))
    pg = util.LAST_PROGRAM
    lib = _lib.load_library()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        lib.sceneego_v2v_run(pg.op_array, 1, pg.buf_ptrs, C.c_void_p(pg.blob.data_ptr()), B, _lib._stream())
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    ms = ts[len(ts) // 2]
    fl = 2 * cin * cout * k ** 3 * S ** 3 * B
    return ms / B * 1000, fl / (ms * 1e-3) / 1e12

if __name__ == "__main__":
    B = 16
    layers = [("stem 33->16 k7", 33, 16, 7, 64, 3), ("conv3 32->32", 32, 32, 3, 64, 1), ("conv3 16->32", 16, 32, 3, 64, 1),
              ("conv3 64->64 S32", 64, 64, 3, 32, 1), ("conv1 32->32", 32, 32, 1, 64, 1)]
    only = sys.argv[1:] 
    for name, cin, cout, k, S, pad in layers:
        if only and not any(o in name for o in only): continue
        for xs in ((1, 2, 4) if cout <= 32 and k > 1 else (1, 2) if k > 1 else (1,)):
            if xs * ((cout + 15) // 16 * 16) > 256: continue
            for tiles, stages in [(0, 0), (8, 2), (4, 2), (4, 3), (4, 4), (2, 3), (2, 4), (2, 2)]:
                try:
                    us, tf = time_layer(cin, cout, k, S, B, pad, xs, tiles, stages)
                    print(f"{name:18s} xs={xs} tiles={tiles or 'auto':>4} stages={stages or 'auto':>4}  {us:8.1f} us/frame  {tf:7.0f} TF", flush=True)
                except Exception as e:
                    print(f"{name:18s} xs={xs} tiles={tiles} stages={stages}  -- {str(e)[:70]}", flush=True)
