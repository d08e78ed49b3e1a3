"""Time the marching conv (csrc/march.cu) at real size under tuning switches (tuning aid, not a test).
usage (GPU box): python tools/tune_march.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch, torch.nn as nn
from tests import util
from sceneego_b200 import _lib


def time_layer(cin, S, B, res, march, env, reps=5, xs=2, pair=2):
    for k in ("SCENEEGO_MARCH_STAGES", "SCENEEGO_MARCH_CTAS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    torch.manual_seed(0)
    conv = nn.Conv3d(cin, 32, 3, padding=1).cuda().eval()
    bn = nn.BatchNorm3d(32).cuda().eval()
    x = torch.randn(B, cin, S, S, S, device="cuda")
    r = torch.randn(B, 32, S, S, S, device="cuda") if res else None
    if march:
        util.run_single_op(x, conv, bn, relu=True, res=r, march=True)
    else:
        util.run_single_op(x, conv, bn, relu=True, res=r, xstack=xs, cta_pair=pair if cin >= 32 else 1)
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
    return ms / B * 1000, 2 * cin * 32 * 27 * S ** 3 * B / (ms * 1e-3) / 1e12


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    for cin, res in ((32, True), (32, False)):
        us, tf = time_layer(cin, 64, B, res, False, {})
        print(f"conv_tc   {cin}->32 res={int(res)}                         {us:7.1f} us/frame {tf:6.0f} TF", flush=True)
        envs = [{}, {"SCENEEGO_MARCH_STAGES": "2"}, {}, {"SCENEEGO_MARCH_STAGES": "2"}, {}, {"SCENEEGO_MARCH_STAGES": "2"}]
        for env in envs:
            us, tf = time_layer(cin, 64, B, res, True, env)
            print(f"march     {cin}->32 res={int(res)} {str(env):36s} {us:7.1f} us/frame {tf:6.0f} TF", flush=True)
