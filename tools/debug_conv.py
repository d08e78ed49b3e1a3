"""Pipeline experiments on single conv layers (SCENEEGO_DEBUG switches in conv_tc_kernel):
1 = epilogue body off, 2 = weight re-streaming off, 4 = window streaming off, 8 = MMAs off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.tune_conv import time_layer
B = 16
layers = [("conv3 32->32 xs2", 32, 32, 3, 64, 1, 2), ("stem 33->16 k7 xs4", 33, 16, 7, 64, 3, 4), ("conv3 64->64 S32", 64, 64, 3, 32, 1, 1),
          ("conv3 128->128 S16", 128, 128, 3, 16, 1, 1)]
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
one = len(sys.argv) > 1 and sys.argv[1] == "one"
for name, cin, cout, k, S, pad, xs in layers:
    if one:
        us, tf = time_layer(cin, cout, k, S, B if S >= 64 else B * 4, pad, xs, 0, 0)
        print(f"{name:22s} DEBUG={os.environ.get('SCENEEGO_DEBUG')} {us:8.1f} us/frame {tf:7.0f} TF", flush=True)
        continue
    for env in ([{}, {"SCENEEGO_WCHUNK": "9"}, {"SCENEEGO_WCHUNK": "7"}, {"SCENEEGO_DEBUG": "6"}] if quick else [{}, {"SCENEEGO_DEBUG": "1"}, {"SCENEEGO_DEBUG": "2"}, {"SCENEEGO_DEBUG": "4"}, {"SCENEEGO_DEBUG": "6"}, {"SCENEEGO_DEBUG": "7"},
                {"SCENEEGO_DEBUG": "8"}, {"SCENEEGO_DEBUG": "9"}, {"SCENEEGO_WSLOTS": "8"}, {"SCENEEGO_WSLOTS": "6"}, {"SCENEEGO_WCHUNK": "1", "SCENEEGO_WSLOTS": "8"},
                {"SCENEEGO_WCHUNK": "9"}, {"SCENEEGO_WCHUNK": "7"}]):
        for kk in ("SCENEEGO_DEBUG", "SCENEEGO_WSLOTS", "SCENEEGO_WCHUNK"):
            os.environ.pop(kk, None)
        os.environ.update(env)
        try:
            us, tf = time_layer(cin, cout, k, S, B if S >= 64 else B * 4, pad, xs, 0, 0)
            print(f"{name:22s} {str(env):60s} {us:8.1f} us/frame {tf:7.0f} TF", flush=True)
        except Exception as e:
            print(f"{name:22s} {str(env):60s} -- {str(e)[:80]}", flush=True)
