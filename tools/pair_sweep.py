"""CTA-pair tuning: single layers, cta_pair x window stages x weight chunking."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.tune_conv import time_layer
B = 32
layers = [("conv3 32->32 xs2", 32, 32, 3, 64, 1, 2), ("conv3 16->32 xs2", 16, 32, 3, 64, 1, 2), ("conv3 64->64 S32", 64, 64, 3, 32, 1, 1)]
for name, cin, cout, k, S, pad, xs in layers:
    for pair in (1, 2):
        for stages in (0, 2, 3, 4):
            for wchunk in (0, 3):
                for kk in ("SCENEEGO_WCHUNK",):
                    os.environ.pop(kk, None)
                os.environ["SCENEEGO_TEST_PAIR"] = str(pair)
                if wchunk: os.environ["SCENEEGO_WCHUNK"] = str(wchunk)
                try:
                    us, tf = time_layer(cin, cout, k, S, B if S >= 64 else B * 4, pad, xs, 0, stages)
                    print(f"{name:20s} pair={pair} stages={stages or 'auto'} wchunk={wchunk or 'auto'} {us:8.1f} us/frame {tf:7.0f} TF", flush=True)
                except Exception as e:
                    print(f"{name:20s} pair={pair} stages={stages} wchunk={wchunk} -- {str(e)[:90]}", flush=True)
