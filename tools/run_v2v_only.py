"""Run only the V2V program (for ncu captures): python tools/run_v2v_only.py [frames] [repeats]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sceneego_b200 import _lib
from sceneego_b200.network.v2v import V2VModel
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
m = V2VModel(33, 15).eval().cuda()
pg = m.program(64, frames, torch.device("cuda"))
x = torch.randn(frames, 33, 64, 64, 64, device="cuda").abs()
_lib.pack_volume(x, pg.buffers[pg.in_buf], pg.lay_in)
out = torch.empty(frames, 15, 64, 64, 64, device="cuda")
for _ in range(reps):
    m.run_chunk(pg, frames, out)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
