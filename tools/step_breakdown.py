"""Where the step time goes (tuning aid): full lift vs lift without output #2 / #3 vs the V2V program alone."""
import os, sys, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sceneego_b200.network.voxel_net_depth import VoxelNetwork_depth
from sceneego_b200.utils import synth
from tests import util

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
with contextlib.redirect_stdout(io.StringIO()):
    net = VoxelNetwork_depth(util.load_config(batch_size=B), device="cuda:0", v2v_chunk=64).eval()
sd = synth.synthetic_state_dict(util.stage_shapes(), seed=0, mode="random_bn")
full = net.state_dict(); full.update(sd); net.load_state_dict(full, strict=True)
feat = synth.synthetic_features(B).cuda()
depth = synth.synthetic_depth_room(B, net.ray).cuda()


def timed(fn, n=8):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def lift():
    with torch.no_grad():
        net.lift(feat, net.grid_coord_proj_batch, net.coord_volumes, depth_map_batch=depth)

print(f"full lift                      {timed(lift):8.3f} ms / {B} frames")
net.materialize_features = False
print(f"  without output #2 (features) {timed(lift):8.3f} ms")
net.materialize_volumes = False
print(f"  without #2 and #3 (softmax)  {timed(lift):8.3f} ms")
vn = net.volume_net
pg = vn.program(64, min(64, B), torch.device("cuda", 0))
logits = torch.empty(min(64, B), 15, 64, 64, 64, device="cuda")
print(f"V2V program alone              {timed(lambda: vn.run_chunk(pg, min(64, B), logits)):8.3f} ms")
ops = vn.profile_chunk(pg, min(64, B), logits)
print(f"  sum of per-op CUDA-event times {sum(ms for _, ms in ops):8.3f} ms")
