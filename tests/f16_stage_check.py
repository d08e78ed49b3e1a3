"""Run with SCENEEGO_ACT_DTYPE=f16 (by tests/test_gpu_f16.py, in a subprocess: the activation dtype is one per process):
the whole stage with fp16 activation / weight storage against the unmodified reference's goldens; prints one JSON line."""
import json
import sys

import numpy as np
import torch

from oracle import sceneego_oracle as orc
from sceneego_b200 import _lib
from sceneego_b200.network.voxel_net_depth import VoxelNetwork_depth
from sceneego_b200.utils import synth
from tests import util


def main():
    assert _lib.act_dtype_name() == "f16" and _lib.load_library().sceneego_act_dtype() == 1
    tabs = orc.StageTables(util.CALIB, 64, 2.0)
    torch.manual_seed(0)
    net = VoxelNetwork_depth(util.load_config(batch_size=4), device="cuda").eval()
    net.keep_logits = True
    g, gl = util.golden("stage_v64.npz"), util.golden("stage_v64_logits.npz")
    out = {"dtype": _lib.act_dtype_name()}
    feat = synth.synthetic_features(2).cuda()
    depth = torch.cat([synth.synthetic_depth_room(1, tabs.ray), synth.synthetic_depth_uniform(1)]).cuda()
    for mode, scale in (("default", 1.0), ("random_bn", 1.0), ("random_bn", 30.0)):
        sd = synth.synthetic_state_dict(util.stage_shapes(), seed=0, mode=mode, logit_scale=scale)
        full = net.state_dict()
        full.update(sd)
        net.load_state_dict(full, strict=True)
        with torch.no_grad():
            kp, _, vol, _ = net.lift(feat, net.grid_coord_proj_batch, net.coord_volumes, depth_map_batch=depth)
        tag = f"{mode}_s{int(scale)}"
        ref = gl[f"logits_{tag}"].astype(np.float64)
        mine = net.last_logits.reshape(2, 15, -1)[:, :, ::257].cpu().numpy().astype(np.float64)
        rng = gl[f"logit_range_{tag}"]
        out[tag] = {"mpjpe_mm": orc.mpjpe(kp.cpu().numpy(), g[f"kp_{tag}"]) * 1000.0,
                    "logits_rel_fro": float(np.linalg.norm(mine - ref) / np.linalg.norm(ref)),
                    "logits_max_over_range": float(np.abs(mine - ref).max() / float(rng[1] - rng[0]))}
    # occupancy stays bit-exact (1.0 is 1.0 in either format) and a saturating store never produces inf
    pg = net.volume_net.program(64, 2, torch.device("cuda", 0))
    occ = _lib.unpack_volume(pg.buffers[pg.in_buf], pg.lay_in, 2, 33)[:, 32]
    out["occupancy_bit_exact"] = bool(np.array_equal(occ[0].cpu().numpy(), orc.voxelize_depth(depth[0].cpu().numpy(), tabs.ray, 64, 2.0)))
    big = torch.full((1, 8, 4, 4, 4), 1.0e6, device="cuda")
    lay = _lib.vol_layout(4, 1, 1)
    buf = _lib.alloc_volume(lay, 8, "cuda")
    _lib.pack_volume(big, buf, lay)
    out["saturates_at"] = float(_lib.unpack_volume(buf, lay, 1, 8).max().item())
    print("F16CHECK " + json.dumps(out))


if __name__ == "__main__":
    sys.exit(main())
