"""Where does the bf16 error of the V2V activation chain come from?  

Emulates the storage roundings of the CUDA path in plain torch fp32 on the GPU (TF32 off): the reference network is
evaluated with a rounding function applied to chosen tensors (weights, every activation that the kernels store in
HBM, only those of some levels, ...), and the resulting poses / logits are compared with the unrounded fp32
evaluation.  Checker-side tool (imports oracle/); prints a table and writes gpurun_out/bf16_attribution.json.

    python tools/bf16_attribution.py [--scale 30] [--mode random_bn]
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import sceneego_oracle as orc  # noqa: E402
from sceneego_b200.utils import synth  # noqa: E402
from tests import util  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def rnd(x, kind):
    if kind == "bf16":
        return x.to(torch.bfloat16).float()
    if kind == "fp16":
        return x.to(torch.float16).float()
    return x


class Net:
    """Functional V2V (network/v2v.py:104-170) with BN folded into the conv like the kernels do, and a rounding policy:
    policy(tag, level) -> 'bf16' | 'fp16' | None for every tensor the CUDA path stores in HBM."""

    def __init__(self, sd, wkind, policy):
        self.sd, self.wkind, self.policy = sd, wkind, policy

    def fold(self, p_conv, p_bn, transposed=False):
        w, b = self.sd[p_conv + ".weight"].double(), self.sd[p_conv + ".bias"].double()
        if p_bn is not None:
            g, bt = self.sd[p_bn + ".weight"].double(), self.sd[p_bn + ".bias"].double()
            mu, var = self.sd[p_bn + ".running_mean"].double(), self.sd[p_bn + ".running_var"].double()
            sc = g / torch.sqrt(var + 1e-5)
            shape = (1, -1, 1, 1, 1) if transposed else (-1, 1, 1, 1, 1)
            w = w * sc.view(shape)
            b = (b - mu) * sc + bt
        return rnd(w.float(), self.wkind), b.float()

    def q(self, x, tag, level):
        return rnd(x, self.policy(tag, level))

    def basic(self, p, x, level, tag):
        w, b = self.fold(p + ".block.0", p + ".block.1")
        return self.q(F.relu(F.conv3d(x, w, b, padding=(w.shape[-1] - 1) // 2)), tag, level)

    def res(self, p, x, level):
        w1, b1 = self.fold(p + ".res_branch.0", p + ".res_branch.1")
        t = self.q(F.relu(F.conv3d(x, w1, b1, padding=1)), "res_mid", level)
        w2, b2 = self.fold(p + ".res_branch.3", p + ".res_branch.4")
        r = F.conv3d(t, w2, b2, padding=1)
        if (p + ".skip_con.0.weight") in self.sd:
            ws, bs = self.fold(p + ".skip_con.0", p + ".skip_con.1")
            x = F.conv3d(x, ws, bs)                       # fused into the same accumulator in the kernels: not stored
        return self.q(F.relu(r + x), "res_out", level)

    def up(self, p, x, skip, level):
        w, b = self.fold(p + ".block.0", p + ".block.1", transposed=True)
        y = F.relu(F.conv_transpose3d(x, w, b, stride=2))
        return self.q(y + skip, "up_out", level)           # the kernels add the skip in the epilogue, one store

    def forward(self, x):
        x = self.q(x, "input", 0)
        x = self.basic("front_layers.0", x, 0, "stem_out")
        for i in (1, 2, 3):
            x = self.res(f"front_layers.{i}", x, 0)
        e = "encoder_decoder."
        skips = []
        for lvl in range(1, 6):
            skips.append(self.res(f"{e}skip_res{lvl}", x, lvl - 1))
            x = F.max_pool3d(x, 2, 2)
            x = self.res(f"{e}encoder_res{lvl}", x, lvl)
        x = self.res(e + "mid_res", x, 5)
        for lvl in range(5, 0, -1):
            x = self.res(f"{e}decoder_res{lvl}", x, lvl)
            x = self.up(f"{e}decoder_upsample{lvl}", x, skips[lvl - 1], lvl - 1)
        x = self.res("back_layers.0", x, 0)
        x = self.basic("back_layers.1", x, 0, "tail_hidden")
        x = self.basic("back_layers.2", x, 0, "tail_hidden")
        w, b = self.fold("output_layer", None)
        return F.conv3d(x, w, b)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=30.0)
    ap.add_argument("--mode", default="random_bn")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bf16_attribution.json"))
    a = ap.parse_args()
    dev = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    tabs = orc.StageTables(util.CALIB, 64, 2.0)
    sd_full = synth.synthetic_state_dict(util.stage_shapes(), seed=0, mode=a.mode, logit_scale=a.scale)
    feat = synth.synthetic_features(2)
    depth = torch.cat([synth.synthetic_depth_room(1, tabs.ray), synth.synthetic_depth_uniform(1)])
    with torch.no_grad():
        feats = orc.process_features(feat, sd_full["process_features.0.weight"], sd_full["process_features.0.bias"])
        lifted = orc.unproject(feats, tabs.grid.unsqueeze(0).expand(2, -1, -1, -1), 64)
        scene = torch.stack([torch.from_numpy(orc.voxelize_depth(d.numpy(), tabs.ray, 64, 2.0)) for d in depth])
    x = torch.cat([lifted, scene.unsqueeze(1)], 1).to(dev)
    sd = {k[len("volume_net."):]: v.to(dev) for k, v in sd_full.items() if k.startswith("volume_net.")}
    coord = tabs.coord_volume.to(dev).unsqueeze(0).expand(2, -1, -1, -1, -1)
    golden = util.golden("stage_v64.npz")
    tag = f"{a.mode}_s{int(a.scale)}"
    kp_gold = golden[f"kp_{tag}"] if f"kp_{tag}" in golden else None

    def run(wkind, policy):
        with torch.no_grad():
            lg = Net(sd, wkind, policy).forward(x)
            kp, _ = orc.soft_argmax(lg.double(), coord.double())
        return lg, kp.float().cpu().numpy()

    base_lg, base_kp = run(None, lambda t, l: None)
    rows = []
    if kp_gold is not None:
        rows.append(("fp32 torch (cuDNN) vs reference golden (CPU)", None, orc.mpjpe(base_kp, kp_gold) * 1000))
    variants = [
        ("weights bf16 only", "bf16", lambda t, l: None),
        ("input bf16 only", None, lambda t, l: "bf16" if t == "input" else None),
        ("all stored activations bf16, weights fp32", None, lambda t, l: "bf16"),
        ("everything bf16 (= the CUDA path's storage)", "bf16", lambda t, l: "bf16"),
        ("everything bf16 except the tail's hidden activations", "bf16", lambda t, l: None if t == "tail_hidden" else "bf16"),
        ("everything bf16 except level-0 (full-res) activations", "bf16", lambda t, l: None if l == 0 else "bf16"),
        ("everything bf16 except levels >= 1", "bf16", lambda t, l: "bf16" if l == 0 else None),
        ("everything bf16 except skip-connection sums (up_out)", "bf16", lambda t, l: None if t == "up_out" else "bf16"),
        ("everything bf16 except Res block outputs", "bf16", lambda t, l: None if t == "res_out" else "bf16"),
        ("everything bf16 except Res block mids", "bf16", lambda t, l: None if t == "res_mid" else "bf16"),
        ("weights bf16, activations fp16", "bf16", lambda t, l: "fp16"),
        ("everything fp16", "fp16", lambda t, l: "fp16"),
        ("weights fp16, activations bf16", "fp16", lambda t, l: "bf16"),
    ]
    for name, wk, pol in variants:
        lg, kp = run(wk, pol)
        rel = ((lg - base_lg).norm() / base_lg.norm()).item()
        mx = ((lg - base_lg).abs().max() / (base_lg.max() - base_lg.min())).item()
        rows.append((name, (rel, mx), orc.mpjpe(kp, base_kp) * 1000))
    print(f"V2V storage-rounding attribution, weights {a.mode}, logit scale x{a.scale:g} (B=2, V=64)")
    for name, lgerr, mm in rows:
        extra = "" if lgerr is None else f"  logits rel-Frobenius {lgerr[0]:.3e}  max-abs/range {lgerr[1]:.3e}"
        print(f"  {name:62s} MPJPE {mm:9.4f} mm{extra}")
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump([{"variant": n, "logits": e, "mpjpe_mm": m} for n, e, m in rows], open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
