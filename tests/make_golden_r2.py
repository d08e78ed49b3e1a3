"""Round-2 golden vectors, again produced by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):
    OPENCV_IO_ENABLE_OPENEXR=1 python tests/make_golden_r2.py [--b64-only|--v128-only|--v64-only|--dataset-only|--forward-only]

Adds to tests/golden/:
  stage_v64_logits.npz   the reference's OWN V2V logits (forward hook on `volume_net`) for the three B=2 stage
                         configurations of stage_v64.npz -- the logits stored there came from the oracle
  stage_v64_b64.npz      BASELINE configs[1] inputs (the bench's 64-frame batch at rank 0): frames 0/21/42/63 run
                         through the reference as one B=4 batch (frames are independent; the reference has no
                         cross-frame state)
  stage_v128.npz         BASELINE configs[3]: V=128, B=1, whole stage (keypoints, sub-sampled logits and softmax)
  voxel_dataset.npz      dataset/real_depth_utils.depth_map_to_voxel (the `voxel_output=True` path), V=64 / 128
  forward_v64.npz        the whole reference forward from `images` (backbone included), B = 2
While generating, the oracle restatement is asserted against the reference in full.
"""
import json
import os
import sys

os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
import numpy as np  # noqa: E402
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import make_golden as mg  # noqa: E402
from oracle import sceneego_oracle as orc  # noqa: E402
from sceneego_b200.utils import synth  # noqa: E402

OUT = mg.OUT
B64_FRAMES = [0, 21, 42, 63]
LOGIT_STRIDE_V64 = 257
LOGIT_STRIDE_V128 = 2053


def _report_update(**kw):
    p = os.path.join(OUT, "report.json")
    r = json.load(open(p))
    r.update(kw)
    json.dump(r, open(p, "w"), indent=1)


def _run_reference(net, feat, depth):
    """Reference forward with the backbone stubbed to return `feat`; returns (kp, softmaxed volumes, V2V logits)."""
    grabbed = []
    h = net.volume_net.register_forward_hook(lambda m, i, o: grabbed.append(o.detach()))
    net.backbone.forward = lambda images, _f=feat: (None, _f[: images.shape[0]])
    b = feat.shape[0]
    try:
        with torch.no_grad():
            kp, _, vol, _ = net(torch.zeros(b, 3, 256, 256), net.grid_coord_proj_batch, net.coord_volumes,
                                depth_map_batch=depth)
    finally:
        h.remove()
    return kp, vol, grabbed[0]


def _load(net, shapes, mode, scale=1.0):
    sd = synth.synthetic_state_dict(shapes, seed=0, mode=mode, logit_scale=scale)
    full = net.state_dict()
    full.update(sd)
    net.load_state_dict(full, strict=True)
    return sd


def _net(config, Net, V, batch):
    config.opt.batch_size = batch
    config.model.volume_size = V
    torch.manual_seed(0)
    net = Net(config, device="cpu").eval()
    shapes = [(k, tuple(v.shape)) for k, v in net.state_dict().items() if not k.startswith("backbone.")]
    return net, shapes


def make_v64_logits(config, Net, calib_path):
    net, shapes = _net(config, Net, 64, 2)
    tabs = orc.StageTables(calib_path, 64, 2.0)
    out, rep = {}, {}
    old = np.load(os.path.join(OUT, "stage_v64.npz"))
    for mode, scale in (("default", 1.0), ("random_bn", 1.0), ("random_bn", 30.0)):
        sd = _load(net, shapes, mode, scale)
        feat = synth.synthetic_features(2)
        depth = torch.cat([synth.synthetic_depth_room(1, tabs.ray), synth.synthetic_depth_uniform(1)])
        kp, vol, logits = _run_reference(net, feat, depth)
        tag = f"{mode}_s{int(scale)}"
        assert np.array_equal(kp.numpy(), old[f"kp_{tag}"]), "stage_v64.npz was generated from other inputs"
        with torch.no_grad():
            _, _, _, inter = orc.stage_forward(tabs, sd, feat, depth_batch=depth, return_intermediates=True)
        e = ((inter["logits"] - logits).norm() / logits.norm()).item()
        assert e <= 1e-5, f"oracle logits differ from the reference's: {e}"
        out[f"logits_{tag}"] = logits.reshape(2, 15, -1)[:, :, ::LOGIT_STRIDE_V64].numpy()
        out[f"logit_range_{tag}"] = np.array([logits.min().item(), logits.max().item(), logits.std().item()])
        rep[f"stage_{tag}_logits_relfro_oracle_vs_ref"] = e
    np.savez_compressed(os.path.join(OUT, "stage_v64_logits.npz"), **out)
    _report_update(**rep)


def make_b64(config, Net, calib_path):
    net, shapes = _net(config, Net, 64, 4)
    tabs = orc.StageTables(calib_path, 64, 2.0)
    sd = _load(net, shapes, "random_bn")
    feat = synth.synthetic_features(64, seed=1234)[B64_FRAMES].contiguous()          # bench.py rank-0 inputs
    depth = synth.synthetic_depth_room(64, tabs.ray, seed=7)[B64_FRAMES].contiguous()
    kp, vol, logits = _run_reference(net, feat, depth)
    with torch.no_grad():
        kp_o, _, _, inter = orc.stage_forward(tabs, sd, feat, depth_batch=depth, return_intermediates=True)
    e = orc.mpjpe(kp_o.numpy(), kp.numpy())
    assert e <= 5e-5, f"B=64 sample: oracle keypoints differ {e}"
    np.savez_compressed(os.path.join(OUT, "stage_v64_b64.npz"), frames=np.array(B64_FRAMES), kp=kp.numpy(),
                        logits=logits.reshape(4, 15, -1)[:, :, ::LOGIT_STRIDE_V64].numpy(),
                        softmax=vol.reshape(4, 15, -1)[:, :, ::LOGIT_STRIDE_V64].numpy(),
                        occupied=np.array([int(s.sum()) for s in inter["scene"]]))
    _report_update(stage_b64_sample_mpjpe_oracle_vs_ref_m=e)


def make_v128(config, Net, calib_path):
    net, shapes = _net(config, Net, 128, 1)
    tabs = orc.StageTables(calib_path, 128, 2.0)
    out, rep = {}, {}
    feat = synth.synthetic_features(1, seed=77)
    depth = synth.synthetic_depth_room(1, tabs.ray, seed=78)
    for mode, scale in (("default", 1.0), ("random_bn", 1.0), ("random_bn", 30.0)):
        sd = _load(net, shapes, mode, scale)
        kp, vol, logits = _run_reference(net, feat, depth)
        tag = f"{mode}_s{int(scale)}"
        if scale == 1.0:                 # the oracle restatement at this size, once per weight mode
            with torch.no_grad():
                kp_o, _, _, inter = orc.stage_forward(tabs, sd, feat, depth_batch=depth, return_intermediates=True)
            e = orc.mpjpe(kp_o.numpy(), kp.numpy())
            el = ((inter["logits"] - logits).norm() / logits.norm()).item()
            assert e <= 1e-4 and el <= 1e-5, f"V=128: oracle differs from the reference: {e} m, logits {el}"
            rep[f"stage_v128_{tag}_mpjpe_oracle_vs_ref_m"] = e
            rep[f"stage_v128_{tag}_logits_relfro_oracle_vs_ref"] = el
            out[f"occupied_{tag}"] = np.array(int(inter["scene"][0].sum()))
        out[f"kp_{tag}"] = kp.numpy()
        out[f"logits_{tag}"] = logits.reshape(1, 15, -1)[:, :, ::LOGIT_STRIDE_V128].numpy()
        out[f"softmax_{tag}"] = vol.reshape(1, 15, -1)[:, :, ::LOGIT_STRIDE_V128].numpy()
        out[f"logit_range_{tag}"] = np.array([logits.min().item(), logits.max().item(), logits.std().item()])
    np.savez_compressed(os.path.join(OUT, "stage_v128.npz"), **out)
    _report_update(**rep)


def make_forward_golden(config, Net, calib_path):
    """The WHOLE reference forward, backbone included (network/voxel_net_depth.py:224-275 from `images`): seeded
    synthetic weights for all 699 state-dict entries, B = 2 images.  Golden for forward() with and without the
    backbone hand-off."""
    net, _ = _net(config, Net, 64, 2)
    tabs = orc.StageTables(calib_path, 64, 2.0)
    shapes = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    sd = synth.synthetic_state_dict(shapes, seed=0, mode="random_bn")
    net.load_state_dict(sd, strict=True)
    img = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(3))
    depth = synth.synthetic_depth_room(2, tabs.ray, seed=4)
    grabbed = []
    h = net.volume_net.register_forward_hook(lambda m, i, o: grabbed.append(o.detach()))
    with torch.no_grad():
        kp, feats, vol, _ = net(img, net.grid_coord_proj_batch, net.coord_volumes, depth_map_batch=depth)
        _, bb_feat = net.backbone(img)
    h.remove()
    logits = grabbed[0]
    # the oracle's post-backbone restatement from the reference backbone's features agrees
    with torch.no_grad():
        kp_o, _, _ = orc.stage_forward(tabs, {k: v for k, v in sd.items() if not k.startswith("backbone.")}, bb_feat,
                                       depth_batch=depth)
    e = orc.mpjpe(kp_o.numpy(), kp.numpy())
    assert e <= 5e-5, f"forward golden: oracle stage differs from the reference forward: {e}"
    np.savez_compressed(os.path.join(OUT, "forward_v64.npz"), kp=kp.numpy(),
                        logits=logits.reshape(2, 15, -1)[:, :, ::LOGIT_STRIDE_V64].numpy(),
                        logit_range=np.array([logits.min().item(), logits.max().item(), logits.std().item()]),
                        backbone_features=bb_feat[:, ::16, ::4, ::4].numpy(),
                        backbone_features_absmax=np.array(bb_feat.abs().max().item()))
    _report_update(forward_v64_mpjpe_oracle_stage_vs_ref_m=e)


def make_dataset_voxel(calib_path):
    """dataset/real_depth_utils.py:29-60 imported unmodified; its last line `voxel_torch[idx.T] = 1` is evaluated with
    the torch-1.13.1 tuple rule like the network's by patching ONLY that function."""
    sys.path.insert(0, mg.REF)
    np.float = float
    np.round_ = np.round
    import importlib
    rdu = importlib.import_module("dataset.real_depth_utils")

    def pinned(point_cloud, cuboid_side, volume_size):
        # dataset/real_depth_utils.py:45-60 line for line, tuple-indexed scatter
        from copy import copy
        p = copy(point_cloud)
        p[:, 0] = (p[:, 0] + cuboid_side / 2) * volume_size / cuboid_side
        p[:, 1] = (p[:, 1] + cuboid_side / 2) * volume_size / cuboid_side
        p[:, 2] = (p[:, 2]) * volume_size / cuboid_side
        p = np.round_(p)
        good = np.all(np.logical_and(volume_size - 1 >= p, p >= 0), axis=1)
        p = p[good]
        voxel = torch.zeros(size=(volume_size, volume_size, volume_size))
        voxel[tuple(torch.from_numpy(p.T).long())] = 1
        return voxel
    rdu.point_cloud_to_voxel_pytorch = pinned
    tabs = orc.StageTables(calib_path, 64, 2.0)
    out, rep = {}, {}
    maps = {}
    for name in ("img_001000", "img_001796", "img_002376"):
        _, maps[name] = mg.load_demo_depth(name)
    maps["room"] = synth.synthetic_depth_room(1, tabs.ray)[0].numpy()
    maps["uniform"] = synth.synthetic_depth_uniform(1)[0].numpy()
    for name, d in maps.items():
        for V in (64, 128):
            ref = rdu.depth_map_to_voxel(tabs.ray, d, 2.0, V).numpy()
            got = orc.voxelize_depth_dataset(d, tabs.ray, V, 2.0)
            assert np.array_equal(ref, got), f"dataset occupancy differs {name} V={V}"
            net_like = orc.voxelize_depth(d, tabs.ray, V, 2.0)
            out[f"{name}_v{V}"] = np.packbits(ref.astype(np.uint8).reshape(-1))
            rep[f"occ_dataset_{name}_v{V}"] = int(ref.sum())
            rep[f"occ_dataset_vs_network_differing_voxels_{name}_v{V}"] = int((ref != net_like).sum())
    np.savez_compressed(os.path.join(OUT, "voxel_dataset.npz"), **out)
    _report_update(**rep)


def main():
    calib_path = os.path.join(ROOT, "sceneego_b200", "data", "fisheye.calibration_05_08.json")
    only = [a for a in sys.argv[1:] if a.endswith("-only")]
    if not only or "--dataset-only" in only:
        make_dataset_voxel(calib_path)
        if only:
            return
    config, Net, cwd = mg.import_reference()        # leaves the cwd at /root/reference (relative calibration path)
    if not only or "--v64-only" in only:
        make_v64_logits(config, Net, calib_path)
    if not only or "--b64-only" in only:
        make_b64(config, Net, calib_path)
    if not only or "--v128-only" in only:
        make_v128(config, Net, calib_path)
    if not only or "--forward-only" in only:
        make_forward_golden(config, Net, calib_path)
    os.chdir(cwd)
    print(open(os.path.join(OUT, "report.json")).read())


if __name__ == "__main__":
    main()
