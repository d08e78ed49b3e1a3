"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):
    OPENCV_IO_ENABLE_OPENEXR=1 python tests/make_golden.py

The reference is imported untouched with four oracle-side shims (SURVEY.md
section 8c): np.float, np.round_, an EasyDict stand-in, and the PyTorch-1.13.1
tuple-index meaning of `voxel[idx.T] = 1` (network/voxel_net_depth.py:221).
While generating, every oracle function is compared IN FULL against the
reference; the committed .npz files hold (sub)samples so that the GPU box,
which has no /root/reference, can re-check both the oracle and the CUDA path.
"""
import os
import sys
import types
import json

os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
import cv2  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import sceneego_oracle as orc  # noqa: E402
from sceneego_b200.utils import synth  # noqa: E402


def import_reference():
    np.float = float
    np.round_ = np.round

    class EasyDict(dict):
        def __init__(self, d=None, **kw):
            super().__init__()
            for k, v in dict(d or {}, **kw).items():
                setattr(self, k, v)

        def __setattr__(self, k, v):
            if isinstance(v, dict) and not isinstance(v, EasyDict):
                v = EasyDict(v)
            super().__setattr__(k, v)
            super().__setitem__(k, v)
        __setitem__ = __setattr__

    m = types.ModuleType("easydict")
    m.EasyDict = EasyDict
    sys.modules["easydict"] = m
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)
    from utils import cfg
    from network.voxel_net_depth import VoxelNetwork_depth

    class Pinned(VoxelNetwork_depth):
        # identical to network/voxel_net_depth.py:207-222 except the last index
        # expression, which is evaluated with the torch-1.13.1 tuple rule.
        def point_cloud_to_voxel_numpy(self, point_cloud):
            from copy import copy
            p = copy(point_cloud)
            p[:, 0] = (p[:, 0] + self.cuboid_side / 2) * self.volume_size / self.cuboid_side
            p[:, 1] = (p[:, 1] + self.cuboid_side / 2) * self.volume_size / self.cuboid_side
            p[:, 2] = (p[:, 2]) * self.volume_size / self.cuboid_side
            p = np.round_(p)
            good = np.logical_and(self.volume_size - 1 >= p, p >= 0)
            good = np.all(good, axis=1)
            p = p[good]
            voxel = torch.zeros(size=(self.volume_size,) * 3).to(self.device)
            voxel[tuple(torch.from_numpy(p.T).long())] = 1
            return voxel

    config = cfg.load_config("experiments/sceneego/test/sceneego.yaml")
    return config, Pinned, cwd


def load_demo_depth(name):
    """dataset/demo_dataset.py:86-91 preprocessing, with cv2 like the reference."""
    d = cv2.imread(f"{REF}/data/demo/depths/{name}.jpg.exr", cv2.IMREAD_ANYCOLOR | cv2.IMREAD_ANYDEPTH)
    raw = d.copy()
    if d.shape[0] != 1024 or d.shape[1] != 1280:
        d = cv2.resize(d, (1280, 1024), interpolation=cv2.INTER_NEAREST)
    if d.ndim == 3:
        d = d[:, :, 0]
    d[d > 10] = 10
    return raw, d


NEAREST_SIZES = [26, 36, 52, 72, 98, 104, 144, 186, 198, 256, 300, 320, 333, 512, 517, 600, 640, 700, 1000, 1024, 1280,
                 1500, 2048, 2560]


def make_nearest_index_golden():
    """Index maps of cv2.resize(INTER_NEAREST) itself (the library the reference calls at
    network/voxel_net_depth.py:197 and dataset/demo_dataset.py:88) for source sizes -> 1024 / 1280, including the
    sizes where the exact-rational rule differs; also asserts the oracle's rule on ~3500 pairs."""
    from oracle import sceneego_oracle as orc

    def cv_idx(n_src, n_dst):
        img = np.arange(n_src, dtype=np.float32)[None, :].repeat(2, 0)
        return cv2.resize(img, (n_dst, 2), interpolation=cv2.INTER_NEAREST)[0].astype(np.int64)
    for ns in list(range(3, 1400)) + [1500, 2048, 2560, 3000]:
        for nd in (1024, 1280):
            assert np.array_equal(cv_idx(ns, nd), orc.nearest_index(ns, nd)), (ns, nd)
    out = {}
    for ns in NEAREST_SIZES:
        for nd in (1024, 1280):
            out[f"{ns}_{nd}"] = cv_idx(ns, nd).astype(np.int16)
    np.savez_compressed(os.path.join(OUT, "nearest_index.npz"), **out)


def make_intersection_golden():
    """with_intersection=true (network/voxel_net_depth.py:66-69,257-260): 65-channel V2V input
    cat([volumes, volumes * scene, scene]).  The UNMODIFIED reference forward (B=1, V=64, CPU) on seeded synthetic
    weights / features / depth; asserts oracle == reference and stores the reference's outputs."""
    from oracle import sceneego_oracle as orc
    from sceneego_b200.utils import synth
    config, Net, cwd = import_reference()
    config.opt.batch_size = 1
    config.model.with_intersection = True
    torch.manual_seed(0)
    net = Net(config, device="cpu").eval()
    shapes = [(k, tuple(v.shape)) for k, v in net.state_dict().items() if not k.startswith("backbone.")]
    assert dict(shapes)["volume_net.front_layers.0.block.0.weight"] == (16, 65, 7, 7, 7)
    tabs = orc.StageTables(os.path.join(REF, "utils/fisheye/fisheye.calibration_05_08.json"), 64, 2.0)
    out = {}
    report = json.load(open(os.path.join(OUT, "report.json")))
    for mode in ("default", "random_bn"):
        sd = synth.synthetic_state_dict(shapes, seed=0, mode=mode)
        full = net.state_dict()
        full.update(sd)
        net.load_state_dict(full, strict=True)
        feat = synth.synthetic_features(1, seed=7)
        depth = synth.synthetic_depth_room(1, tabs.ray, seed=5)
        net.backbone.forward = lambda images, _f=feat: (None, _f[: images.shape[0]])
        with torch.no_grad():
            kp_ref, _, vol_ref, _ = net(torch.zeros(1, 3, 256, 256), net.grid_coord_proj_batch, net.coord_volumes,
                                        depth_map_batch=depth)
            kp, _, vol, inter = orc.stage_forward(tabs, sd, feat, depth_batch=depth, with_intersection=True,
                                                  return_intermediates=True)
        e = orc.mpjpe(kp.numpy(), kp_ref.numpy())
        assert e <= 5e-5, f"with_intersection stage keypoints differ {e}"
        out[f"kp_{mode}"] = kp_ref.numpy()
        out[f"logits_{mode}"] = inter["logits"].reshape(1, 15, -1)[:, :, ::257].numpy()
        out[f"softmax_{mode}"] = vol_ref.reshape(1, 15, -1)[:, :, ::257].numpy()
        report[f"stage_intersection_{mode}_mpjpe_oracle_vs_ref_m"] = e
    out["state_dict_shapes"] = np.array(json.dumps([[k, list(sh)] for k, sh in shapes]))
    np.savez_compressed(os.path.join(OUT, "stage_intersection_v64.npz"), **out)
    json.dump(report, open(os.path.join(OUT, "report.json"), "w"), indent=1)
    os.chdir(cwd)


def make_eval_golden():
    """Evaluation math: the reference's own umeyama (utils/rigid_transform_with_scale.py, imported unmodified) on
    seeded pose pairs; calculate_error / align_skeleton (utils/calculate_errors.py imports the absent `utils_proj`
    package, so those two are restated in the oracle and run here AROUND the reference's umeyama)."""
    from oracle import sceneego_oracle as orc
    sys.path.insert(0, REF)
    import importlib
    ref_rt = importlib.import_module("utils.rigid_transform_with_scale")
    rng = np.random.default_rng(21)
    B, J = 12, 15
    gt = rng.normal(0, 0.4, (B, J, 3))                                   # float64 like the pickled ground truth
    rot = np.linalg.qr(rng.normal(size=(B, 3, 3)))[0]
    pred = (np.einsum("bjk,bkl->bjl", gt, rot) * rng.uniform(0.7, 1.3, (B, 1, 1)) + rng.normal(0, 0.3, (B, 1, 3))
            + rng.normal(0, 0.02, (B, J, 3))).astype(np.float32)
    pred[3] = (gt[3] * np.array([1.0, 1.0, -1.0])).astype(np.float32)    # a reflected pose: the det < 0 branch
    gt[5, :, 2] = 0.25                                                   # planar ground truth
    # exactly what the reference computes: float32 predictions (network output .cpu().numpy()), float64 ground truth,
    # aligned poses stored back into a float32 array (np.zeros_like(estimated_seq), calculate_errors.py:73)
    T = np.zeros((B, 13))
    aligned = np.zeros_like(pred)
    for b in range(B):
        c, R, t = ref_rt.umeyama(pred[b], gt[b])
        c2, R2, t2 = orc.umeyama(pred[b], gt[b])
        assert abs(c - c2) <= 1e-12 and np.abs(R - R2).max() <= 1e-12 and np.abs(t - t2).max() <= 1e-12
        T[b, 0], T[b, 1:10], T[b, 10:] = c, R.reshape(-1), t
        aligned[b] = pred[b].dot(R) * c + t
    al2, _ = orc.align_skeleton(pred, gt)
    assert al2.dtype == np.float32 and np.array_equal(al2, aligned)
    mp = float(np.mean(np.linalg.norm(pred - gt, axis=2)))
    pa = float(np.mean(np.linalg.norm(aligned - gt, axis=2)))
    assert orc.evaluate_mpjpe(pred, gt) == (mp, pa)
    # the same in float64 throughout (what the device kernel computes from the float32 inputs)
    al64, _ = orc.align_skeleton(pred.astype(np.float64), gt)
    np.savez_compressed(os.path.join(OUT, "eval_poses.npz"), pred=pred, gt=gt, transform=T, aligned=aligned,
                        mpjpe=np.array(mp), pampjpe=np.array(pa), aligned_f64=al64,
                        pampjpe_f64=np.array(orc.calculate_error(al64, gt)))


def make_v2v_simple_golden():
    """V2VModelSimple (network/v2v.py:224-257) at V=32, B=1: the reference class itself on seeded weights."""
    from oracle import sceneego_oracle as orc
    from sceneego_b200.utils import synth
    sys.path.insert(0, REF)
    from network.v2v import V2VModelSimple
    g = torch.Generator().manual_seed(6)
    x = torch.randn(1, 33, 32, 32, 32, generator=g).abs()
    x[:, 32] = (x[:, 32] > 1.0).float()
    out = {}
    report = json.load(open(os.path.join(OUT, "report.json")))
    m = V2VModelSimple(33, 15).eval()
    shapes = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    for mode in ("default", "random_bn"):
        vs = synth.synthetic_state_dict(shapes, seed=2, mode=mode)
        m.load_state_dict(vs, strict=True)
        with torch.no_grad():
            ref = m(x)
            got = orc.v2v_simple_forward(vs, x)
        e = (ref - got).abs().max().item()
        assert e <= 1e-5 * max(1.0, ref.abs().max().item()), f"v2v simple differs {e}"
        out[mode] = ref.reshape(15, -1)[:, ::13].numpy()
        report[f"v2v_simple32_{mode}_max_abs_vs_ref"] = e
    out["state_dict_shapes"] = np.array(json.dumps([[k, list(sh)] for k, sh in shapes]))
    np.savez_compressed(os.path.join(OUT, "v2v_simple_v32.npz"), **out)
    json.dump(report, open(os.path.join(OUT, "report.json"), "w"), indent=1)


def main():
    if "--simple-only" in sys.argv:
        make_v2v_simple_golden()
        return
    if "--eval-only" in sys.argv:
        make_eval_golden()
        return
    if "--intersection-only" in sys.argv:
        make_intersection_golden()
        return
    if "--nearest-only" in sys.argv:
        os.makedirs(OUT, exist_ok=True)
        make_nearest_index_golden()
        return
    config, Net, cwd = import_reference()
    os.makedirs(OUT, exist_ok=True)
    report = {}
    nets = {}
    for V in (64, 128):
        config.opt.batch_size = 2
        config.model.volume_size = V
        torch.manual_seed(0)
        nets[V] = Net(config, device="cpu").eval()
    os.chdir(cwd)
    calib_path = os.path.join(ROOT, "sceneego_b200", "data", "fisheye.calibration_05_08.json")

    # ---- state-dict manifest (names + shapes are part of the drop-in boundary)
    manifest = [(k, list(v.shape)) for k, v in nets[64].state_dict().items()]
    json.dump(manifest, open(os.path.join(OUT, "state_dict_manifest.json"), "w"))
    report["state_dict_keys"] = len(manifest)

    # ---- tables: oracle vs reference, in full
    tabs = {}
    for V in (64, 128):
        net = nets[V]
        t = orc.StageTables(calib_path, V, 2.0)
        tabs[V] = t
        assert np.array_equal(t.ray, net.ray), "ray table differs"
        assert torch.equal(t.coord_volume, net.coord_volume), "coord volume differs"
        dpx = (t.grid_px - net.grid_coord_proj).abs().max().item()
        assert dpx == 0.0, f"grid px differs {dpx}"
        assert torch.equal(t.grid.unsqueeze(0).expand(2, -1, -1, -1), net.grid_coord_proj_batch)
        idx = np.arange(0, V ** 3, 997)
        np.savez_compressed(os.path.join(OUT, f"tables_v{V}.npz"),
                            ray_idx=np.arange(0, t.ray.shape[0], 9973), ray=net.ray[::9973],
                            ray_xor=np.bitwise_xor.reduce(np.ascontiguousarray(net.ray).view(np.uint64), axis=0), vox_idx=idx,
                            grid_px=net.grid_coord_proj.numpy()[idx],
                            coord=net.coord_volume.reshape(-1, 3).numpy()[idx])
    report["tables"] = "exact"
    make_nearest_index_golden()
    report["nearest_index"] = "oracle rule == cv2.resize(INTER_NEAREST) for 2802 (n_src, n_dst) pairs"

    # ---- voxelisation: demo EXRs + synthetic, V=64 and 128, bit-exact
    vox = {}
    for name in ("img_001000", "img_001796", "img_002376"):
        raw, d = load_demo_depth(name)
        if raw.ndim == 3:
            raw = raw[:, :, 0]
        vox[f"{name}_raw"] = raw
        # our restatement of the dataset-side nearest resize must equal cv2's
        mine = orc.resize_nearest(raw, 1024, 1280).copy()
        mine[mine > 10] = 10
        assert np.array_equal(mine, d), "dataset nearest-resize restatement differs"
        for V in (64, 128):
            ref = nets[V].depth_map_to_voxel_numpy(torch.from_numpy(d)).numpy()
            got = orc.voxelize_depth(d, tabs[V].ray, V, 2.0)
            assert np.array_equal(ref, got), f"occupancy differs {name} V={V}"
            assert ref[V // 2, V // 2, 0] == 1
            vox[f"{name}_v{V}"] = np.packbits(ref.astype(np.uint8).reshape(-1))
            report[f"occ_{name}_v{V}"] = int(ref.sum())
    for tag, depth in (("uniform", synth.synthetic_depth_uniform(1)[0].numpy()),
                       ("room", synth.synthetic_depth_room(1, tabs[64].ray)[0].numpy()),
                       ("uniform1024", synth.synthetic_depth_uniform(1, h=1024, w=1024)[0].numpy())):
        for V in (64, 128):
            ref = nets[V].depth_map_to_voxel_numpy(torch.from_numpy(depth)).numpy()
            got = orc.voxelize_depth(depth, tabs[V].ray, V, 2.0)
            assert np.array_equal(ref, got), f"occupancy differs {tag} V={V}"
            vox[f"{tag}_v{V}"] = np.packbits(ref.astype(np.uint8).reshape(-1))
            report[f"occ_{tag}_v{V}"] = int(ref.sum())
    np.savez_compressed(os.path.join(OUT, "voxel.npz"), **vox)

    # ---- process_features + unprojection (B=2, V=64)
    net = nets[64]
    shapes = [(k, tuple(s)) for k, s in manifest if not k.startswith("backbone.")]
    sd = synth.synthetic_state_dict(shapes, seed=0, mode="random_bn")
    full = net.state_dict()
    full.update(sd)
    net.load_state_dict(full, strict=True)
    feat = synth.synthetic_features(2)
    with torch.no_grad():
        pf_ref = net.process_features(feat)
        from utils import op as ref_op
        lift_ref = ref_op.unproject_heatmaps_one_view_batch(pf_ref, net.grid_coord_proj_batch, 64)
        pf = orc.process_features(feat, sd["process_features.0.weight"], sd["process_features.0.bias"])
        assert torch.equal(pf, pf_ref), "process_features differs"
        lift = orc.unproject(pf, net.grid_coord_proj_batch, 64)
    err = (lift - lift_ref).abs().max().item()
    assert err <= 2e-6, f"unproject differs {err}"
    report["unproject_max_abs_vs_ref"] = err
    idx = np.arange(0, 64 ** 3, 61)
    np.savez_compressed(os.path.join(OUT, "unproject_v64.npz"), vox_idx=idx,
                        lifted=lift_ref.reshape(2, 32, -1)[:, :, idx].numpy())

    # ---- V2V alone at V=32 (B=1) with both weight modes
    from network.v2v import V2VModel
    g = torch.Generator().manual_seed(5)
    x32 = torch.randn(1, 33, 32, 32, 32, generator=g).abs()
    x32[:, 32] = (x32[:, 32] > 1.0).float()
    v2v_out = {}
    for mode in ("default", "random_bn"):
        m = V2VModel(33, 15).eval()
        vs = synth.synthetic_state_dict([(k, tuple(v.shape)) for k, v in m.state_dict().items()], seed=1, mode=mode)
        m.load_state_dict(vs, strict=True)
        with torch.no_grad():
            ref = m(x32)
            got = orc.v2v_forward(vs, x32)
        e = (ref - got).abs().max().item()
        assert e <= 1e-5 * max(1.0, ref.abs().max().item()), f"v2v differs {e}"
        v2v_out[mode] = ref.reshape(15, -1)[:, ::13].numpy()
        report[f"v2v32_{mode}_max_abs_vs_ref"] = e
    np.savez_compressed(os.path.join(OUT, "v2v_v32.npz"), **v2v_out)

    # ---- whole stage, B=1 (plus B=2 keypoints), V=64, through the reference forward
    stage = {}
    for mode, scale in (("default", 1.0), ("random_bn", 1.0), ("random_bn", 30.0)):
        sd = synth.synthetic_state_dict(shapes, seed=0, mode=mode, logit_scale=scale)
        full = net.state_dict()
        full.update(sd)
        net.load_state_dict(full, strict=True)
        feat = synth.synthetic_features(2)
        depth = torch.cat([synth.synthetic_depth_room(1, tabs[64].ray), synth.synthetic_depth_uniform(1)])
        net.backbone.forward = lambda images, _f=feat: (None, _f[: images.shape[0]])
        with torch.no_grad():
            kp_ref, _, vol_ref, _ = net(torch.zeros(2, 3, 256, 256), net.grid_coord_proj_batch, net.coord_volumes,
                                        depth_map_batch=depth)
            kp, _, vol, inter = orc.stage_forward(tabs[64], sd, feat, depth_batch=depth, return_intermediates=True)
        e = orc.mpjpe(kp.numpy(), kp_ref.numpy())
        assert e <= 5e-5, f"stage keypoints differ {e}"  # fp32 summation-order noise only (metres)
        tag = f"{mode}_s{int(scale)}"
        stage[f"kp_{tag}"] = kp_ref.numpy()
        stage[f"logits_{tag}"] = inter["logits"].reshape(2, 15, -1)[:, :, ::257].numpy()
        stage[f"softmax_{tag}"] = vol_ref.reshape(2, 15, -1)[:, :, ::257].numpy()
        report[f"stage_{tag}_mpjpe_oracle_vs_ref_m"] = e
    np.savez_compressed(os.path.join(OUT, "stage_v64.npz"), **stage)

    json.dump(report, open(os.path.join(OUT, "report.json"), "w"), indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
