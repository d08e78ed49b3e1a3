"""CPU: the oracle restatement vs vectors produced by the UNMODIFIED reference
(tests/make_golden.py, run in the build container).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

from oracle import sceneego_oracle as orc
from sceneego_b200.utils import synth
from tests import util


@pytest.fixture(scope="module")
def tabs():
    return {V: orc.StageTables(util.CALIB, V, 2.0) for V in (64, 128)}


@pytest.mark.parametrize("V", [64, 128])
def test_tables(tabs, V):
    g, t = util.golden(f"tables_v{V}.npz"), tabs[V]
    assert np.array_equal(t.ray[g["ray_idx"]], g["ray"])                 # fp64, bit-exact
    assert np.array_equal(np.bitwise_xor.reduce(np.ascontiguousarray(t.ray).view(np.uint64), axis=0), g["ray_xor"])  # order-free checksum of all 1,310,720 rays
    assert np.array_equal(t.grid_px.numpy()[g["vox_idx"]], g["grid_px"])
    assert np.array_equal(t.coord_volume.reshape(-1, 3).numpy()[g["vox_idx"]], g["coord"])


def test_world2camera_raises_on_axis():
    calib = orc.load_calibration(util.CALIB)
    with pytest.raises(Exception, match="norm is zero"):
        orc.world2camera_f32(calib, orc.build_coord_volume(33, 2.0).reshape(-1, 3))


@pytest.mark.parametrize("V", [64, 128])
def test_voxel_occupancy_bit_exact(tabs, V):
    g = util.golden("voxel.npz")
    counts = {}
    for name in ("img_001000", "img_001796", "img_002376"):
        d = orc.resize_nearest(g[f"{name}_raw"], 1024, 1280).copy()
        d[d > 10] = 10
        got = orc.voxelize_depth(d, tabs[V].ray, V, 2.0)
        assert np.array_equal(got, util.unpack_bits(g[f"{name}_v{V}"], V))
        counts[name] = int(got.sum())
    if V == 64:   # fingerprints recorded by the survey
        assert counts == {"img_001000": 8331, "img_001796": 6787, "img_002376": 7450}
    d = synth.synthetic_depth_uniform(1, h=1024, w=1024)[0].numpy()
    assert np.array_equal(orc.voxelize_depth(d, tabs[V].ray, V, 2.0), util.unpack_bits(g[f"uniform1024_v{V}"], V))


def test_unproject(tabs):
    g = util.golden("unproject_v64.npz")
    sd = synth.synthetic_state_dict(util.stage_shapes(), seed=0, mode="random_bn")
    pf = orc.process_features(synth.synthetic_features(2), sd["process_features.0.weight"],
                              sd["process_features.0.bias"])
    assert pf.shape == (2, 32, 1024, 1280) and float(pf[..., :128].abs().max()) == 0.0
    lifted = orc.unproject(pf, tabs[64].grid.unsqueeze(0).expand(2, -1, -1, -1), 64)
    got = lifted.reshape(2, 32, -1)[:, :, g["vox_idx"]].numpy()
    assert np.abs(got - g["lifted"]).max() <= 2e-6


@pytest.mark.parametrize("mode", ["default", "random_bn"])
def test_v2v(mode):
    g = util.golden("v2v_v32.npz")
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(1, 33, 32, 32, 32, generator=gen).abs()
    x[:, 32] = (x[:, 32] > 1.0).float()
    shapes = [(k[len("volume_net."):], s) for k, s in util.manifest() if k.startswith("volume_net.")]
    sd = synth.synthetic_state_dict(shapes, seed=1, mode=mode)
    with torch.no_grad():
        out = orc.v2v_forward(sd, x)
    assert np.abs(out.reshape(15, -1)[:, ::13].numpy() - g[mode]).max() <= 1e-5


def test_stage_end_to_end(tabs):
    g = util.golden("stage_v64.npz")
    sd = synth.synthetic_state_dict(util.stage_shapes(), seed=0, mode="random_bn", logit_scale=30.0)
    feat = synth.synthetic_features(2)
    depth = torch.cat([synth.synthetic_depth_room(1, tabs[64].ray), synth.synthetic_depth_uniform(1)])
    with torch.no_grad():
        kp, _, vol, inter = orc.stage_forward(tabs[64], sd, feat, depth_batch=depth, return_intermediates=True)
    assert orc.mpjpe(kp.numpy(), g["kp_random_bn_s30"]) <= 5e-5          # metres: fp32 summation order only
    assert np.allclose(inter["logits"].reshape(2, 15, -1)[:, :, ::257].numpy(), g["logits_random_bn_s30"],
                       rtol=1e-4, atol=1e-4)
    assert float(inter["scene"][:, 32, 32, 0].min()) == 1.0


def test_nearest_index_matches_cv2_golden():
    """cv2.resize(INTER_NEAREST) index maps recorded from OpenCV itself (tests/make_golden.py): the oracle's rule
    reproduces all of them, including the source sizes where floor(dst * n_src / n_dst) in exact arithmetic does not."""
    g = util.golden("nearest_index.npz")
    differs = 0
    for key in g.files:
        ns, nd = (int(t) for t in key.split("_"))
        ref = g[key].astype(np.int64)
        assert np.array_equal(orc.nearest_index(ns, nd), ref), key
        differs += int(not np.array_equal(np.minimum(np.arange(nd) * ns // nd, ns - 1), ref))
    assert differs >= 5          # the fixture does cover the cases that tell the two rules apart


def test_dataset_preprocessing_restated(tabs):
    """dataset/demo_dataset.py:86-91 on the raw demo EXRs (512x640): resize + clamp, then the occupancy golden."""
    g = util.golden("voxel.npz")
    raw = g["img_001796_raw"]
    d = orc.preprocess_depth(raw)
    assert d.shape == (1024, 1280) and d.dtype == np.float32 and d.max() <= 10.0
    assert np.array_equal(orc.voxelize_depth(d, tabs[64].ray, 64, 2.0), util.unpack_bits(g["img_001796_v64"], 64))
    already = np.full((1024, 1280), 12.5, np.float32)
    assert np.array_equal(orc.preprocess_depth(already), np.full((1024, 1280), 10.0, np.float32))


def test_evaluation_math_restated():
    """umeyama / align_skeleton / calculate_error against the reference's own umeyama outputs
    (tests/golden/eval_poses.npz, utils/rigid_transform_with_scale.py imported unmodified by make_golden)."""
    g = util.golden("eval_poses.npz")
    pred, gt = g["pred"], g["gt"]
    for b in range(pred.shape[0]):
        c, R, t = orc.umeyama(pred[b], gt[b])
        assert abs(c - g["transform"][b, 0]) <= 1e-12
        assert np.abs(R.reshape(-1) - g["transform"][b, 1:10]).max() <= 1e-12
        assert np.abs(t - g["transform"][b, 10:]).max() <= 1e-12
        assert abs(np.linalg.det(R) - 1.0) <= 1e-9                      # a rotation, also for the reflected pose
    aligned, gt_out = orc.align_skeleton(pred, gt)
    assert np.array_equal(aligned, g["aligned"]) and np.array_equal(gt_out, gt)
    assert orc.evaluate_mpjpe(pred, gt) == (float(g["mpjpe"]), float(g["pampjpe"]))
    a0, g0 = orc.align_skeleton(pred, gt, scale=False)                   # centred, rotation + translation only
    assert np.abs(g0.mean(axis=1)).max() <= 1e-12 and np.abs(a0.mean(axis=1)).max() <= 1e-5


@pytest.mark.parametrize("V", [64, 128])
def test_dataset_voxel_occupancy_bit_exact(tabs, V):
    """dataset/real_depth_utils.depth_map_to_voxel (voxel_output=True path): the oracle restatement vs grids produced
    by the reference's own function (tests/make_golden_r2.py)."""
    g, gd = util.golden("voxel.npz"), util.golden("voxel_dataset.npz")
    for name in ("img_001000", "img_002376"):
        d = orc.preprocess_depth(g[f"{name}_raw"])
        got = orc.voxelize_depth_dataset(d, tabs[64].ray, V, 2.0)
        assert np.array_equal(got, util.unpack_bits(gd[f"{name}_v{V}"], V))
        assert not np.array_equal(got, util.unpack_bits(g[f"{name}_v{V}"], V))     # not the network's semantics
    d = synth.synthetic_depth_room(1, tabs[64].ray)[0].numpy()
    assert np.array_equal(orc.voxelize_depth_dataset(d, tabs[64].ray, V, 2.0), util.unpack_bits(gd[f"room_v{V}"], V))


def test_stage_b64_sample_vs_reference(tabs):
    """One of the four sampled frames of the bench's 64-frame batch (BASELINE configs[1]) through the oracle vs the
    reference's keypoints and its own V2V logits (forward hook)."""
    g = util.golden("stage_v64_b64.npz")
    f = int(g["frames"][1])
    sd = synth.synthetic_state_dict(util.stage_shapes(), seed=0, mode="random_bn")
    feat = synth.synthetic_features(64, seed=1234)[f:f + 1].contiguous()
    depth = synth.synthetic_depth_room(64, tabs[64].ray, seed=7)[f:f + 1].contiguous()
    with torch.no_grad():
        kp, _, vol, inter = orc.stage_forward(tabs[64], sd, feat, depth_batch=depth, return_intermediates=True)
    assert orc.mpjpe(kp.numpy(), g["kp"][1:2]) <= 5e-5
    lg = inter["logits"].reshape(1, 15, -1)[:, :, ::257].numpy()
    assert np.linalg.norm(lg - g["logits"][1:2]) / np.linalg.norm(g["logits"][1:2]) <= 1e-5
    assert int(inter["scene"][0].sum()) == int(g["occupied"][1])


def test_stage_forward_device_equals_stage_forward(tabs):
    """`stage_forward_device` (the reference's own torch op sequence -- F.interpolate, F.grid_sample -- that
    `bench.py --impl reference-gpu` times on the B200) against the reference's keypoints, here on the CPU."""
    g = util.golden("stage_v64.npz")
    sd = synth.synthetic_state_dict(util.stage_shapes(), seed=0, mode="random_bn")
    feat = synth.synthetic_features(2)[:1]
    depth = synth.synthetic_depth_room(1, tabs[64].ray)
    with torch.no_grad():
        kp, features, vol, tim = orc.stage_forward_device(tabs[64], sd, feat, depth, torch.device("cpu"), timings=True)
    assert orc.mpjpe(kp.numpy(), g["kp_random_bn_s1"][:1]) <= 5e-5
    assert features.shape == (1, 32, 1024, 1280) and tim["voxel_loop_s"] > 0
