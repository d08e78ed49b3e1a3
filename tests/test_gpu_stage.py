"""Whole volumetric stage through the drop-in module vs the unmodified reference's outputs."""
import numpy as np
import pytest
import torch

from oracle import sceneego_oracle as orc
from sceneego_b200.utils import synth
from tests import util

pytestmark = pytest.mark.gpu

# bf16 bounds of the V2V logits against the reference's fp32 logits (measured values are printed by the tests and
# recorded in profiles/r02_parity.txt): relative Frobenius error and largest absolute error over the logit range
LOGIT_REL_FRO = 1.2e-2          # measured 5.9e-3 .. 7.9e-3 over all configurations (V = 64 / 128, B = 2 / 64, x1 / x30)
LOGIT_MAX_OVER_RANGE = 8.0e-3   # measured 3.3e-3 .. 4.7e-3
SOFTMAX_RTOL = 5.0e-2         # softmaxed volume (output #3), was 0.2 in round 1
S30_GATE_MM = 50.0            # output layer x30 (sharp softmax), tracked: measured 33.9 mm (V=64), 27.5 mm (V=128);
                              # attribution of the bf16 storage error in profiles/r02_bf16_attribution.txt


def _logit_errors(logits, gold_sub, gold_range, stride):
    """GPU logits (B,J,V,V,V) vs the sub-sampled reference logits (B,J,n): (rel-Frobenius, max-abs / range)."""
    b, j = gold_sub.shape[:2]
    mine = logits.reshape(b, j, -1)[:, :, ::stride].cpu().numpy().astype(np.float64)
    ref = gold_sub.astype(np.float64)
    rel = float(np.linalg.norm(mine - ref) / np.linalg.norm(ref))
    mx = float(np.abs(mine - ref).max() / float(gold_range[1] - gold_range[0]))
    return rel, mx


@pytest.fixture(scope="module")
def tables64():
    return orc.StageTables(util.CALIB, 64, 2.0)


@pytest.fixture(scope="module")
def net():
    from sceneego_b200.network.voxel_net_depth import VoxelNetwork_depth
    torch.manual_seed(0)
    return VoxelNetwork_depth(util.load_config(batch_size=4), device="cuda").eval()


def _load(net, mode, scale):
    sd = synth.synthetic_state_dict(util.stage_shapes(), seed=0, mode=mode, logit_scale=scale)
    full = net.state_dict()
    full.update(sd)
    net.load_state_dict(full, strict=True)
    return sd


def test_state_dict_matches_reference_manifest(net):
    mine = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    assert mine == util.manifest()                      # 699 keys, same order, same shapes


def test_attributes_like_reference(net, tables64):
    assert tuple(net.grid_coord_proj_batch.shape) == (4, 64 ** 3, 1, 2) and net.grid_coord_proj_batch.is_cuda
    assert tuple(net.coord_volumes.shape) == (4, 64, 64, 64, 3)
    assert torch.equal(net.coord_volume.cpu(), tables64.coord_volume)
    assert np.array_equal(net.ray, tables64.ray)        # fp64 bit-exact, x-major like the reference
    assert (net.grid_coord_proj.cpu() - tables64.grid_px).abs().max().item() <= 1e-3


@pytest.mark.parametrize("mode,scale,tol_mm", [("default", 1.0, 0.5), ("random_bn", 1.0, 0.5), ("random_bn", 30.0, None)])
def test_stage_keypoints_vs_reference(net, tables64, mode, scale, tol_mm):
    """North-star gate: per-joint 3D output within 0.5 mm MPJPE of the reference PyTorch path; the V2V logits
    themselves against the reference's own (forward hook on its `volume_net`, tests/make_golden_r2.py) within the
    bf16 bound LOGIT_* below; the softmaxed volume within what that logit bound implies.
    The sharpened case (output layer x30) is the regime a trained checkpoint lives in: tracked with its own gate."""
    _load(net, mode, scale)
    feat = synth.synthetic_features(2)
    depth = torch.cat([synth.synthetic_depth_room(1, tables64.ray), synth.synthetic_depth_uniform(1)])
    net.keep_logits = True
    try:
        with torch.no_grad():
            kp, features, volumes, coord = net.lift(feat.cuda(), net.grid_coord_proj_batch, net.coord_volumes,
                                                    depth_map_batch=depth.cuda())
        logits = net.last_logits
    finally:
        net.keep_logits = False
    g, gl = util.golden("stage_v64.npz"), util.golden("stage_v64_logits.npz")
    tag = f"{mode}_s{int(scale)}"
    err_mm = orc.mpjpe(kp.cpu().numpy(), g[f"kp_{tag}"]) * 1000.0
    rel, mx = _logit_errors(logits, gl[f"logits_{tag}"], gl[f"logit_range_{tag}"], 257)
    print(f"[{tag}] MPJPE vs reference {err_mm:.4f} mm; logits rel-Frobenius {rel:.3e}, max-abs/range {mx:.3e}")
    assert err_mm <= (tol_mm if tol_mm is not None else S30_GATE_MM)
    assert rel <= LOGIT_REL_FRO and mx <= LOGIT_MAX_OVER_RANGE
    assert features.shape == (2, 32, 1024, 1280) and volumes.shape == (2, 15, 64, 64, 64)
    assert coord is net.coord_volumes
    sm = volumes.reshape(2, 15, -1)[:, :, ::257].cpu().numpy()
    if tol_mm is not None:
        # p = exp(l) / Z: a logit error of at most d changes p by at most a factor exp(2 d)
        d = mx * float(gl[f"logit_range_{tag}"][1] - gl[f"logit_range_{tag}"][0])
        assert np.allclose(sm, g[f"softmax_{tag}"], rtol=float(np.expm1(2.2 * d)) + 1e-4, atol=1e-9)
        assert np.allclose(sm, g[f"softmax_{tag}"], rtol=SOFTMAX_RTOL, atol=1e-9)


def test_stage_b64_bench_configuration(tables64):
    """BASELINE configs[1] as bench.py runs it: 64 frames, one 64-frame V2V chunk (2,112 marching items on 296 CTAs),
    the bench's rank-0 inputs.  Frames 0 / 21 / 42 / 63 against the unmodified reference's outputs for those frames
    (tests/golden/stage_v64_b64.npz): keypoints, V2V logits, softmaxed volumes, occupancy counts."""
    from sceneego_b200.network.voxel_net_depth import VoxelNetwork_depth
    torch.manual_seed(0)
    big = VoxelNetwork_depth(util.load_config(batch_size=64), device="cuda", v2v_chunk=64).eval()
    _load(big, "random_bn", 1.0)
    g = util.golden("stage_v64_b64.npz")
    frames = [int(f) for f in g["frames"]]
    feat = synth.synthetic_features(64, seed=1234).cuda()
    depth = synth.synthetic_depth_room(64, tables64.ray, seed=7).cuda()
    big.keep_logits = True
    with torch.no_grad():
        kp, features, volumes, _ = big.lift(feat, big.grid_coord_proj_batch, big.coord_volumes, depth_map_batch=depth)
    assert kp.shape == (64, 15, 3) and features.shape == (64, 32, 1024, 1280)
    err_mm = orc.mpjpe(kp[frames].cpu().numpy(), g["kp"]) * 1000.0
    lg = big.last_logits[frames]
    rng_ = np.array([g["logits"].min(), g["logits"].max()])
    rel, mx = _logit_errors(lg, g["logits"], rng_, 257)
    print(f"[B=64 chunk 64] MPJPE vs reference {err_mm:.4f} mm; logits rel-Frobenius {rel:.3e}, max-abs/range {mx:.3e}")
    assert err_mm <= 0.5 and rel <= LOGIT_REL_FRO and mx <= LOGIT_MAX_OVER_RANGE
    sm = volumes[frames].reshape(4, 15, -1)[:, :, ::257].cpu().numpy()
    assert np.allclose(sm, g["softmax"], rtol=SOFTMAX_RTOL, atol=1e-9)
    pg = big.volume_net.program(64, 64, torch.device("cuda", 0))
    from sceneego_b200 import _lib
    occ = _lib.unpack_volume(pg.buffers[pg.in_buf], pg.lay_in, 64, 33)[frames, 32]
    assert [int(o.sum().item()) for o in occ] == [int(c) for c in g["occupied"]]
    # every frame of the batch equals the same frame run alone (no cross-frame state at the bench's launch shapes)
    with torch.no_grad():
        solo = big.lift(feat[21:22], big.grid_coord_proj_batch, big.coord_volumes, depth_map_batch=depth[21:22])[0]
    assert torch.allclose(solo[0], kp[21], atol=1e-6)
    del big


@pytest.mark.parametrize("mode,scale,tol_mm", [("default", 1.0, 0.5), ("random_bn", 1.0, 0.5), ("random_bn", 30.0, None)])
def test_stage_v128_vs_reference(mode, scale, tol_mm):
    """BASELINE configs[3]: 128^3 voxel cube, B=1, whole stage against the UNMODIFIED reference run at that size
    (tests/golden/stage_v128.npz): keypoints, sub-sampled V2V logits and softmaxed volume."""
    from sceneego_b200.network.voxel_net_depth import VoxelNetwork_depth
    torch.manual_seed(0)
    net128 = VoxelNetwork_depth(util.load_config(batch_size=1, volume_size=128), device="cuda", v2v_chunk=1).eval()
    _load(net128, mode, scale)
    t128 = orc.StageTables(util.CALIB, 128, 2.0)
    feat = synth.synthetic_features(1, seed=77)
    depth = synth.synthetic_depth_room(1, t128.ray, seed=78)
    net128.keep_logits = True
    with torch.no_grad():
        kp, _, volumes, _ = net128.lift(feat.cuda(), net128.grid_coord_proj_batch, net128.coord_volumes,
                                        depth_map_batch=depth.cuda())
    g = util.golden("stage_v128.npz")
    tag = f"{mode}_s{int(scale)}"
    err_mm = orc.mpjpe(kp.cpu().numpy(), g[f"kp_{tag}"]) * 1000.0
    rel, mx = _logit_errors(net128.last_logits, g[f"logits_{tag}"], g[f"logit_range_{tag}"], 2053)
    print(f"[V=128 {tag}] MPJPE vs reference {err_mm:.4f} mm; logits rel-Frobenius {rel:.3e}, max-abs/range {mx:.3e}")
    assert err_mm <= (tol_mm if tol_mm is not None else S30_GATE_MM)
    assert rel <= LOGIT_REL_FRO and mx <= LOGIT_MAX_OVER_RANGE
    if tol_mm is not None:
        sm = volumes.reshape(1, 15, -1)[:, :, ::2053].cpu().numpy()
        assert np.allclose(sm, g[f"softmax_{tag}"], rtol=SOFTMAX_RTOL, atol=1e-10)
        if scale == 1.0 and f"occupied_{tag}" in g:
            pg = net128.volume_net.program(128, 1, torch.device("cuda", 0))
            from sceneego_b200 import _lib
            occ = _lib.unpack_volume(pg.buffers[pg.in_buf], pg.lay_in, 1, 33)[0, 32]
            assert int(occ.sum().item()) == int(g[f"occupied_{tag}"])
    del net128


def test_forward_signature_variants(net, tables64):
    _load(net, "random_bn", 1.0)
    feat = synth.synthetic_features(3, seed=5).cuda()
    depth = synth.synthetic_depth_room(3, tables64.ray, seed=3)
    with torch.no_grad():
        a = net.lift(feat, net.grid_coord_proj_batch, net.coord_volumes, depth_map_batch=depth.cuda())[0]
        # scene_volumes= path (voxel_net_depth.py:246-249) with the oracle's voxel grids
        sv = torch.stack([torch.from_numpy(orc.voxelize_depth(d.numpy(), tables64.ray, 64, 2.0)) for d in depth])
        b = net.lift(feat, net.grid_coord_proj_batch, net.coord_volumes, scene_volumes=sv.cuda())[0]
        assert torch.equal(a, b)
        assert net.lift(feat, net.grid_coord_proj_batch, net.coord_volumes) is None     # prints, returns None
        # fused in-kernel projection gives the same poses within 0.5 mm
        net.fused_projection = True
        c = net.lift(feat, net.grid_coord_proj_batch, net.coord_volumes, depth_map_batch=depth.cuda())[0]
        net.fused_projection = False
        assert orc.mpjpe(c.cpu().numpy(), a.cpu().numpy()) * 1000 <= 0.5
        # a frame alone equals the frame inside a batch (frames are independent)
        d = net.lift(feat[1:2], net.grid_coord_proj_batch, net.coord_volumes, depth_map_batch=depth[1:2].cuda())[0]
        assert torch.allclose(d[0], a[1], atol=1e-6)
        # full forward with the stock backbone runs and has the reference's output shapes
        img = torch.randn(2, 3, 256, 256, device="cuda")
        out = net(img, net.grid_coord_proj_batch, net.coord_volumes, depth_map_batch=depth[:2].cuda())
        assert out[0].shape == (2, 15, 3) and torch.isfinite(out[0]).all()


def test_cpu_device_is_rejected():
    from sceneego_b200 import _lib
    from sceneego_b200.network.voxel_net_depth import VoxelNetwork_depth
    with pytest.raises(_lib.SceneEgoError):
        VoxelNetwork_depth(util.load_config(), device="cpu")
    with pytest.raises(_lib.SceneEgoError):
        _lib.softargmax3d(torch.zeros(1, 1, 4, 4, 4), 1.0, True, torch.zeros(3, 4), None, False)


def test_ragged_chunks(tables64):
    """Batches that do not divide the V2V chunk (5 frames in chunks of 2, 2, 1) give the poses of the same frames
    run one chunk at a time."""
    from sceneego_b200.network.voxel_net_depth import VoxelNetwork_depth
    torch.manual_seed(0)
    small = VoxelNetwork_depth(util.load_config(batch_size=2), device="cuda", v2v_chunk=2).eval()
    _load(small, "random_bn", 1.0)
    feat = synth.synthetic_features(5, seed=9).cuda()
    depth = synth.synthetic_depth_room(5, tables64.ray, seed=2).cuda()
    with torch.no_grad():
        all5 = small.lift(feat, small.grid_coord_proj_batch, small.coord_volumes, depth_map_batch=depth)[0]
        last = small.lift(feat[4:5], small.grid_coord_proj_batch, small.coord_volumes, depth_map_batch=depth[4:5])[0]
    assert all5.shape == (5, 15, 3) and torch.allclose(all5[4], last[0], atol=1e-6)
    del small


@pytest.mark.parametrize("mode", ["default", "random_bn"])
def test_with_intersection_vs_reference(tables64, mode):
    """with_intersection=true (network/voxel_net_depth.py:66-69,257-260): 65-channel V2V input
    cat([volumes, volumes * scene, scene]) against the UNMODIFIED reference's outputs
    (tests/golden/stage_intersection_v64.npz).  Same 0.5 mm MPJPE gate as the default configuration."""
    import json
    from sceneego_b200 import _lib
    from sceneego_b200.network.voxel_net_depth import VoxelNetwork_depth
    cfg = util.load_config(batch_size=1)
    cfg.model.with_intersection = True
    net = VoxelNetwork_depth(cfg, device="cuda").eval()
    g = util.golden("stage_intersection_v64.npz")
    ref_shapes = [(k, tuple(s)) for k, s in json.loads(str(g["state_dict_shapes"]))]
    mine = [(k, tuple(v.shape)) for k, v in net.state_dict().items() if not k.startswith("backbone.")]
    assert mine == ref_shapes and dict(mine)["volume_net.front_layers.0.block.0.weight"] == (16, 65, 7, 7, 7)
    sd = synth.synthetic_state_dict(ref_shapes, seed=0, mode=mode)
    full = net.state_dict()
    full.update(sd)
    net.load_state_dict(full, strict=True)
    feat = synth.synthetic_features(1, seed=7)
    depth = synth.synthetic_depth_room(1, tables64.ray, seed=5)
    with torch.no_grad():
        kp, _, volumes, _ = net.lift(feat.cuda(), net.grid_coord_proj_batch, net.coord_volumes, depth_map_batch=depth.cuda())
    err_mm = orc.mpjpe(kp.cpu().numpy(), g[f"kp_{mode}"]) * 1000.0
    print(f"MPJPE vs reference [with_intersection, {mode}]: {err_mm:.4f} mm")
    assert err_mm <= 0.5
    sm = volumes.reshape(1, 15, -1)[:, :, ::257].cpu().numpy()
    assert np.allclose(sm, g[f"softmax_{mode}"], rtol=0.2, atol=1e-7)
    # the V2V input the kernels built: channels 32..63 are channels 0..31 times the occupancy channel, bit for bit
    pg = net.volume_net.program(64, 1, torch.device("cuda", 0))
    x = _lib.unpack_volume(pg.buffers[pg.in_buf], pg.lay_in, 1, 65)
    occ = torch.from_numpy(orc.voxelize_depth(depth[0].numpy(), tables64.ray, 64, 2.0)).cuda()
    assert torch.equal(x[0, 64], occ)
    assert torch.equal(x[0, 32:64], x[0, :32] * occ)
    with pytest.raises(_lib.SceneEgoError):             # the reference's scene_volumes path feeds 33 channels to this net
        net.lift(feat.cuda(), net.grid_coord_proj_batch, net.coord_volumes, scene_volumes=occ[None])


def test_host_pipeline_matches_direct_lift(net, tables64):
    """HostStagePipeline.run (pinned host buffers -> poses on the host; the first batch of a call is ramped through
    the copy stream in parts) returns exactly what lift() returns on the same frames, for ramped and plain batches."""
    from sceneego_b200.pipeline import HostStagePipeline
    _load(net, "random_bn", 1.0)
    B = 8
    feats = [synth.synthetic_features(B, seed=40 + i).pin_memory() for i in range(3)]
    depths = [synth.synthetic_depth_room(B, tables64.ray, seed=50 + i).pin_memory() for i in range(3)]
    pipe = HostStagePipeline(net)
    pipe.ramp_min_batch = 8                                  # ramp already at this small test batch: parts of 2, 2, 4 frames
    outs = [o.clone() for o in pipe.run(list(zip(feats, depths)))]
    assert pipe.h2d_bytes == sum(f.numel() * 4 + d.numel() * 4 for f, d in zip(feats, depths)) and pipe.d2h_bytes == 3 * B * 15 * 3 * 4
    again = [o.clone() for o in pipe.run(list(zip(feats, depths)))]          # buffers reused, same results
    with torch.no_grad():
        for i in range(3):
            ref = net.lift(feats[i].cuda(), net.grid_coord_proj_batch, net.coord_volumes, depth_map_batch=depths[i].cuda())[0].cpu()
            # no cross-frame state: cutting the batch only changes how many partial sums the soft-argmax splits a
            # joint volume into (fp32 summation order): 1e-6 m
            assert torch.equal(outs[i], again[i]), i
            assert (outs[i] - ref).abs().max().item() <= 1e-6, i
    with pytest.raises(ValueError):
        pipe.run([(feats[0].clone(), depths[0])])            # pageable host memory is refused


def test_lift_with_fused_depth_preprocessing(net, tables64):
    """net.depth_preprocess = (1024, 1280, 10.0): depth_map_batch holds the RAW decoded maps (512x640 demo EXRs, values
    above 10 m) and the dataset's resize + clamp (dataset/demo_dataset.py:86-91) happen inside the voxelisation kernel.
    Same poses, bit for bit, as preprocessing on the host first."""
    _load(net, "random_bn", 1.0)
    g = util.golden("voxel.npz")
    raws = np.stack([g[f"{n}_raw"] for n in ("img_001000", "img_001796", "img_002376")]).astype(np.float32)
    raws[1, 100:140, 200:260] = 25.0                                     # beyond the clamp
    pre = np.stack([orc.preprocess_depth(r) for r in raws])
    feat = synth.synthetic_features(3, seed=9).cuda()
    with torch.no_grad():
        a = net.lift(feat, net.grid_coord_proj_batch, net.coord_volumes, depth_map_batch=torch.from_numpy(pre).cuda())[0]
        net.depth_preprocess = (1024, 1280, 10.0)
        try:
            b = net.lift(feat, net.grid_coord_proj_batch, net.coord_volumes, depth_map_batch=torch.from_numpy(raws).cuda())[0]
        finally:
            net.depth_preprocess = None
    assert torch.equal(a, b)


def test_module_semantics_like_reference(net, tables64):
    """Boundary hygiene (VERDICT r1 'Boundary / semantics'): output #2 is a fresh tensor per call; a caller-supplied
    `coord_volumes` is honoured; materialised (non-expanded) grid / coordinate tables are accepted when their rows are
    equal and refused when they differ; in-place parameter edits invalidate the packed weights; training mode is
    refused instead of silently folding running statistics."""
    from sceneego_b200 import _lib
    _load(net, "random_bn", 1.0)
    feat = synth.synthetic_features(2, seed=21).cuda()
    depth = synth.synthetic_depth_room(2, tables64.ray, seed=22).cuda()
    with torch.no_grad():
        kp1, f1, v1, _ = net.lift(feat, net.grid_coord_proj_batch, net.coord_volumes, depth_map_batch=depth)
        f1_copy = f1.clone()
        kp2, f2, v2, _ = net.lift(feat.flip(0), net.grid_coord_proj_batch, net.coord_volumes, depth_map_batch=depth.flip(0))
        torch.cuda.synchronize()
        assert f2.data_ptr() != f1.data_ptr() and torch.equal(f1, f1_copy)          # the first call's output survives
        assert torch.equal(f2, f1.flip(0)) and torch.allclose(kp2, kp1.flip(0), atol=1e-6)
        # coord_volumes: another table (metres -> millimetres, shifted) changes the poses accordingly
        other = (net.coord_volume * 1000.0 + 5.0).unsqueeze(0).expand(2, -1, -1, -1, -1)
        kp_mm = net.lift(feat, net.grid_coord_proj_batch, other, depth_map_batch=depth)[0]
        assert torch.allclose(kp_mm, kp1 * 1000.0 + 5.0, rtol=1e-4, atol=5e-2)
        # materialised copies of the module's own tables: same result; rows that differ: refused
        g_mat = net.grid_coord_proj_batch[:2].clone()
        c_mat = net.coord_volumes[:2].clone()
        kp_mat = net.lift(feat, g_mat, c_mat, depth_map_batch=depth)[0]
        assert torch.allclose(kp_mat, kp1, atol=2e-6)
        g_bad = g_mat.clone()
        g_bad[1, 0, 0, 0] += 0.25
        with pytest.raises(_lib.SceneEgoError):
            net.lift(feat, g_bad, net.coord_volumes, depth_map_batch=depth)
        c_bad = c_mat.clone()
        c_bad[1, 0, 0, 0, 0] += 1.0
        with pytest.raises(_lib.SceneEgoError):
            net.lift(feat, net.grid_coord_proj_batch, c_bad, depth_map_batch=depth)
        # in-place edit of a parameter: the next call repacks (no invalidate() needed)
        w = net.volume_net.output_layer.weight
        w.mul_(2.0)
        lg_scale = net.lift(feat, net.grid_coord_proj_batch, net.coord_volumes, depth_map_batch=depth)[2]
        w.mul_(0.5)
        back = net.lift(feat, net.grid_coord_proj_batch, net.coord_volumes, depth_map_batch=depth)
        assert not torch.allclose(lg_scale, v1, rtol=1e-3, atol=0) and torch.equal(back[0], kp1) and torch.equal(back[2], v1)
    net.train()
    try:
        with pytest.raises(_lib.SceneEgoError, match="eval"):
            net.lift(feat, net.grid_coord_proj_batch, net.coord_volumes, depth_map_batch=depth)
    finally:
        net.eval()


def test_one_program_for_ragged_batches(tables64):
    """A smaller batch (the ragged last DataLoader batch, the pipeline's ramp) runs on the existing buffer pool; only
    a larger one replaces it: there is never a second multi-GB program."""
    from sceneego_b200.network.voxel_net_depth import VoxelNetwork_depth
    torch.manual_seed(0)
    m = VoxelNetwork_depth(util.load_config(batch_size=8), device="cuda", v2v_chunk=8, materialize_features=False).eval()
    _load(m, "random_bn", 1.0)
    feat = synth.synthetic_features(8, seed=31).cuda()
    depth = synth.synthetic_depth_room(8, tables64.ray, seed=32).cuda()
    with torch.no_grad():
        k8 = m.lift(feat, m.grid_coord_proj_batch, m.coord_volumes, depth_map_batch=depth)[0]
        pg = m.volume_net._programs[64]
        for n in (3, 1, 5):
            kn = m.lift(feat[:n], m.grid_coord_proj_batch, m.coord_volumes, depth_map_batch=depth[:n])[0]
            assert torch.allclose(kn, k8[:n], atol=1e-6)
            assert m.volume_net._programs[64] is pg and len(m.volume_net._programs) == 1
    del m


@pytest.mark.parametrize("b", [1, 3])
def test_cuda_graph_lift_equals_eager(tables64, b):
    """graph_max_batch > 0: the gather -> V2V launches of a small batch replay as one captured CUDA graph
    (demo.py:44-60 runs batch 1).  Same outputs as the eager path, bit for bit; a second call with other inputs
    re-uses the capture; scene_volumes= gets its own capture."""
    from sceneego_b200.network.voxel_net_depth import VoxelNetwork_depth
    torch.manual_seed(0)
    eager = VoxelNetwork_depth(util.load_config(batch_size=4), device="cuda", v2v_chunk=4).eval()
    graphed = VoxelNetwork_depth(util.load_config(batch_size=4), device="cuda", v2v_chunk=4, graph_max_batch=4).eval()
    _load(eager, "random_bn", 1.0)
    _load(graphed, "random_bn", 1.0)
    with torch.no_grad():
        for seed in (41, 42):
            feat = synth.synthetic_features(b, seed=seed).cuda()
            depth = synth.synthetic_depth_room(b, tables64.ray, seed=seed + 10).cuda()
            e = eager.lift(feat, eager.grid_coord_proj_batch, eager.coord_volumes, depth_map_batch=depth)
            g = graphed.lift(feat, graphed.grid_coord_proj_batch, graphed.coord_volumes, depth_map_batch=depth)
            assert torch.equal(e[0], g[0]) and torch.equal(e[1], g[1]) and torch.equal(e[2], g[2])
        assert len(graphed._graphs) == 1
        sv = torch.stack([torch.from_numpy(orc.voxelize_depth(d.cpu().numpy(), tables64.ray, 64, 2.0)) for d in depth]).cuda()
        g2 = graphed.lift(feat, graphed.grid_coord_proj_batch, graphed.coord_volumes, scene_volumes=sv)
        assert torch.equal(g2[0], e[0]) and len(graphed._graphs) == 2
        assert graphed.lift(feat, graphed.grid_coord_proj_batch, graphed.coord_volumes) is None
    del eager, graphed


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_second_device_without_set_device(tables64):
    """VoxelNetwork_depth(config, device='cuda:1') without torch.cuda.set_device(1): the binding selects the tensors'
    device and its current stream for every launch and sets the >48 KB shared-memory attribute per device."""
    from sceneego_b200.network.voxel_net_depth import VoxelNetwork_depth
    assert torch.cuda.current_device() == 0
    torch.manual_seed(0)
    n0 = VoxelNetwork_depth(util.load_config(batch_size=2), device="cuda:0", v2v_chunk=2, materialize_features=False).eval()
    n1 = VoxelNetwork_depth(util.load_config(batch_size=2), device="cuda:1", v2v_chunk=2, materialize_features=False).eval()
    _load(n0, "random_bn", 1.0)
    _load(n1, "random_bn", 1.0)
    feat = synth.synthetic_features(2, seed=51)
    depth = synth.synthetic_depth_room(2, tables64.ray, seed=52)
    with torch.no_grad():
        a = n0.lift(feat.to("cuda:0"), n0.grid_coord_proj_batch, n0.coord_volumes, depth_map_batch=depth.to("cuda:0"))[0]
        b = n1.lift(feat.to("cuda:1"), n1.grid_coord_proj_batch, n1.coord_volumes, depth_map_batch=depth.to("cuda:1"))[0]
    assert b.device == torch.device("cuda", 1) and torch.cuda.current_device() == 0
    assert torch.equal(a.cpu(), b.cpu())


def test_stage_v96_non_power_of_two_cube_vs_oracle():
    """volume_size = 96 (divisible by 32 as the five poolings need, NOT a power of two: the index decodes of the gather,
    the soft-argmax and the marching stem's x-segments take their generic branches): whole stage, B = 2, against the
    CPU oracle run here (no reference-generated golden at this size; the oracle is pinned at 64 and 128)."""
    from sceneego_b200.network.voxel_net_depth import VoxelNetwork_depth
    torch.manual_seed(0)
    net96 = VoxelNetwork_depth(util.load_config(batch_size=2, volume_size=96), device="cuda", v2v_chunk=2).eval()
    sd = _load(net96, "random_bn", 1.0)
    t96 = orc.StageTables(util.CALIB, 96, 2.0)
    feat = synth.synthetic_features(2, seed=61)
    depth = synth.synthetic_depth_room(2, t96.ray, seed=62)
    net96.keep_logits = True
    with torch.no_grad():
        kp, _, volumes, _ = net96.lift(feat.cuda(), net96.grid_coord_proj_batch, net96.coord_volumes, depth_map_batch=depth.cuda())
        kp_ref, _, vol_ref, inter = orc.stage_forward(t96, sd, feat[:1], depth_batch=depth[:1], return_intermediates=True)
    err_mm = orc.mpjpe(kp[:1].cpu().numpy(), kp_ref.numpy()) * 1000.0
    lg = inter["logits"]
    rel = ((net96.last_logits[:1].cpu() - lg).norm() / lg.norm()).item()
    print(f"[V=96] MPJPE vs oracle {err_mm:.4f} mm; logits rel-Frobenius {rel:.3e}")
    assert err_mm <= 0.5 and rel <= LOGIT_REL_FRO
    pg = net96.volume_net.program(96, 2, torch.device("cuda", 0))
    from sceneego_b200 import _lib
    occ = _lib.unpack_volume(pg.buffers[pg.in_buf], pg.lay_in, 2, 33)[0, 32]
    assert torch.equal(occ.cpu(), inter["scene"][0])
    del net96
