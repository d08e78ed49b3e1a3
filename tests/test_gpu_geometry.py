"""Tables, voxelisation (bit-exact) and unprojection (1e-3 relative) on the GPU, through the C-ABI."""
import numpy as np
import pytest
import torch

from oracle import sceneego_oracle as orc
from sceneego_b200 import _lib
from sceneego_b200.utils import synth
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cam():
    from sceneego_b200.utils.fisheye.FishEyeCalibrated import FishEyeCameraCalibrated
    return FishEyeCameraCalibrated(util.CALIB)


@pytest.fixture(scope="module")
def tables64():
    return orc.StageTables(util.CALIB, 64, 2.0)


def test_ray_table_bit_exact(cam, tables64):
    ray = cam.ray_table_device(1280, 1024).cpu().numpy()           # (H, W, 3) row-major
    ref = tables64.ray.reshape(1280, 1024, 3).transpose(1, 0, 2)   # oracle is x-major
    assert np.array_equal(ray, ref)                                # every one of 1,310,720 pixels, fp64 bits
    g = util.golden("tables_v64.npz")                              # vectors from the unmodified reference
    flat = ray.transpose(1, 0, 2).reshape(-1, 3)
    assert np.array_equal(flat[g["ray_idx"]], g["ray"])


@pytest.mark.parametrize("V", [64, 128])
def test_project_voxels(cam, V):
    from sceneego_b200 import _lib
    px, grid = _lib.project_voxels(cam.calib_struct(1280, 1024), V, 2.0, (1024, 1280), "cuda")
    t = orc.StageTables(util.CALIB, V, 2.0)
    # fp32 atan/sqrt differ from the CPU libm by an ulp or two: 1e-3 px on a 1280 px image
    assert (px.cpu() - t.grid_px).abs().max().item() <= 1e-3
    assert (grid.cpu() - t.grid.reshape(-1, 2)).abs().max().item() <= 2e-6
    g = util.golden(f"tables_v{V}.npz")
    assert np.abs(px.cpu().numpy()[g["vox_idx"]] - g["grid_px"]).max() <= 1e-3
    # arbitrary-point entry agrees with the fused-coordinate entry
    px2 = cam.world2camera_pytorch(t.coord_volume.reshape(-1, 3).cuda())
    assert (px2 - px).abs().max().item() <= 1e-3


def test_project_voxels_raises_on_axis(cam):
    from sceneego_b200 import _lib
    with pytest.raises(Exception, match="norm is zero"):
        _lib.project_voxels(cam.calib_struct(1280, 1024), 33, 2.0, (1024, 1280), "cuda")   # odd V: voxel on the axis


def _voxelize(cam, depth, V, side=2.0):
    from sceneego_b200 import _lib
    ray = cam.ray_table_device(1280, 1024)
    d = torch.as_tensor(depth).float().cuda()
    if d.dim() == 2:
        d = d[None]
    occ = torch.zeros(d.shape[0], V, V, V, device="cuda")
    lay = _lib.vol_layout(V, 3, d.shape[0])
    buf = _lib.alloc_volume(lay, 48, "cuda")
    _lib.voxelize_depth(d.contiguous(), ray, 1024, 1280, V, side, occ, buf, lay, channel=32)
    planar = _lib.unpack_volume(buf, lay, d.shape[0], 40)[:, 32]
    assert torch.equal(planar, occ), "planar bf16 scene channel differs from the f32 grid"
    # the stem's space-to-depth input: the 2x2x2 block's occupancy bits share one 16-byte cell
    lay2 = _lib.vol_layout_s2d(V, d.shape[0])
    buf2 = _lib.alloc_volume(lay2, 33 * 8, "cuda")
    _lib.voxelize_depth(d.contiguous(), ray, 1024, 1280, V, side, None, buf2, lay2, channel=32)
    assert torch.equal(_lib.unpack_volume(buf2, lay2, d.shape[0], 33)[:, 32], occ), "s2d scene channel differs"
    assert buf2[:32].abs().max().item() == 0.0, "s2d voxelisation touched a feature plane"
    # the marching stem's input: occupancy as a z-window plane (cell (x,y,z) = occ[x][y][z-3 .. z+4])
    lay3 = _lib.vol_layout_zwin(V, d.shape[0])
    buf3 = _lib.alloc_volume(lay3, 40, "cuda")
    _lib.voxelize_depth(d.contiguous(), ray, 1024, 1280, V, side, None, buf3, lay3, channel=32)
    assert torch.equal(_lib.unpack_volume(buf3, lay3, d.shape[0], 33)[:, 32], occ), "z-window scene channel differs"
    assert buf3[:4].abs().max().item() == 0.0, "z-window voxelisation touched a feature plane"
    ref3 = _lib.alloc_volume(lay3, 40, "cuda")
    _lib.pack_volume(occ.unsqueeze(1).contiguous(), ref3, lay3, c_offset=32)      # gather form of the same plane
    assert torch.equal(buf3[4], ref3[4]), "scattered z-window cells differ from the packed ones"
    # the path lift() takes: one store per pixel into a plain f32 grid, then the whole plane from that grid in one
    # pass that also leaves the grid all-zero for the next batch
    scratch = occ.clone()
    buf4 = _lib.alloc_volume(lay3, 40, "cuda")
    buf4[4].fill_(7.0)                                             # stale cells must all be overwritten
    _lib.occ_expand_zwin(scratch, buf4, lay3, 32)
    written = buf4[4] != 7.0                                       # all eight entries of every real cell, nothing else
    assert int(written.sum().item()) == d.shape[0] * V ** 3 * 8, "expand: not exactly the real cells were written"
    assert torch.equal(buf4[4][written], buf3[4][written]), "expanded z-window cells differ from the scattered ones"
    assert scratch.abs().max().item() == 0.0, "expand: the grid is not all-zero afterwards"
    assert buf4[:4].abs().max().item() == 0.0, "expand touched a feature plane"
    return occ.cpu().numpy()


@pytest.mark.parametrize("V", [64, 128])
def test_voxelize_demo_frames_bit_exact(cam, V):
    g = util.golden("voxel.npz")
    for name in ("img_001000", "img_001796", "img_002376"):
        raw = g[f"{name}_raw"]
        d = orc.resize_nearest(raw, 1024, 1280).copy()   # dataset/demo_dataset.py:86-91
        d[d > 10] = 10
        got = _voxelize(cam, d, V)[0]
        ref = util.unpack_bits(g[f"{name}_v{V}"], V)
        assert np.array_equal(got, ref), name
        assert got[V // 2, V // 2, 0] == 1.0


@pytest.mark.parametrize("V", [64, 128])
def test_voxelize_synthetic_bit_exact(cam, tables64, V):
    g = util.golden("voxel.npz")
    cases = {"uniform": synth.synthetic_depth_uniform(1)[0].numpy(),
             "room": synth.synthetic_depth_room(1, tables64.ray)[0].numpy(),
             "uniform1024": synth.synthetic_depth_uniform(1, h=1024, w=1024)[0].numpy()}
    for tag, d in cases.items():
        assert np.array_equal(_voxelize(cam, d, V)[0], util.unpack_bits(g[f"{tag}_v{V}"], V)), tag


def test_voxelize_edge_cases_vs_oracle(cam, tables64):
    rng = np.random.default_rng(3)
    cases = [np.zeros((1024, 1280), np.float32),                         # empty: only voxel (V/2,V/2,0)
             np.full((1024, 1280), 10.0, np.float32),                    # clamp value everywhere
             (rng.random((256, 320), dtype=np.float32) * 3),             # ragged small input, nearest upsample
             (rng.random((1024, 1280), dtype=np.float32) * 6 - 1)]       # negative depths
    weird = rng.random((1024, 1280), dtype=np.float32) * 2
    weird[::7, ::5] = np.nan
    weird[::11, ::3] = np.inf
    cases.append(weird)
    for d in cases:
        ref = orc.voxelize_depth(d, tables64.ray, 64, 2.0)
        assert np.array_equal(_voxelize(cam, d, 64)[0], ref)
    assert _voxelize(cam, cases[0], 64)[0].sum() == 1.0
    # cuboid sides that are not a power of two take the fp64-divide path (2.0 takes the exact-multiply path)
    for side in (3.0, 1.7, 4.0):
        for d in (cases[3], cases[4]):
            assert np.array_equal(_voxelize(cam, d, 64, side)[0], orc.voxelize_depth(d, tables64.ray, 64, side)), side


def test_voxelize_raw_depth_fused_preprocessing(cam, tables64):
    """sceneego_voxelize_depth_raw_f64: the dataset's nearest resize to 1280x1024 and the 10 m clamp
    (dataset/demo_dataset.py:86-91) fused into the voxelisation load.  Raw demo EXRs (512x640) against the
    reference-generated occupancy; odd source sizes (where OpenCV's index map differs from the exact-rational one),
    values above the clamp, NaN / Inf against the oracle."""
    g = util.golden("voxel.npz")
    ray = cam.ray_table_device(1280, 1024, "cuda")
    for V in (64, 128):
        raws = np.stack([g[f"{n}_raw"] for n in ("img_001000", "img_001796", "img_002376")])
        occ = torch.zeros(3, V, V, V, device="cuda")
        _lib.voxelize_depth_raw(torch.from_numpy(raws).cuda(), (1024, 1280), 10.0, ray, 1024, 1280, V, 2.0, occ, None, None)
        for i, n in enumerate(("img_001000", "img_001796", "img_002376")):
            assert np.array_equal(occ[i].cpu().numpy(), util.unpack_bits(g[f"{n}_v{V}"], V)), n
    rng = np.random.default_rng(11)
    for h, w in ((104, 144), (26, 36), (333, 517), (1024, 1280), (2048, 2560)):
        raw = rng.random((2, h, w), dtype=np.float32) * 14 - 1          # above the clamp and below zero
        raw[0, ::5, ::3] = np.nan
        raw[1, ::7, ::2] = np.inf
        occ = torch.zeros(2, 64, 64, 64, device="cuda")
        _lib.voxelize_depth_raw(torch.from_numpy(raw).cuda(), (1024, 1280), 10.0, ray, 1024, 1280, 64, 2.0, occ, None, None)
        for i in range(2):
            ref = orc.voxelize_depth(orc.preprocess_depth(raw[i]), tables64.ray, 64, 2.0)
            assert np.array_equal(occ[i].cpu().numpy(), ref), (h, w, i)
    # network-semantics batch helper: raw maps -> the grids the NETWORK builds (for scene_volumes=)
    from sceneego_b200.dataset import real_depth_utils as rdu
    vox2 = rdu.depth_maps_to_voxels(ray, torch.from_numpy(g["img_001000_raw"][None]).cuda(), 2.0, 64,
                                    network_semantics=True)
    assert np.array_equal(vox2[0].cpu().numpy(), util.unpack_bits(g["img_001000_v64"], 64))


def test_dataset_depth_map_to_voxel_vs_reference(cam, tables64):
    """dataset/real_depth_utils.depth_map_to_voxel (the voxel_output=True path of demo_dataset.py:93-94 /
    test_dataset.py:145-146): ray table x the full 1280-wide map, pixel for pixel -- NOT the network's squash-and-pad.
    Against occupancy grids produced by the reference's own function (tests/golden/voxel_dataset.npz), V = 64 / 128,
    bit-exact; plus raw maps with the dataset's resize + clamp fused, NaN / Inf / negative values vs the oracle."""
    from sceneego_b200.dataset import real_depth_utils as rdu
    g, gd = util.golden("voxel.npz"), util.golden("voxel_dataset.npz")
    maps = {n: orc.preprocess_depth(g[f"{n}_raw"]) for n in ("img_001000", "img_001796", "img_002376")}
    maps["room"] = synth.synthetic_depth_room(1, tables64.ray)[0].numpy()
    maps["uniform"] = synth.synthetic_depth_uniform(1)[0].numpy()
    for V in (64, 128):
        for name, d in maps.items():
            want = util.unpack_bits(gd[f"{name}_v{V}"], V)
            vox = rdu.depth_map_to_voxel(tables64.ray, torch.from_numpy(d), 2.0, V)     # reference-order NumPy ray table
            assert np.array_equal(vox.cpu().numpy(), want), (name, V)
            assert np.array_equal(orc.voxelize_depth_dataset(d, tables64.ray, V, 2.0), want), (name, V)
    # the two semantics really differ (7782 vs 8331 occupied voxels on img_001000)
    assert int(util.unpack_bits(gd["img_001000_v64"], 64).sum()) == 7782
    assert int(util.unpack_bits(g["img_001000_v64"], 64).sum()) == 8331
    # raw maps: dataset resize + clamp fused into the same launch, batch of 3
    raws = np.stack([g[f"{n}_raw"] for n in ("img_001000", "img_001796", "img_002376")])
    ray = cam.ray_table_device(1280, 1024, "cuda")
    occ = rdu.depth_maps_to_voxels(ray, torch.from_numpy(raws).cuda(), 2.0, 64)
    for i, n in enumerate(("img_001000", "img_001796", "img_002376")):
        assert np.array_equal(occ[i].cpu().numpy(), util.unpack_bits(gd[f"{n}_v64"], 64)), n
    rng = np.random.default_rng(13)
    for h, w in ((104, 144), (333, 517), (1024, 1280)):
        raw = rng.random((2, h, w), dtype=np.float32) * 14 - 1
        raw[0, ::5, ::3] = np.nan
        raw[1, ::7, ::2] = np.inf
        occ = rdu.depth_maps_to_voxels(ray, torch.from_numpy(raw).cuda(), 2.0, 64)
        for i in range(2):
            ref = orc.voxelize_depth_dataset(orc.preprocess_depth(raw[i]), tables64.ray, 64, 2.0)
            assert np.array_equal(occ[i].cpu().numpy(), ref), (h, w, i)
    with pytest.raises(_lib.SceneEgoError):
        rdu.depth_maps_to_voxels(ray, torch.zeros(1, 512, 640, device="cuda"), 2.0, 64, preprocess=False)


def test_voxelize_odd_source_size_matches_cv2_rule(cam, tables64):
    """Model-side resize (network/voxel_net_depth.py:197) of a depth map whose size is one of those where the
    exact-rational nearest map differs from OpenCV's."""
    rng = np.random.default_rng(12)
    d = rng.random((198, 186), dtype=np.float32) * 4
    assert not np.array_equal(orc.nearest_index(198, 1024), np.minimum(np.arange(1024) * 198 // 1024, 197))
    assert np.array_equal(_voxelize(cam, d, 64)[0], orc.voxelize_depth(d, tables64.ray, 64, 2.0))


def test_voxelize_batch_consistency_full_size(cam, tables64):
    """BASELINE config sizes (B=64 frames of 1024x1280): frame i of a batch equals the frame alone,
    every frame marks voxel (V/2,V/2,0), occupancy is {0,1}."""
    d = torch.cat([synth.synthetic_depth_room(32, tables64.ray, seed=1), synth.synthetic_depth_uniform(32, seed=2)])
    occ = _voxelize(cam, d, 64)
    assert set(np.unique(occ)) <= {0.0, 1.0}
    assert (occ[:, 32, 32, 0] == 1).all()
    for i in (0, 31, 32, 63):
        assert np.array_equal(occ[i], _voxelize(cam, d[i], 64)[0])
    assert np.array_equal(occ[5], orc.voxelize_depth(d[5].numpy(), tables64.ray, 64, 2.0))


def _pf(sd, feat):
    return orc.process_features(feat, sd["process_features.0.weight"], sd["process_features.0.bias"])


def test_unproject_vs_reference_golden(cam, tables64):
    from sceneego_b200 import _lib
    sd = synth.synthetic_state_dict(util.stage_shapes(), seed=0, mode="random_bn")
    feat = synth.synthetic_features(2)
    w, b = sd["process_features.0.weight"].cuda(), sd["process_features.0.bias"].cuda()
    feat32 = _lib.feature_conv1x1(feat.cuda(), w, b)
    ref32 = torch.nn.functional.conv2d(feat, sd["process_features.0.weight"], sd["process_features.0.bias"])
    assert (feat32.cpu().permute(0, 3, 1, 2) - ref32).abs().max().item() <= 1e-4 * ref32.abs().max().item()
    g = util.golden("unproject_v64.npz")
    ref = torch.from_numpy(g["lifted"])                      # from the unmodified reference
    idx = torch.from_numpy(g["vox_idx"])
    lay = _lib.vol_layout(64, 3, 2)
    for fused in (False, True):
        out = torch.empty(2, 32, 64, 64, 64, device="cuda")
        buf = _lib.alloc_volume(lay, 48, "cuda")
        grid = None if fused else tables64.grid.reshape(-1, 2).cuda().contiguous()
        _lib.unproject(feat32, grid, cam.calib_struct(1280, 1024) if fused else None, 64, 2.0, 1024, 1280, out, buf,
                       lay, extra_zero_planes=2)
        got = out.reshape(2, 32, -1)[:, :, idx].cpu()
        rel = ((got - ref).norm() / ref.norm()).item()
        assert rel <= 1e-3, (fused, rel)                    # north-star tolerance
        if not fused:
            assert torch.allclose(got, ref, rtol=1e-4, atol=1e-4)
        else:
            assert torch.allclose(got, ref, rtol=1e-3, atol=2e-3)
        planar = _lib.unpack_volume(buf, lay, 2, 48)
        assert torch.equal(planar[:, :32], out.to(torch.bfloat16).float())   # same values, bf16-rounded
        assert planar[:, 32:].abs().max().item() == 0.0
        # same gather written in the stem's space-to-depth layout; the occupancy plane is cleared
        lay2 = _lib.vol_layout_s2d(64, 2)
        buf2 = _lib.alloc_volume(lay2, 33 * 8, "cuda")
        buf2[32].fill_(1.0)
        _lib.unproject(feat32, grid, cam.calib_struct(1280, 1024) if fused else None, 64, 2.0, 1024, 1280, None, buf2,
                       lay2, extra_zero_planes=1)
        s2d = _lib.unpack_volume(buf2, lay2, 2, 33)
        assert torch.equal(s2d[:, :32], planar[:, :32]) and s2d[:, 32].abs().max().item() == 0.0
        # ... and in the marching stem's z-window layout (one occupancy plane cleared behind the features)
        lay3 = _lib.vol_layout_zwin(64, 2)
        buf3 = _lib.alloc_volume(lay3, 40, "cuda")
        buf3[4].fill_(1.0)
        _lib.unproject(feat32, grid, cam.calib_struct(1280, 1024) if fused else None, 64, 2.0, 1024, 1280, None, buf3,
                       lay3, extra_zero_planes=1)
        zw = _lib.unpack_volume(buf3, lay3, 2, 33)
        assert torch.equal(zw[:, :32], planar[:, :32]) and zw[:, 32].abs().max().item() == 0.0


@pytest.mark.parametrize("B,h,w", [(2, 64, 64), (3, 20, 27), (1, 8, 8)])
def test_feature_conv1x1_tensor_core_split_vs_fp64(B, h, w):
    """process_features[0] (Conv2d(256,32,1), network/voxel_net_depth.py:58-63) on tcgen05 with operands split into two
    bf16 parts (three MMAs per product): fp32-grade accuracy -- within 2e-5 of the output range of an fp64 evaluation,
    where the plain fp32 CUDA-core kernel (SCENEEGO_FEATURE_CONV_SIMT=1) sits at ~1e-6 -- incl. a ragged last tile."""
    import os
    from sceneego_b200 import _lib
    g = torch.Generator().manual_seed(B * 100 + h)
    x = (torch.randn(B, 256, h, w, generator=g).abs() * 3.0).cuda()
    wt = (torch.randn(32, 256, 1, 1, generator=g) * 0.08).cuda()
    bs = (torch.randn(32, generator=g) * 0.1).cuda()
    ref = torch.nn.functional.conv2d(x.double(), wt.double(), bs.double()).permute(0, 2, 3, 1)
    rng = (ref.max() - ref.min()).item()
    got = _lib.feature_conv1x1(x, wt, bs)
    os.environ["SCENEEGO_FEATURE_CONV_SIMT"] = "1"
    try:
        simt = _lib.feature_conv1x1(x, wt, bs)
    finally:
        del os.environ["SCENEEGO_FEATURE_CONV_SIMT"]
    e_tc = (got.double() - ref).abs().max().item() / rng
    e_simt = (simt.double() - ref).abs().max().item() / rng
    print(f"feature_conv1x1 B={B} {h}x{w}: max-abs/range vs fp64: tensor-core split {e_tc:.2e}, fp32 CUDA-core {e_simt:.2e}")
    assert got.shape == (B, h, w, 32) and e_tc <= 2e-5 and e_simt <= 5e-6
    assert torch.equal(_lib.feature_conv1x1(x, wt, bs), got)            # deterministic


def test_materialised_features_and_generic_grid_sample(tables64):
    """Output #2 of the reference forward and the op-level drop-in: gathering from the
    materialised 1024x1280 map with the generic kernel equals the fused gather (loop == batch)."""
    from sceneego_b200 import _lib
    from sceneego_b200.utils import op
    sd = synth.synthetic_state_dict(util.stage_shapes(), seed=0, mode="random_bn")
    feat = synth.synthetic_features(2, seed=9)
    feat32 = _lib.feature_conv1x1(feat.cuda(), sd["process_features.0.weight"].cuda(),
                                  sd["process_features.0.bias"].cuda())
    full = _lib.features_upsample_pad(feat32, 1024, 128)
    ref_full = _pf(sd, feat)
    assert full.shape == (2, 32, 1024, 1280)
    assert (full.cpu() - ref_full).abs().max().item() <= 1e-4 * ref_full.abs().max().item()
    grid_b = op.get_grid_coord_proj_batch(tables64.grid_px.cuda(), 4, (1024, 1280))
    assert grid_b.shape == (4, 64 ** 3, 1, 2) and grid_b.stride(0) == 0
    assert torch.equal(grid_b[0].cpu(), tables64.grid)       # bit-identical to the reference's CPU table
    lifted = op.unproject_heatmaps_one_view_batch(full, grid_b, 64)
    ref = orc.unproject(ref_full, tables64.grid.unsqueeze(0).expand(2, -1, -1, -1), 64)
    # inputs differ by the 256-term fp32 summation order of the 1x1 conv (GPU vs CPU)
    assert ((lifted.cpu() - ref).norm() / ref.norm()).item() <= 1e-5
    assert torch.allclose(lifted.cpu(), ref, rtol=1e-3, atol=1e-3)
    one = op.unproject_heatmaps_one_view(full[1:2], tables64.grid_px.cuda(), 64)
    assert (one - lifted[1:2]).abs().sum().item() == 0.0     # the reference's own loop==batch smoke check
    fused = torch.empty(2, 32, 64, 64, 64, device="cuda")
    _lib.unproject(feat32, tables64.grid.reshape(-1, 2).cuda().contiguous(), None, 64, 2.0, 1024, 1280, fused, None, None)
    # same grid, same values, different association of the four products (fma chain vs sum)
    d = (fused - lifted).abs().max().item()
    assert ((fused - lifted).norm() / lifted.norm()).item() <= 1e-6 and d <= 1e-5 * lifted.abs().max().item(), d


def test_generic_grid_sample_out_of_bounds():
    from sceneego_b200 import _lib
    g = torch.Generator().manual_seed(0)
    img = torch.randn(2, 5, 37, 53, generator=g)
    grid = (torch.rand(2, 4001, 1, 2, generator=g) * 2.6 - 1.3)          # 15% outside [-1,1]: zeros padding
    ref = orc.grid_sample_bilinear(img, grid)
    ref2 = torch.nn.functional.grid_sample(img, grid, align_corners=True).reshape(2, 5, -1)
    assert torch.allclose(ref, ref2, atol=1e-5)
    gd = grid.cuda().contiguous()
    got = _lib.grid_sample(img.cuda(), gd, gd.stride(0))
    assert torch.allclose(got.cpu(), ref, rtol=1e-5, atol=1e-5)
