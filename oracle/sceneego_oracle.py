"""CPU oracle for the SceneEgo volumetric lifting stage.

TEST INFRASTRUCTURE ONLY.  Nothing under ``sceneego_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker / the
timed CPU baseline -- never as the product path.

It restates, in NumPy / CPU-torch, the algorithm of the reference hot path
(``/root/reference/network/voxel_net_depth.py:237-273``) and of the callers either
side of it that the product also covers (dataset depth preprocessing, ``V2VModelSimple``,
the MPJPE / PA-MPJPE evaluation).  Each function cites the reference lines it follows.  Parity pin: the reference ships no tests or
golden vectors, so this oracle is pinned against outputs
of the unmodified reference imported in the build container -- see
``tests/make_golden.py`` (generator) and ``tests/golden/*.npz`` (committed
vectors), checked by ``tests/test_oracle_golden.py``.

Scatter semantics: ``voxel[idx.T] = 1`` is evaluated with the PyTorch 1.13.1
"sequence is a tuple" rule, i.e. one (x, y, z) write per point (SURVEY.md
appendix C); the literal line under torch >= 2.9 fills whole x-planes.
"""
from __future__ import annotations

import json
import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------
# camera model (utils/fisheye/FishEyeCalibrated.py)
# ----------------------------------------------------------------------------
def load_calibration(path: str) -> Dict[str, np.ndarray]:
    """FishEyeCalibrated.py:8-16 -- JSON -> centre + the two polynomials."""
    with open(path) as f:
        d = json.load(f)
    intr = np.array(d["intrinsic"], dtype=np.float64)
    return {
        "size": np.array(d["size"], dtype=np.int64),           # (w, h)
        "center": np.array([intr[0][2], intr[1][2]]),           # (cx, cy)
        "c2w": np.array(d["polynomialC2W"], dtype=np.float64),  # ascending powers
        "w2c": np.array(d["polynomialW2C"], dtype=np.float64),  # ascending powers
    }


def ray_table(calib, width: int, height: int) -> np.ndarray:
    """Unit ray per pixel, fp64, row index = x*height + y.

    voxel_net_depth.py:147-155 (pixel grid, x-major) + FishEyeCalibrated.py:36-51
    (np.polyval Horner with separate mul/add; norm = sqrt((x^2+y^2)+z^2)).
    """
    cx, cy = calib["center"]
    xs = np.repeat(np.arange(width, dtype=np.float64), height)
    ys = np.tile(np.arange(height, dtype=np.float64), width)
    xc = xs - cx
    yc = ys - cy
    d = np.sqrt(xc * xc + yc * yc)
    z = np.zeros_like(d)
    for c in calib["c2w"][::-1]:          # highest power first
        z = z * d + c
    nz = -z
    norm = np.sqrt((xc * xc + yc * yc) + nz * nz)
    return np.stack([xc / norm, yc / norm, nz / norm], axis=1)


def build_coord_volume(volume_size: int, cuboid_side: float) -> torch.Tensor:
    """Voxel-centre coordinates (V,V,V,3) f32 -- voxel_net_depth.py:110-134.

    The reference multiplies a NumPy float64 scalar into an fp32 torch tensor;
    torch treats it as a Python scalar, so the arithmetic is fp32:
    fl32(fl32(step) * i) then + fl32(lo), two roundings (checked against the
    reference in tests/make_golden.py).
    """
    V = volume_size
    idx = torch.arange(V).float()
    step = float(cuboid_side / (V - 1))
    axis_xy = float(-cuboid_side / 2) + step * idx
    axis_z = 0.0 + step * idx
    gx, gy, gz = torch.meshgrid(axis_xy, axis_xy, axis_z, indexing="ij")
    return torch.stack([gx, gy, gz], dim=-1).contiguous()


def world2camera_f32(calib, points: torch.Tensor) -> torch.Tensor:
    """3D -> pixel through the Scaramuzza inverse polynomial, fp32.

    FishEyeCalibrated.py:137-174: z <- -z, r = |(x,y)|, theta = atan(z/r),
    rho = sum a_n theta^n (running power, ascending), p = (x,y)/r*rho + centre.
    Raises like the reference when any r == 0.
    """
    p = points.clone().float()
    p[:, 2] = points[:, 2] * -1
    pt = p.transpose(0, 1)            # (3,N) strided view, like the reference: torch.norm's
    x, y, z = pt[0], pt[1], pt[2]     # rounding depends on the memory layout (1 ulp)
    norm = torch.norm(pt[:2], dim=0)
    if not bool((norm != 0).all()):
        raise Exception("norm is zero!")
    theta = torch.atan(z / norm)
    inv = 1.0 / norm
    coef = calib["w2c"]
    rho = float(coef[0])
    t_i = 1.0
    for i in range(1, len(coef)):
        t_i = t_i * theta
        rho = rho + t_i * float(coef[i])
    cx = torch.tensor([calib["center"][0]], dtype=torch.float32)
    cy = torch.tensor([calib["center"][1]], dtype=torch.float32)
    return torch.stack([x * inv * rho + cx, y * inv * rho + cy], dim=1)


def normalise_grid(grid_px: torch.Tensor, heatmap_shape) -> torch.Tensor:
    """utils/op.py:177-184 -- g = 2*(p/[W,H] - 0.5); returns (N,1,2)."""
    g = torch.zeros_like(grid_px)
    g[:, 0] = 2 * (grid_px[:, 0] / heatmap_shape[1] - 0.5)
    g[:, 1] = 2 * (grid_px[:, 1] / heatmap_shape[0] - 0.5)
    return g.unsqueeze(1)


# ----------------------------------------------------------------------------
# a1 / a2: process_features + unprojection
# ----------------------------------------------------------------------------
def process_features(feat256: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor,
                     up: int = 1024, pad: int = 128) -> torch.Tensor:
    """voxel_net_depth.py:58-63 -- 1x1 conv, nearest upsample to up x up
    (src = dst * in // up), zero-pad `pad` columns left and right."""
    x = F.conv2d(feat256, weight, bias)
    h, w = x.shape[-2:]
    iy = (torch.arange(up) * h) // up
    ix = (torch.arange(up) * w) // up
    x = x[:, :, iy][:, :, :, ix]
    return F.pad(x, (pad, pad, 0, 0))


def grid_sample_bilinear(img: torch.Tensor, grid: torch.Tensor) -> torch.Tensor:
    """ATen grid_sampler_2d, bilinear / zeros padding / align_corners=True,
    as called at utils/op.py:209.  img (B,C,H,W); grid (B,N,1,2) in [-1,1]
    -> (B,C,N).  Restated tap by tap so the CUDA kernel has a reference for
    every intermediate."""
    B, C, H, W = img.shape
    gx = grid[..., 0].reshape(B, -1)
    gy = grid[..., 1].reshape(B, -1)
    ix = ((gx + 1) / 2) * (W - 1)
    iy = ((gy + 1) / 2) * (H - 1)
    x0 = torch.floor(ix)
    y0 = torch.floor(iy)
    x1 = x0 + 1
    y1 = y0 + 1
    w_nw = (x1 - ix) * (y1 - iy)
    w_ne = (ix - x0) * (y1 - iy)
    w_sw = (x1 - ix) * (iy - y0)
    w_se = (ix - x0) * (iy - y0)
    flat = img.reshape(B, C, H * W)
    out = torch.zeros(B, C, gx.shape[1], dtype=img.dtype)

    def tap(xs, ys, wgt):
        ok = (xs >= 0) & (xs <= W - 1) & (ys >= 0) & (ys <= H - 1)
        lin = (ys.clamp(0, H - 1) * W + xs.clamp(0, W - 1)).long()
        val = torch.gather(flat, 2, lin.unsqueeze(1).expand(B, C, -1))
        return val * (wgt * ok).unsqueeze(1)

    out = tap(x0, y0, w_nw) + tap(x1, y0, w_ne) + tap(x0, y1, w_sw) + tap(x1, y1, w_se)
    return out


def unproject(features: torch.Tensor, grid_batch: torch.Tensor, volume_size: int) -> torch.Tensor:
    """utils/op.py:194-214 -- gather then view (B,C,V,V,V), flat index x*V*V+y*V+z."""
    B, C = features.shape[:2]
    V = volume_size
    return grid_sample_bilinear(features, grid_batch[:B]).reshape(B, C, V, V, V)


# ----------------------------------------------------------------------------
# a5: depth map -> occupancy grid
# ----------------------------------------------------------------------------
def nearest_index(n_src: int, n_dst: int) -> np.ndarray:
    """Index map of cv2.resize(..., INTER_NEAREST) (OpenCV resizeNN): src = min(cvFloor(dst * ifx), n_src - 1)
    with ifx = 1. / ((double)n_dst / n_src) in fp64 -- pinned against cv2 itself for 3500 size pairs by
    tests/make_golden.py (tests/golden/nearest_index.npz); floor(dst * n_src / n_dst) in exact arithmetic is
    NOT the same map (115 source sizes below 1400 differ)."""
    ifx = 1.0 / (float(n_dst) / float(n_src))
    return np.minimum(np.floor(np.arange(n_dst, dtype=np.float64) * ifx).astype(np.int64), n_src - 1)


def resize_nearest(depth: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """cv2.resize(depth, (out_w, out_h), interpolation=cv2.INTER_NEAREST)."""
    h, w = depth.shape
    return depth[nearest_index(h, out_h)][:, nearest_index(w, out_w)]


def preprocess_depth(depth_raw: np.ndarray, pre_h: int = 1024, pre_w: int = 1280, clamp_max: float = 10.0) -> np.ndarray:
    """dataset/demo_dataset.py:86-91 (= dataset/test_dataset.py:138-143): nearest resize to 1280 x 1024 when the
    decoded map has another size, then depth_map[depth_map > 10] = 10."""
    d = np.asarray(depth_raw, dtype=np.float32)
    if d.shape[0] != pre_h or d.shape[1] != pre_w:
        d = resize_nearest(d, pre_h, pre_w)
    d = d.copy()
    d[d > clamp_max] = clamp_max
    return d


def voxelize_depth(depth: np.ndarray, ray: np.ndarray, volume_size: int, cuboid_side: float,
                   image_height: int = 1024, pad: int = 128) -> np.ndarray:
    """voxel_net_depth.py:194-222 -- depth (h,w) f32 -> (V,V,V) f32 {0,1}.

    nearest-resize to H x H, pad `pad` zero columns each side, flatten x-major,
    P = ray * depth (fp64), q_xy = (P + s/2) * V / s, q_z = P_z * V / s,
    round half-to-even, keep iff 0 <= q <= V-1 on all axes, voxel[q] = 1.
    """
    V, s = volume_size, cuboid_side
    d = resize_nearest(np.asarray(depth, dtype=np.float32), image_height, image_height)
    d = np.pad(d, ((0, 0), (pad, pad)), "constant", constant_values=0)
    flat = d.T.reshape(-1)                       # fp32, index x*H + y
    pc = (ray.T * flat).T                         # fp64
    q = np.empty_like(pc)
    q[:, 0] = (pc[:, 0] + s / 2) * V / s
    q[:, 1] = (pc[:, 1] + s / 2) * V / s
    q[:, 2] = pc[:, 2] * V / s
    q = np.round(q)
    ok = np.all(np.logical_and(V - 1 >= q, q >= 0), axis=1)
    qi = q[ok].astype(np.int64)
    vox = np.zeros((V, V, V), dtype=np.float32)
    vox[qi[:, 0], qi[:, 1], qi[:, 2]] = 1.0
    return vox


def voxelize_depth_dataset(depth: np.ndarray, ray: np.ndarray, volume_size: int, cuboid_side: float) -> np.ndarray:
    """dataset/real_depth_utils.py:29-60 (`depth_map_to_voxel` + `point_cloud_to_voxel_pytorch`), the function the
    datasets call with `voxel_output=True` (dataset/demo_dataset.py:93-94, dataset/test_dataset.py:145-146).

    NOT the network's `depth_map_to_voxel_numpy`: the preprocessed (image_height, image_width) map is multiplied by
    the ray table pixel for pixel -- no nearest squash to H x H and no zero-padded columns (:31-33).  The
    quantisation (:46-56) is the same as voxel_net_depth.py:207-222.
    """
    V, s = volume_size, cuboid_side
    flat = np.asarray(depth, dtype=np.float32).T.reshape(-1)   # index x*H + y, like the x-major ray table
    pc = (ray.T * flat).T                                       # fp64
    q = np.empty_like(pc)
    q[:, 0] = (pc[:, 0] + s / 2) * V / s
    q[:, 1] = (pc[:, 1] + s / 2) * V / s
    q[:, 2] = pc[:, 2] * V / s
    q = np.round(q)
    ok = np.all(np.logical_and(V - 1 >= q, q >= 0), axis=1)
    qi = q[ok].astype(np.int64)
    vox = np.zeros((V, V, V), dtype=np.float32)
    vox[qi[:, 0], qi[:, 1], qi[:, 2]] = 1.0
    return vox


# ----------------------------------------------------------------------------
# a7: V2V encoder-decoder (network/v2v.py), functional, fp32
# ----------------------------------------------------------------------------
def _bn(sd, p, x, eps=1e-5):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"],
                        sd[p + ".weight"], sd[p + ".bias"], False, 0.0, eps)


def _basic(sd, p, x):
    """v2v.py:8-18 conv(k, same) - BN - ReLU."""
    w = sd[p + ".block.0.weight"]
    x = F.conv3d(x, w, sd[p + ".block.0.bias"], padding=(w.shape[-1] - 1) // 2)
    return F.relu(_bn(sd, p + ".block.1", x))


def _res(sd, p, x):
    """v2v.py:21-43 conv3-BN-ReLU-conv3-BN (+ conv1-BN skip), add, ReLU."""
    r = F.conv3d(x, sd[p + ".res_branch.0.weight"], sd[p + ".res_branch.0.bias"], padding=1)
    r = F.relu(_bn(sd, p + ".res_branch.1", r))
    r = F.conv3d(r, sd[p + ".res_branch.3.weight"], sd[p + ".res_branch.3.bias"], padding=1)
    r = _bn(sd, p + ".res_branch.4", r)
    if (p + ".skip_con.0.weight") in sd:
        x = _bn(sd, p + ".skip_con.1", F.conv3d(x, sd[p + ".skip_con.0.weight"], sd[p + ".skip_con.0.bias"]))
    return F.relu(r + x)


def _up(sd, p, x):
    """v2v.py:55-67 ConvTranspose3d(k2,s2) - BN - ReLU."""
    x = F.conv_transpose3d(x, sd[p + ".block.0.weight"], sd[p + ".block.0.bias"], stride=2)
    return F.relu(_bn(sd, p + ".block.1", x))


def v2v_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, prefix: str = "") -> torch.Tensor:
    """network/v2v.py:104-139,165-170 with a state dict keyed like the module
    (optionally under `prefix`, e.g. 'volume_net.')."""
    if prefix:
        sd = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    x = _basic(sd, "front_layers.0", x)
    for i in (1, 2, 3):
        x = _res(sd, f"front_layers.{i}", x)
    e = "encoder_decoder."
    skips = []
    for lvl in range(1, 6):
        skips.append(_res(sd, f"{e}skip_res{lvl}", x))
        x = F.max_pool3d(x, 2, 2)
        x = _res(sd, f"{e}encoder_res{lvl}", x)
    x = _res(sd, e + "mid_res", x)
    for lvl in range(5, 0, -1):
        x = _res(sd, f"{e}decoder_res{lvl}", x)
        x = _up(sd, f"{e}decoder_upsample{lvl}", x)
        x = x + skips[lvl - 1]
    x = _res(sd, "back_layers.0", x)
    x = _basic(sd, "back_layers.1", x)
    x = _basic(sd, "back_layers.2", x)
    return F.conv3d(x, sd["output_layer.weight"], sd["output_layer.bias"])


def v2v_simple_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor) -> torch.Tensor:
    """network/v2v.py:184-241 -- V2VModelSimple / EncoderDecoderSimple (two pooling levels, one 1x1 back layer)."""
    e = "encoder_decoder."
    x = _basic(sd, "front_layers.0", x)
    skip1 = _res(sd, e + "skip_res1", x)
    x = _res(sd, e + "encoder_res1", F.max_pool3d(x, 2, 2))
    skip2 = _res(sd, e + "skip_res2", x)
    x = _res(sd, e + "encoder_res2", F.max_pool3d(x, 2, 2))
    x = _res(sd, e + "mid_res", x)
    x = _res(sd, e + "decoder_res2", x)
    x = _up(sd, e + "decoder_upsample2", x) + skip2
    x = _res(sd, e + "decoder_res1", x)
    x = _up(sd, e + "decoder_upsample1", x) + skip1
    x = _basic(sd, "back_layers.0", x)
    return F.conv3d(x, sd["output_layer.weight"], sd["output_layer.bias"])


# ----------------------------------------------------------------------------
# a8: soft-argmax
# ----------------------------------------------------------------------------
def soft_argmax(volumes: torch.Tensor, coord_volumes: torch.Tensor, softmax: bool = True
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """utils/op.py:83-96."""
    B, J = volumes.shape[:2]
    flat = volumes.reshape(B, J, -1)
    flat = F.softmax(flat, dim=2) if softmax else F.relu(flat)
    vol = flat.reshape(volumes.shape)
    kp = torch.einsum("bnxyz,bxyzc->bnc", vol, coord_volumes[:B])
    return kp, vol


# ----------------------------------------------------------------------------
# whole stage (voxel_net_depth.py:237-273), post-backbone
# ----------------------------------------------------------------------------
class StageTables:
    """Frame-invariant tables built once (voxel_net_depth.py:79-105)."""

    def __init__(self, calibration_path: str, volume_size: int = 64, cuboid_side: float = 2.0,
                 heatmap_shape=(1024, 1280), image_width: int = 1280, image_height: int = 1024):
        self.calib = load_calibration(calibration_path)
        self.V, self.side = volume_size, cuboid_side
        self.image_width, self.image_height = image_width, image_height
        self.coord_volume = build_coord_volume(volume_size, cuboid_side)
        self.grid_px = world2camera_f32(self.calib, self.coord_volume.reshape(-1, 3))
        self.grid = normalise_grid(self.grid_px, heatmap_shape)          # (N,1,2)
        self.ray = ray_table(self.calib, image_width, image_height)


def stage_forward(tables: StageTables, sd: Dict[str, torch.Tensor], feat256: torch.Tensor,
                  depth_batch: Optional[torch.Tensor] = None, scene_volumes: Optional[torch.Tensor] = None,
                  with_intersection: bool = False, volume_multiplier: float = 1.0, softmax: bool = True,
                  return_intermediates: bool = False):
    """Everything VoxelNetwork_depth.forward does after the backbone, fp32 CPU.

    `sd` is the full-module state dict (keys 'process_features.0.*',
    'volume_net.*').  Returns (keypoints (B,15,3), features, volumes) like the
    reference's first three outputs.
    """
    B = feat256.shape[0]
    V = tables.V
    features = process_features(feat256, sd["process_features.0.weight"], sd["process_features.0.bias"],
                                up=tables.image_height, pad=(tables.image_width - tables.image_height) // 2)
    grid_b = tables.grid.unsqueeze(0).expand(B, -1, -1, -1)
    lifted = unproject(features, grid_b, V)
    if scene_volumes is None and depth_batch is not None:
        scene_volumes = torch.stack([
            torch.from_numpy(voxelize_depth(d.numpy(), tables.ray, V, tables.side, tables.image_height,
                                            (tables.image_width - tables.image_height) // 2))
            for d in depth_batch])
    if scene_volumes is not None:
        sv = scene_volumes.unsqueeze(1)
        vol_in = torch.cat([lifted, lifted * sv, sv], 1) if with_intersection else torch.cat([lifted, sv], 1)
    else:
        vol_in = lifted
    logits = v2v_forward(sd, vol_in, prefix="volume_net.")
    coord = tables.coord_volume.unsqueeze(0).expand(B, -1, -1, -1, -1)
    kp, vol = soft_argmax(logits * volume_multiplier, coord, softmax)
    if return_intermediates:
        return kp, features, vol, {"lifted": lifted, "scene": scene_volumes, "logits": logits}
    return kp, features, vol


def stage_forward_device(tables: StageTables, sd: Dict[str, torch.Tensor], feat256: torch.Tensor,
                         depth_batch: torch.Tensor, device, timings: bool = False):
    """The reference forward after the backbone with the reference's OWN op sequence on `device`
    (network/voxel_net_depth.py:237-273 constructed with device='cuda'): nn.Conv2d + nn.Upsample + ConstantPad2d
    (:58-63), F.grid_sample(align_corners=True) (utils/op.py:209), the per-frame host loop
    `depth_map.cpu().numpy() -> NumPy voxelisation -> .to(device)` (:251-257), cuDNN Conv3d V2V, softmax + einsum
    (utils/op.py:88-94).  fp32 throughout.  `sd`, `feat256`, `depth_batch` live on `device`.  Used by
    `bench.py --impl reference-gpu` (the reference's single-GPU PyTorch stage) and checked against `stage_forward`
    in tests/test_oracle_golden.py on the CPU."""
    import time as _time
    B, V = feat256.shape[0], tables.V
    up, pad = tables.image_height, (tables.image_width - tables.image_height) // 2
    x = F.conv2d(feat256, sd["process_features.0.weight"], sd["process_features.0.bias"])
    x = F.interpolate(x, size=(up, up))                                  # nn.Upsample(size=...) default mode 'nearest'
    features = F.pad(x, (pad, pad, 0, 0), value=0.0)
    grid = tables.grid.to(device).unsqueeze(0).expand(B, -1, -1, -1)
    lifted = F.grid_sample(features, grid, align_corners=True).view(B, features.shape[1], V, V, V)
    t0 = _time.perf_counter()
    scene = []
    for i in range(B):                                                   # voxel_net_depth.py:251-257
        d = depth_batch[i].cpu().numpy()
        scene.append(torch.from_numpy(voxelize_depth(d, tables.ray, V, tables.side, tables.image_height, pad)).to(device))
    scene = torch.stack(scene, dim=0).unsqueeze(1)
    t_vox = _time.perf_counter() - t0
    vol_in = torch.cat([lifted, scene], dim=1)
    logits = v2v_forward(sd, vol_in, prefix="volume_net.")
    coord = tables.coord_volume.to(device).unsqueeze(0).expand(B, -1, -1, -1, -1)
    kp, vol = soft_argmax(logits, coord, True)
    if timings:
        return kp, features, vol, {"voxel_loop_s": t_vox}
    return kp, features, vol


# ----------------------------------------------------------------------------
# evaluation math
# ----------------------------------------------------------------------------
def umeyama(P: np.ndarray, Q: np.ndarray):
    """utils/rigid_transform_with_scale.py:18-43 -- (c, R, t) minimising sum |P c R + t - Q|^2."""
    n = P.shape[0]
    cP, cQ = P - P.mean(axis=0), Q - Q.mean(axis=0)
    C = cP.T.dot(cQ) / n
    V, S, W = np.linalg.svd(C)
    if np.linalg.det(V) * np.linalg.det(W) < 0.0:
        S[-1] = -S[-1]
        V[:, -1] = -V[:, -1]
    R = V.dot(W)
    c = 1 / np.var(P, axis=0).sum() * np.sum(S)
    t = Q.mean(axis=0) - P.mean(axis=0).dot(c * R)
    return c, R, t


def calculate_error(estimated_seq, gt_seq) -> float:
    """utils/calculate_errors.py:22-28."""
    d = np.linalg.norm(np.asarray(estimated_seq) - np.asarray(gt_seq), axis=2)
    return float(np.mean(d))


def align_skeleton(estimated_seq, gt_seq, scale: bool = True):
    """utils/calculate_errors.py:60-91 with skeleton_model=None (what dataset/test_dataset.py:108 passes)."""
    est = np.array(estimated_seq, copy=True)      # dtype kept: float32 network output stays float32, like deepcopy(np.asarray())
    gt = np.array(gt_seq, copy=True)
    out = np.zeros_like(est)
    for s in range(est.shape[0]):
        p, g = est[s], gt[s]
        if scale is False:
            p -= np.mean(p, axis=0)
            g -= np.mean(g, axis=0)
        c, R, t = umeyama(p, g)
        out[s] = p.dot(R) * c + t if scale else p.dot(R) + t
    return out, gt


def evaluate_mpjpe(pred, gt):
    """dataset/test_dataset.py:102-112."""
    aligned, g = align_skeleton(pred, gt)
    return calculate_error(pred, gt), calculate_error(aligned, g)


def mpjpe(pred: np.ndarray, gt: np.ndarray) -> float:
    """utils/calculate_errors.py:22-28 semantics: mean per-joint L2 distance."""
    return float(np.mean(np.linalg.norm(np.asarray(pred) - np.asarray(gt), axis=-1)))
