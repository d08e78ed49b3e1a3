"""Drop-in `VoxelNetwork_depth` (reference: network/voxel_net_depth.py:19-275).

Same constructor, config keys, attributes, state-dict keys, forward signature and
4-tuple return as the reference, so `demo.py` / `test.py` can use it unchanged.
Everything after the backbone runs hand-written sm_100a kernels through the
C-ABI of include/sceneego_b200.h:

    process_features (1x1 conv; the x16 upsample + pad are never materialised
                      unless `materialize_features`)      -> sceneego_feature_conv1x1_f32
    unproject_heatmaps_one_view_batch + torch.cat         -> sceneego_unproject_f32
    depth_map_to_voxel_numpy loop (host NumPy in the ref) -> sceneego_voxelize_depth_f64
    volume_net (V2VModel)                                 -> sceneego_v2v_run
    integrate_tensor_3d_with_coordinates                  -> sceneego_softargmax3d_f32

There is no CPU path: constructing on a non-CUDA device raises.
"""
import os

import numpy as np
import torch
from torch import nn

from .. import _lib
from ..utils import op
from ..utils.fisheye.FishEyeCalibrated import FishEyeCameraCalibrated
from . import pose_resnet
from .v2v import V2VModel

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEFAULT_CALIBRATION = os.path.join(_PKG, "data", "fisheye.calibration_05_08.json")


def _resolve_calibration(path):
    if os.path.exists(path):
        return path
    alt = os.path.join(_PKG, "data", os.path.basename(path))
    if os.path.exists(alt):
        return alt
    raise FileNotFoundError(path)


class VoxelNetwork_depth(nn.Module):
    def __init__(self, config, device='cuda', materialize_features=True, materialize_volumes=True,
                 fused_projection=False, v2v_chunk=32):
        """Extra keyword switches (all default to reference-identical outputs):
        materialize_features / materialize_volumes: build outputs #2 / #3 of the reference
            forward (168 MB and 15.7 MB per frame, ignored by demo.py:57 / test.py:54);
            False returns None in their place.
        fused_projection: project voxel centres inside the gather kernel (Scaramuzza in-kernel)
            instead of reading the `grid_coord_proj_batch` argument."""
        super().__init__()
        dev = torch.device(device)
        if dev.type != 'cuda':
            raise _lib.SceneEgoError("sceneego_b200.VoxelNetwork_depth needs a CUDA device (no CPU fallback)")
        _lib.load_library()
        self.device = device
        self.num_joints = config.model.backbone.num_joints
        self.volume_softmax = config.model.volume_softmax
        self.volume_multiplier = config.model.volume_multiplier
        self.volume_size = config.model.volume_size
        self.cuboid_side = config.model.cuboid_side
        self.kind = config.model.kind
        self.heatmap_softmax = config.model.heatmap_softmax
        self.heatmap_multiplier = config.model.heatmap_multiplier
        self.materialize_features = materialize_features
        self.materialize_volumes = materialize_volumes
        self.fused_projection = fused_projection

        if config.model.backbone.local_checkpoint:
            loads = torch.load(config.model.backbone.checkpoint)
            self.backbone = pose_resnet.get_pose_net(state_dict=loads['state_dict'])
        else:
            print('Do not load checkpoint')
            self.backbone = pose_resnet.get_pose_net(None)
        self.backbone = self.backbone.to(device)
        if config.opt.train_2d is False:
            for p in self.backbone.parameters():
                p.requires_grad = False

        self.heatmap_shape = tuple(config.heatmap_shape)                     # (1024, 1280)
        self.image_width = config.dataset.image_width
        self.image_height = config.dataset.image_height
        # parameter container only (state-dict keys process_features.0.{weight,bias});
        # the Upsample / ConstantPad2d children hold no parameters
        self.process_features = nn.Sequential(
            nn.Conv2d(256, 32, 1),
            nn.Upsample(size=(self.image_height, self.image_height)),
            nn.ConstantPad2d(padding=((self.image_width - self.image_height) // 2,
                                      (self.image_width - self.image_height) // 2, 0, 0), value=0.0),
        ).to(device)

        self.with_scene = config.model.with_scene
        self.with_intersection = False
        if self.with_scene is True:
            if config.model.with_intersection is True:
                volume_input_channel_num = 32 + 1 + 32       # volumes, volumes * scene, scene (:66-68, :257-260)
                self.with_intersection = True
            else:
                volume_input_channel_num = 32 + 1
                self.with_intersection = False
        else:
            volume_input_channel_num = 32
        self.volume_net = V2VModel(volume_input_channel_num, self.num_joints, max_chunk=v2v_chunk).to(device)

        print('build coord volume')
        self.fisheye_camera_model = FishEyeCameraCalibrated(
            calibration_file_path=_resolve_calibration(config.dataset.camera_calibration_path))
        self._calib = self.fisheye_camera_model.calib_struct(self.image_width, self.image_height)
        self.coord_volume = self.build_coord_volume()
        self.coord_volumes = self.coord_volume.unsqueeze(0).expand(config.opt.batch_size, -1, -1, -1, -1)
        print('build reprojected grid coord')
        self.grid_coord_proj = op.get_projected_2d_points_with_coord_volumes(
            fisheye_model=self.fisheye_camera_model, coord_volume=self.coord_volume)
        self.grid_coord_proj_batch = op.get_grid_coord_proj_batch(
            self.grid_coord_proj, batch_size=config.opt.batch_size, heatmap_shape=config.heatmap_shape)
        # ray table: device copy row-major for the kernel, NumPy x-major copy for API parity
        self._ray_dev = self.fisheye_camera_model.ray_table_device(self.image_width, self.image_height, dev)
        self.ray = self._ray_dev.permute(1, 0, 2).reshape(-1, 3).cpu().numpy()
        cv = self.coord_volume
        self._axis = torch.stack([cv[:, 0, 0, 0], cv[0, :, 0, 1], cv[0, 0, :, 2]]).contiguous()
        self.last_launches = 0
        self._side_stream = None
        self._features_buf = None
        # (pre_h, pre_w, clamp_max): depth_map_batch holds RAW decoded maps and the dataset's preprocessing
        # (dataset/demo_dataset.py:86-91: nearest resize to 1280x1024, depth > 10 -> 10) is fused into the
        # voxelisation kernel's load; None: the caller has preprocessed them, like the reference's datasets do
        self.depth_preprocess = None

    def build_coord_volume(self):
        """network/voxel_net_depth.py:110-134, on the device (fp32, mul then add)."""
        v, s = self.volume_size, self.cuboid_side
        idx = torch.arange(v, dtype=torch.float32, device=self.device)
        step = float(s / (v - 1))
        axy = float(-s / 2) + step * idx
        az = 0.0 + step * idx
        gx, gy, gz = torch.meshgrid(axy, axy, az, indexing='ij')
        return torch.stack([gx, gy, gz], dim=-1).contiguous()

    # ------------------------------------------------------------------ hot path
    def lift(self, backbone_features, grid_coord_proj_batch, coord_volumes=None, scene_volumes=None,
             depth_map_batch=None):
        """Everything the reference forward does after `self.backbone(images)`
        (network/voxel_net_depth.py:237-273)."""
        feat = backbone_features.contiguous().float()
        b = feat.shape[0]
        v = self.volume_size
        conv = self.process_features[0]
        feat32 = _lib.feature_conv1x1(feat, conv.weight, conv.bias)           # (B,64,64,32) channel-last
        launches = 1
        features = None
        feat_done = None
        if self.materialize_features:
            # output #2 of the reference is a pure HBM write (168 MB/frame) nothing downstream reads: it runs on a
            # side stream and shares the SMs (and the idle HBM bandwidth) with the tensor-bound V2V kernels
            main = torch.cuda.current_stream(feat.device)
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream(device=feat.device)
            side = self._side_stream
            side.wait_stream(main)
            # one persistent buffer per batch size (168 MB per frame): a fresh 10 GB tensor per call, kept alive across
            # two streams by record_stream, made the caching allocator run out of reusable blocks as soon as the host
            # ran a few steps ahead of the GPU and fall back to synchronising cudaFree / cudaMalloc cycles.  The
            # returned tensor is therefore overwritten by the next call (clone it to keep it); consumers on the
            # current stream are ordered after the writer, and the next call's writer after them.
            shape = (b, 32, self.image_height, self.image_width)
            if self._features_buf is None:
                self._features_buf = {}
            if b not in self._features_buf:
                if len(self._features_buf) >= 4:            # a few batch sizes at most (e.g. the pipeline's ramp-up)
                    self._features_buf.clear()
                self._features_buf[b] = torch.empty(shape, dtype=torch.float32, device=feat.device)
            with torch.cuda.stream(side):
                features = _lib.features_upsample_pad(feat32, self.image_height,
                                                      (self.image_width - self.image_height) // 2,
                                                      out=self._features_buf[b])
                feat_done = torch.cuda.Event()
                feat_done.record(side)
            feat32.record_stream(side)
            launches += 1
        if self.with_scene is True and scene_volumes is None and depth_map_batch is None:
            print("no scene volume or depth input!")
            if feat_done is not None:
                torch.cuda.current_stream(feat.device).wait_event(feat_done)
            return None
        grid = None
        if not self.fused_projection:
            if grid_coord_proj_batch.shape[0] > 1 and grid_coord_proj_batch.stride(0) != 0:
                raise _lib.SceneEgoError("per-frame projection grids are not supported (the reference expands one table)")
            grid = grid_coord_proj_batch[0].reshape(-1, 2).contiguous()
        vn = self.volume_net
        chunk = min(vn.max_chunk, b)
        pg = vn.program(v, chunk, feat.device)
        logits = torch.empty(b, self.num_joints, v, v, v, dtype=torch.float32, device=feat.device)
        in_buf = pg.buffers[pg.in_buf]
        img_h, img_w = self.heatmap_shape
        for s in range(0, b, chunk):
            n = min(chunk, b - s)
            _lib.unproject(feat32[s:s + n], grid, self._calib if self.fused_projection else None, v,
                           float(self.cuboid_side), img_h, img_w, None, in_buf, pg.lay_in,
                           extra_zero_planes=(pg.in_pad - 32) // 8)
            launches += 1
            if self.with_scene is True:
                scene_ch = 64 if self.with_intersection else 32
                if scene_volumes is not None:
                    if self.with_intersection:
                        # the reference concatenates 33 channels here and its 65-channel V2V then fails (:246-249)
                        raise _lib.SceneEgoError("with_intersection=true takes depth_map_batch, not scene_volumes "
                                                 "(the reference builds a 33-channel input on this path)")
                    sv = scene_volumes[s:s + n].contiguous().float().unsqueeze(1)
                    _lib.pack_volume(sv, in_buf, pg.lay_in, c_offset=32)
                else:
                    d = depth_map_batch[s:s + n]
                    d = d.reshape(n, d.shape[-2], d.shape[-1]).contiguous().float()
                    if self.depth_preprocess is not None:
                        ph, pw, cm = self.depth_preprocess
                        _lib.voxelize_depth_raw(d, (ph, pw), float(cm), self._ray_dev, self.image_height,
                                                self.image_width, v, float(self.cuboid_side), None, in_buf, pg.lay_in,
                                                channel=scene_ch)
                    else:
                        _lib.voxelize_depth(d, self._ray_dev, self.image_height, self.image_width, v,
                                            float(self.cuboid_side), None, in_buf, pg.lay_in, channel=scene_ch)
                    if self.with_intersection:
                        _lib.intersect(in_buf, pg.lay_in, n, 32)
                        launches += 1
                launches += 1
            launches += vn.run_chunk(pg, n, logits[s:s + n])
        kp, volumes = _lib.softargmax3d(logits, float(self.volume_multiplier), bool(self.volume_softmax),
                                        self._axis, None, self.materialize_volumes)
        launches += 3 if self.materialize_volumes else 2
        if feat_done is not None:
            torch.cuda.current_stream(feat.device).wait_event(feat_done)
        self.last_launches = launches
        return kp, features, volumes, self.coord_volumes

    def forward(self, images, grid_coord_proj_batch, coord_volumes, scene_volumes=None, depth_map_batch=None):
        heatmaps, features = self.backbone(images)
        return self.lift(features, grid_coord_proj_batch, coord_volumes, scene_volumes, depth_map_batch)

    # kept for API parity with the reference (network/voxel_net_depth.py:194-205): one frame
    def depth_map_to_voxel_numpy(self, depth):
        d = depth.reshape(1, depth.shape[-2], depth.shape[-1]).contiguous().float().to(self.device)
        occ = torch.zeros(1, self.volume_size, self.volume_size, self.volume_size, dtype=torch.float32,
                          device=d.device)
        _lib.voxelize_depth(d, self._ray_dev, self.image_height, self.image_width, self.volume_size,
                            float(self.cuboid_side), occ, None, None)
        return occ[0]
