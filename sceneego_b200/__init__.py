"""sceneego_b200: B200-native volumetric lifting stage of SceneEgo.

Public surface mirrors the reference:
    sceneego_b200.network.voxel_net_depth.VoxelNetwork_depth
    sceneego_b200.network.v2v.V2VModel
    sceneego_b200.utils.op.{unproject_heatmaps_one_view_batch, integrate_tensor_3d_with_coordinates, ...}
    sceneego_b200.utils.cfg.load_config
"""
import os

__version__ = "0.1.0"
PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))
DEFAULT_CONFIG = os.path.join(PACKAGE_DIR, "data", "sceneego.yaml")
DEFAULT_CALIBRATION = os.path.join(PACKAGE_DIR, "data", "fisheye.calibration_05_08.json")
