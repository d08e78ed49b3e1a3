"""CPU: the plain-C restatement (oracle/voxelize_oracle.c) vs the NumPy oracle and the
reference-generated golden vectors."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import sceneego_oracle as orc
from tests import util

ODIR = os.path.join(util.ROOT, "oracle")


@pytest.fixture(scope="module")
def clib():
    subprocess.run(["make", "-C", ODIR, "-s"], check=True)
    lib = C.CDLL(os.path.join(ODIR, "libvoxelize_oracle.so"))
    lib.oracle_voxelize.restype = C.c_long
    return lib


def test_c_oracle_matches_numpy_oracle_and_golden(clib):
    calib = orc.load_calibration(util.CALIB)
    ray = np.empty((1280 * 1024, 3), np.float64)
    c2w = np.ascontiguousarray(calib["c2w"])
    clib.oracle_ray_table(C.c_double(calib["center"][0]), C.c_double(calib["center"][1]),
                          c2w.ctypes.data_as(C.c_void_p), 1280, 1024, ray.ctypes.data_as(C.c_void_p))
    assert np.array_equal(ray, orc.ray_table(calib, 1280, 1024))          # all 1,310,720 rays, fp64 bits
    g = util.golden("voxel.npz")
    for name, V in (("img_001000", 64), ("img_002376", 128)):
        d = orc.resize_nearest(g[f"{name}_raw"], 1024, 1280).copy()
        d[d > 10] = 10
        d = np.ascontiguousarray(d, dtype=np.float32)
        occ = np.empty((V, V, V), np.float32)
        clib.oracle_voxelize(d.ctypes.data_as(C.c_void_p), 1024, 1280, ray.ctypes.data_as(C.c_void_p), 1024, 128, V,
                             C.c_double(2.0), occ.ctypes.data_as(C.c_void_p))
        assert np.array_equal(occ, util.unpack_bits(g[f"{name}_v{V}"], V))
    small = (np.random.default_rng(1).random((300, 200), dtype=np.float32) * 5 - 0.5)
    occ = np.empty((64, 64, 64), np.float32)
    clib.oracle_voxelize(small.ctypes.data_as(C.c_void_p), 300, 200, ray.ctypes.data_as(C.c_void_p), 1024, 128, 64,
                         C.c_double(2.0), occ.ctypes.data_as(C.c_void_p))
    assert np.array_equal(occ, orc.voxelize_depth(small, orc.ray_table(calib, 1280, 1024), 64, 2.0))
