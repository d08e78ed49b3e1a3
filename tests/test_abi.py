"""CPU: the C-ABI library loads and exports every function include/sceneego_b200.h declares."""
import ctypes
import os
import re

from sceneego_b200 import _lib
from tests import util


def _declared():
    src = open(os.path.join(util.ROOT, "include", "sceneego_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sceneego_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load_library()
    names = _declared()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.SYMBOLS) == names, "python binding list out of sync with the header"
    hdr = open(os.path.join(util.ROOT, "include", "sceneego_b200.h")).read()
    assert lib.sceneego_abi_version() == int(re.search(r"#define SCENEEGO_ABI_VERSION (\d+)", hdr).group(1)) == 5


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.Calib) == 8 * 20 + 8          # 20 doubles + 2 int32
    assert ctypes.sizeof(_lib.VolLayout) == 40               # 6 int32, int64, s2d + reserved
    assert ctypes.sizeof(_lib.V2VOp) == 12 * 4 + 16 + 8 + 2 * 40  # 12 int32, 2 int64, src2 + cin2, 2 layouts


def test_host_only_entry_points():
    lay = _lib.vol_layout(64, 1, 4)
    assert (lay.pitch_y, lay.pitch_x) == (65, 4225)
    assert lay.guard >= 4225 + 65 + 1 and lay.guard % 8 == 0 and lay.frame_pitch % 8 == 0
    assert lay.plane_stride >= 4 * lay.frame_pitch + lay.guard + 1024
    lay7 = _lib.vol_layout(64, 3, 1)
    assert lay7.guard >= 3 * (67 * 67 + 67 + 1) and lay7.s2d == 0
    s2d = _lib.vol_layout_s2d(64, 2)                          # stem input: 8 parity sub-volumes of side 32, pad 2
    assert (s2d.s2d, s2d.side, s2d.pad, s2d.pitch_y, s2d.pitch_x) == (1, 32, 2, 34, 34 * 34)
    assert s2d.guard >= 2 * s2d.pitch_x + 2 * (s2d.pitch_y + 1)
    assert _lib.load_library().sceneego_v2v_stem_s2d_weight_bytes() == 276 * 16384


def test_missing_library_fails_loudly(tmp_path):
    import pytest
    with pytest.raises(_lib.SceneEgoError, match="no CPU or PyTorch fallback"):
        _lib.load_library(str(tmp_path / "nope.so"))


def test_cpu_tensors_are_rejected():
    import pytest
    import torch
    with pytest.raises(_lib.SceneEgoError):
        _lib._ptr(torch.zeros(4))


def test_fp16_storage_library_loads_and_packs():
    """libsceneego_b200_f16.so (the same sources with -DSCENEEGO_ACT_F16) exports the same symbols, reports its storage
    type, and its host-side packers -- fold, fp16 round-to-nearest-even with saturation, the marching / stem / hand-off
    layouts walked like the kernels walk them -- pass the host-logic suite (one dtype per process: a subprocess)."""
    import subprocess
    import sys
    f16 = ctypes.CDLL(_lib.LIB_PATHS["f16"])
    bf = ctypes.CDLL(_lib.LIB_PATHS["bf16"])
    for n in _declared():
        assert hasattr(f16, n), n
    assert f16.sceneego_act_dtype() == 1 and bf.sceneego_act_dtype() == 0
    env = dict(os.environ, SCENEEGO_ACT_DTYPE="f16", PYTHONPATH=util.ROOT)
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_host_logic.py", "-q", "-x"], cwd=util.ROOT, env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-1500:]
