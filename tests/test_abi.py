"""The C-ABI library loads on a CPU box, exports every symbol include/otslam_b200.h declares, and
fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "otslam_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(otslam_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    from otslam_b200 import _lib
    syms = header_symbols()
    assert len(syms) >= 25
    lib = C.CDLL(_lib.SO_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/otslam_b200.h but not exported"
    assert not _lib.MISSING
    assert set(_lib.EXPORTS) == set(syms), set(_lib.EXPORTS) ^ set(syms)      # the binding covers the whole header


def test_header_cites_reference_call_sites():
    txt = open(os.path.join(ROOT, "include", "otslam_b200.h")).read()
    for cite in ("reconstruct_rgbd.py:79-83", "reconstruct_rgbd.py:99-107", "reconstruct_rgbd.py:112", "reconstruct_rgbd_filter.py:123",
                 "reconstruct_rgbd_filter.py:126-132", "hybrid_map.py:45-55", "check_one_frame.py:27", "check_one_frame.py:28"):
        assert cite in txt, cite


def test_version_and_error_string():
    from otslam_b200 import _lib
    assert _lib.lib.otslam_version() >= 100
    assert _lib.lib.otslam_launch_count() >= 0


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from otslam_b200 import _lib
    from otslam_b200.volume import TSDFVolume
    with pytest.raises(RuntimeError) as e:
        TSDFVolume(0.01, 0.04)
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)
    out = np.empty(4, np.float32)
    d = np.arange(4, dtype=np.uint16)
    assert _lib.lib.otslam_depth_convert(_lib.ptr(d), 4, 1000.0, 3.0, _lib.ptr(out), 0) == _lib.ERR_CUDA


def test_argument_validation_before_cuda():
    from otslam_b200 import _lib
    h = C.c_void_p()
    assert _lib.lib.otslam_volume_create(-1.0, 0.04, 1, 0, None, C.byref(h)) == _lib.ERR_INVALID
    assert "voxel_length" in _lib.last_error()
    assert _lib.lib.otslam_volume_create(0.01, 0.04, 7, 0, None, C.byref(h)) == _lib.ERR_INVALID
    bad = _lib.SlabSpec(4, 8, 2, 0)
    assert _lib.lib.otslam_volume_create(0.01, 0.04, 1, 0, C.byref(bad), C.byref(h)) == _lib.ERR_INVALID
    diag_with_halo = _lib.SlabSpec(3, 1, 2, 0, 1)                 # diagonal slabs need the exchange mode
    assert _lib.lib.otslam_volume_create(0.01, 0.04, 1, 0, C.byref(diag_with_halo), C.byref(h)) == _lib.ERR_INVALID
    assert "halo = 0" in _lib.last_error()
    assert _lib.lib.otslam_volume_reset(None) == _lib.ERR_INVALID
    n = C.c_int64(0)
    assert _lib.lib.otslam_cloud_voxel_down_sample(None, None, 0, -1.0, None, None, None, None, C.byref(n), 0) == _lib.ERR_INVALID
    assert _lib.lib.otslam_cloud_remove_statistical_outlier(None, 0, 0, 1.0, None, C.byref(n), None, 0) == _lib.ERR_INVALID
    assert _lib.lib.otslam_mesh_sample_uniform(None, None, None, 0, None, 0, 10, 0, None, None, None, 0) == _lib.ERR_INVALID


def test_oracle_is_not_linked_into_the_product():
    """The product library must not depend on the oracle (it is test infrastructure)."""
    from otslam_b200 import _lib
    blob = open(_lib.SO_PATH, "rb").read()
    assert b"liboracle" not in blob and b"oracle_volume" not in blob
    pkg = os.path.join(ROOT, "object-triggered-3d-slam_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert "from oracle" not in src and "import oracle" not in src and "liboracle" not in src, os.path.join(dp, f)
