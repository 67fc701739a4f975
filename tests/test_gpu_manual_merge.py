"""the rigid-body edits of fusion/hybrid_map_manual.py (transform / get_center / rotate) and
fusion/2d_selective_merge.py::smart_paste, CUDA path vs the oracle -- bit-exact (FP64 copies and integer bytes)."""
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle

pytestmark = pytest.mark.gpu


def _cloud(n, seed, normals=True):
    import otslam_b200.o3d_compat as o3d
    rng = np.random.default_rng(seed)
    pc = o3d.geometry.PointCloud()
    pc.points = rng.normal(size=(n, 3)) * [2.0, 1.0, 0.5] + [3.0, -1.0, 0.4]
    if normals:
        nr = rng.normal(size=(n, 3))
        pc.normals = nr / np.linalg.norm(nr, axis=1, keepdims=True)
    return pc


@pytest.mark.parametrize("n", [1, 1000, 300_000])
def test_transform_center_rotate_match_oracle(n):
    pc = _cloud(n, n)
    p0, n0 = pc.points.copy(), pc.normals.copy()
    T = np.eye(4)
    T[:3, :3] = pc.get_rotation_matrix_from_xyz((0.1, -0.2, 0.3))
    T[:3, 3] = [0.05, -0.1, 0.2]
    pc.transform(T)
    op, on = oracle.transform(p0, n0, T)
    assert (pc.points == op).all() and (pc.normals == on).all()
    c = pc.get_center()
    assert (c == oracle.center(op)).all()                 # negative coordinates: summed in index order all the same
    R = pc.get_rotation_matrix_from_xyz((0, 0, np.radians(2.0)))
    pc.rotate(R, center=c)
    rp, rn = oracle.rotate(op, on, R, c)
    assert (pc.points == rp).all() and (pc.normals == rn).all()


def test_smart_paste_matches_reference_function():
    spec = importlib.util.spec_from_file_location("selective_merge", os.path.join(ROOT, "fusion", "2d_selective_merge.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(0)
    base = rng.choice(np.array([0, 205, 254], np.uint8), size=(300, 400))
    over = rng.integers(190, 221, size=(300, 400)).astype(np.uint8)     # straddles the 200..210 "unknown" band
    for rect in ((10, 20, 100, 50), (0, 0, 400, 300), (399, 299, 1, 1), (350, 10, 100, 10), (-1, 0, 5, 5), (5, 5, 0, 7)):
        want = oracle.smart_paste(base, over, *rect)
        got = mod.smart_paste(base.copy(), over, *rect)
        assert (got == want).all(), rect


def test_manual_script_replays_recorded_keys(tmp_path):
    """hybrid_map_manual.py in batch mode == the reference's key callbacks applied with the oracle."""
    import cv2
    import otslam_b200.o3d_compat as o3d
    from otslam_b200 import synth
    mapdir, objdir = tmp_path / "map", tmp_path / "obj"
    mapdir.mkdir(); objdir.mkdir()
    img = synth.occupancy_map(120, 90, 0.05, 1)
    cv2.imwrite(str(mapdir / "map_selective.pgm"), img)
    (mapdir / "map_selective.yaml").write_text("resolution: 0.05\norigin: [-3.0, -2.25, 0.0]\n")
    a = _cloud(5000, 7, normals=False)
    a.paint_uniform_color([0.1, 0.9, 0.1])
    o3d.io.write_point_cloud(str(objdir / "a.ply"), a)
    out = tmp_path / "hyb" / "adjusted.ply"
    env = dict(os.environ, OTSLAM_MAP_BASE=str(mapdir), OTSLAM_OBJ_DIR=str(objdir), OTSLAM_HYBRID_SAVE=str(out),
               OTSLAM_MANUAL_KEYS=json.dumps({"a.ply": "WWAZZCSDQW"}), OTSLAM_HEADLESS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "fusion", "hybrid_map_manual.py")], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    got = o3d.io.read_point_cloud(str(out))
    # expected: map points (grey 0.3) + the object after the same edits done with the oracle
    rows, cols = np.where(img < 100)
    mp = np.stack([-3.0 + cols * 0.05, -2.25 + (img.shape[0] - 1 - rows) * 0.05, np.zeros(len(rows))], 1)
    p = np.asarray(o3d.io.read_point_cloud(str(objdir / "a.ply")).points)
    step = {"W": (0, 0.05), "S": (0, -0.05), "A": (1, 0.05), "D": (1, -0.05)}
    for k in "WWAZZCSD":
        if k in step:
            T = np.eye(4); T[step[k][0], 3] = step[k][1]
            p, _ = oracle.transform(p, None, T)
        else:
            yaw = np.radians(2.0 if k == "Z" else -2.0)
            R = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
            p, _ = oracle.rotate(p, None, R, oracle.center(p))
    want = np.concatenate([mp, p])
    assert np.asarray(got.points).shape == want.shape and (np.asarray(got.points) == want).all()
    cols_got = np.asarray(got.colors)
    assert np.allclose(cols_got[:len(mp)], 0.3, atol=1 / 255) and np.allclose(cols_got[len(mp):], [1, 0, 0], atol=1 / 255)
