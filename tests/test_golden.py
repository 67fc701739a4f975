"""Golden fixture tests/golden/small_sequence.npz (made by tools/make_golden.py): committed inputs +
digests of the oracle's outputs.  CPU leg pins the oracle; the GPU leg (marked gpu) pins the CUDA
path against the same digests through the C ABI."""
import importlib.util
import json
import os

import numpy as np
import pytest

from conftest import ROOT


def load():
    z = np.load(os.path.join(ROOT, "tests", "golden", "small_sequence.npz"))
    exp = json.loads(bytes(z["expected"]).decode())
    return z, exp


def make_golden():
    spec = importlib.util.spec_from_file_location("mg", os.path.join(ROOT, "tools", "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_oracle_matches_golden():
    z, exp = load()
    mg = make_golden()
    got = mg.compute(z["depth"], z["rgb"], tuple(z["intr"]), z["extrinsic"], float(z["voxel"][0]), float(z["voxel"][1]))
    for k in exp:
        assert got[k] == exp[k], k


@pytest.mark.gpu
def test_gpu_matches_golden():
    from otslam_b200.volume import TSDFVolume
    z, exp = load()
    mg = make_golden()
    v = TSDFVolume(float(z["voxel"][0]), float(z["voxel"][1]))
    v.integrate_batch(z["depth"], z["rgb"], tuple(z["intr"]), z["extrinsic"])
    keys, tsdf, w, col = v.export_blocks()
    assert len(keys) == exp["n_blocks"]
    assert mg.digest(keys) == exp["keys"]
    assert mg.digest(w.astype(np.uint16)) == exp["weight"]
    assert mg.digest(tsdf) == exp["tsdf"]                       # the f32 running mean is reproduced bit for bit
    assert mg.digest(np.floor(col.astype(np.float64) + 0.5).astype(np.uint8)) == exp["color_u8"]
    verts, cols, nrm, faces, ek = v.extract_triangle_mesh()
    o = mg.lexorder(ek)
    assert (len(verts), len(faces)) == (exp["mesh_nv"], exp["mesh_nf"])
    assert mg.digest(ek[o]) == exp["mesh_ekeys"] and mg.digest(verts[o]) == exp["mesh_verts"]
    pts, pcols, pek = v.extract_point_cloud()
    po = mg.lexorder(pek)
    assert len(pts) == exp["pc_n"] and mg.digest(pek[po]) == exp["pc_ekeys"] and mg.digest(pts[po]) == exp["pc_pts"]
    # post stage pins (voxel_down_sample, remove_statistical_outlier, transform) on the canonically ordered cloud
    import otslam_b200.o3d_compat as o3d
    pc = o3d.geometry.PointCloud()
    pc.points, pc.colors = np.ascontiguousarray(pts[po]), np.ascontiguousarray(pcols[po])
    ds = pc.voxel_down_sample(2.5 * float(z["voxel"][0]))
    assert len(ds.points) == exp["vds_n"]
    assert mg.digest(np.asarray(ds.points)) == exp["vds_pts"]      # (colours: exact integer sums here vs the FP64 running mean, <= 1e-9 apart)
    _, idx = pc.remove_statistical_outlier(20, 2.0)
    assert len(idx) == exp["sor_n"] and mg.digest(np.asarray(idx, np.int64)) == exp["sor_idx"]
    moved = o3d.geometry.PointCloud(np.ascontiguousarray(pts[po])).transform(mg.POST_T)
    assert mg.digest(np.asarray(moved.points)) == exp["xform_pts"]

