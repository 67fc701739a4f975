"""Cross-check of the ORACLE against the real backend of the reference, wherever it exists.

The reference scripts call `open3d` (/root/reference/3d_model/reconstruct_rgbd.py:79-113,
reconstruct_rgbd_filter.py:123, check_one_frame.py:27-28).  The wheel is not installable in the build
container, so every test here is skipped there (`pytest.importorskip("open3d")`) and the oracle stays
"parity unpinned"; on any box that has the wheel these tests run the reference's own call sequence on
the committed golden inputs and compare with the oracle at the north_star tolerances:
  voxel set + TSDF (what extract_voxel_point_cloud exposes)  bit-exact / <= 1e-4
  mesh vertex set, triangle count, colours                    chamfer <= 0.25 voxel (expected 0), <= 1/255
  voxel_down_sample output, remove_statistical_outlier indices  exact
No GPU is needed: this pins the checker, and the GPU path is pinned against the checker elsewhere.
"""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import open3d_ref, oracle

pytest.importorskip("open3d", reason="open3d (the reference's backend) is not installed on this box: parity stays unpinned")

GOLDEN = ("small_sequence.npz", "hd_sequence.npz")


def load(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name))
    depth, rgb, extr = z["depth"], z["rgb"], z["extrinsic"]
    fxfycxcy = tuple(float(x) for x in z["intr"])
    H, W = depth.shape[1:]
    vl, trunc = float(z["voxel"][0]), float(z["voxel"][1])
    return depth, rgb, extr, (W, H) + fxfycxcy, fxfycxcy, vl, trunc


def oracle_volume(depth, rgb, extr, fxfycxcy, vl, trunc):
    v = oracle.Volume(vl, trunc)
    for k in range(len(depth)):
        v.integrate(oracle.depth_convert(depth[k], 1000.0, 3.0), rgb[k], fxfycxcy, extr[k])
    return v


def oracle_near_surface(v, vl):
    keys, tsdf, w, _ = v.export_blocks(color=False)
    sel = (w != 0) & (tsdf < 0.98) & (tsdf >= -0.98)
    b, i = np.nonzero(sel)
    x, y, z = i >> 8, (i >> 4) & 15, i & 15                      # export index = x*256 + y*16 + z
    idx = np.stack([keys[b, 0].astype(np.int64) * 16 + x, keys[b, 1].astype(np.int64) * 16 + y,
                    keys[b, 2].astype(np.int64) * 16 + z], 1)
    order = np.lexsort((idx[:, 2], idx[:, 1], idx[:, 0]))
    return idx[order], tsdf[b, i][order]


def rows_sorted(a):
    a = np.ascontiguousarray(a, np.float64)
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]


@pytest.mark.parametrize("name", GOLDEN)
def test_voxels_and_tsdf(name):
    depth, rgb, extr, whk, k, vl, trunc = load(name)
    ref = open3d_ref.integrate_sequence(depth, rgb, whk, extr, vl, trunc)
    ridx, rtsdf = open3d_ref.near_surface_voxels(ref, vl)
    oidx, otsdf = oracle_near_surface(oracle_volume(depth, rgb, extr, k, vl, trunc), vl)
    assert ridx.shape == oidx.shape and (ridx == oidx).all(), "set of observed near-surface voxels differs from open3d"
    assert np.abs(rtsdf - otsdf).max() <= 1e-4
    print(f"{name}: {len(ridx)} voxels, TSDF max |diff| {np.abs(rtsdf - otsdf).max():.3g}, bit-equal {(rtsdf == otsdf).mean():.4f}")


@pytest.mark.parametrize("name", GOLDEN)
def test_mesh(name):
    from scipy.spatial import cKDTree
    depth, rgb, extr, whk, k, vl, trunc = load(name)
    rv, rc, rn, rf = open3d_ref.mesh_arrays(open3d_ref.integrate_sequence(depth, rgb, whk, extr, vl, trunc))
    ov, oc, of, _ = oracle_volume(depth, rgb, extr, k, vl, trunc).extract_triangle_mesh()
    assert len(rv) == len(ov) and len(rf) == len(of)
    d, nn = cKDTree(ov).query(rv)
    assert d.max() <= 0.25 * vl
    assert np.abs(rc - oc[nn]).max() <= 1.0 / 255
    on = oracle.vertex_normals(ov, of)
    assert np.abs(rn - on[nn]).max() < 1e-6
    print(f"{name}: {len(rv)} vertices, max vertex distance {d.max():.3g} m, exact matches {(d == 0).mean():.4f}")


def test_extract_point_cloud():
    from scipy.spatial import cKDTree
    depth, rgb, extr, whk, k, vl, trunc = load(GOLDEN[0])
    rp, rc, _ = open3d_ref.point_cloud_arrays(open3d_ref.integrate_sequence(depth, rgb, whk, extr, vl, trunc))
    op, oc, _ = oracle_volume(depth, rgb, extr, k, vl, trunc).extract_point_cloud()
    assert len(rp) == len(op)
    d, nn = cKDTree(op).query(rp)
    assert d.max() <= 0.25 * vl and np.abs(rc - oc[nn]).max() <= 1.0 / 255


def test_filters_exact():
    """voxel_down_sample (check_one_frame.py:28) and remove_statistical_outlier (north_star) on the oracle's
    extracted cloud: occupied voxels / means and kept indices."""
    depth, rgb, extr, whk, k, vl, trunc = load(GOLDEN[0])
    pts, cols, ek = oracle_volume(depth, rgb, extr, k, vl, trunc).extract_point_cloud()
    o = np.lexsort((ek[:, 3], ek[:, 2], ek[:, 1], ek[:, 0]))
    pts, cols = np.ascontiguousarray(pts[o]), np.ascontiguousarray(cols[o])
    pc = open3d_ref.cloud(pts, cols)
    rds = pc.voxel_down_sample(2.5 * vl)
    op, oc, _, _ = oracle.voxel_down_sample(pts, cols, 2.5 * vl)
    assert len(rds.points) == len(op)
    assert np.abs(rows_sorted(np.asarray(rds.points)) - rows_sorted(op)).max() <= 1e-12
    _, ridx = pc.remove_statistical_outlier(20, 2.0)
    kept, _ = oracle.remove_statistical_outlier(pts, 20, 2.0)
    assert list(ridx) == [int(i) for i in kept]


def test_create_from_rgbd_image():
    """check_one_frame.py:27 (dense back-projection, A.12)."""
    import open3d as o3d
    depth, rgb, extr, whk, k, vl, trunc = load(GOLDEN[0])
    color = o3d.geometry.Image(np.ascontiguousarray(rgb[0]))
    dimg = o3d.geometry.Image(np.ascontiguousarray(depth[0]))
    rgbd = o3d.geometry.RGBDImage.create_from_color_and_depth(color, dimg, depth_scale=1000.0, depth_trunc=3.0,
                                                              convert_rgb_to_intensity=False)
    assert (np.asarray(rgbd.depth) == oracle.depth_convert(depth[0], 1000.0, 3.0)).all()
    pc = o3d.geometry.PointCloud.create_from_rgbd_image(rgbd, open3d_ref.intrinsic(*whk))
    op, oc = oracle.backproject_rgbd(oracle.depth_convert(depth[0], 1000.0, 3.0), rgb[0], k)[:2]
    assert len(pc.points) == len(op)
    assert np.abs(np.asarray(pc.points) - op).max() <= 1e-12 and np.abs(np.asarray(pc.colors) - oc).max() <= 1e-12
