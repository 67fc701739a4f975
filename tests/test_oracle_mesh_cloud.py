"""Pins for the oracle's extraction and cloud operators."""
import hashlib
import os
import re

import numpy as np
import pytest
from scipy.spatial import cKDTree

from conftest import ROOT, lexorder
from oracle import oracle


def parse_tri_table(path):
    txt = open(path).read()
    body = txt[txt.index("MC_TRI_TABLE[256][16] = {"):]
    rows = re.findall(r"\{([-\d,]+)\},", body)
    return [[int(x) for x in r.split(",") if int(x) != -1] for r in rows[:256]]


@pytest.mark.parametrize("rel", ["oracle/mc_tables.h", "object-triggered-3d-slam_b200/csrc/mc_tables.h"])
def test_mc_table_sha1_pin(rel):
    """sha1 of the machine-validated tri table."""
    tri = parse_tri_table(os.path.join(ROOT, rel))
    assert hashlib.sha1(repr(tri).encode()).hexdigest() == "1749c38ecb4960fc8d20bc47673822b171cfd169"


def test_mc_table_structure():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen", os.path.join(ROOT, "tools", "gen_mc_tables.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    gen.check(parse_tri_table(os.path.join(ROOT, "oracle", "mc_tables.h")))     # checks (1)-(6)


def sphere_volume(vl=0.02):
    """Integrate a synthetic sphere (r = 0.3 at 1.5 m) from one view."""
    W, H, f = 320, 240, 280.0
    cx, cy = 160.5, 120.5
    j, i = np.meshgrid(np.arange(W), np.arange(H))
    dx, dy = (j - cx) / f, (i - cy) / f
    # ray o + t*(dx,dy,1) hits sphere centre (0,0,1.5) radius 0.3
    a = dx * dx + dy * dy + 1
    b = -2 * 1.5
    c = 1.5 * 1.5 - 0.09
    disc = b * b - 4 * a * c
    t = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), 0.0)
    depth = np.round(t * 1000).astype(np.uint16)
    rgb = np.zeros((H, W, 3), np.uint8)
    rgb[..., 0] = 200
    rgb[..., 1] = (j % 256).astype(np.uint8)
    v = oracle.Volume(vl, 4 * vl)
    v.integrate(oracle.depth_convert(depth), rgb, (f, f, cx, cy), np.eye(4))
    return v


def test_mesh_lies_on_surface_and_is_consistent():
    vl = 0.02
    v = sphere_volume(vl)
    verts, cols, faces, ek = v.extract_triangle_mesh()
    assert len(verts) > 500 and len(faces) > 500
    r = np.linalg.norm(verts - np.array([0, 0, 1.5]), axis=1)
    assert np.abs(r - 0.3).max() < 0.75 * vl                 # zero crossings within a voxel of the analytic surface
    assert len(np.unique(ek, axis=0)) == len(ek)             # one vertex per lattice edge
    assert faces.min() >= 0 and faces.max() < len(verts)
    # vertex position is on its lattice edge: two coordinates at voxel centres, one within [0, vl)
    base = 0.5 * vl + vl * ek[:, :3]
    off = verts - base
    ax = ek[:, 3]
    for a in range(3):
        m = ax == a
        assert np.abs(np.delete(off[m], a, axis=1)).max() < 1e-12
        assert (off[m][:, a] >= 0).all() and (off[m][:, a] <= vl).all()
    # orientation: normals point to the positive-TSDF (camera / outside) side
    n = oracle.vertex_normals(verts, faces)
    outward = verts - np.array([0, 0, 1.5])
    assert (np.einsum("ij,ij->i", n, outward) > 0).mean() > 0.97
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0)
    assert cols.min() >= 0 and cols.max() <= 1 and abs(cols[:, 0].mean() - 200 / 255) < 1e-6


def test_point_cloud_subset_of_mesh_edges():
    v = sphere_volume()
    verts, cols, faces, ek = v.extract_triangle_mesh()
    pts, pcols, pek = v.extract_point_cloud()
    assert len(pts) > 0
    mesh_edges = {tuple(k) for k in ek.tolist()}
    inside = [tuple(k) in mesh_edges for k in pek.tolist()]
    assert np.mean(inside) > 0.9          # same zero crossings wherever the surrounding cubes are complete
    o1, o2 = lexorder(ek), lexorder(pek)
    lookup = {tuple(k): i for i, k in enumerate(ek.tolist())}
    idx = [lookup[tuple(k)] for k in pek.tolist() if tuple(k) in lookup]
    sel = [i for i, k in enumerate(pek.tolist()) if tuple(k) in lookup]
    assert np.abs(pts[sel] - verts[idx]).max() < 5e-7      # A.5 (FP64) and A.6 (denominator |f0| + |f1| rounded to FP32: 6e-8 relative on the coordinate) interpolate the same crossing
    # the normals extract_point_cloud attaches (GetNormalAt): normalised TSDF gradient = the sphere's outward radial direction
    n = v.point_normals(pts)
    radial = pts - np.array([0, 0, 1.5])
    radial /= np.linalg.norm(radial, axis=1, keepdims=True)
    d = np.einsum("ij,ij->i", n, radial)      # (a projective TSDF's gradient leans towards the view ray near the silhouette)
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0) and (d > 0.7).all() and np.median(d) > 0.99


def test_vertex_normals_known_mesh():
    verts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], float)
    faces = np.array([[0, 1, 2], [0, 2, 3]], np.int32)
    n = oracle.vertex_normals(verts, faces)
    assert np.allclose(n[1], [0, 0, 1]) and np.allclose(n[3], [1, 0, 0])
    assert np.allclose(n[0], np.array([1, 0, 1]) / np.sqrt(2))
    assert np.allclose(oracle.vertex_normals(verts, faces[:0]), [[0, 0, 1]] * 4)   # isolated -> (0,0,1)


def test_sample_uniform_counts_and_barycentric():
    rng = np.random.default_rng(1)
    verts = rng.random((50, 3))
    faces = rng.integers(0, 50, (80, 3)).astype(np.int32)
    faces = faces[(faces[:, 0] != faces[:, 1]) & (faces[:, 1] != faces[:, 2]) & (faces[:, 0] != faces[:, 2])]
    cols = rng.random((50, 3))
    n = 5000
    p, c, _, tri = oracle.sample_uniform(verts, cols, None, faces, n, seed=3)
    area = 0.5 * np.linalg.norm(np.cross(verts[faces[:, 1]] - verts[faces[:, 0]], verts[faces[:, 2]] - verts[faces[:, 0]]), axis=1)
    cdf = np.cumsum(area / area.sum())
    ends = np.round(cdf * n).astype(int)
    ends[-1] = n
    counts = np.diff(np.concatenate([[0], ends]))
    assert (np.bincount(tri, minlength=len(faces)) == counts).all()      # A.10: per-triangle counts are deterministic
    assert (np.diff(tri) >= 0).all()
    # every sample lies in its triangle's plane and inside it
    a, b, cc = verts[faces[tri, 0]], verts[faces[tri, 1]], verts[faces[tri, 2]]
    M = np.stack([b - a, cc - a], -1)
    uv = np.einsum("nij,nj->ni", np.linalg.pinv(M), p - a)
    assert (uv >= -1e-9).all() and (uv.sum(1) <= 1 + 1e-9).all()
    p2, _, _, _ = oracle.sample_uniform(verts, cols, None, faces, n, seed=3)
    assert (p == p2).all()
    with pytest.raises(RuntimeError):
        oracle.sample_uniform(verts, None, None, faces[:0], 10)


def test_voxel_down_sample_vs_numpy():
    rng = np.random.default_rng(2)
    pts = rng.random((3000, 3)) * [1.0, 0.5, 0.2]
    cols = rng.random((3000, 3))
    v = 0.05
    op, oc, keys, counts = oracle.voxel_down_sample(pts, cols, v)
    vmin = pts.min(0) - v / 2
    k = np.floor((pts - vmin) / v).astype(int)
    uk, inv, cnt = np.unique(k, axis=0, return_inverse=True, return_counts=True)
    assert (uk == keys).all() and (cnt == counts).all()
    for m in range(0, len(uk), 37):
        sel = np.nonzero(inv.reshape(-1) == m)[0]
        s = np.zeros(3)
        for i in sel:                     # index-order sum, A.7
            s += pts[i]
        assert (op[m] == s / len(sel)).all()
    assert np.allclose(oc, np.stack([np.bincount(inv.reshape(-1), cols[:, j]) for j in range(3)], 1) / cnt[:, None])
    with pytest.raises(RuntimeError):
        oracle.voxel_down_sample(pts, cols, 0.0)


@pytest.mark.parametrize("n,k", [(2000, 20), (500, 5), (50, 64)])
def test_sor_vs_ckdtree(n, k):
    rng = np.random.default_rng(n)
    pts = rng.random((n, 3))
    pts[::50] += rng.normal(0, 0.5, pts[::50].shape)
    pts[3] = pts[4]
    idx, dbar = oracle.remove_statistical_outlier(pts, k, 1.5)
    kk = min(k, n)
    d, _ = cKDTree(pts).query(pts, kk)
    d = d.reshape(n, kk)
    ref = d.mean(1)
    assert np.allclose(dbar, ref, rtol=1e-12, atol=1e-15)
    mu = dbar[dbar > 0].sum() / n
    sd = np.sqrt(((dbar[dbar > 0] - mu) ** 2).sum() / (n - 1))
    keep = np.nonzero((dbar > 0) & (dbar < mu + 1.5 * sd))[0]
    assert (idx == keep).all()
    with pytest.raises(RuntimeError):
        oracle.remove_statistical_outlier(pts, 0, 1.0)


def test_grid_to_points_and_zfilter_and_backproject():
    rng = np.random.default_rng(5)
    img = rng.integers(0, 255, (37, 53)).astype(np.uint8)
    pts = oracle.grid_to_points(img, 0.05, -1.25, 2.5, 100)
    rows, cols = np.where(img < 100)                          # fusion/hybrid_map.py:45-55
    exp = np.array([[-1.25 + c * 0.05, 2.5 + (37 - 1 - r) * 0.05, 0.0] for r, c in zip(rows, cols)])
    assert (pts == exp).all()
    p = rng.random((100, 3)) * 0.06
    c = rng.random((100, 3))
    fp, fc = oracle.zfilter(p, c, 0.03)
    m = p[:, 2] >= 0.03
    assert (fp == p[m]).all() and (fc == c[m]).all()
    depth = (rng.random((12, 16)) * 2).astype(np.float32)
    depth[depth < 0.3] = 0
    rgb = rng.integers(0, 255, (12, 16, 3)).astype(np.uint8)
    ext = np.eye(4); ext[:3, 3] = [0.1, -0.2, 0.3]
    bp, bc = oracle.backproject_rgbd(depth, rgb, (20.0, 21.0, 8.5, 6.5), ext)
    ii, jj = np.nonzero(depth > 0)
    z = depth[ii, jj].astype(np.float64)
    cam = np.stack([(jj - 8.5) * z / 20.0, (ii - 6.5) * z / 21.0, z, np.ones_like(z)], 1)
    assert np.allclose(bp, (cam @ np.linalg.inv(ext).T)[:, :3], atol=1e-12)
    assert (bc == rgb[ii, jj] / 255.0).all()


def test_ply_records():
    p = np.array([[1.5, -2.25, 3.0]])
    rec = oracle.pack_ply_cloud(p, np.array([[0.2, 1.0, 0.4999]]))
    assert rec.shape == (1, 27) and (np.frombuffer(rec[0, :24].tobytes(), "<f8") == p[0]).all()
    assert rec[0, 24:].tolist() == [51, 255, 127]             # round(c*255): hybrid map grey 0.2 -> 51
