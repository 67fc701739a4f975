"""GPU check: stateless cloud / mesh operators vs the oracle (dev tool)."""
import sys, time, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np
from otslam_b200 import synth, _lib
from oracle import oracle
L = _lib.lib; P = _lib.ptr
rng = np.random.default_rng(0)

# depth convert
d = rng.integers(0, 6000, (480, 640)).astype(np.uint16)
out = np.empty(d.shape, np.float32); _lib.check(L.otslam_depth_convert(P(d), d.size, 1000.0, 3.0, P(out), 0))
print("depth_convert equal", bool((out == oracle.depth_convert(d)).all()))

# backproject
seq = synth.make_sequence("table", 300, subsample=(0, 150)); dd, cc = seq.numpy()
df = oracle.depth_convert(dd[0], 1000.0, 5.0)
op, oc = oracle.backproject_rgbd(df, cc[0], seq.fxfycxcy, seq.extrinsic[0])
gp = np.empty((df.size, 3)); gc = np.empty((df.size, 3)); n = C.c_int64(0)
k = np.array(seq.fxfycxcy, np.float64); e = np.ascontiguousarray(seq.extrinsic[0])
_lib.check(L.otslam_backproject_rgbd(P(df), P(cc[0]), 640, 480, P(k), P(e), P(gp), P(gc), C.byref(n), 0))
print("backproject n", n.value, len(op), "pts equal", bool((gp[:n.value] == op).all()), "cols equal", bool((gc[:n.value] == oc).all()))

# mesh ops on an oracle mesh
ov = oracle.Volume(0.01, 0.04)
for i in range(len(seq)): ov.integrate(oracle.depth_convert(dd[i]), cc[i], seq.fxfycxcy, seq.extrinsic[i])
v, col, f, ek = ov.extract_triangle_mesh()
on = oracle.vertex_normals(v, f)
gn = np.empty_like(v); _lib.check(L.otslam_mesh_vertex_normals(P(v), len(v), P(f), len(f), P(gn), 0))
print("normals maxabs", np.abs(on - gn).max())
ns = 100000
osp, osc, osn, tri = oracle.sample_uniform(v, col, on, f, ns, seed=7)
gsp = np.empty((ns, 3)); gsc = np.empty((ns, 3)); gsn = np.empty((ns, 3))
t = time.time(); _lib.check(L.otslam_mesh_sample_uniform(P(v), P(col), P(on), len(v), P(f), len(f), ns, 7, P(gsp), P(gsc), P(gsn), 0)); print("sample t", time.time() - t)
print("sample pts equal", bool((gsp == osp).all()), "maxabs", np.abs(gsp - osp).max(), "cols equal", bool((gsc == osc).all()), "normals equal", bool((gsn == osn).all()))

# zfilter
zp, zc = oracle.zfilter(osp, osc, 0.03)
gzp = np.empty_like(osp); gzc = np.empty_like(osc); m = C.c_int64(0)
_lib.check(L.otslam_cloud_zfilter(P(osp), P(osc), ns, 0.03, P(gzp), P(gzc), C.byref(m), 0))
print("zfilter n", m.value, len(zp), "equal", bool((gzp[:m.value] == zp).all() and (gzc[:m.value] == zc).all()))

# voxel down sample
for vs in (0.01, 0.05):
    a = oracle.voxel_down_sample(osp, osc, vs)
    gp2 = np.empty_like(osp); gc2 = np.empty_like(osc); gk = np.empty((ns, 3), np.int32); gn2 = np.empty(ns, np.int32)
    _lib.check(L.otslam_cloud_voxel_down_sample(P(osp), P(osc), ns, vs, P(gp2), P(gc2), P(gk), P(gn2), C.byref(m), 0))
    mm = m.value
    print("vds", vs, mm, len(a[0]), "keys", bool(mm == len(a[0]) and (gk[:mm] == a[2]).all()), "counts", bool((gn2[:mm] == a[3]).all()),
          "pts equal", bool((gp2[:mm] == a[0]).all()), "cols equal", bool((gc2[:mm] == a[1]).all()))

# SOR
for npts, kk in ((20000, 20), (100000, 20), (5000, 64)):
    pts = osp[:npts].copy(); pts[::97] += rng.normal(0, 0.05, pts[::97].shape)   # outliers
    pts[5] = pts[6]                                                           # duplicate
    t = time.time(); oi, odb = oracle.remove_statistical_outlier(pts, kk, 2.0); to = time.time() - t
    gi = np.empty(npts, np.int64); gdb = np.empty(npts)
    t = time.time(); _lib.check(L.otslam_cloud_remove_statistical_outlier(P(pts), npts, kk, 2.0, P(gi), C.byref(m), P(gdb), 0)); tg = time.time() - t
    print("sor", npts, kk, "kept", m.value, len(oi), "idx equal", bool(m.value == len(oi) and (gi[:m.value] == oi).all()), "dbar equal", bool((gdb == odb).all()),
          "maxabs", np.abs(gdb - odb).max(), "t_oracle %.3f t_gpu %.3f" % (to, tg))

# grid to points + merge
img = synth.occupancy_map(500, 400, 0.02, 1)
og = oracle.grid_to_points(img, 0.05, -10.0, -7.5, 100)
gg = np.empty((img.size, 3)); _lib.check(L.otslam_grid_to_points(P(img), 500, 400, 0.05, -10.0, -7.5, 100, P(gg), C.byref(m), 0))
print("grid n", m.value, len(og), "equal", bool((gg[:m.value] == og).all()))
clouds = [og, osp[:1000], osp[1000:1777]]
paint = np.array([[0.2, 0.2, 0.2], [1, 0, 0], [1, 0, 0]], np.float64)
tot = sum(len(c) for c in clouds)
ref = np.concatenate([oracle.pack_ply_cloud(c, np.tile(p, (len(c), 1))) for c, p in zip(clouds, paint)])
outb = np.empty((tot, 27), np.uint8)
pp = (C.c_void_p * 3)(*[c.ctypes.data for c in clouds]); cnt = np.array([len(c) for c in clouds], np.int64)
_lib.check(L.otslam_cloud_merge_pack(3, pp, None, P(cnt), P(paint), P(outb), 0))
print("merge paint equal", bool((outb == ref).all()))
cols = [np.tile(p, (len(c), 1)) * 0.77 for c, p in zip(clouds, paint)]
ref2 = np.concatenate([oracle.pack_ply_cloud(c, k_) for c, k_ in zip(clouds, cols)])
cp = (C.c_void_p * 3)(*[c.ctypes.data for c in cols])
_lib.check(L.otslam_cloud_merge_pack(3, pp, cp, P(cnt), None, P(outb), 0))
print("merge colors equal", bool((outb == ref2).all()))
