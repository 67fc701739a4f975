"""Timings of the other kernels of the path at BASELINE.json config sizes (configs 2, 3, 5):
surface extraction, mesh sampling + z-filter, voxel_down_sample, statistical outlier removal,
hybrid-map merge, and the per-frame (non-batched) integrate API.  Wall-clock through the C ABI
(host buffers in, host buffers out) next to the oracle on the host cores; kernel-only times come
from the ncu launch list of this same script (profiles/launches_ops_*.md)."""
import ctypes as C
import json
import sys
import time

sys.path.insert(0, "/root/repo")
import numpy as np

from oracle import oracle
from otslam_b200 import _lib, synth
from otslam_b200.volume import TSDFVolume
import otslam_b200.o3d_compat as o3d

small = "--small" in sys.argv
out = {}


def timed(fn, reps=3):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    return (time.perf_counter() - t0) / reps, r


# ---- integration: per-frame API vs batched, config-1 table scene at 5 mm
n_frames = 32 if small else 300
seq = synth.make_sequence("table", n_frames)
d, c = seq.numpy()
vol = TSDFVolume(0.005, 0.02)
t0 = time.perf_counter()
for k in range(len(seq)):
    vol.integrate_u16(d[k], c[k], seq.fxfycxcy, seq.extrinsic[k])
out["integrate_per_frame_api_fps"] = len(seq) / (time.perf_counter() - t0)
vol.reset()
t0 = time.perf_counter()
vol.integrate_batch(d, c, seq.fxfycxcy, seq.extrinsic)
out["integrate_batch_api_pageable_fps"] = len(seq) / (time.perf_counter() - t0)
st = vol.stats()
out["volume"] = st

# ---- config 3: 4 objects, each its own volume (reference defaults 1 cm / 4 cm), sequential vs concurrent
from concurrent.futures import ThreadPoolExecutor
import torch
nf3 = 16 if small else 150
objs3 = []
for scene in ("table", "chair", "cone", "cardboard"):
    sq = synth.make_sequence(scene, nf3, device="cuda")
    hd_ = torch.empty(sq.depth.shape, dtype=sq.depth.dtype, pin_memory=True); hd_.copy_(sq.depth)
    hc_ = torch.empty(sq.rgb.shape, dtype=sq.rgb.dtype, pin_memory=True); hc_.copy_(sq.rgb)
    objs3.append((TSDFVolume(0.01, 0.04), hd_, hc_, sq))
torch.cuda.synchronize()
def run_obj(o):
    o[0].reset(); o[0].integrate_batch(o[1], o[2], o[3].fxfycxcy, o[3].extrinsic); return o[0].stats()["weight_sum"]
for o in objs3: run_obj(o)
t0 = time.perf_counter(); seq_sums = [run_obj(o) for o in objs3]; t_seq = time.perf_counter() - t0
with ThreadPoolExecutor(4) as ex:
    t0 = time.perf_counter(); par_sums = list(ex.map(run_obj, objs3)); t_par = time.perf_counter() - t0
out["config3_4objects_sequential_fps"] = 4 * nf3 / t_seq
out["config3_4objects_concurrent_fps"] = 4 * nf3 / t_par
out["config3_results_equal"] = seq_sums == par_sums
for o in objs3: o[0].close()

# ---- extraction (K5/K6/K7)
t, mesh = timed(lambda: vol.extract_triangle_mesh(), 2)
verts, cols, nrm, faces, ek = mesh
out["extract_mesh_s"] = t
out["mesh"] = {"vertices": len(verts), "faces": len(faces)}
out["extract_mesh_algorithmic_MB"] = (8 * 4096 * st["n_blocks"] + 48 * len(verts) + 12 * len(faces)) / 1e6
t, pc = timed(lambda: vol.extract_point_cloud(), 2)
out["extract_points_s"] = t
out["points"] = len(pc[0])

# ---- sampling + z filter (K8/K9), config 2 post stage
m = o3d.geometry.TriangleMesh()
m.vertices, m.vertex_colors, m.vertex_normals, m.triangles = verts, cols, nrm, faces
t, pcd = timed(lambda: m.sample_points_uniformly(100000, seed=0))
out["sample_100k_s"] = t
t0 = time.perf_counter()
osp, osc, _, _ = oracle.sample_uniform(verts, cols, None, faces, 100000, 0)
out["sample_100k_oracle_s"] = time.perf_counter() - t0

# ---- filters at 1 M points (K10/K11)
N = 100000 if small else 1000000
pts, cl, _, _ = oracle.sample_uniform(verts, cols, None, faces, N, 1)
cloud = o3d.geometry.PointCloud()
cloud.points, cloud.colors = pts, cl
t, ds = timed(lambda: cloud.voxel_down_sample(0.01))
out["voxel_down_sample_1M_s"] = t
out["voxel_down_sample_out"] = len(ds.points)
out["voxel_down_sample_algorithmic_MB"] = 36 * (N + len(ds.points)) / 1e6
t0 = time.perf_counter()
oracle.voxel_down_sample(pts, cl, 0.01)
out["voxel_down_sample_1M_oracle_s"] = time.perf_counter() - t0
t, (sel, idx) = timed(lambda: cloud.remove_statistical_outlier(20, 2.0), 2)
out["sor_1M_k20_s"] = t
out["sor_kept"] = len(idx)
out["sor_algorithmic_MB"] = (24 * N * 2 + 8 * N + 36 * len(idx)) / 1e6
t0 = time.perf_counter()
oi, _ = oracle.remove_statistical_outlier(pts, 20, 2.0)
out["sor_1M_k20_oracle_s"] = time.perf_counter() - t0
out["sor_indices_equal"] = bool(len(oi) == len(idx) and (np.array(idx) == oi).all())

# ---- hybrid map merge (K12/K13), config 5: 2000x2000 map + 20 x 1 M-point objects
img = synth.occupancy_map(500 if small else 2000, 500 if small else 2000, 0.02, 0)
n_obj, per = (4, 100000) if small else (20, 1000000)
rng = np.random.default_rng(0)
objs = [rng.normal(size=(per, 3)) for _ in range(n_obj)]
mp = np.empty((img.size, 3))
nmap = C.c_int64(0)
t, _ = timed(lambda: _lib.check(_lib.lib.otslam_grid_to_points(_lib.ptr(img), img.shape[1], img.shape[0], 0.05, -50.0, -50.0, 100,
                                                                _lib.ptr(mp), C.byref(nmap), 0)))
out["grid_to_points_s"] = t
mp = mp[:nmap.value]
t0 = time.perf_counter()
rows, colsi = np.where(img < 100)
ref_pts = [[-50.0 + cc * 0.05, -50.0 + (img.shape[0] - 1 - r) * 0.05, 0.0] for r, cc in zip(rows, colsi)]   # the reference's Python loop
out["grid_to_points_reference_loop_s"] = time.perf_counter() - t0
paint = [[0.2, 0.2, 0.2]] + [[1, 0, 0]] * n_obj
t, rec = timed(lambda: o3d.io.pack_cloud_records([mp] + objs, paint=paint), 2)
out["merge_pack_s"] = t
out["merge_points"] = int(len(rec))
out["merge_algorithmic_MB"] = (img.size + 24 * (len(rec) - len(mp)) + 27 * len(rec)) / 1e6
t0 = time.perf_counter()
ref = np.concatenate([oracle.pack_ply_cloud(p, np.tile(q, (len(p), 1))) for p, q in zip([mp] + objs, paint)])
out["merge_pack_oracle_s"] = time.perf_counter() - t0
out["merge_bytes_equal"] = bool((rec == ref).all())
print(json.dumps(out, indent=1))
