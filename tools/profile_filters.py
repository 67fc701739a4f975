"""Workload for the ncu captures of the filter kernels (voxel_down_sample K10, remove_statistical_outlier
K11) at the config-5 object size: 1 M points sampled on the surfaces of a table-like solid, one call of
each filter through the C ABI.  Usage (see profiles/README.md):
    ncu --set full --clock-control none -k regex:'knn_mean_dist|radix_scatter|voxel_mean' -c 12 \
        -o gpurun_out/prof_filters python tools/profile_filters.py
Without ncu it prints the operators' wall / device times and HBM figures as one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import otslam_b200.o3d_compat as o3d
from otslam_b200 import _lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(0)
# points on the faces of a 1.2 x 0.8 x 0.75 box (table envelope) with 1 mm noise + 1 % far outliers
face = rng.integers(0, 6, N)
uvw = rng.random((N, 3)) * np.array([1.2, 0.8, 0.75])
ax = face // 2
uvw[np.arange(N), ax] = np.where(face % 2 == 0, 0.0, np.array([1.2, 0.8, 0.75])[ax])
pts = uvw + rng.normal(scale=1e-3, size=(N, 3))
out = rng.random(N) < float(os.environ.get("OUTLIER_FRAC", "0.01"))
pts[out] += rng.normal(scale=0.2, size=(int(out.sum()), 3))
cols = rng.random((N, 3))
pc = o3d.geometry.PointCloud()
pc.points, pc.colors = pts, cols

warm = o3d.geometry.PointCloud()           # CUDA context + library load outside the timed calls
warm.points = pts[:1000]
warm.voxel_down_sample(0.05)
res = {"points": N}
for name, fn, alg in (("voxel_down_sample", lambda: pc.voxel_down_sample(0.01), lambda r: 36 * (N + len(r.points))),
                      ("remove_statistical_outlier", lambda: pc.remove_statistical_outlier(20, 2.0),
                       lambda r: 24 * N * 2 + 8 * N + 36 * len(r[1]))):
    wall = dev = 1e30
    for _ in range(1 if os.environ.get("OTSLAM_PROFILE_ONCE") else 3):     # best of 3: the first call also fills the scratch cache
        t0 = time.perf_counter()
        r = fn()
        wall = min(wall, time.perf_counter() - t0)
        dev = min(dev, _lib.last_op_device_ms())
    b = alg(r)
    res[name] = {"wall_ms": 1e3 * wall, "device_ms": dev, "algorithmic_MB": b / 1e6, "device_GBps": b / dev / 1e6,
                 "out": len(r.points) if name == "voxel_down_sample" else len(r[1])}
print(json.dumps(res))
