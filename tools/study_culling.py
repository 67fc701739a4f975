"""Design study for round 2 (CPU, NumPy; not product code): how much of the integration kernel's work could a
conservative depth-range test remove?  For sampled frames of the default bench workload it rebuilds, per block the
frame touches, the kernel's per-voxel outcome (updated / behind the surface / outside the image or invalid depth)
and asks how many 8-voxel z-groups of a column (the kernel's gather group) lie entirely behind
    d_max(footprint of the block in the image) + sdf_trunc
-- the test one thread could do per group from one per-(block, frame) number.  Prints fractions of voxel-frame
pairs.  Usage: python tools/study_culling.py [n_frames]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import oracle
from otslam_b200 import synth

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 6
vl, trunc = 0.005, 0.02
unit = 16 * vl
seq = synth.make_sequence("chair_table", 1000, subsample=(0, 1000 // n_frames))
depth, rgb = seq.numpy()
fx, fy, cx, cy = seq.fxfycxcy
H, W = depth.shape[1:]
tot = dict(pairs=0, updated=0, behind=0, outside=0, culled_groups_pairs=0, culled_blocks_pairs=0)
ii, jj = np.meshgrid(np.arange(0, H, 4), np.arange(0, W, 4), indexing="ij")
g = (np.arange(16) + 0.5) * vl
vx, vy, vz = np.meshgrid(g, g, g, indexing="ij")                      # voxel centres inside a block, index [x][y][z]
for k in range(len(seq)):
    d = oracle.depth_convert(depth[k], 1000.0, 3.0).astype(np.float64)
    E = seq.extrinsic[k]
    P = np.linalg.inv(E)
    # allocation: stride-4 samples, +-trunc box -> block keys
    z = d[ii, jj]
    m = z > 0
    x = (jj[m] - cx) * z[m] / fx
    y = (ii[m] - cy) * z[m] / fy
    pw = (P[:3, :3] @ np.stack([x, y, z[m]])).T + P[:3, 3]
    keys = set()
    for dx in (-trunc, trunc):
        for dy in (-trunc, trunc):
            for dz in (-trunc, trunc):
                keys |= set(map(tuple, np.floor((pw + [dx, dy, dz]) / unit).astype(int)))
    for key in keys:
        o = np.array(key) * unit
        pc = E[:3, :3] @ np.stack([vx.ravel() + o[0], vy.ravel() + o[1], vz.ravel() + o[2]]) + E[:3, 3:4]
        zc = pc[2]
        u = np.floor(pc[0] * fx / zc + cx + 0.5).astype(int)
        v = np.floor(pc[1] * fy / zc + cy + 0.5).astype(int)
        inside = (zc > 0) & (u >= 0) & (u < W) & (v >= 0) & (v < H)
        dd = np.zeros_like(zc)
        dd[inside] = d[v[inside], u[inside]]
        valid = inside & (dd > 0)
        upd = valid & ((dd - zc) > -trunc)          # the multiplier (>= 1) only makes the band wider along the ray: ignored here
        behind = valid & ~upd
        tot["pairs"] += 4096
        tot["updated"] += int(upd.sum()); tot["behind"] += int(behind.sum()); tot["outside"] += int((~valid).sum())
        if inside.any():
            u0, u1, v0, v1 = u[inside].min(), u[inside].max(), v[inside].min(), v[inside].max()
            dmax = d[v0:v1 + 1, u0:u1 + 1].max()
        else:
            dmax = 0.0
        zc3 = zc.reshape(16, 16, 16)                # [x][y][z]
        grp_min = zc3.reshape(16, 16, 2, 8).min(axis=3)
        culled = grp_min > dmax + trunc
        tot["culled_groups_pairs"] += int(culled.sum()) * 8
        if zc.min() > dmax + trunc:
            tot["culled_blocks_pairs"] += 4096
p = tot["pairs"]
print(f"frames {len(seq)}  voxel-frame pairs {p}")
for k in ("updated", "behind", "outside", "culled_groups_pairs", "culled_blocks_pairs"):
    print(f"  {k:22s} {tot[k] / p:6.3f}")
