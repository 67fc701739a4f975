"""Design study for round 2 (CPU, NumPy; not product code): how many voxel-frame pairs of the integration kernel could a conservative
per-group depth test skip?  The judge asked for a per-column early-out (42 % of the visited pairs lie more than the truncation
distance behind the surface).  A safe test needs an upper bound of the depth over every pixel a group of voxels can gather; this
script evaluates the tightest cheap one: per frame, a max-depth image over T x T pixel tiles, and per z-group (4 / 8 / 16 voxels of a
column) the max over the tiles of the bounding box of its two projected end points (+-1 px).  A group is skippable when
min(z_cam) - d_max >= sdf_trunc (then (d - z) * mult <= -trunc for every voxel, mult >= 1).  Reported: fraction of ALL pairs that
are skippable per THREAD (thrG) and per WARP (warpG: all 32 columns of a warp agree -- what a SIMT kernel can actually skip).
Result on the default workload (3 frames, T = 8):  upd 0.417  thr4 0.387 warp4 0.237  thr8 0.324 warp8 0.180  thr16 0.150 warp16 0.050.
A warp-level 4-voxel test costs about 160 instructions per (column, frame) (5 approximate projections, tile indexing, 4-9 tile
loads) against 0.237 x 16 x 50 = 190 saved: not worth a kernel.  Usage: python tools/study_culling_tiles.py [n_frames] [tile_px]"""
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle
from otslam_b200 import synth
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T = int(sys.argv[2]) if len(sys.argv) > 2 else 8          # tile size in px
vl, trunc = 0.005, 0.02
unit = 16 * vl
seq = synth.make_sequence("chair_table", 1000, subsample=(37, 1000 // n_frames))
depth, rgb = seq.numpy()
fx, fy, cx, cy = seq.fxfycxcy
H, W = depth.shape[1:]
ii, jj = np.meshgrid(np.arange(0, H, 4), np.arange(0, W, 4), indexing="ij")
g = (np.arange(16) + 0.5) * vl
vx, vy, vz = np.meshgrid(g, g, g, indexing="ij")
tot = dict(pairs=0, upd=0)
for G in (4, 8, 16):
    tot[f"thr{G}"] = 0; tot[f"warp{G}"] = 0
for k in range(len(seq)):
    d = oracle.depth_convert(depth[k], 1000.0, 3.0).astype(np.float64)
    th, tw = (H + T - 1) // T, (W + T - 1) // T
    dp = np.zeros((th * T, tw * T)); dp[:H, :W] = d
    M = dp.reshape(th, T, tw, T).max(axis=(1, 3))
    E = seq.extrinsic[k]; P = np.linalg.inv(E)
    z = d[ii, jj]; m = z > 0
    x = (jj[m] - cx) * z[m] / fx; y = (ii[m] - cy) * z[m] / fy
    pw = (P[:3, :3] @ np.stack([x, y, z[m]])).T + P[:3, 3]
    keys = set()
    for dx in (-trunc, trunc):
        for dy in (-trunc, trunc):
            for dz in (-trunc, trunc):
                keys |= set(map(tuple, np.floor((pw + [dx, dy, dz]) / unit).astype(int)))
    keys = list(keys)[::3]
    for key in keys:
        o = np.array(key) * unit
        pc = E[:3, :3] @ np.stack([vx.ravel() + o[0], vy.ravel() + o[1], vz.ravel() + o[2]]) + E[:3, 3:4]
        zc = pc[2].reshape(16, 16, 16)
        uf = (pc[0] * fx / pc[2] + cx + 0.5).reshape(16, 16, 16)
        vf = (pc[1] * fy / pc[2] + cy + 0.5).reshape(16, 16, 16)
        u = np.floor(uf).astype(int); v = np.floor(vf).astype(int)
        inside = (zc > 0) & (u >= 0) & (u < W) & (v >= 0) & (v < H)
        dd = np.zeros_like(zc); dd[inside] = d[v[inside], u[inside]]
        upd = inside & (dd > 0) & ((dd - zc) > -trunc)
        tot["pairs"] += 4096; tot["upd"] += int(upd.sum())
        for G in (4, 8, 16):
            ng = 16 // G
            zg = zc.reshape(16, 16, ng, G)
            ug = uf.reshape(16, 16, ng, G); vg = vf.reshape(16, 16, ng, G)
            # bbox from group endpoints with 1 px margin, in tiles
            u0 = np.floor((np.minimum(ug[..., 0], ug[..., -1]) - 1) / T).astype(int); u1 = np.floor((np.maximum(ug[..., 0], ug[..., -1]) + 1) / T).astype(int)
            v0 = np.floor((np.minimum(vg[..., 0], vg[..., -1]) - 1) / T).astype(int); v1 = np.floor((np.maximum(vg[..., 0], vg[..., -1]) + 1) / T).astype(int)
            zmin = np.minimum(zg[..., 0], zg[..., -1])
            need = np.ones((16, 16, ng), bool)
            for a in range(16):
                for b in range(16):
                    for c in range(ng):
                        if zmin[a, b, c] <= 0:
                            continue
                        x0, x1, y0, y1 = max(u0[a, b, c], 0), min(u1[a, b, c], tw - 1), max(v0[a, b, c], 0), min(v1[a, b, c], th - 1)
                        if x0 > x1 or y0 > y1:
                            need[a, b, c] = False; continue
                        if (x1 - x0 + 1) * (y1 - y0 + 1) > 16:
                            continue
                        dmax = M[y0:y1 + 1, x0:x1 + 1].max()
                        need[a, b, c] = zmin[a, b, c] - dmax < trunc
            # sanity: culled groups contain no update
            ug_upd = upd.reshape(16, 16, ng, G).any(axis=3)
            assert not (ug_upd & ~need).any()
            tot[f"thr{G}"] += int((~need).sum()) * G
            # warp = 2 consecutive x rows
            nw = need.reshape(8, 2, 16, ng).any(axis=(1, 2))
            tot[f"warp{G}"] += int((~nw).sum()) * G * 32
p = tot["pairs"]
print("tile", T, {k: round(v / p, 3) for k, v in tot.items()})
