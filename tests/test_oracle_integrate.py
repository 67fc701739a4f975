"""Pins for the CPU oracle's integration path.  The reference holds no tests or
golden vectors for this path and its backend (open3d) cannot run here -> PARITY UNPINNED against
the reference itself; these closed-form checks are what anchors the oracle."""
import numpy as np
import pytest

from oracle import oracle

K = (565.6009, 565.6009, 320.5, 240.5)


def test_inverse4_matches_numpy():
    rng = np.random.default_rng(0)
    for _ in range(20):
        m = rng.normal(size=(4, 4))
        assert np.allclose(oracle.inverse4(m), np.linalg.inv(m), rtol=1e-9, atol=1e-11)
    with pytest.raises(RuntimeError):
        oracle.inverse4(np.zeros((4, 4)))


def test_depth_convert_semantics():
    # u16 / 1000 as float, >= trunc -> 0, zero stays zero
    d = np.array([[0, 1, 999, 2999, 3000, 3001, 65535]], np.uint16)
    out = oracle.depth_convert(d, 1000.0, 3.0)
    exp = d.astype(np.float32) / np.float32(1000.0)
    exp[exp.astype(np.float64) >= 3.0] = 0
    assert out.dtype == np.float32 and (out == exp).all()
    assert out[0, 3] > 0 and out[0, 4] == 0 and out[0, 0] == 0


def plane_frame(z_mm, W=640, H=480, color=(10, 20, 30)):
    depth = np.full((H, W), z_mm, np.uint16)
    rgb = np.broadcast_to(np.array(color, np.uint8), (H, W, 3)).copy()
    return depth, rgb


def test_plane_tsdf_closed_form():
    """A fronto-parallel plane at depth d seen from the origin (extrinsic = I): every updated voxel
    must carry tsdf = min(1, (d - z) * mult / trunc) with mult = sqrt(xx^2 + yy^2 + 1) at its pixel,
    weight 1 and the plane colour."""
    vl, trunc = 0.01, 0.04
    depth, rgb = plane_frame(1500)
    v = oracle.Volume(vl, trunc)
    nt, nu = v.integrate(oracle.depth_convert(depth), rgb, K, np.eye(4))
    keys, tsdf, w, col = v.export_blocks()
    assert nt == len(keys) and nu == int(w.sum())
    # voxel centres (FP64 is plenty to predict which side of the band a voxel is on, away from ties)
    idx = np.arange(4096)
    lx, ly, lz = idx >> 8, (idx >> 4) & 15, idx & 15
    for b in range(0, len(keys), max(1, len(keys) // 25)):
        cx = (keys[b, 0] * 16 + lx + 0.5) * vl
        cy = (keys[b, 1] * 16 + ly + 0.5) * vl
        cz = (keys[b, 2] * 16 + lz + 0.5) * vl
        u = np.floor(cx * K[0] / np.maximum(cz, 1e-9) + K[2] + 0.5)
        vv = np.floor(cy * K[1] / np.maximum(cz, 1e-9) + K[3] + 0.5)
        mult = np.sqrt(((u - K[2]) / K[0]) ** 2 + ((vv - K[3]) / K[1]) ** 2 + 1.0)
        sdf = (1.5 - cz) * mult
        pred = np.minimum(1.0, sdf / trunc)
        upd = w[b] > 0
        # interior of the update region (not within 1e-4 of a decision boundary)
        sure = (cz > 1e-3) & (u >= 1) & (u < 639) & (vv >= 1) & (vv < 479) & (sdf > -trunc + 1e-4)
        assert (upd[sure]).all()
        assert np.abs(tsdf[b][sure] - pred[sure]).max() < 2e-4
        assert not upd[(sdf < -trunc - 1e-4)].any()
        assert (w[b][upd] == 1).all()
        assert np.allclose(col[b][upd], [10, 20, 30])


def test_weights_count_frames_and_running_mean():
    vl, trunc = 0.01, 0.04
    v = oracle.Volume(vl, trunc)
    frames = [(1500, (10, 20, 30)), (1500, (30, 60, 90)), (1500, (20, 10, 60))]
    for z, c in frames:
        d, rgb = plane_frame(z, color=c)
        v.integrate(oracle.depth_convert(d), rgb, K, np.eye(4))
    keys, tsdf, w, col = v.export_blocks()
    assert set(np.unique(w)) <= {0.0, 3.0}                 # identical views: every seen voxel seen 3 times
    seen = w > 0
    assert np.allclose(col[seen], np.mean([f[1] for f in frames], axis=0), atol=1e-9)
    # the f32 running mean of three equal samples reproduces the sample to rounding
    v1 = oracle.Volume(vl, trunc)
    d, rgb = plane_frame(1500)
    v1.integrate(oracle.depth_convert(d), rgb, K, np.eye(4))
    _, t1, w1, _ = v1.export_blocks()
    assert np.abs(tsdf[seen] - t1[w1 > 0]).max() < 1e-6


def test_invalid_depth_and_trunc_are_skipped():
    vl, trunc = 0.01, 0.04
    d, rgb = plane_frame(1500)
    d[:, :320] = 0                      # invalid half
    d[:240, 320:] = 3500                # beyond depth_trunc = 3.0 -> masked by A.1
    v = oracle.Volume(vl, trunc)
    v.integrate(oracle.depth_convert(d), rgb, K, np.eye(4))
    keys, tsdf, w, col = v.export_blocks()
    # only the lower-right quadrant (x > 0, y > 0 in camera coordinates) can hold observations
    obs = np.argwhere(w > 0)
    gx = keys[obs[:, 0], 0] * 16 + (obs[:, 1] >> 8)
    gy = keys[obs[:, 0], 1] * 16 + ((obs[:, 1] >> 4) & 15)
    assert (gx >= -2).all() and (gy >= -2).all()
    empty = oracle.Volume(vl, trunc)
    nt, nu = empty.integrate(np.zeros((480, 640), np.float32), rgb, K, np.eye(4))
    assert (nt, nu) == (0, 0) and empty.num_blocks() == 0


def test_allocation_rule(small_seq):
    """Touched blocks = union over stride-4 samples of the +-trunc box, independent
    restatement in NumPy FP64."""
    seq, d, c = small_seq
    vl, trunc = 0.02, 0.08
    v = oracle.Volume(vl, trunc)
    df = oracle.depth_convert(d[0])
    v.integrate(df, c[0], seq.fxfycxcy, seq.extrinsic[0])
    keys, _, _, _ = v.export_blocks()
    fx, fy, cx, cy = seq.fxfycxcy
    pose = np.linalg.inv(seq.extrinsic[0])
    ii, jj = np.meshgrid(np.arange(0, df.shape[0], 4), np.arange(0, df.shape[1], 4), indexing="ij")
    z = df[ii, jj].astype(np.float64)
    ok = z > 0
    P = np.stack([(jj - cx) * z / fx, (ii - cy) * z / fy, z, np.ones_like(z)], -1)[ok] @ pose.T
    unit = vl * 16
    lo = np.floor((P[:, :3] - trunc) / unit).astype(int)
    hi = np.floor((P[:, :3] + trunc) / unit).astype(int)
    want = set()
    for l, h in zip(lo, hi):
        for x in range(l[0], h[0] + 1):
            for y in range(l[1], h[1] + 1):
                for zz in range(l[2], h[2] + 1):
                    want.add((x, y, zz))
    assert want == set(map(tuple, keys.tolist()))


def test_slab_partition_covers_volume(small_seq):
    seq, d, c = small_seq
    vl, trunc = 0.02, 0.08
    full = oracle.Volume(vl, trunc)
    parts = [oracle.Volume(vl, trunc, slab=(0, 2, 3, r)) for r in range(3)]
    for k in range(len(seq)):
        df = oracle.depth_convert(d[k])
        for v in [full] + parts:
            v.integrate(df, c[k], seq.fxfycxcy, seq.extrinsic[k])
    fk, ft, fw, _ = full.export_blocks()
    fmap = {tuple(k): i for i, k in enumerate(fk.tolist())}
    owned_total = 0
    for r, v in enumerate(parts):
        pk, pt, pw, _ = v.export_blocks()
        for i, k in enumerate(pk.tolist()):
            j = fmap[tuple(k)]
            assert (pw[i] == fw[j]).all() and (pt[i] == ft[j]).all()     # replicated halo blocks are bit-identical
            owned_total += ((k[0] // 2) % 3 == r)
    assert owned_total == len(fk)
    # extraction: union of the per-rank point clouds == the full cloud
    fp, fc, fe = full.extract_point_cloud()
    pp = [v.extract_point_cloud() for v in parts]
    pe = np.concatenate([p[2] for p in pp])
    assert len(pe) == len(fe)
    from conftest import lexorder
    a, b = lexorder(fe), lexorder(pe)
    assert (fe[a] == pe[b]).all() and (fp[a] == np.concatenate([p[0] for p in pp])[b]).all()
