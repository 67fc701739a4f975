"""GPU parity on the configurations bench.py actually measures (BASELINE.json configs[1] and configs[3]), against the
CPU oracle / the committed HD golden (tests/golden/hd_sequence.npz, made by tools/make_golden.py):
  * configs[1]: a CONTIGUOUS 64-frame window of the 1000-frame chair+table sequence at 5 mm -- two full 32-frame
    fused batches whose every voxel is oracle-checked;
  * configs[3]: 1280x720, K x2, 2 mm voxels on one GPU, and as 8 diagonal-slab ranks (emulated on one GPU) with
    the device-resident halo exchange, whose merged extraction must equal the single-volume one;
  * the stream contract of OTSLAM_MEM_DEVICE inputs (frames produced on a torch stream).
Tolerances (north_star): keys / weights bit-exact, TSDF <= 1e-4 (expected: bit-exact), colour <= 1/255.
"""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, canon_mesh, lexorder
from oracle import oracle

pytestmark = pytest.mark.gpu


def digest(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_hd():
    z = np.load(os.path.join(ROOT, "tests", "golden", "hd_sequence.npz"))
    return z, json.loads(bytes(z["expected"]).decode())


def test_config1_contiguous_64_frame_window_5mm():
    """Frames 468..531 of configs[1] (straddles the switch from the low to the high camera ring at frame 500): two
    full 32-frame batches, every block / weight / TSDF / colour compared with the oracle."""
    from otslam_b200 import synth
    from otslam_b200.volume import TSDFVolume
    seq = synth.make_sequence("chair_table", 1000, poses=synth.trajectory("chair_table", 1000)[468:532], device="cuda")
    d, c = seq.numpy()
    ext = seq.extrinsic
    assert len(d) == 64
    vl = 0.005
    ov = oracle.Volume(vl, 4 * vl)
    nupd = 0
    for k in range(64):
        nupd += ov.integrate(oracle.depth_convert(d[k], 1000.0, 3.0), c[k], seq.fxfycxcy, ext[k])[1]
    gv = TSDFVolume(vl, 4 * vl)
    gv.integrate_batch(d, c, seq.fxfycxcy, ext)
    gk, gt, gw, gc = gv.export_blocks()
    ok, ot, ow, oc = ov.export_blocks()
    assert gk.shape == ok.shape and (gk == ok).all()
    assert (gw == ow).all() and int(gw.max()) > 32, "a voxel must have been updated by frames of both batches"
    assert (gt == ot).all(), "TSDF running mean must be reproduced bit for bit (tolerance 1e-4)"
    assert np.abs(gc - oc)[ow > 0].max() <= 1.0
    assert gv.stats()["weight_sum"] == nupd
    gv.close()


def test_hd_single_gpu_matches_golden():
    """configs[3] geometry on one GPU: 8 frames of the large room at 1280x720 / 2 mm vs the oracle's digests."""
    from otslam_b200.volume import TSDFVolume
    z, exp = load_hd()
    v = TSDFVolume(float(z["voxel"][0]), float(z["voxel"][1]))
    v.integrate_batch(z["depth"], z["rgb"], tuple(z["intr"]), z["extrinsic"])
    assert v.stats()["weight_sum"] == sum(u for _, u in exp["touched_updated"])
    keys, tsdf, w, col = v.export_blocks()
    assert len(keys) == exp["n_blocks"] and digest(keys) == exp["keys"]
    assert digest(w.astype(np.uint16)) == exp["weight"]
    assert digest(tsdf) == exp["tsdf"]
    assert digest(np.floor(col.astype(np.float64) + 0.5).astype(np.uint8)) == exp["color_u8"]
    del tsdf, w, col
    verts, cols, nrm, faces, ek = v.extract_triangle_mesh()
    o = lexorder(ek)
    assert (len(verts), len(faces)) == (exp["mesh_nv"], exp["mesh_nf"])
    assert digest(ek[o]) == exp["mesh_ekeys"] and digest(verts[o]) == exp["mesh_verts"]
    pts, pcols, pek = v.extract_point_cloud()
    po = lexorder(pek)
    assert len(pts) == exp["pc_n"] and digest(pek[po]) == exp["pc_ekeys"] and digest(pts[po]) == exp["pc_pts"]
    v.close()


def exchange_on_one_gpu(parts):
    """The halo exchange of slab.exchange_halo with the ranks emulated as volumes on one GPU: device-packed pieces
    (otslam_volume_halo_pack / _fetch into CUDA tensors) handed to the destination volume's import as device pointers."""
    packed = [v.halo_pack_tensors() for v in parts]
    received = [0] * len(parts)
    for src, (keys, planes, counts) in enumerate(packed):
        assert keys.is_cuda and planes.is_cuda and int(counts.sum()) == len(keys) and counts[src] == 0
        off = 0
        for dst, n in enumerate(int(x) for x in counts):
            if n:
                parts[dst].halo_import(keys[off:off + n].contiguous(), planes[off:off + n].contiguous())
                received[dst] += n
            off += n
    return received


def test_hd_eight_diagonal_slab_ranks_with_device_halo_exchange():
    """configs[3] as bench.py --gpus 8 --hd shards it: diagonal slabs of one block, owned blocks only, boundary pieces
    exchanged in HBM, per-rank extraction merged -- must reproduce the golden (= single-volume oracle) mesh and points."""
    import torch
    from otslam_b200 import slab as slabmod
    from otslam_b200.volume import TSDFVolume
    z, exp = load_hd()
    vl, trunc = float(z["voxel"][0]), float(z["voxel"][1])
    dd, cc = torch.from_numpy(z["depth"]).cuda(), torch.from_numpy(z["rgb"]).cuda()
    parts = []
    for r in range(8):
        v = TSDFVolume(vl, trunc, slab=(3, 1, 8, r, 0))
        v.integrate_batch(dd, cc, tuple(z["intr"]), z["extrinsic"])
        parts.append(v)
    owned = [v.num_blocks() for v in parts]
    assert sum(owned) == exp["n_blocks"], "diagonal slabs must partition the blocks (no replicated integration)"
    assert max(owned) < 1.15 * min(owned), "diagonal slabs balance the room's walls and floor across ranks"
    assert sum(v.stats()["weight_sum"] for v in parts) == sum(u for _, u in exp["touched_updated"])
    got = exchange_on_one_gpu(parts)
    assert all(g > 0 for g in got)
    # points: concatenation of the ranks' extractions == golden
    pe = [v.extract_point_cloud_tensors() for v in parts]
    pts = torch.cat([p[0] for p in pe]).cpu().numpy()
    pek = torch.cat([p[2] for p in pe]).cpu().numpy()
    del pe
    po = lexorder(pek)
    assert len(pts) == exp["pc_n"] and digest(pek[po]) == exp["pc_ekeys"] and digest(pts[po]) == exp["pc_pts"]
    del pts, pek, po
    # mesh: per-rank meshes merged on the device (duplicates on slab boundaries unified by edge key) == golden
    me = [v.extract_mesh_tensors() for v in parts]
    base, faces = 0, []
    for m in me:
        faces.append(m[2].to(torch.int64) + base)
        base += m[0].shape[0]
    verts, cols, f, ek = slabmod.merge_mesh_tensors(torch.cat([m[0] for m in me]), torch.cat([m[1] for m in me]), torch.cat(faces),
                                                    torch.cat([m[3] for m in me]))
    assert (verts.shape[0], f.shape[0]) == (exp["mesh_nv"], exp["mesh_nf"])
    ek, verts = ek.cpu().numpy(), verts.cpu().numpy()
    o = lexorder(ek)
    assert digest(ek[o]) == exp["mesh_ekeys"] and digest(verts[o]) == exp["mesh_verts"]
    for v in parts:
        v.close()


@pytest.mark.parametrize("n_ranks,thickness,axis", [(2, 1, 0), (3, 2, 1), (4, 1, 3), (8, 1, 3)])
def test_device_halo_exchange_equals_host_exchange(table_seq, n_ranks, thickness, axis):
    """halo_pack_tensors (HBM-resident, grouped by destination) carries exactly the pieces of the host-side
    halo_export, and importing them from device pointers gives the same volumes."""
    from otslam_b200.volume import TSDFVolume
    seq, d, c = table_seq

    def build():
        vs = []
        for r in range(n_ranks):
            v = TSDFVolume(0.01, 0.04, slab=(axis, thickness, n_ranks, r, 0))
            v.integrate_batch(d, c, seq.fxfycxcy, seq.extrinsic)
            vs.append(v)
        return vs
    A, B = build(), build()
    for r, v in enumerate(A):
        keys, dest, planes = v.halo_export()                         # host form
        tk, tp, counts = B[r].halo_pack_tensors()                    # device form
        assert (np.bincount(dest, minlength=n_ranks) == counts).all()
        assert (np.repeat(np.arange(n_ranks), counts) == dest).all(), "pieces are grouped by destination rank"
        assert (tk.cpu().numpy() == keys).all() and (tp.cpu().numpy() == planes).all()
    exports = [v.halo_export() for v in A]
    for r, v in enumerate(A):
        for keys, dest, planes in exports:
            if (dest == r).any():
                v.halo_import(keys[dest == r], planes[dest == r])
    exchange_on_one_gpu(B)
    for a, b in zip(A, B):
        ea, eb = a.export_blocks(), b.export_blocks()
        assert all((x == y).all() for x, y in zip(ea, eb))
        a.close(); b.close()


def test_device_frames_from_a_busy_torch_stream():
    """ADVICE r1 (slab.py:210): OTSLAM_MEM_DEVICE frames that a torch stream is still producing when integrate_batch is
    called.  The volume runs on its own non-blocking streams; integrate_batch must order them after torch's current
    stream (otslam_volume_wait_stream), otherwise it integrates whatever the buffer held before."""
    import torch
    from otslam_b200 import synth
    from otslam_b200.volume import TSDFVolume
    seq = synth.make_sequence("table", 300, subsample=(0, 30))
    d, c = seq.numpy()
    ref = TSDFVolume(0.01, 0.04)
    ref.integrate_batch(d, c, seq.fxfycxcy, seq.extrinsic)
    want = ref.stats()
    hd, hc = torch.from_numpy(d).pin_memory(), torch.from_numpy(c).pin_memory()
    gd = torch.zeros(hd.shape, dtype=hd.dtype, device="cuda")
    gc = torch.zeros(hc.shape, dtype=hc.dtype, device="cuda")
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    v = TSDFVolume(0.01, 0.04)                                       # no set_stream: the volume's own streams
    with torch.cuda.stream(side):
        torch.cuda._sleep(int(4e8))                                  # ~0.2 s of device time ahead of the copies
        gd.copy_(hd, non_blocking=True)
        gc.copy_(hc, non_blocking=True)
        v.integrate_batch(gd, gc, seq.fxfycxcy, seq.extrinsic)       # current stream = side
    assert v.stats() == want and want["weight_sum"] > 0
    ref.close(); v.close()


def test_config3_four_objects_in_one_arena_match_the_oracle():
    """BASELINE configs[2] (multi_reconstruct_rgbd_filter.py): table, chair, cone and cardboard, each its own volume in the
    reference; here one multi-object arena (object id in the block key) fed with the frames of all four interleaved the
    way pipeline.integrate_many does -- one work list / one integration launch per batch.  Every object must equal the
    oracle's separate volume: keys, weights, TSDF bit-exact, colour <= 1/255, mesh vertices identical."""
    from otslam_b200 import pipeline, synth
    from otslam_b200.volume import ArenaView, TSDFVolume
    scenes = ("table", "chair", "cone", "cardboard")
    n_per = (14, 11, 9, 12)                                   # unequal counts: the round-robin runs end at different times
    seqs = [synth.make_sequence(s, 48, subsample=(i, 48 // n), device="cuda") for i, (s, n) in enumerate(zip(scenes, n_per))]
    data = [s.numpy() for s in seqs]
    counts = [len(s) for s in seqs]
    order = pipeline.interleave_plan(counts)
    assert sorted(order) == sorted((o, k) for o, c in enumerate(counts) for k in range(c))
    for o in range(4):
        assert [k for oo, k in order if oo == o] == list(range(counts[o])), "an object's own frame order must be preserved"
    depth = np.stack([data[o][0][k] for o, k in order])
    rgb = np.stack([data[o][1][k] for o, k in order])
    ext = np.stack([seqs[o].extrinsic[k] for o, k in order])
    ids = np.array([o for o, _ in order], np.int32)
    arena = TSDFVolume(0.01, 0.04)
    arena.set_objects(4)
    arena.integrate_batch(depth, rgb, seqs[0].fxfycxcy, ext, object_ids=ids)
    for o in range(4):
        ov = oracle.Volume(0.01, 0.04)
        nupd = 0
        for k in range(counts[o]):
            nupd += ov.integrate(oracle.depth_convert(data[o][0][k]), data[o][1][k], seqs[o].fxfycxcy, seqs[o].extrinsic[k])[1]
        view = ArenaView(arena, o)
        gk, gt, gw, gc = view.export_blocks()
        ok, ot, ow, oc = ov.export_blocks()
        assert gk.shape == ok.shape and (gk == ok).all(), f"object {o}: block keys"
        assert (gw == ow).all() and (gt == ot).all() and np.abs(gc - oc)[ow > 0].max() <= 1.0
        st = view.stats()
        assert st["weight_sum"] == nupd and st["n_blocks"] == len(ok) == view.num_blocks()
        verts, cols, nrm, faces, ek = view.extract_triangle_mesh()
        overts, ocols, ofaces, oek = ov.extract_triangle_mesh()
        A, B = canon_mesh(verts, cols, faces, ek), canon_mesh(overts, ocols, ofaces, oek)
        assert len(A[0]) == len(B[0]) > 500 and (A[3] == B[3]).all() and (A[0] == B[0]).all() and (A[2] == B[2]).all()
        pts, pcols, pek = view.extract_point_cloud()
        op, opc, opek = ov.extract_point_cloud()
        a, b = lexorder(pek), lexorder(opek)
        assert len(pts) == len(op) and (pek[a] == opek[b]).all() and (pts[a] == op[b]).all()
    arena.select_object(-1)
    with pytest.raises(RuntimeError, match="select_object"):
        arena.extract_triangle_mesh()
    with pytest.raises(RuntimeError, match="integrate_batch_objects"):
        arena.integrate_batch(depth[:1], rgb[:1], seqs[0].fxfycxcy, ext[:1])
    with pytest.raises(RuntimeError, match="object id"):
        arena.integrate_batch(depth[:1], rgb[:1], seqs[0].fxfycxcy, ext[:1], object_ids=[7])
    arena.close()
