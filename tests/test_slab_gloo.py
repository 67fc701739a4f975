"""Multi-rank path on CPU: world_size-2 gloo processes run the gather / merge host logic of
otslam_b200.slab on per-rank slab outputs (produced by the oracle, since there is no GPU here) and
rank 0 must reassemble exactly the unsharded result."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT, canon_mesh, lexorder


class OracleSlabVolume:
    """Duck-types the extraction methods of otslam_b200.volume.TSDFVolume on top of the oracle."""

    def __init__(self, vol):
        self.v = vol

    def extract_point_cloud(self):
        return self.v.extract_point_cloud()

    def halo_export(self):
        return self.v.halo_export()

    def halo_import(self, keys, planes):
        return self.v.halo_import(keys, planes)

    def extract_triangle_mesh(self):
        return self.v.extract_triangle_mesh()


def _worker(rank, world, port, out_path, halo=1, thickness=2, axis=0):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import oracle
    from otslam_b200 import slab, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    intr = (160, 120, 565.6009 / 4, 565.6009 / 4, 80.5, 60.5)
    seq = synth.make_sequence("chair_table", 40, intr=intr, subsample=(0, 10))
    d, c = seq.numpy()
    vol = oracle.Volume(0.02, 0.08, slab=slab.slab_spec(rank, world, axis=axis, thickness=thickness, halo=halo))
    for k in range(len(seq)):                                   # every rank receives every frame
        vol.integrate(oracle.depth_convert(d[k]), c[k], seq.fxfycxcy, seq.extrinsic[k])
    if not halo:
        n_owned = vol.num_blocks()
        got = slab.exchange_halo(OracleSlabVolume(vol), rank, world)
        assert got > 0 and n_owned < vol.num_blocks() <= n_owned + got     # diagonal slabs receive up to 3 pieces per block
    pts = slab.extract_and_gather_points(OracleSlabVolume(vol), rank, world)
    mesh = slab.extract_and_gather_mesh(OracleSlabVolume(vol), rank, world)
    if rank == 0:
        np.savez(out_path, pts=pts[0], pcols=pts[1], pek=pts[2], verts=mesh[0], cols=mesh[1], faces=mesh[2], ek=mesh[3])
    else:
        assert pts is None and mesh is None
    dist.barrier()
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("world,halo,thickness,axis", [(2, 1, 2, 0), (2, 0, 1, 0), (3, 0, 2, 0), (3, 0, 1, 3), (2, 0, 2, 3)])
def test_world_size_n_gather_reassembles_full_result(tmp_path, world, halo, thickness, axis):
    """halo=1: replicated +1 blocks, no exchange; halo=0: owned blocks only + boundary-piece exchange; axis 3 =
    diagonal slabs (ownership by kx + ky: x plane, y plane and the x = y = 0 column travel)."""
    from oracle import oracle
    from otslam_b200 import synth
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "rank0.npz")
    mp.spawn(_worker, args=(world, port, out, halo, thickness, axis), nprocs=world, join=True)
    z = np.load(out)
    intr = (160, 120, 565.6009 / 4, 565.6009 / 4, 80.5, 60.5)
    seq = synth.make_sequence("chair_table", 40, intr=intr, subsample=(0, 10))
    d, c = seq.numpy()
    full = oracle.Volume(0.02, 0.08)
    for k in range(len(seq)):
        full.integrate(oracle.depth_convert(d[k]), c[k], seq.fxfycxcy, seq.extrinsic[k])
    fp, fc, fe = full.extract_point_cloud()
    a, b = lexorder(fe), lexorder(z["pek"])
    assert len(fe) == len(z["pek"]) and (fe[a] == z["pek"][b]).all()
    assert (fp[a] == z["pts"][b]).all() and (fc[a] == z["pcols"][b]).all()
    fv, fcol, ff, fek = full.extract_triangle_mesh()
    A = canon_mesh(fv, fcol, ff, fek)
    B = canon_mesh(z["verts"], z["cols"], z["faces"], z["ek"])
    assert len(A[0]) == len(B[0]) and (A[3] == B[3]).all() and (A[0] == B[0]).all() and (A[1] == B[1]).all()
    assert A[2].shape == B[2].shape and (A[2] == B[2]).all()


def test_merge_mesh_parts_unifies_duplicates():
    from otslam_b200 import slab
    v = np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0]])
    ek = np.array([[0, 0, 0, 0], [1, 0, 0, 1], [0, 1, 0, 2]], np.int32)
    p0 = (v, v * 0.5, np.array([[0, 1, 2]], np.int32), ek)
    v1 = np.array([[1.0, 0, 0], [0, 1, 0], [1, 1, 0]])
    ek1 = np.array([[1, 0, 0, 1], [0, 1, 0, 2], [1, 1, 0, 0]], np.int32)
    p1 = (v1, v1 * 0.5, np.array([[0, 2, 1]], np.int32), ek1)
    verts, cols, faces, keys = slab.merge_mesh_parts([p0, p1])
    assert len(verts) == 4 and len(faces) == 2 and faces.max() == 3
    tri = verts[faces]
    assert {tuple(map(tuple, t)) for t in tri.tolist()} == {tuple(map(tuple, v[[0, 1, 2]].tolist())), tuple(map(tuple, v1[[0, 2, 1]].tolist()))}
    assert slab.slab_spec(0, 1) is None and slab.slab_spec(3, 8) == (0, 8, 8, 3)


def test_shard_plan_covers_every_frame_once_in_order():
    """Host ingest for slab volumes (slab.integrate_host_sharded): chunks are split into per-rank blocks whose
    concatenation in rank order is the chunk in frame order."""
    from otslam_b200 import slab
    for n, world, chunk in ((1000, 8, 256), (200, 8, 256), (7, 2, 256), (33, 4, 32), (1, 8, 256), (513, 3, 100)):
        seen = []
        for c0, nk, per in slab.shard_plan(n, world, chunk):
            assert nk <= per * world and per * world >= min(chunk, 1)
            for r in range(world):
                lo, hi = min(c0 + r * per, c0 + nk), min(c0 + (r + 1) * per, c0 + nk)
                seen += list(range(lo, hi))
        assert seen == list(range(n))


def test_rank_shards_hold_exactly_the_frames_a_rank_uploads():
    """slab.rank_shards packs, per rank, the frames integrate_host_sharded(shards=True) reads, in the order it reads them."""
    import torch
    from otslam_b200 import slab
    for n, world, chunk in ((37, 4, 16), (100, 8, 32), (5, 2, 256)):
        depth = torch.arange(n, dtype=torch.int16).view(n, 1, 1).repeat(1, 2, 3)
        rgb = torch.arange(n, dtype=torch.uint8).view(n, 1, 1, 1).repeat(1, 2, 3, 3)
        seen = {}
        for r in range(world):
            hd, hc = slab.rank_shards(depth, rgb, r, world, chunk, pin=False)
            assert hd.shape[1:] == depth.shape[1:] and hc.shape[1:] == rgb.shape[1:] and (hc[:, 0, 0, 0] == hd[:, 0, 0].to(torch.uint8)).all()
            off = 0
            for c0, nk, per in slab.shard_plan(n, world, chunk):
                lo, hi = min(c0 + r * per, c0 + nk), min(c0 + (r + 1) * per, c0 + nk)
                assert hd[off:off + hi - lo, 0, 0].tolist() == list(range(lo, hi))
                for f in range(lo, hi):
                    seen[f] = r
                off += hi - lo
            assert off == len(hd)
        assert sorted(seen) == list(range(n))
