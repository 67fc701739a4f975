"""Slab-sharded volumes across the GPUs of one box.

Every rank receives every frame and integrates only the blocks of its own x-slabs, so there is NO
collective during integration.  Extraction needs the +1 neighbour voxels of owned blocks; two modes:
  halo=1  the +1 neighbour blocks are integrated redundantly on this rank (no exchange, (T+1)/T work);
  halo=0  owned blocks only (perfectly partitioned work, thin slabs balance well) and ONE exchange
          of the 256-voxel boundary planes before extraction (`exchange_halo`).
The other exchange is the final gather of the extracted geometry to rank 0.  Both use
torch.distributed point-to-point ops (ncclSend/ncclRecv over NVLink with the nccl backend; gloo in
the CPU tests).

Vertices on a slab boundary can be produced by two ranks (the edge is owned by a halo block of one
of them); both compute them from bit-identical replicated voxels, and rank 0 unifies them by their
global lattice-edge key (X, Y, Z, axis).
"""
import numpy as np
import torch
import torch.distributed as dist

DEFAULT_THICKNESS = 8


def slab_spec(rank, world, axis=0, thickness=DEFAULT_THICKNESS, halo=1):
    """(axis, thickness, n_ranks, rank[, halo]) for TSDFVolume(slab=...); None on a single GPU."""
    if world <= 1:
        return None
    return (axis, thickness, world, rank) if halo else (axis, thickness, world, rank, 0)


def exchange_halo(vol, rank, world, device="cpu"):
    """halo=0 mode: send each owned boundary plane to the rank that owns the -axis neighbour block and
    insert the received planes as non-owned blocks.  Must run after integration and before
    extraction.  Returns the number of planes received."""
    if world <= 1:
        return 0
    keys, dest, planes = vol.halo_export()
    rec = planes.shape[1] if planes.ndim == 2 else 0
    counts = torch.tensor([int((dest == r).sum()) for r in range(world)] + [rec], dtype=torch.int64, device=device)
    allc = [torch.zeros(world + 1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(allc, counts)
    allc = [[int(x) for x in c.tolist()] for c in allc]          # allc[src][dst]
    rec = max(c[world] for c in allc)
    ops, send_keep, recv = [], [], []
    for r in range(world):
        if r == rank:
            continue
        n_out = allc[rank][r]
        if n_out:
            sel = dest == r
            tk = torch.from_numpy(np.ascontiguousarray(keys[sel])).to(device)
            tp = torch.from_numpy(np.ascontiguousarray(planes[sel])).to(device)
            send_keep += [tk, tp]
            ops += [dist.P2POp(dist.isend, tk, r), dist.P2POp(dist.isend, tp, r)]
        n_in = allc[r][rank]
        if n_in:
            rk = torch.empty((n_in, 3), dtype=torch.int32, device=device)
            rp = torch.empty((n_in, rec), dtype=torch.uint8, device=device)
            recv.append((rk, rp))
            ops += [dist.P2POp(dist.irecv, rk, r), dist.P2POp(dist.irecv, rp, r)]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    got = 0
    for rk, rp in recv:
        vol.halo_import(rk.cpu().numpy(), rp.cpu().numpy())
        got += len(rk)
    return got


def _gatherv(arrays, rank, world, device):
    """Gather a list of variable-length numpy arrays (same dtype / trailing shape on every rank) to
    rank 0.  Returns, on rank 0, one list per input with the parts of ranks 0..world-1."""
    counts = torch.tensor([len(a) for a in arrays], dtype=torch.int64, device=device)
    all_counts = [torch.zeros(len(arrays), dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    all_counts = [[int(x) for x in c.tolist()] for c in all_counts]
    out = [[None] * world for _ in arrays]
    if rank == 0:
        ops, bufs = [], []
        for r in range(1, world):
            for ai, a in enumerate(arrays):
                t = torch.empty((all_counts[r][ai],) + a.shape[1:], dtype=torch.from_numpy(a[:0]).dtype, device=device)
                bufs.append((ai, r, t))
                if all_counts[r][ai] > 0:
                    ops.append(dist.P2POp(dist.irecv, t, r))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for ai, a in enumerate(arrays):
            out[ai][0] = a
        for ai, r, t in bufs:
            out[ai][r] = t.cpu().numpy()
        return out
    ops = []
    keep = []
    for a in arrays:
        if len(a) > 0:
            t = torch.from_numpy(np.ascontiguousarray(a)).to(device)
            keep.append(t)
            ops.append(dist.P2POp(dist.isend, t, 0))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return None


def merge_mesh_parts(parts):
    """Host logic of the mesh gather: concatenate per-rank (vertices, colors, faces, edge_keys),
    rebase the face indices, and unify boundary vertices that two ranks emitted (same edge key).
    Returns (vertices, colors, faces, edge_keys) with vertices ordered by edge key."""
    verts = np.concatenate([p[0] for p in parts]) if parts else np.zeros((0, 3))
    cols = np.concatenate([p[1] for p in parts]) if parts else np.zeros((0, 3))
    ek = np.concatenate([p[3] for p in parts]) if parts else np.zeros((0, 4), np.int32)
    faces, base = [], 0
    for p in parts:
        faces.append(p[2].astype(np.int64) + base)
        base += len(p[0])
    faces = np.concatenate(faces) if faces else np.zeros((0, 3), np.int64)
    if len(verts) == 0:
        return verts, cols, faces.astype(np.int32), ek
    uniq, first, inverse = np.unique(ek, axis=0, return_index=True, return_inverse=True)
    inverse = inverse.reshape(-1)
    return verts[first], cols[first], inverse[faces].astype(np.int32), uniq.astype(np.int32)


def extract_and_gather_points(vol, rank, world, device="cpu"):
    """volume.extract_point_cloud() on every rank's slab, gathered to rank 0 (None elsewhere)."""
    pts, cols, ek = vol.extract_point_cloud()
    if world <= 1:
        return pts, cols, ek
    g = _gatherv([pts, cols, ek], rank, world, device)
    if rank != 0:
        return None
    return tuple(np.concatenate(x) for x in g)


def extract_and_gather_mesh(vol, rank, world, device="cpu"):
    """volume.extract_triangle_mesh() on every rank's slab, merged on rank 0 (None elsewhere).
    Vertex normals must be recomputed on the merged mesh (they depend on faces of both sides)."""
    r = vol.extract_triangle_mesh(normals=False) if hasattr(vol, "set_batch") else vol.extract_triangle_mesh()
    verts, cols, faces, ek = (r[0], r[1], r[3], r[4]) if len(r) == 5 else r
    if world <= 1:
        return merge_mesh_parts([(verts, cols, faces, ek)])
    g = _gatherv([verts, cols, faces, ek], rank, world, device)
    if rank != 0:
        return None
    return merge_mesh_parts([(g[0][r_], g[1][r_], g[2][r_], g[3][r_]) for r_ in range(world)])
