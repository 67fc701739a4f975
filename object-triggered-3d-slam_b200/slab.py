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
            rk = torch.empty((n_in, 4), dtype=torch.int32, device=device)
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


# ---------------------------------------------------------------------------------------------
# Frame ingest for slab-sharded volumes: every rank needs every frame, but the frames only have to
# cross PCIe ONCE per box.  Each rank uploads 1/world of a chunk from its host buffer and the ranks
# all-gather the chunk over NVLink (ncclAllGather); uploading the whole sequence on every rank costs
# world x the host-memory / PCIe traffic and made the 8-GPU host path slower than one GPU.
# ---------------------------------------------------------------------------------------------
def shard_plan(n_frames, world, chunk_frames=128):
    """[(first frame, frames in the chunk, frames per rank)]: chunks of `chunk_frames` (rounded up to a multiple
    of world) frames; inside a chunk rank r owns frames [first + r*per, first + (r+1)*per) -- blocked, so
    the gathered buffer is in frame order and a short last chunk is a prefix of it."""
    per = max(1, -(-chunk_frames // world))
    out, c0 = [], 0
    while c0 < n_frames:
        n = min(per * world, n_frames - c0)
        out.append((c0, n, per))
        c0 += n
    return out


def integrate_host_sharded(vol, depth, rgb, intr, extrinsics, rank, world, device, depth_scale=1000.0, depth_trunc=3.0,
                           chunk_frames=128, stream=None):
    """The frame loop from HOST buffers (pinned CPU torch tensors depth [n,H,W] u16, rgb [n,H,W,3] u8, identical
    on every rank or at least valid for this rank's shards) into a slab-sharded volume: per chunk, H2D of this
    rank's 1/world share -> all_gather over NVLink -> integrate_batch on the resident chunk.  The upload + gather
    of chunk k+1 is queued on a side stream before chunk k integrates.  Frame order is preserved; results are
    identical to vol.integrate_batch(depth, rgb, ...)."""
    n, H, W = int(depth.shape[0]), int(depth.shape[1]), int(depth.shape[2])
    if world <= 1:
        vol.integrate_batch(depth, rgb, intr, extrinsics, depth_scale, depth_trunc)
        return
    plan = shard_plan(n, world, chunk_frames)
    per = plan[0][2]
    side = torch.cuda.Stream(device=device)
    main = stream if stream is not None else torch.cuda.current_stream(device)
    bufs = [(torch.empty((per * world, H, W), dtype=depth.dtype, device=device),
             torch.empty((per * world, H, W, 3), dtype=rgb.dtype, device=device),
             torch.cuda.Event(), torch.cuda.Event()) for _ in range(2)]

    def issue(k):
        c0, nk, _ = plan[k]
        gd, gc, ready, free = bufs[k & 1]
        with torch.cuda.stream(side):
            side.wait_event(free)                                   # the integration that last read this buffer
            lo, hi = min(c0 + rank * per, c0 + nk), min(c0 + (rank + 1) * per, c0 + nk)
            md, mc = gd[rank * per:(rank + 1) * per], gc[rank * per:(rank + 1) * per]
            if hi > lo:
                md[:hi - lo].copy_(depth[lo:hi], non_blocking=True)
                mc[:hi - lo].copy_(rgb[lo:hi], non_blocking=True)
            # in-place all-gather: every rank's share lands at its slot of the chunk buffer
            dist.all_gather_into_tensor(gd.view(torch.uint8), md.view(torch.uint8))      # NCCL in torch has no 16-bit integer type
            dist.all_gather_into_tensor(gc, mc)
            ready.record(side)

    for _, _, _, free in bufs:
        free.record(main)
    issue(0)
    for k, (c0, nk, _) in enumerate(plan):
        gd, gc, ready, free = bufs[k & 1]
        if k + 1 < len(plan):
            issue(k + 1)
        main.wait_event(ready)
        vol.integrate_batch(gd[:nk], gc[:nk], intr, extrinsics[c0:c0 + nk], depth_scale, depth_trunc)   # returns when done
        free.record(main)
