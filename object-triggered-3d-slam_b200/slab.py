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


def _t(a, device):
    """numpy array -> tensor on `device` (the CPU / gloo tests and host-side backends); tensors pass through."""
    if isinstance(a, torch.Tensor):
        return a
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


_PINNED = {}


def to_host(tensors, copy=True):
    """Device tensors -> numpy arrays through cached page-locked staging buffers (pageable .cpu() copies run at ~2 GB/s,
    pinned ones at PCIe speed; a 1 GB gathered cloud is 0.4 s against 20 ms).  copy=True returns arrays the caller owns;
    copy=False returns views of the staging buffers, valid until the next to_host call."""
    out = []
    for t in tensors:
        if not t.is_cuda:
            out.append(t.numpy())
            continue
        nbytes = t.numel() * t.element_size()
        key = (t.device.index, len(out))
        buf = _PINNED.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
            _PINNED[key] = buf
        view = buf[:nbytes].view(t.dtype).view(t.shape)
        view.copy_(t, non_blocking=True)
        out.append(view)
    if any(t.is_cuda for t in tensors):
        torch.cuda.current_stream().synchronize()
    out = tuple(v.numpy() if isinstance(v, torch.Tensor) else v for v in out)
    return tuple(np.array(v) for v in out) if copy else out


def _halo_pack(vol, world, device):
    """(keys [n,4] i32, pieces [n,rec] u8, pieces per destination rank) as tensors on `device`, grouped by
    destination.  A GPU TSDFVolume packs them inside HBM (otslam_volume_halo_pack: no host copy of voxel data);
    host-side backends (the oracle in the gloo tests) return numpy arrays sorted by key, grouped here."""
    if hasattr(vol, "halo_pack_tensors"):
        keys, planes, counts = vol.halo_pack_tensors()
        return keys, planes, [int(c) for c in counts[:world]]
    keys, dest, planes = vol.halo_export()
    order = np.argsort(dest, kind="stable")
    counts = [int((dest == r).sum()) for r in range(world)]
    planes = planes.reshape(len(keys), -1) if len(keys) else planes.reshape(0, planes.shape[1] if planes.ndim == 2 else 0)
    return _t(keys[order], device), _t(planes[order], device), counts


def exchange_halo(vol, rank, world, device="cpu"):
    """halo=0 mode: send each owned boundary piece to the rank that owns the -axis neighbour block and insert the
    received pieces as non-owned blocks.  Must run after integration and before extraction.  On GPUs the pieces
    never leave HBM: pack kernel -> ncclSend/ncclRecv on device buffers -> import kernel.  Returns the number of
    pieces received."""
    if world <= 1:
        return 0
    keys, planes, counts = _halo_pack(vol, world, device)
    rec = int(planes.shape[1]) if planes.ndim == 2 else 0
    mine = torch.tensor(counts + [rec], dtype=torch.int64, device=device)
    allc = torch.empty(world * (world + 1), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allc, mine)
    allc = allc.view(world, world + 1).tolist()                   # allc[src][dst]
    rec = max(int(c[world]) for c in allc)
    n_in = [int(allc[r][rank]) if r != rank else 0 for r in range(world)]
    total_in = sum(n_in)
    rk = torch.empty((total_in, 4), dtype=torch.int32, device=device)
    rp = torch.empty((total_in, rec), dtype=torch.uint8, device=device)
    ops, off_out, off_in = [], 0, 0
    for r in range(world):
        n_out = counts[r]
        if n_out and r != rank:
            ops += [dist.P2POp(dist.isend, keys[off_out:off_out + n_out], r), dist.P2POp(dist.isend, planes[off_out:off_out + n_out], r)]
        off_out += n_out
        if n_in[r]:
            ops += [dist.P2POp(dist.irecv, rk[off_in:off_in + n_in[r]], r), dist.P2POp(dist.irecv, rp[off_in:off_in + n_in[r]], r)]
            off_in += n_in[r]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    if total_in:
        if hasattr(vol, "halo_pack_tensors"):
            vol.halo_import(rk, rp)                               # device pointers; ordered after the recv on torch's stream
        else:
            vol.halo_import(rk.cpu().numpy(), rp.cpu().numpy())
    return total_in


def _gatherv(tensors, rank, world, device):
    """Gather variable-length tensors (same dtype / trailing shape on every rank) to rank 0 over send/recv.  Rank 0
    receives every part straight into its slice of one destination tensor per input (no staging, no host copy) and
    returns (list of gathered tensors, counts[r][i]); other ranks return (None, counts)."""
    counts = torch.tensor([int(t.shape[0]) for t in tensors], dtype=torch.int64, device=device)
    allc = torch.empty(world * len(tensors), dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allc, counts)
    allc = allc.view(world, len(tensors)).tolist()
    if rank == 0:
        out, ops = [], []
        for i, t in enumerate(tensors):
            total = sum(int(allc[r][i]) for r in range(world))
            dst = torch.empty((total,) + tuple(t.shape[1:]), dtype=t.dtype, device=device)
            n0 = int(allc[0][i])
            dst[:n0].copy_(t)
            off = n0
            for r in range(1, world):
                n = int(allc[r][i])
                if n:
                    ops.append(dist.P2POp(dist.irecv, dst[off:off + n], r))
                off += n
            out.append(dst)
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        return out, allc
    ops = [dist.P2POp(dist.isend, t.contiguous(), 0) for t in tensors if t.shape[0] > 0]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return None, allc


def merge_mesh_tensors(verts, cols, faces, ek):
    """Unify vertices that two ranks emitted for the same lattice edge (same edge key): returns (vertices, colors,
    faces, edge_keys) with vertices ordered by edge key.  Tensor code, runs where the tensors live (HBM on rank 0)."""
    if verts.shape[0] == 0:
        return verts, cols, faces.to(torch.int32), ek
    uniq, inverse = torch.unique(ek, dim=0, return_inverse=True)
    inverse = inverse.reshape(-1)
    first = torch.full((uniq.shape[0],), verts.shape[0], dtype=torch.int64, device=verts.device)
    first.scatter_reduce_(0, inverse, torch.arange(verts.shape[0], dtype=torch.int64, device=verts.device), reduce="amin")
    return verts[first], cols[first], inverse[faces.to(torch.int64)].to(torch.int32), uniq.to(torch.int32)


def merge_mesh_parts(parts):
    """Host form of the mesh merge (numpy in / out): concatenate per-rank (vertices, colors, faces, edge_keys),
    rebase the face indices, and unify boundary vertices that two ranks emitted (same edge key)."""
    if not parts:
        return np.zeros((0, 3)), np.zeros((0, 3)), np.zeros((0, 3), np.int32), np.zeros((0, 4), np.int32)
    verts = torch.from_numpy(np.concatenate([np.asarray(p[0], np.float64).reshape(-1, 3) for p in parts]))
    cols = torch.from_numpy(np.concatenate([np.asarray(p[1], np.float64).reshape(-1, 3) for p in parts]))
    ek = torch.from_numpy(np.concatenate([np.asarray(p[3], np.int32).reshape(-1, 4) for p in parts]))
    faces, base = [], 0
    for p in parts:
        faces.append(np.asarray(p[2]).reshape(-1, 3).astype(np.int64) + base)
        base += len(p[0])
    faces = torch.from_numpy(np.concatenate(faces))
    return tuple(x.numpy() for x in merge_mesh_tensors(verts, cols, faces, ek))


def extract_and_gather_points(vol, rank, world, device="cpu", as_numpy=True):
    """volume.extract_point_cloud() on every rank's slab, gathered to rank 0 (None elsewhere).  With a GPU volume the
    extracted points go HBM -> NVLink -> rank 0's HBM; as_numpy=False returns the device tensors there."""
    if hasattr(vol, "extract_point_cloud_tensors"):
        parts = list(vol.extract_point_cloud_tensors())
    else:
        parts = [_t(a, device) for a in vol.extract_point_cloud()]
    if world > 1:
        parts, _ = _gatherv(parts, rank, world, device)
        if rank != 0:
            return None
    return to_host(parts) if as_numpy else tuple(parts)


def extract_and_gather_mesh(vol, rank, world, device="cpu", as_numpy=True):
    """volume.extract_triangle_mesh() on every rank's slab, merged on rank 0 (None elsewhere).
    Vertex normals must be recomputed on the merged mesh (they depend on faces of both sides)."""
    if hasattr(vol, "extract_mesh_tensors"):
        verts, cols, faces, ek = vol.extract_mesh_tensors()
    else:
        r = vol.extract_triangle_mesh()
        verts, cols, faces, ek = (_t(a, device) for a in ((r[0], r[1], r[3], r[4]) if len(r) == 5 else r))
    faces = faces.to(torch.int64)
    if world > 1:
        g, allc = _gatherv([verts, cols, faces, ek], rank, world, device)
        if rank != 0:
            return None
        verts, cols, faces, ek = g
        off_f, base = 0, 0
        for r in range(world):                                    # rebase every rank's face indices by its vertex offset
            nf, nv = int(allc[r][2]), int(allc[r][0])
            faces[off_f:off_f + nf] += base
            off_f += nf
            base += nv
    out = merge_mesh_tensors(verts, cols, faces, ek)
    return to_host(out) if as_numpy else out


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


def rank_shards(depth, rgb, rank, world, chunk_frames=128, pin=True):
    """This rank's share of every chunk of shard_plan, packed back to back into (pinned) host tensors: what a rank has
    to hold in host memory to feed integrate_host_sharded(..., shards=True).  depth / rgb: full-sequence tensors on any
    device (or anything sliceable into tensors)."""
    n = int(depth.shape[0])
    spans = []
    for c0, nk, per in shard_plan(n, world, chunk_frames):
        lo, hi = min(c0 + rank * per, c0 + nk), min(c0 + (rank + 1) * per, c0 + nk)
        spans.append((lo, hi))
    total = sum(hi - lo for lo, hi in spans)
    hd = torch.empty((total,) + tuple(depth.shape[1:]), dtype=depth.dtype, pin_memory=pin)
    hc = torch.empty((total,) + tuple(rgb.shape[1:]), dtype=rgb.dtype, pin_memory=pin)
    off = 0
    for lo, hi in spans:
        if hi > lo:
            hd[off:off + hi - lo].copy_(depth[lo:hi]); hc[off:off + hi - lo].copy_(rgb[lo:hi])
            off += hi - lo
    return hd, hc


def integrate_host_sharded(vol, depth, rgb, intr, extrinsics, rank, world, device, depth_scale=1000.0, depth_trunc=3.0,
                           chunk_frames=128, stream=None, shards=False):
    """The frame loop from HOST buffers (pinned CPU torch tensors depth [n,H,W] u16, rgb [n,H,W,3] u8, identical
    on every rank or at least valid for this rank's shards; with shards=True they hold ONLY this rank's shares, packed
    by rank_shards with the same chunk_frames, and n = len(extrinsics)) into a slab-sharded volume: per chunk, H2D of this
    rank's 1/world share -> all_gather over NVLink -> integrate_batch on the resident chunk.  The upload + gather
    of chunk k+1 is queued on a side stream before chunk k integrates.  Frame order is preserved; results are
    identical to vol.integrate_batch(depth, rgb, ...)."""
    n, H, W = int(len(extrinsics)) if shards else int(depth.shape[0]), int(depth.shape[1]), int(depth.shape[2])
    if world <= 1:
        vol.integrate_batch(depth, rgb, intr, extrinsics, depth_scale, depth_trunc)
        return
    plan = shard_plan(n, world, chunk_frames)
    per = plan[0][2]
    copy_s = torch.cuda.Stream(device=device)       # H2D of this rank's share of chunk k+1 ...
    side = torch.cuda.Stream(device=device)         # ... runs under the all-gather of chunk k (separate streams: the copy
    #                                                 must not queue behind the previous chunk's collective)
    main = stream if stream is not None else torch.cuda.current_stream(device)
    bufs = [(torch.empty((per * world, H, W), dtype=depth.dtype, device=device),
             torch.empty((per * world, H, W, 3), dtype=rgb.dtype, device=device),
             torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()) for _ in range(2)]
    local_off = [0]
    import os as _os, time as _time
    dbg = [] if _os.environ.get("OTSLAM_INGEST_DEBUG") else None
    t_origin = _time.perf_counter()

    def issue(k):
        c0, nk, _ = plan[k]
        gd, gc, ready, free, copied = bufs[k & 1]
        lo, hi = min(c0 + rank * per, c0 + nk), min(c0 + (rank + 1) * per, c0 + nk)
        md, mc = gd[rank * per:(rank + 1) * per], gc[rank * per:(rank + 1) * per]
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(free)                                 # the integration that last read this buffer
            if hi > lo:
                src = local_off[0] if shards else lo
                md[:hi - lo].copy_(depth[src:src + hi - lo], non_blocking=True)
                mc[:hi - lo].copy_(rgb[src:src + hi - lo], non_blocking=True)
                local_off[0] += hi - lo
            copied.record(copy_s)
        with torch.cuda.stream(side):
            side.wait_event(copied)
            # in-place all-gather: every rank's share lands at its slot of the chunk buffer
            dist.all_gather_into_tensor(gd.view(torch.uint8), md.view(torch.uint8))      # NCCL in torch has no 16-bit integer type
            dist.all_gather_into_tensor(gc, mc)
            ready.record(side)

    for b in bufs:
        b[3].record(main)
    issue(0)
    for k, (c0, nk, _) in enumerate(plan):
        gd, gc, ready, free, _ = bufs[k & 1]
        if k + 1 < len(plan):
            issue(k + 1)
        main.wait_event(ready)
        if dbg is not None:
            t_a = _time.perf_counter()
            ready.synchronize()
            t_b = _time.perf_counter()
        with torch.cuda.stream(main):       # integrate_batch orders the volume's own streams after torch's CURRENT stream
            vol.integrate_batch(gd[:nk], gc[:nk], intr, extrinsics[c0:c0 + nk], depth_scale, depth_trunc)   # returns when done
        free.record(main)
        if dbg is not None:
            dbg.append((k, 1e3 * (t_a - t_origin), 1e3 * (t_b - t_a), 1e3 * (_time.perf_counter() - t_b)))
    if dbg is not None and rank == 0:
        print("ingest chunks (k, issued at ms, waited for gather ms, integrate ms):", [tuple(round(x, 2) for x in d) for d in dbg], flush=True)
