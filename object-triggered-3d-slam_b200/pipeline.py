"""Shared frame loop of the drop-in scripts.

The reference integrates frame by frame (/root/reference/3d_model/reconstruct_rgbd.py:86-109:
read colour, read depth, loadtxt pose, pose @ T_fix, inv, create_from_color_and_depth, integrate).
Here the files are handled in chunks: their compressed bytes are uploaded and decoded on the GPU (decoder.py; or decoded by
OpenCV threads on the host, OTSLAM_GPU_DECODE=0) and each chunk is integrated in ONE call (frame order preserved, up to
32 frames fused per block residency);
per-frame failures keep the reference's semantics: abort (reconstruct_rgbd.py, no try) or
print-and-skip (reconstruct_rgbd_filter.py:108-109, multi_reconstruct_rgbd_filter.py:102-103).
"""
import os
import sys

import numpy as np

from .o3d_compat import io as o3d_io

CHUNK_FRAMES = 256
DECODE_AHEAD = 4              # GPU decode: chunks being read / uploaded / decoded while one integrates
_FMT = "[ScalableTSDFVolume::Integrate] Unsupported image format."


def load_frame(color_path, depth_path, pose_path, intrinsics, T_fix):
    """Decode one capture triple; returns (rgb u8 HxWx3, depth u16 HxW, extrinsic 4x4) or raises the
    exception the reference's per-frame body would have raised."""
    color = np.asarray(o3d_io.read_image(color_path))
    depth = np.asarray(o3d_io.read_image(depth_path))
    pose_ros = np.loadtxt(pose_path)
    pose_optical = pose_ros @ T_fix
    extrinsic = np.linalg.inv(pose_optical)
    if color.size == 0 or depth.size == 0 or color.shape[:2] != depth.shape[:2]:
        raise RuntimeError("[CreateFromColorAndDepth] Unsupported image format.")
    if depth.ndim != 2 or depth.dtype != np.uint16 or color.ndim != 3 or color.shape[2] != 3 or color.dtype != np.uint8 \
            or depth.shape != (intrinsics.height, intrinsics.width):
        raise RuntimeError(_FMT)
    return color, depth, extrinsic


def _decode_workers():
    n = int(os.environ.get("OTSLAM_DECODE_THREADS", "0"))
    if n > 0:
        return min(n, 64)
    try:
        cores = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        cores = os.cpu_count() or 1
    return max(1, min(32 if sidecar_enabled() else 16, cores))      # side-car reads are I/O + memcpy (GIL released): more threads pay


def read_pose(path):
    """np.loadtxt(pose_path) for the 4x4 text the capture nodes write (scanner_node.cpp:294-298), without loadtxt's
    ~0.3 ms of interpreter time per file (it runs under the GIL and capped the threaded decode).  Same correctly
    rounded decimal -> binary64 conversion; anything but 16 plain numbers goes through np.loadtxt itself."""
    try:
        with open(path) as f:
            txt = f.read()
        vals = txt.split()
        if len(vals) == 16 and "#" not in txt and "," not in txt:
            return np.array([float(v) for v in vals], np.float64).reshape(4, 4)
    except (OSError, ValueError):
        pass
    return np.loadtxt(path)


def read_poses(paths):
    """read_pose for many files at once through the library's host threads (otslam_read_pose_files: no interpreter time per
    file).  Returns (poses [n,4,4] f64, status [n] i32); status != 0 = not the plain 16-number text: use read_pose."""
    import ctypes as C
    from . import _lib
    n = len(paths)
    poses = np.zeros((n, 4, 4), np.float64)
    status = np.ones(n, np.int32)
    if n:
        arr = (C.c_char_p * n)(*[os.fsencode(p) for p in paths])
        _lib.check(_lib.lib.otslam_read_pose_files(n, arr, _lib.ptr(poses), _lib.ptr(status)))
    return poses, status


# ---------------------------------------------------------------------------------------------------------------------
# Raw side-car of a capture tree ("streaming ingest"): JPEG / PNG decoding is what bounds the drop-in
# scripts end to end (~1.8 k frames/s on 16 host threads against ~50 k frames/s of GPU integration).  With
# OTSLAM_SIDECAR=1 the first pass over a tree stores every decoded frame pair next to it
#     <base>/.otslam_raw/<depth file stem>.raw = header | depth u16 [H][W] | rgb u8 [H][W][3]
# (exactly the arrays the decoders produced, so the volumes are identical), keyed by the size and mtime of the two source
# files; later passes read the raw bytes straight into the staging buffers.  Poses stay in their text files.
# ---------------------------------------------------------------------------------------------------------------------
_SIDECAR_MAGIC = b"OTSLAMRAW2\n\0\0\0\0\0"      # 16 bytes
_SIDECAR_HDR = 256                                  # magic | 8 x int64 tag | 16 x float64 pose_ros | padding


def sidecar_enabled():
    return os.environ.get("OTSLAM_SIDECAR", "0") not in ("", "0")


def _sidecar_path(depth_path):
    d = os.path.dirname(os.path.abspath(depth_path))
    return os.path.join(os.path.dirname(d), ".otslam_raw", os.path.splitext(os.path.basename(depth_path))[0] + ".raw")


def _sidecar_tag(cp, dp, pp, H, W):
    sc, sd, sp = os.stat(cp), os.stat(dp), os.stat(pp)
    return np.array([H, W, sc.st_size, sc.st_mtime_ns, sd.st_size, sd.st_mtime_ns, sp.st_size, sp.st_mtime_ns], np.int64)


def _sidecar_load(cp, dp, pp, depth_out, color_out):
    """pose_ros (4x4) when a valid side-car filled the two staging slots, else None."""
    try:
        with open(_sidecar_path(dp), "rb", buffering=0) as f:
            hdr = f.read(_SIDECAR_HDR)
            if len(hdr) != _SIDECAR_HDR or hdr[:16] != _SIDECAR_MAGIC:
                return None
            H, W = depth_out.shape
            if not (np.frombuffer(hdr, np.int64, 8, 16) == _sidecar_tag(cp, dp, pp, H, W)).all():
                return None                                      # a source file changed (or another image size): decode again
            if f.readinto(memoryview(depth_out).cast("B")) != H * W * 2 or f.readinto(memoryview(color_out).cast("B")) != H * W * 3:
                return None
            return np.frombuffer(hdr, np.float64, 16, 80).reshape(4, 4)
    except OSError:
        return None


def _sidecar_store(cp, dp, pp, depth, color, pose_ros):
    path = _sidecar_path(dp)
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        H, W = depth.shape
        hdr = _SIDECAR_MAGIC + _sidecar_tag(cp, dp, pp, H, W).tobytes() + np.ascontiguousarray(pose_ros, np.float64).tobytes()
        tmp = f"{path}.{os.getpid()}.tmp"
        with open(tmp, "wb") as f:
            f.write(hdr + b"\0" * (_SIDECAR_HDR - len(hdr)))
            f.write(memoryview(np.ascontiguousarray(depth)).cast("B"))
            f.write(memoryview(np.ascontiguousarray(color)).cast("B"))
        os.replace(tmp, path)
    except (OSError, ValueError):
        pass                                                     # read-only dataset: keep decoding


def _decode_into(triple, intrinsics, T_fix, depth_out, color_out):
    """load_frame, decoding straight into one slot of the chunk's staging buffers (no per-frame arrays, no np.stack)."""
    import cv2
    cp, dp, pp, _ = triple
    try:
        if sidecar_enabled() and depth_out.shape == (intrinsics.height, intrinsics.width):
            pose_ros = _sidecar_load(cp, dp, pp, depth_out, color_out)
            if pose_ros is not None:
                return np.linalg.inv(pose_ros @ T_fix), None
        c = cv2.imread(cp, cv2.IMREAD_UNCHANGED) if os.path.exists(cp) else None
        d = cv2.imread(dp, cv2.IMREAD_UNCHANGED) if os.path.exists(dp) else None
        for path, a in ((cp, c), (dp, d)):
            if a is None:
                print(f"[Open3D WARNING] Read image failed: unable to open file: {path}")
        pose_ros = read_pose(pp)
        extrinsic = np.linalg.inv(pose_ros @ T_fix)
        if c is None or d is None or c.size == 0 or d.size == 0 or c.shape[:2] != d.shape[:2]:
            raise RuntimeError("[CreateFromColorAndDepth] Unsupported image format.")
        if d.ndim != 2 or d.dtype != np.uint16 or c.ndim != 3 or c.shape[2] not in (3, 4) or c.dtype != np.uint8 \
                or d.shape != (intrinsics.height, intrinsics.width):
            raise RuntimeError(_FMT)
        cv2.cvtColor(c, cv2.COLOR_BGRA2RGB if c.shape[2] == 4 else cv2.COLOR_BGR2RGB, dst=color_out)
        np.copyto(depth_out, d)
        if sidecar_enabled():
            _sidecar_store(cp, dp, pp, depth_out, color_out, pose_ros)
        return extrinsic, None
    except Exception as err:  # noqa: BLE001
        return None, err


class _Staging:
    """Two sets of chunk buffers: chunk k+1 is decoded into one while chunk k is copied to the GPU from the other.
    Kept in a pool for the life of the process (first-touch page faults are not free either).  Ordinary pageable
    memory on purpose: page-locking 2 x 390 MB costs ~0.5 s up front (measured: it halved the throughput of a
    128-frame run), while the copy of a decoded chunk is three orders of magnitude faster than decoding it."""

    _free = {}          # (height, width) -> idle staging objects
    _lock = None

    def __init__(self, frames, height, width, n_sets):
        self.frames, self.key = frames, (height, width)
        self.sets = [(np.empty((frames, height, width), np.uint16), np.empty((frames, height, width, 3), np.uint8))
                     for _ in range(n_sets)]

    @classmethod
    def acquire(cls, frames, height, width, n_sets):
        import threading
        if cls._lock is None:
            cls._lock = threading.Lock()
        with cls._lock:
            idle = cls._free.setdefault((height, width), [])
            for i, st in enumerate(idle):
                if st.frames >= frames and len(st.sets) >= n_sets:
                    return idle.pop(i)
            idle.clear()                                      # too small: drop them, allocate the larger one
        return cls(frames, height, width, n_sets)

    def release(self):
        with _Staging._lock:
            _Staging._free.setdefault(self.key, []).append(self)


last_decode_profile = []      # per chunk of the last GPU-decoded integrate_files call: decoder.FrameDecoder.profile() + counts


def gpu_decode_enabled():
    """OTSLAM_GPU_DECODE=0 keeps JPEG / PNG decoding on the host (OpenCV threads); default: the GPU decoders."""
    return os.environ.get("OTSLAM_GPU_DECODE", "1") not in ("", "0")


_decoder_pool = {}            # (height, width, frames, device) -> idle FrameDecoders (device + pinned buffers are kept)
_decoder_lock = None


def _acquire_decoders(count, height, width, frames, device):
    import threading
    from .decoder import FrameDecoder
    global _decoder_lock
    if _decoder_lock is None:
        _decoder_lock = threading.Lock()
    key = (height, width, frames, device)
    with _decoder_lock:
        idle = _decoder_pool.setdefault(key, [])
        got = [idle.pop() for _ in range(min(count, len(idle)))]
    return got + [FrameDecoder(height, width, frames, device) for _ in range(count - len(got))], key


def release_decoders():
    """Free the pooled GPU decoders (their device and page-locked staging buffers)."""
    for idle in _decoder_pool.values():
        while idle:
            idle.pop().close()


import atexit                                                     # noqa: E402
atexit.register(release_decoders)                                 # before the CUDA runtime is torn down


class GpuDecodeSelfCheckFailed(RuntimeError):
    pass


gpu_decode_selfcheck = {"done": False, "failed": None}    # once per process: see _selfcheck_first_frame


def _selfcheck_first_frame(dec, chunk, cstat, dstat):
    """Once per process, the first frame pair the GPU decoders produced is downloaded and compared with the stock decoders'
    output for the same two files (7 ms).  The decoders are bit-exact by construction and by test; this is the production
    guard for the case nobody tested (a driver / hardware oddity): a mismatch never reaches a volume -- the loop raises
    GpuDecodeSelfCheckFailed before anything is integrated and integrate_files re-runs on the host decoders, loudly."""
    import cv2
    for k in range(len(chunk)):
        if cstat[k] == 0 and dstat[k] == 0:
            d, c = dec.fetch(k, 1)
            cc, dd = cv2.imread(chunk[k][0], cv2.IMREAD_UNCHANGED), cv2.imread(chunk[k][1], cv2.IMREAD_UNCHANGED)
            if cc is None or dd is None or cc.ndim != 3:
                continue                                          # the stock decoder disagrees about the KIND of file: not this check's business
            ref = cv2.cvtColor(cc, cv2.COLOR_BGRA2RGB if cc.shape[2] == 4 else cv2.COLOR_BGR2RGB)
            if dd.shape != d[0].shape or ref.shape != c[0].shape or not (dd == d[0]).all() or not (ref == c[0]).all():
                raise GpuDecodeSelfCheckFailed(f"GPU-decoded frame differs from the stock decoders: {chunk[k][0]}, {chunk[k][1]}")
            return


def _integrate_files_gpu(volume, triples, intrinsics, T_fix, depth_scale, depth_trunc, skip_errors, progress, on_error):
    """integrate_files with the decoders on the GPU (decoder.py): a chunk's files are read by the library's host threads,
    the compressed bytes are uploaded and decoded in HBM, and the decoded slots go straight into the volume.  DECODE_AHEAD
    chunks are in preparation (read + upload + decode: one worker thread and one decoder with its own streams each) while
    another integrates; the pose files are parsed by the library's host threads (read_poses) meanwhile.  Frames the GPU decoders pass on
    (status != 0: progressive JPEG, another size, a damaged or missing file ...) go through the stock decoders exactly as in
    the host path, so warnings, exceptions and skip semantics are the host path's."""
    from concurrent.futures import ThreadPoolExecutor
    H, W = intrinsics.height, intrinsics.width
    n = len(triples)
    chunks = [triples[c0:c0 + CHUNK_FRAMES] for c0 in range(0, n, CHUNK_FRAMES)]
    vol = volume._vol
    del last_decode_profile[:]
    # Chunks in preparation beside the one integrating.  One file = one sequential bit stream = one warp at ~0.16
    # instructions per cycle (ncu, profiles/decode_r02a.md): the GPU is nowhere near busy with one chunk's 2 x 256 warps,
    # so throughput comes from the number of files in flight.
    ahead = min(DECODE_AHEAD, len(chunks))
    decs, pool_key = _acquire_decoders(min(ahead + 1, len(chunks)), H, W, CHUNK_FRAMES, vol.device)
    done = 0
    try:
        with ThreadPoolExecutor(max_workers=_decode_workers()) as pool, ThreadPoolExecutor(max_workers=ahead) as stage:
            def extrinsics(chunk):
                """extrinsic of every frame of a chunk with the host loop's arithmetic (inv(pose @ T_fix), frame by frame);
                the text files are parsed by the library's threads -- a pool of interpreter threads doing it contended for
                the GIL (160 us per file instead of 27) and was what bounded the loop"""
                poses, pstat = read_poses([t[2] for t in chunk])
                out = []
                for k in range(len(chunk)):
                    try:
                        out.append((np.linalg.inv((poses[k] if pstat[k] == 0 else read_pose(chunk[k][2])) @ T_fix), None))
                    except Exception as err:  # noqa: BLE001
                        out.append((None, err))
                return out

            def stock(k, chunk, dec):
                """the reference's per-frame body on the host for a frame the GPU decoders passed on"""
                d, c = np.empty((H, W), np.uint16), np.empty((H, W, 3), np.uint8)
                ext, err = _decode_into(chunk[k], intrinsics, T_fix, d, c)
                if err is None:
                    dec.put(k, d, c)
                return ext, err

            def prepare(ci):
                dec, chunk = decs[ci % len(decs)], chunks[ci]
                exts = pool.submit(extrinsics, chunk)                                     # while the call below decodes
                cstat, dstat = dec.decode_files([t[0] for t in chunk], [t[1] for t in chunk])
                if ci == 0 and not gpu_decode_selfcheck["done"]:         # chunk 0: nothing has been integrated yet
                    _selfcheck_first_frame(dec, chunk, cstat, dstat)
                    gpu_decode_selfcheck["done"] = True
                prof = dec.profile()
                prof["frames"], prof["passed_on"] = len(chunk), int(np.count_nonzero(cstat | dstat))
                last_decode_profile.append(prof)
                redo = {int(k): pool.submit(stock, int(k), chunk, dec) for k in np.nonzero(cstat | dstat)[0]}
                res = exts.result()
                for k, f in redo.items():
                    res[k] = f.result()
                return res

            futs = {ci: stage.submit(prepare, ci) for ci in range(ahead)}
            for ci, chunk in enumerate(chunks):
                res = futs.pop(ci).result()
                if ci + ahead < len(chunks):                       # its decoder's previous chunk (ci - 1) has been integrated
                    futs[ci + ahead] = stage.submit(prepare, ci + ahead)
                slots, exts = [], []
                for k, ((ext, err), triple) in enumerate(zip(res, chunk)):
                    if err is not None:
                        if not skip_errors:
                            raise err
                        if on_error:
                            on_error(triple[3], err)
                        continue
                    slots.append(k); exts.append(ext)
                    if progress:
                        progress(triple[3], ci * CHUNK_FRAMES + k + 1, n)
                if exts:
                    decs[ci % len(decs)].integrate(vol, slots, intrinsics.fxfycxcy(), np.stack(exts), depth_scale, depth_trunc)
                    done += len(slots)
    finally:
        with _decoder_lock:
            _decoder_pool.setdefault(pool_key, []).extend(decs)
    return done


def integrate_files(volume, triples, intrinsics, T_fix, depth_scale=1000.0, depth_trunc=3.0, skip_errors=False,
                    progress=None, on_error=None):
    """Integrate capture triples [(color, depth, pose, label)] in order. Returns frames integrated.

    JPEG / PNG decoding is what bounds this loop once integration runs on the GPU (~6 ms per frame pair on one
    core against ~0.02 ms of GPU work), so the files of a chunk are decoded by a thread pool (OpenCV releases the
    GIL) straight into pooled chunk buffers, and chunk k+1 is decoded while the GPU integrates chunk k (the
    C-ABI call releases the GIL as well).  Results are consumed in file order, so the per-frame semantics -- abort
    on the first bad frame, or print-and-skip -- and the frame order seen by the volume are exactly the
    sequential loop's.

    Default (OTSLAM_GPU_DECODE unset or 1, a plain volume, no side-car): the decoding itself runs on the GPU as well --
    _integrate_files_gpu above; the host-decode loop below remains for OTSLAM_GPU_DECODE=0, for the raw side-car and
    for volumes that are objects of a multi-object arena."""
    from concurrent.futures import ThreadPoolExecutor
    from .volume import TSDFVolume
    done = 0
    n = len(triples)
    chunks = [triples[c0:c0 + CHUNK_FRAMES] for c0 in range(0, n, CHUNK_FRAMES)]
    if not chunks:
        return 0
    if gpu_decode_enabled() and not sidecar_enabled() and isinstance(getattr(volume, "_vol", None), TSDFVolume) \
            and gpu_decode_selfcheck["failed"] is None:
        try:
            return _integrate_files_gpu(volume, triples, intrinsics, T_fix, depth_scale, depth_trunc, skip_errors, progress, on_error)
        except GpuDecodeSelfCheckFailed as err:                       # raised before any frame reached the volume
            gpu_decode_selfcheck["failed"] = str(err)
            sys.stderr.write(f"[otslam_b200] {err}; GPU decoding is OFF for this process (host decoders take over)\n")
    staging = _Staging.acquire(min(CHUNK_FRAMES, n), intrinsics.height, intrinsics.width, 2 if len(chunks) > 1 else 1)
    try:
        with ThreadPoolExecutor(max_workers=_decode_workers()) as pool:
            def submit(ci):
                dbuf, cbuf = staging.sets[ci & 1 if len(staging.sets) > 1 else 0]
                return [pool.submit(_decode_into, t, intrinsics, T_fix, dbuf[k], cbuf[k]) for k, t in enumerate(chunks[ci])]

            pending = submit(0)
            for ci, chunk in enumerate(chunks):
                futures = pending
                dbuf, cbuf = staging.sets[ci & 1 if len(staging.sets) > 1 else 0]
                slots, exts = [], []
                for k, (fut, triple) in enumerate(zip(futures, chunk)):
                    ext, err = fut.result()
                    label = triple[3]
                    if err is not None:
                        if not skip_errors:
                            for f in futures[k + 1:]:
                                f.cancel()
                            raise err
                        if on_error:
                            on_error(label, err)
                        continue
                    slots.append(k); exts.append(ext)
                    if progress:
                        progress(label, ci * CHUNK_FRAMES + k + 1, n)
                pending = submit(ci + 1) if ci + 1 < len(chunks) else []     # decoded while this chunk integrates
                if exts:
                    m = len(slots)
                    if slots == list(range(m)):
                        deps, cols = dbuf[:m], cbuf[:m]
                    else:                                                     # skipped frames left holes: close them up
                        deps, cols = np.ascontiguousarray(dbuf[slots]), np.ascontiguousarray(cbuf[slots])
                    volume.integrate_sequence(deps, cols, intrinsics, np.stack(exts), depth_scale, depth_trunc)
                    done += m
    finally:
        staging.release()
    return done


def _decode_job(pool, triples, intrinsics, T_fix, skip_errors, on_error, progress):
    """Decode every frame of one job (thread pool) and apply the per-frame error semantics in file order; returns
    (depth [m,H,W] u16, rgb [m,H,W,3] u8, extrinsics [m,4,4]) of the frames that survive."""
    n = len(triples)
    dbuf = np.empty((n, intrinsics.height, intrinsics.width), np.uint16)
    cbuf = np.empty((n, intrinsics.height, intrinsics.width, 3), np.uint8)
    futs = [pool.submit(_decode_into, t, intrinsics, T_fix, dbuf[k], cbuf[k]) for k, t in enumerate(triples)]
    slots, exts = [], []
    for k, (fut, triple) in enumerate(zip(futs, triples)):
        ext, err = fut.result()
        if err is not None:
            if not skip_errors:
                for f in futs[k + 1:]:
                    f.cancel()
                raise err
            if on_error:
                on_error(triple[3], err)
            continue
        slots.append(k); exts.append(ext)
        if progress:
            progress(triple[3], k + 1, n)
    if len(slots) != n:
        dbuf, cbuf = np.ascontiguousarray(dbuf[slots]), np.ascontiguousarray(cbuf[slots])
    return dbuf, cbuf, (np.stack(exts) if exts else np.zeros((0, 4, 4)))


def interleave_plan(counts, batch=32):
    """Order in which the frames of several objects enter an arena: batches of `batch` frames drawn round-robin in runs of
    batch // n_active frames per object, each object's own order preserved.  Returns [(object, frame index)]."""
    nxt = [0] * len(counts)
    order = []
    while True:
        active = [o for o, c in enumerate(counts) if nxt[o] < c]
        if not active:
            return order
        run = max(1, batch // len(active))
        for o in active:
            take = min(run, counts[o] - nxt[o])
            order += [(o, nxt[o] + k) for k in range(take)]
            nxt[o] += take


def integrate_many(jobs, max_workers=None):
    """Config 3 (multi_reconstruct_rgbd_filter.py:139-145: several objects, each its own volume, one after the other):
    the objects' frame loops run TOGETHER through one multi-object arena -- one block hash with the object id in the block
    key, one work list and ONE integration launch per batch over the union of the objects' touched blocks (a single small
    object cannot fill 148 SMs).  jobs = [(compat ScalableTSDFVolume, triples, kwargs of integrate_files)]; all jobs share
    the camera and the voxel parameters.  Afterwards every job's volume addresses its own object of the arena (extraction,
    export and statistics are per object); each object's result is bit-identical to a volume of its own.  Returns the frame
    counts in job order."""
    from concurrent.futures import ThreadPoolExecutor
    from .volume import ArenaView, TSDFVolume
    if not jobs:
        return []
    if len(jobs) == 1 or len(jobs) > 8:
        return [integrate_files(vol, triples, **kw) for vol, triples, kw in jobs]
    v0, kw0 = jobs[0][0], jobs[0][2]
    intr = kw0["intrinsics"]
    for vol, _, kw in jobs:
        if (vol.voxel_length, vol.sdf_trunc, vol.color_type) != (v0.voxel_length, v0.sdf_trunc, v0.color_type) or \
                kw["intrinsics"].fxfycxcy() != intr.fxfycxcy() or (kw["intrinsics"].width, kw["intrinsics"].height) != (intr.width, intr.height):
            return [integrate_files(vol, triples, **kw) for vol, triples, kw in jobs]      # not one arena's worth: run them apart
    with ThreadPoolExecutor(max_workers=max_workers or _decode_workers()) as pool:
        decoded = [_decode_job(pool, triples, kw["intrinsics"], kw["T_fix"], kw.get("skip_errors", False), kw.get("on_error"),
                               kw.get("progress")) for _, triples, kw in jobs]
    counts = [len(d[2]) for d in decoded]
    color = v0._vol._h is not None and v0.color_type.name == "RGB8"
    arena = TSDFVolume(v0.voxel_length, v0.sdf_trunc, color=color, device=v0._vol.device)
    arena.set_objects(len(jobs))
    order = interleave_plan(counts)
    if order:
        objs = np.array([o for o, _ in order], np.int32)
        depth = np.empty((len(order),) + decoded[0][0].shape[1:], np.uint16)
        rgb = np.empty((len(order),) + decoded[0][1].shape[1:], np.uint8)
        ext = np.empty((len(order), 4, 4), np.float64)
        for o in range(len(jobs)):
            sel = np.nonzero(objs == o)[0]
            depth[sel], rgb[sel], ext[sel] = decoded[o][0][:len(sel)], decoded[o][1][:len(sel)], decoded[o][2][:len(sel)]
        arena.integrate_batch(depth, rgb if color else None, intr.fxfycxcy(), ext, kw0.get("depth_scale", 1000.0),
                              kw0.get("depth_trunc", 3.0), object_ids=objs)
    for o, (vol, _, _) in enumerate(jobs):
        vol._vol.close()
        vol._vol = ArenaView(arena, o)
    return counts


def stdout_progress(fmt):
    def cb(label, i, n):
        sys.stdout.write(fmt.format(label=label, i=i, n=n))
        sys.stdout.flush()
    return cb
