"""Shared frame loop of the drop-in scripts.

The reference integrates frame by frame (/root/reference/3d_model/reconstruct_rgbd.py:86-109:
read colour, read depth, loadtxt pose, pose @ T_fix, inv, create_from_color_and_depth, integrate).
Here the files are decoded on the host in chunks and each chunk goes to the GPU in ONE
`integrate_sequence` call (frame order preserved, up to 32 frames fused per block residency);
per-frame failures keep the reference's semantics: abort (reconstruct_rgbd.py, no try) or
print-and-skip (reconstruct_rgbd_filter.py:108-109, multi_reconstruct_rgbd_filter.py:102-103).
"""
import os
import sys

import numpy as np

from .o3d_compat import io as o3d_io

CHUNK_FRAMES = 256
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
    return max(1, min(16, int(os.environ.get("OTSLAM_DECODE_THREADS", "0")) or (os.cpu_count() or 1)))


def _try_load(triple, intrinsics, T_fix):
    cp, dp, pp, _ = triple
    try:
        return load_frame(cp, dp, pp, intrinsics, T_fix), None
    except Exception as err:  # noqa: BLE001
        return None, err


def integrate_files(volume, triples, intrinsics, T_fix, depth_scale=1000.0, depth_trunc=3.0, skip_errors=False,
                    progress=None, on_error=None):
    """Integrate capture triples [(color, depth, pose, label)] in order. Returns frames integrated.

    JPEG / PNG decoding is what bounds this loop once integration runs on the GPU (~5 ms per frame pair on
    one core against ~0.02 ms of GPU work), so the files of a chunk are decoded by a thread pool (OpenCV
    releases the GIL) and chunk k+1 is decoded while the GPU integrates chunk k (the C-ABI call releases the
    GIL as well).  Results are consumed in file order, so the per-frame semantics -- abort on the first bad
    frame, or print-and-skip -- and the frame order seen by the volume are exactly the sequential loop's."""
    from concurrent.futures import ThreadPoolExecutor
    done = 0
    n = len(triples)
    chunks = [triples[c0:c0 + CHUNK_FRAMES] for c0 in range(0, n, CHUNK_FRAMES)]
    if not chunks:
        return 0
    with ThreadPoolExecutor(max_workers=_decode_workers()) as pool:
        def submit(chunk):
            return [pool.submit(_try_load, t, intrinsics, T_fix) for t in chunk]

        pending = submit(chunks[0])
        for ci, chunk in enumerate(chunks):
            futures = pending
            cols, deps, exts = [], [], []
            for k, (fut, triple) in enumerate(zip(futures, chunk)):
                frame, err = fut.result()
                label = triple[3]
                if err is not None:
                    if not skip_errors:
                        for f in futures[k + 1:]:
                            f.cancel()
                        raise err
                    if on_error:
                        on_error(label, err)
                    continue
                c, d, e = frame
                cols.append(c); deps.append(d); exts.append(e)
                if progress:
                    progress(label, ci * CHUNK_FRAMES + k + 1, n)
            pending = submit(chunks[ci + 1]) if ci + 1 < len(chunks) else []     # decoded while this chunk integrates
            if exts:
                volume.integrate_sequence(np.stack(deps), np.stack(cols), intrinsics, np.stack(exts), depth_scale, depth_trunc)
                done += len(exts)
    return done


def integrate_many(jobs, max_workers=None):
    """Config 3 (multi_reconstruct_rgbd_filter.py: several objects, each its own volume): run the
    independent per-object frame loops CONCURRENTLY on one GPU.  Each volume owns its CUDA streams and
    the C-ABI calls release the GIL, so the kernels of small per-object volumes (which alone cannot
    fill 148 SMs) overlap on the device.  jobs = [(volume, triples, kwargs)]; returns the frame counts
    in job order.  Results are identical to running the jobs one after the other."""
    from concurrent.futures import ThreadPoolExecutor
    if not jobs:
        return []
    with ThreadPoolExecutor(max_workers=max_workers or len(jobs)) as ex:
        futs = [ex.submit(integrate_files, vol, triples, **kw) for vol, triples, kw in jobs]
        return [f.result() for f in futs]


def stdout_progress(fmt):
    def cb(label, i, n):
        sys.stdout.write(fmt.format(label=label, i=i, n=n))
        sys.stdout.flush()
    return cb
