#!/usr/bin/env python3
"""Drop-in for /root/reference/3d_model/multi_reconstruct_rgbd_filter.py: objects are manual frame
ranges of ONE file prefix, each range its own volume (OBJECT_RANGES; override with
OTSLAM_OBJECT_RANGES='{"object_0": [1, 16], ...}')."""
import json
import os

from _common import (DEPTH_SCALE, DEPTH_TRUNC, SDF_TRUNC, T_fix, VOXEL_LENGTH, cx, cy, fx, fy, height, o3d, scan_dirs,
                     width)
from otslam_b200 import pipeline

base_dir, _d = scan_dirs("/home/ros2_env/taki/otslam/3d_model/object_scan_update")
color_dir, depth_dir, pose_dir, save_dir = _d["color_dir"], _d["depth_dir"], _d["pose_dir"], _d["save_dir"]

FILE_PREFIX = os.environ.get("OTSLAM_FILE_PREFIX", "Object_0")
OBJECT_RANGES = {
    # name of the .ply : (start, end) inclusive -- the reference ships one active entry (:25-32)
    "object_0": (1, 16),
}
if os.environ.get("OTSLAM_OBJECT_RANGES"):
    OBJECT_RANGES = {k: tuple(v) for k, v in json.loads(os.environ["OTSLAM_OBJECT_RANGES"]).items()}

intrinsics = o3d.camera.PinholeCameraIntrinsic(width, height, fx, fy, cx, cy)

import _common  # noqa: E402  (shared sampling / z-filter / save tail)

Z_FILTER_THRESHOLD = _common.Z_FILTER_THRESHOLD


def _range_triples(start_frame, end_frame):
    triples = []
    for i in range(start_frame, end_frame + 1):
        c_path = os.path.join(color_dir, f"{FILE_PREFIX}_{i}.jpg")
        if not os.path.exists(c_path):                      # reference :78-80
            print(f"   ⚠️ Warning: File missing {FILE_PREFIX}_{i}.jpg, skipping...")
            continue
        triples.append((c_path, os.path.join(depth_dir, f"{FILE_PREFIX}_{i}.png"),
                        os.path.join(pose_dir, f"{FILE_PREFIX}_{i}.txt"), i))
    return triples


def _new_volume():
    return o3d.pipelines.integration.ScalableTSDFVolume(
        voxel_length=VOXEL_LENGTH, sdf_trunc=SDF_TRUNC, color_type=o3d.pipelines.integration.TSDFVolumeColorType.RGB8)


def _finish(obj_name, volume, frames_processed):
    if frames_processed == 0:
        print(f"\n❌ No frames were integrated for {obj_name}. Check your ranges or file paths.")
        return
    print("\n   Extracting mesh...")
    mesh = volume.extract_triangle_mesh()
    mesh.compute_vertex_normals()
    _common.filter_and_save(mesh, obj_name, save_dir)


def reconstruct_range(obj_name, start_frame, end_frame):
    print("\n========================================")
    print(f"🛠️  Processing: {obj_name} (Frames {start_frame} -> {end_frame})")
    print("========================================")
    volume = _new_volume()
    frames_processed = pipeline.integrate_files(
        volume, _range_triples(start_frame, end_frame), intrinsics, T_fix, DEPTH_SCALE, DEPTH_TRUNC, skip_errors=True,
        progress=lambda label, i, n: print(f"\r   Integrate: Frame {label}", end="", flush=True),
        on_error=lambda label, e: print(f"\n   ⚠️ Error on frame {label}: {e}"))
    _finish(obj_name, volume, frames_processed)


def main():
    print(f"Starting reconstruction for {len(OBJECT_RANGES)} objects...")
    if os.environ.get("OTSLAM_PARALLEL_OBJECTS", "1") not in ("", "0") and 1 < len(OBJECT_RANGES) <= 8:
        # config 3: every object is an independent volume -> their frame loops go through ONE multi-object arena (one work
        # list and one integration launch per batch over the union of the objects' blocks); results per object are
        # bit-identical to the sequential loop below (OTSLAM_PARALLEL_OBJECTS=0), only the order of the messages differs
        jobs, names = [], []
        for name, (start, end) in OBJECT_RANGES.items():
            print("\n========================================")
            print(f"🛠️  Processing: {name} (Frames {start} -> {end})")
            print("========================================")
            jobs.append((_new_volume(), _range_triples(start, end),
                         dict(intrinsics=intrinsics, T_fix=T_fix, depth_scale=DEPTH_SCALE, depth_trunc=DEPTH_TRUNC, skip_errors=True,
                              progress=lambda label, i, n: print(f"\r   Integrate: Frame {label}", end="", flush=True),
                              on_error=lambda label, e: print(f"\n   ⚠️ Error on frame {label}: {e}"))))
            names.append(name)
        counts = pipeline.integrate_many(jobs)
        for name, (vol, _, _), n in zip(names, jobs, counts):
            print(f"\n🛠️  {name}: {n} frames integrated")
            _finish(name, vol, n)
    else:
        for name, (start, end) in OBJECT_RANGES.items():
            reconstruct_range(name, start, end)
    print("\n🎉 All reconstructions finished!")


if __name__ == "__main__":
    main()
