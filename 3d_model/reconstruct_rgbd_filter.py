#!/usr/bin/env python3
"""Drop-in for /root/reference/3d_model/reconstruct_rgbd_filter.py: per-object TSDF fusion on the
B200, mesh -> 100000 uniformly sampled points -> floor removal (z >= Z_FILTER_THRESHOLD) -> point
cloud PLY.  Optional north_star stage (off by default, the reference does not run it):
OTSLAM_POST_VOXEL=<m> voxel_down_sample and OTSLAM_POST_SOR="<k>,<ratio>" statistical outlier
removal on the filtered cloud."""
import glob
import os

import _common
from _common import (Z_FILTER_THRESHOLD, NUMBER_OF_POINTS)  # noqa: F401  (the reference's module constants)
from _common import (DEPTH_SCALE, DEPTH_TRUNC, SDF_TRUNC, T_fix, VOXEL_LENGTH, cx, cy, fx, fy, height, o3d, scan_dirs,
                     width)
from otslam_b200 import pipeline

base_dir, _d = scan_dirs("/home/ros2_env/taki/otslam/3d_model/object_scan_2")
color_dir, depth_dir, pose_dir, save_dir = _d["color_dir"], _d["depth_dir"], _d["pose_dir"], _d["save_dir"]

intrinsics = o3d.camera.PinholeCameraIntrinsic(width, height, fx, fy, cx, cy)


def get_unique_object_names():
    names = set()
    for path in glob.glob(os.path.join(color_dir, "*.jpg")):
        parts = os.path.basename(path).split("_")
        if len(parts) >= 2:
            names.add("_".join(parts[:-1]))
    return sorted(names)


def filter_and_save(mesh, obj_name):
    """Tail of reconstruct_object (reference :112-140); shared with the multi-range script through _common."""
    return _common.filter_and_save(mesh, obj_name, save_dir)


def reconstruct_object(obj_name):
    print("\n========================================")
    print(f"🛠️  Processing: {obj_name}")
    print("========================================")
    cf = sorted(glob.glob(os.path.join(color_dir, f"{obj_name}_*.jpg")))
    df = sorted(glob.glob(os.path.join(depth_dir, f"{obj_name}_*.png")))
    pf = sorted(glob.glob(os.path.join(pose_dir, f"{obj_name}_*.txt")))
    n_frames = len(cf)
    if n_frames == 0:
        print(f"❌ Error: No files found for {obj_name}")
        return
    print(f"📸 Found {n_frames} images for reconstruction.")
    volume = o3d.pipelines.integration.ScalableTSDFVolume(
        voxel_length=VOXEL_LENGTH, sdf_trunc=SDF_TRUNC, color_type=o3d.pipelines.integration.TSDFVolumeColorType.RGB8)
    # a frame whose files are missing/short raises inside the reference's try and is skipped
    triples = [(cf[i], df[i] if i < len(df) else "", pf[i] if i < len(pf) else "", i + 1) for i in range(n_frames)]
    pipeline.integrate_files(
        volume, triples, intrinsics, T_fix, DEPTH_SCALE, DEPTH_TRUNC, skip_errors=True,
        progress=pipeline.stdout_progress("\r   Integrate: {i}/{n}"),
        on_error=lambda label, e: print(f"\n   ⚠️ Skipping frame {label} due to error: {e}"))
    print("\n   Extracting mesh...")
    mesh = volume.extract_triangle_mesh()
    mesh.compute_vertex_normals()
    filter_and_save(mesh, obj_name)


def main():
    objects = get_unique_object_names()
    if not objects:
        print("No objects found in directory!")
        return
    print(f"Found {len(objects)} unique objects: {objects}")
    for obj in objects:
        reconstruct_object(obj)
    print("\n🎉 All reconstructions finished!")


if __name__ == "__main__":
    main()
