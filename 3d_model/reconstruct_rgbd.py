#!/usr/bin/env python3
"""Drop-in for /root/reference/3d_model/reconstruct_rgbd.py: one mesh PLY per object prefix found
in color/, integrated on the B200 (`python3 3d_model/reconstruct_rgbd.py`, no arguments;
OTSLAM_BASE_DIR / OTSLAM_VOXEL_LENGTH / OTSLAM_SDF_TRUNC override the module constants)."""
import glob
import os

from _common import (DEPTH_SCALE, DEPTH_TRUNC, SDF_TRUNC, T_fix, VOXEL_LENGTH, cx, cy, fx, fy, height, o3d, scan_dirs,
                     width)
from otslam_b200 import pipeline

base_dir, _d = scan_dirs("/home/ros2_env/taki/otslam/3d_model/object_scan_2")
color_dir, depth_dir, pose_dir, save_dir = _d["color_dir"], _d["depth_dir"], _d["pose_dir"], _d["save_dir"]
intrinsics = o3d.camera.PinholeCameraIntrinsic(width, height, fx, fy, cx, cy)


def get_unique_object_names():
    """Object prefix = colour file name minus its last '_' token (reference :36-58)."""
    names = set()
    for path in glob.glob(os.path.join(color_dir, "*.jpg")):
        parts = os.path.basename(path).split("_")
        if len(parts) >= 2:
            names.add("_".join(parts[:-1]))
    return sorted(names)


def frame_triples(obj_name):
    """Lexicographically sorted colour/depth/pose lists, paired by position (reference :67-69)."""
    cf = sorted(glob.glob(os.path.join(color_dir, f"{obj_name}_*.jpg")))
    df = sorted(glob.glob(os.path.join(depth_dir, f"{obj_name}_*.png")))
    pf = sorted(glob.glob(os.path.join(pose_dir, f"{obj_name}_*.txt")))
    # the reference indexes depth_files[i] / pose_files[i] for i < len(color_files): a shorter list
    # is an IndexError there, so it is one here too
    return [(cf[i], df[i], pf[i], i + 1) for i in range(len(cf))]


def new_volume():
    return o3d.pipelines.integration.ScalableTSDFVolume(
        voxel_length=VOXEL_LENGTH, sdf_trunc=SDF_TRUNC, color_type=o3d.pipelines.integration.TSDFVolumeColorType.RGB8)


def reconstruct_object(obj_name):
    print("\n========================================")
    print(f"🛠️  Processing: {obj_name}")
    print("========================================")
    triples = frame_triples(obj_name)
    if not triples:
        print(f"❌ Error: No files found for {obj_name}")
        return
    print(f"   Found {len(triples)} frames.")
    volume = new_volume()
    # no try/except in the reference: a bad frame aborts the run
    pipeline.integrate_files(volume, triples, intrinsics, T_fix, DEPTH_SCALE, DEPTH_TRUNC, skip_errors=False,
                             progress=pipeline.stdout_progress("\r   Integrate: {i}/{n}"))
    print("\n   Extracting mesh...")
    mesh = volume.extract_triangle_mesh()
    mesh.compute_vertex_normals()
    output_path = os.path.join(save_dir, f"{obj_name}.ply")
    o3d.io.write_triangle_mesh(output_path, mesh)
    print(f"✅ Saved: {output_path}")


def main():
    objects = get_unique_object_names()
    if not objects:
        print("No objects found in directory!")
        return
    print(f"Found {len(objects)} objects: {objects}")
    for obj in objects:
        reconstruct_object(obj)
    print("\n🎉 All reconstructions finished!")


if __name__ == "__main__":
    main()
