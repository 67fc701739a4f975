#!/usr/bin/env python3
"""Drop-in for /root/reference/3d_model/reconstruct_rgbd_filter.py: per-object TSDF fusion on the
B200, mesh -> 100000 uniformly sampled points -> floor removal (z >= Z_FILTER_THRESHOLD) -> point
cloud PLY.  Optional north_star stage (off by default, the reference does not run it):
OTSLAM_POST_VOXEL=<m> voxel_down_sample and OTSLAM_POST_SOR="<k>,<ratio>" statistical outlier
removal on the filtered cloud."""
import glob
import os

import numpy as np

from _common import (DEPTH_SCALE, DEPTH_TRUNC, SDF_TRUNC, T_fix, VOXEL_LENGTH, cx, cy, fx, fy, height, o3d, scan_dirs,
                     width)
from otslam_b200 import pipeline

base_dir, _d = scan_dirs("/home/ros2_env/taki/otslam/3d_model/object_scan_2")
color_dir, depth_dir, pose_dir, save_dir = _d["color_dir"], _d["depth_dir"], _d["pose_dir"], _d["save_dir"]

Z_FILTER_THRESHOLD = 0.03          # removes floor points (reference :22)
NUMBER_OF_POINTS = 100000          # reference :123
POST_VOXEL = float(os.environ.get("OTSLAM_POST_VOXEL", "0") or 0)
POST_SOR = os.environ.get("OTSLAM_POST_SOR", "")
SAMPLE_SEED = os.environ.get("OTSLAM_SAMPLE_SEED")

intrinsics = o3d.camera.PinholeCameraIntrinsic(width, height, fx, fy, cx, cy)


def get_unique_object_names():
    names = set()
    for path in glob.glob(os.path.join(color_dir, "*.jpg")):
        parts = os.path.basename(path).split("_")
        if len(parts) >= 2:
            names.add("_".join(parts[:-1]))
    return sorted(names)


def filter_and_save(mesh, obj_name):
    """Shared tail of the two filter scripts (reference :112-140)."""
    if len(mesh.vertices) == 0:
        print("❌ Warning: Mesh is empty! Check poses or depth scale.")
        return None
    print(f"   Filtering points below Z < {Z_FILTER_THRESHOLD:.2f}m...")
    seed = None if SAMPLE_SEED is None else int(SAMPLE_SEED)
    pcd = mesh.sample_points_uniformly(number_of_points=NUMBER_OF_POINTS, seed=seed)
    points, colors = np.asarray(pcd.points), np.asarray(pcd.colors)
    mask = points[:, 2] >= Z_FILTER_THRESHOLD
    filtered_pcd = o3d.geometry.PointCloud()
    filtered_pcd.points = o3d.utility.Vector3dVector(points[mask])
    filtered_pcd.colors = o3d.utility.Vector3dVector(colors[mask])
    if POST_VOXEL > 0:
        filtered_pcd = filtered_pcd.voxel_down_sample(POST_VOXEL)
    if POST_SOR:
        k, ratio = POST_SOR.split(",")
        filtered_pcd, _ = filtered_pcd.remove_statistical_outlier(int(k), float(ratio))
    print(f"   Points remaining: {len(filtered_pcd.points)}")
    output_path = os.path.join(save_dir, f"{obj_name}.ply")
    o3d.io.write_point_cloud(output_path, filtered_pcd)
    print(f"✅ Saved 3D Model: {output_path}")
    return output_path


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
