#!/usr/bin/env python3
"""Drop-in for /root/reference/3d_model/reconstruct_rgbd_gt.py (SURVEY 8f "next" row 2): single
object captured with ground-truth odometry (`gt_color_*.jpg`, `gt_depth_*.png`, `gt_pose_*.txt`),
the STANDARD body->optical matrix (transpose of the main scripts' T_fix, reference :52-57), mesh
output, (headless) viewer at the end."""
import glob
import os

import numpy as np

from _common import DEPTH_SCALE, DEPTH_TRUNC, SDF_TRUNC, VOXEL_LENGTH, cx, cy, fx, fy, height, o3d, width
from otslam_b200 import pipeline

base_dir = os.environ.get("OTSLAM_BASE_DIR", "/home/ros2_env/taki/otslam/3d_model/object_scan")     # reference :12
color_dir = os.path.join(base_dir, "color")
depth_dir = os.path.join(base_dir, "depth")
pose_dir = os.path.join(base_dir, "poses")
save_dir = os.path.join(base_dir, "3d_reconst")
try:
    os.makedirs(save_dir, exist_ok=True)          # the reference does this at import time (:18-19)
except OSError as e:
    print(f"[otslam_b200] cannot create {save_dir}: {e} (set OTSLAM_BASE_DIR)")

intrinsics = o3d.camera.PinholeCameraIntrinsic(width, height, fx, fy, cx, cy)
T_fix = np.array([[0, 0, 1, 0], [-1, 0, 0, 0], [0, -1, 0, 0], [0, 0, 0, 1]])


def main():
    color_files = sorted(glob.glob(os.path.join(color_dir, "gt_color*.jpg")))
    if not color_files:      # rgbd_capture_node_gt.cpp:126 writes the colour image as PNG; the reference script globs *.jpg
        color_files = sorted(glob.glob(os.path.join(color_dir, "gt_color*.png")))
    depth_files = sorted(glob.glob(os.path.join(depth_dir, "gt_depth*.png")))
    pose_files = sorted(glob.glob(os.path.join(pose_dir, "gt_pose*.txt")))
    n_frames = len(color_files)
    if n_frames == 0:
        print(f"❌ Error: No .png files found in {color_dir}")
        return
    print(f"✅ Found {n_frames} frames. Starting reconstruction...")
    volume = o3d.pipelines.integration.ScalableTSDFVolume(
        voxel_length=VOXEL_LENGTH, sdf_trunc=SDF_TRUNC, color_type=o3d.pipelines.integration.TSDFVolumeColorType.RGB8)
    triples = [(color_files[i], depth_files[i], pose_files[i], i + 1) for i in range(n_frames)]
    pipeline.integrate_files(volume, triples, intrinsics, T_fix, DEPTH_SCALE, DEPTH_TRUNC, skip_errors=False,
                             progress=pipeline.stdout_progress("\rProcessing frame {i}/{n}..."))
    print("\nExtracting mesh...")
    mesh = volume.extract_triangle_mesh()
    mesh.compute_vertex_normals()
    output_path = os.path.join(save_dir, "object_reconst_gt.ply")
    o3d.io.write_triangle_mesh(output_path, mesh)
    print(f"✅ Success! Saved to: {output_path}")
    origin = o3d.geometry.TriangleMesh.create_coordinate_frame(size=0.3)
    o3d.visualization.draw_geometries([mesh, origin])


if __name__ == "__main__":
    main()
