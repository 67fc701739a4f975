#!/usr/bin/env python3
"""Drop-in for /root/reference/3d_model/check_one_frame.py (SURVEY 8f "next" row 2): one RGB-D frame ->
dense back-projection (PointCloud.create_from_rgbd_image) -> voxel_down_sample(0.01) -> viewer
(headless here; the down-sampled cloud is also written next to the inputs)."""
import os

from _common import cx, cy, fx, fy, height, o3d, width

base_dir = os.environ.get("OTSLAM_BASE_DIR", "/home/ros2_env/taki/otslam/3d_model/object_scan")
color_path = os.path.join(base_dir, "color/color_0000.png")
depth_path = os.path.join(base_dir, "depth/depth_0000.png")
intrinsics = o3d.camera.PinholeCameraIntrinsic(width, height, fx, fy, cx, cy)
depth_scale = 1000.0   # depth saved as uint16 millimetres


def main():
    color_raw = o3d.io.read_image(color_path)
    depth_raw = o3d.io.read_image(depth_path)
    rgbd = o3d.geometry.RGBDImage.create_from_color_and_depth(
        color_raw, depth_raw, depth_scale=depth_scale, depth_trunc=5.0, convert_rgb_to_intensity=False)
    pcd = o3d.geometry.PointCloud.create_from_rgbd_image(rgbd, intrinsics)
    pcd = pcd.voxel_down_sample(0.01)
    out = os.path.join(base_dir, "one_frame_cloud.ply")
    o3d.io.write_point_cloud(out, pcd)
    o3d.visualization.draw_geometries([pcd])
    print(f"✅ Displayed single-frame point cloud ({len(pcd.points)} points, saved to {out}).")
    return pcd


if __name__ == "__main__":
    main()
