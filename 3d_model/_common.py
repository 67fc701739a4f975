"""Module-level configuration shared by the drop-in 3d_model scripts: same constant names and
defaults as the reference (/root/reference/3d_model/reconstruct_rgbd.py:11-34), plus environment
overrides (the reference is configured by editing the file)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

import otslam_b200.o3d_compat as o3d  # noqa: E402,F401


def env_float(name, default):
    return float(os.environ.get(name, default))


def scan_dirs(default_base):
    base_dir = os.environ.get("OTSLAM_BASE_DIR", default_base)
    dirs = {k: os.path.join(base_dir, v) for k, v in
            (("color_dir", "color"), ("depth_dir", "depth"), ("pose_dir", "poses"), ("save_dir", "3d_reconst"))}
    try:
        os.makedirs(dirs["save_dir"], exist_ok=True)       # the reference does this at import time
    except OSError as e:
        print(f"[otslam_b200] cannot create {dirs['save_dir']}: {e} (set OTSLAM_BASE_DIR)")
    return base_dir, dirs


# Camera (Gazebo RealSense R200 model) and ROS-body -> optical fix, reconstruct_rgbd.py:22-34
fx, fy = 565.6009, 565.6009
cx, cy = 320.5, 240.5
width, height = 640, 480
T_fix = np.array([[0, -1, 0, 0], [0, 0, -1, 0], [1, 0, 0, 0], [0, 0, 0, 1]])
VOXEL_LENGTH = env_float("OTSLAM_VOXEL_LENGTH", 0.01)      # reference: voxel_length=0.01
SDF_TRUNC = env_float("OTSLAM_SDF_TRUNC", 0.04)            # reference: sdf_trunc=0.04
DEPTH_SCALE, DEPTH_TRUNC = 1000.0, 3.0


# ---- shared tail of the two filter scripts (reference reconstruct_rgbd_filter.py:112-140 ==
#      multi_reconstruct_rgbd_filter.py:110-137): lives here so that neither script has to import the other
#      (importing reconstruct_rgbd_filter would run ITS module-level directory set-up for a different dataset)
Z_FILTER_THRESHOLD = 0.03          # removes floor points (reference :22)
NUMBER_OF_POINTS = 100000          # reference :123
POST_VOXEL = float(os.environ.get("OTSLAM_POST_VOXEL", "0") or 0)
POST_SOR = os.environ.get("OTSLAM_POST_SOR", "")
SAMPLE_SEED = os.environ.get("OTSLAM_SAMPLE_SEED")


def filter_and_save(mesh, obj_name, save_dir):
    if len(mesh.vertices) == 0:
        print("❌ Warning: Mesh is empty! Check poses or depth scale.")
        return None
    print(f"   Filtering points below Z < {Z_FILTER_THRESHOLD:.2f}m...")
    seed = None if SAMPLE_SEED is None else int(SAMPLE_SEED)
    if getattr(mesh, "_res", None) is not None and os.environ.get("OTSLAM_HOST_POST", "0") in ("", "0"):
        # the mesh is still in HBM: sample -> z mask -> (voxel_down_sample -> remove_statistical_outlier) stay there too and the
        # final cloud is downloaded once (otslam_b200.cloud.DeviceCloud; same kernels, bit-identical to the host-array calls)
        from otslam_b200.cloud import DeviceCloud
        from otslam_b200.o3d_compat.geometry import _next_seed
        dc = DeviceCloud.sample_mesh(mesh._res[0], NUMBER_OF_POINTS, _next_seed(seed), colors=mesh.has_vertex_colors(), normals=False)
        dc = dc.zfilter(Z_FILTER_THRESHOLD)
        if POST_VOXEL > 0:
            dc = dc.voxel_down_sample(POST_VOXEL)
        if POST_SOR:
            k, ratio = POST_SOR.split(",")
            dc, _ = dc.remove_statistical_outlier(int(k), float(ratio))
        filtered_pcd = dc.to_pointcloud()
    else:
        pcd = mesh.sample_points_uniformly(number_of_points=NUMBER_OF_POINTS, seed=seed)
        points, colors = np.asarray(pcd.points), np.asarray(pcd.colors)
        mask = points[:, 2] >= Z_FILTER_THRESHOLD
        filtered_pcd = o3d.geometry.PointCloud()
        filtered_pcd.points = o3d.utility.Vector3dVector(points[mask])
        filtered_pcd.colors = o3d.utility.Vector3dVector(colors[mask])
        if POST_VOXEL > 0:                                     # north_star stage, off by default (the reference does not run it)
            filtered_pcd = filtered_pcd.voxel_down_sample(POST_VOXEL)
        if POST_SOR:
            k, ratio = POST_SOR.split(",")
            filtered_pcd, _ = filtered_pcd.remove_statistical_outlier(int(k), float(ratio))
    print(f"   Points remaining: {len(filtered_pcd.points)}")
    output_path = os.path.join(save_dir, f"{obj_name}.ply")
    o3d.io.write_point_cloud(output_path, filtered_pcd)
    print(f"✅ Saved 3D Model: {output_path}")
    return output_path
