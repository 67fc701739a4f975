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
