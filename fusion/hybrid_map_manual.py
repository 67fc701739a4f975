#!/usr/bin/env python3
"""Batch (non-interactive) drop-in for /root/reference/fusion/hybrid_map_manual.py: the 2-D map becomes grey Z=0
points, every object PLY is painted red, moved by a RECORDED sequence of the reference's key presses and merged
into one hybrid-map PLY.  The reference applies those edits in an Open3D window (ManualAligner, :38-119); here the
same operations run as CUDA kernels, replayed from OTSLAM_MANUAL_KEYS, a JSON object
    {"<object file name>.ply": "WWAZZQ", ...}
with the reference's bindings (:67-77): W/S = +-TRANS_STEP along X, A/D = +-TRANS_STEP along Y (obj_pcd.transform),
Z/C = +-ROT_STEP degrees of yaw about the object's centre (get_center + get_rotation_matrix_from_xyz + rotate),
Q = finish.  Objects without an entry are merged unmoved.  Paths: OTSLAM_MAP_BASE, OTSLAM_OBJ_DIR,
OTSLAM_HYBRID_SAVE override the constants."""
import copy
import glob
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

import otslam_b200.o3d_compat as o3d  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import hybrid_map as _hm  # noqa: E402  (create_map_cloud: same PGM + YAML -> points kernel path)

map_base = os.environ.get("OTSLAM_MAP_BASE", "/home/ros2_env/taki/otslam/2d_map")
yaml_path = os.path.join(map_base, "map_selective.yaml")
pgm_path = os.path.join(map_base, "map_selective.pgm")
obj_dir = os.environ.get("OTSLAM_OBJ_DIR", "/home/ros2_env/taki/otslam/3d_model/object_scan_update/3d_reconst")
save_path = os.environ.get("OTSLAM_HYBRID_SAVE", "/home/ros2_env/taki/otslam/fusion/hybrid_maps/hybrid_map_selective_adjusted.ply")

TRANS_STEP = 0.05  # Translation step size (meters)
ROT_STEP = 2.0     # Rotation step size (degrees)


class ManualAligner:
    """The reference's aligner with the window replaced by a key string."""

    def __init__(self, static_map_pcd, target_obj_pcd, obj_name, keys=""):
        self.map_pcd, self.obj_pcd, self.obj_name, self.keys = static_map_pcd, target_obj_pcd, obj_name, keys

    def run(self):
        print(f"\n=== Adjusting Object: {self.obj_name} === (recorded keys: {self.keys or '-'})")
        ops = {"W": self.move_x_pos, "S": self.move_x_neg, "A": self.move_y_pos, "D": self.move_y_neg,
               "Z": self.rot_yaw_pos, "C": self.rot_yaw_neg}
        for k in self.keys.upper():
            if k == "Q":
                break
            if k in ops:
                ops[k](None)
        print(f"  -> Object {self.obj_name} Confirmed.")
        return self.obj_pcd

    def apply_trans(self, trans_matrix):
        self.obj_pcd.transform(trans_matrix)

    def move_x_pos(self, vis):
        T = np.eye(4); T[0, 3] = TRANS_STEP
        self.apply_trans(T)

    def move_x_neg(self, vis):
        T = np.eye(4); T[0, 3] = -TRANS_STEP
        self.apply_trans(T)

    def move_y_pos(self, vis):
        T = np.eye(4); T[1, 3] = TRANS_STEP
        self.apply_trans(T)

    def move_y_neg(self, vis):
        T = np.eye(4); T[1, 3] = -TRANS_STEP
        self.apply_trans(T)

    def rot_yaw_pos(self, vis):
        center = self.obj_pcd.get_center()
        R = self.obj_pcd.get_rotation_matrix_from_xyz((0, 0, np.radians(ROT_STEP)))
        self.obj_pcd.rotate(R, center=center)

    def rot_yaw_neg(self, vis):
        center = self.obj_pcd.get_center()
        R = self.obj_pcd.get_rotation_matrix_from_xyz((0, 0, np.radians(-ROT_STEP)))
        self.obj_pcd.rotate(R, center=center)


def create_map_cloud(yaml_file, pgm_file):
    pcd = _hm.create_map_cloud(yaml_file, pgm_file)
    if pcd is not None:
        pcd.paint_uniform_color([0.3, 0.3, 0.3])  # Gray walls (reference :145)
    return pcd


def main():
    keys = json.loads(os.environ.get("OTSLAM_MANUAL_KEYS", "{}"))
    print("--- 1. Loading 2D Map ---")
    map_pcd = create_map_cloud(yaml_path, pgm_path)
    if map_pcd is None:
        return
    print("\n--- 2. Object Adjustment (recorded) ---")
    ply_files = sorted(glob.glob(os.path.join(obj_dir, "*.ply")))
    final_merged_pcd = copy.deepcopy(map_pcd)
    if len(ply_files) == 0:
        print("No objects found.")
        return
    for f in ply_files:
        obj_name = os.path.basename(f)
        temp_pcd = o3d.io.read_point_cloud(f)
        if len(temp_pcd.points) == 0:
            mesh = o3d.io.read_triangle_mesh(f)
            temp_pcd = mesh.sample_points_uniformly(number_of_points=15000)
        temp_pcd.paint_uniform_color([1.0, 0.0, 0.0])
        adjusted_obj = ManualAligner(map_pcd, temp_pcd, obj_name, keys.get(obj_name, "")).run()
        final_merged_pcd += adjusted_obj
    print("\n--- 3. Saving Final Map ---")
    if not os.path.exists(os.path.dirname(save_path)):
        os.makedirs(os.path.dirname(save_path))
    o3d.io.write_point_cloud(save_path, final_merged_pcd)
    print(f"Saved to: {save_path}")
    o3d.visualization.draw_geometries([final_merged_pcd], window_name="Final Result")


if __name__ == "__main__":
    main()
