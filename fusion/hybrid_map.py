#!/usr/bin/env python3
"""Drop-in for /root/reference/fusion/hybrid_map.py: 2-D occupancy map (PGM + YAML) -> grey Z=0
points, every object PLY painted red, merged into one hybrid-map PLY.  The per-pixel Python loop
(reference :45-55) and the paint + concatenate + write (reference :59,88-91,115,121) run as CUDA
kernels.  Paths: OTSLAM_MAP_BASE, OTSLAM_OBJ_DIR, OTSLAM_HYBRID_SAVE override the constants."""
import glob
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import ctypes as C  # noqa: E402

import cv2  # noqa: E402
import numpy as np  # noqa: E402
import yaml  # noqa: E402

import otslam_b200.o3d_compat as o3d  # noqa: E402
from otslam_b200 import _lib  # noqa: E402

map_base = os.environ.get("OTSLAM_MAP_BASE", "/home/ros2_env/taki/otslam/2d_map")
yaml_path = os.path.join(map_base, "map_selective.yaml")
pgm_path = os.path.join(map_base, "map_selective.pgm")
obj_dir = os.environ.get("OTSLAM_OBJ_DIR", "/home/ros2_env/taki/otslam/3d_model/object_scan_update/3d_reconst")
save_path = os.environ.get("OTSLAM_HYBRID_SAVE", "/home/ros2_env/taki/otslam/fusion/hybrid_maps/hybrid_map_selective.ply")

OCCUPIED_BELOW = 100               # "Black < 100" (reference :45)
MAP_COLOR = [0.2, 0.2, 0.2]
OBJECT_COLOR = [1.0, 0.0, 0.0]


def create_map_cloud(yaml_file, pgm_file):
    print(f"   -> Loading Map: {pgm_file}")
    if not os.path.exists(yaml_file):
        print(f"❌ Error: YAML not found: {yaml_file}")
        return None
    with open(yaml_file, "r") as f:
        data = yaml.safe_load(f)
    res, origin = data["resolution"], data["origin"]
    img = cv2.imread(pgm_file, cv2.IMREAD_GRAYSCALE)
    if img is None:
        print(f"❌ Error: PGM not found: {pgm_file}")
        return None
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    n = C.c_int64(0)
    pts = np.empty((h * w, 3), np.float64)
    _lib.check(_lib.lib.otslam_grid_to_points(_lib.ptr(img), w, h, float(res), float(origin[0]), float(origin[1]),
                                              OCCUPIED_BELOW, _lib.ptr(pts), C.byref(n), 0))
    print(f"   -> Generating 3D Walls from {n.value} pixels...")
    pcd = o3d.geometry.PointCloud()
    pcd.points = o3d.utility.Vector3dVector(pts[:n.value])
    pcd.paint_uniform_color(MAP_COLOR)
    return pcd


def load_all_objects(directory):
    ply_files = sorted(glob.glob(os.path.join(directory, "*.ply")))
    if len(ply_files) == 0:
        print(f"❌ Error: No .ply files found in {directory}")
        return None
    print(f"   -> Found {len(ply_files)} objects: {[os.path.basename(f) for f in ply_files]}")
    combined_objects = o3d.geometry.PointCloud()
    for f in ply_files:
        print(f"      Loading: {os.path.basename(f)}...")
        try:
            temp_pcd = o3d.io.read_point_cloud(f)
            if len(temp_pcd.points) == 0:               # vertex-less file: sample the mesh (reference :82-84)
                mesh = o3d.io.read_triangle_mesh(f)
                temp_pcd = mesh.sample_points_uniformly(number_of_points=15000)
            temp_pcd.paint_uniform_color(OBJECT_COLOR)
            combined_objects += temp_pcd
        except Exception as e:  # noqa: BLE001
            print(f"❌ Error loading {f}: {e}")
    return combined_objects


def main():
    print("--- 1. Processing Global Map ---")
    map_pcd = create_map_cloud(yaml_path, pgm_path)
    if map_pcd is None:
        return
    print("\n--- 2. Processing Local Objects ---")
    all_objs_pcd = load_all_objects(obj_dir)
    if all_objs_pcd is None or len(all_objs_pcd.points) == 0:
        print(" CRITICAL WARNING: No objects loaded.")
        print("    Continuing with Map Only...")
        clouds, paint = [map_pcd.points], [MAP_COLOR]
    else:
        print(f"   -> Total object points: {len(all_objs_pcd.points)}")
        print("\n--- 3. Merging & Saving ---")
        clouds, paint = [map_pcd.points, all_objs_pcd.points], [MAP_COLOR, OBJECT_COLOR]
    os.makedirs(os.path.dirname(save_path), exist_ok=True)
    # map + objects: one merge/paint/pack launch writes the 27-byte PLY records (== map_pcd + all_objs_pcd)
    records = o3d.io.pack_cloud_records(clouds, paint=paint)
    o3d.io.write_cloud_records(save_path, records)
    print("✅ SUCCESS! Hybrid map saved to:")
    print(f"   {save_path}")
    print("\nOpening Visualizer... (Gray=Map, Red=Objects)")
    origin = o3d.geometry.TriangleMesh.create_coordinate_frame(size=1.0)
    combined_pcd = o3d.geometry.PointCloud()            # lightweight handle for the (headless) viewer
    o3d.visualization.draw_geometries([combined_pcd, origin], window_name="Hybrid Map")


if __name__ == "__main__":
    main()
