#!/usr/bin/env python3
"""Batch (non-interactive) drop-in for /root/reference/fusion/2d_selective_merge.py: rectangles of the NEW
occupancy map are pasted into the OLD one wherever the new map holds data (smart_paste, reference :58-69).
The reference collects the rectangles with the mouse; here they come from OTSLAM_MERGE_RECTS, a JSON list of
[x, y, w, h], applied in order by a CUDA kernel.  Paths: OTSLAM_OLD_MAP, OTSLAM_NEW_MAP, OTSLAM_MERGED_MAP."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import cv2  # noqa: E402
import numpy as np  # noqa: E402

from otslam_b200 import _lib  # noqa: E402

OLD_MAP_PATH = os.environ.get("OTSLAM_OLD_MAP", "/home/ros2_env/taki/otslam/2d_map/map_check_nov30.pgm")
NEW_MAP_PATH = os.environ.get("OTSLAM_NEW_MAP", "/home/ros2_env/taki/otslam/2d_map/map_dec_11_copy.pgm")
OUTPUT_PATH = os.environ.get("OTSLAM_MERGED_MAP", "/home/ros2_env/taki/otslam/2d_map/map_selective.pgm")


def smart_paste(base_img, overlay_img, x, y, w, h):
    """Same contract as the reference's function: returns base_img with the rectangle updated (in place)."""
    base_img = np.ascontiguousarray(base_img, np.uint8)
    overlay_img = np.ascontiguousarray(overlay_img, np.uint8)
    h_img, w_img = base_img.shape
    _lib.check(_lib.lib.otslam_grid_smart_paste(_lib.ptr(base_img), _lib.ptr(overlay_img), w_img, h_img, int(x), int(y), int(w), int(h),
                                                205, 5, 0))
    return base_img


def main():
    old_img = cv2.imread(OLD_MAP_PATH, cv2.IMREAD_GRAYSCALE)
    new_img = cv2.imread(NEW_MAP_PATH, cv2.IMREAD_GRAYSCALE)
    if old_img is None or new_img is None:
        print("Error: Images not found.")
        return
    if old_img.shape != new_img.shape:
        new_img = cv2.resize(new_img, (old_img.shape[1], old_img.shape[0]))
    result_map = old_img.copy()
    for x, y, w, h in json.loads(os.environ.get("OTSLAM_MERGE_RECTS", "[]")):
        print(f"Applying update to area: x={x}, y={y}, w={w}, h={h}")
        result_map = smart_paste(result_map, new_img, x, y, w, h)
    cv2.imwrite(OUTPUT_PATH, result_map)
    print(f"\nSaved: {OUTPUT_PATH}")


if __name__ == "__main__":
    main()
