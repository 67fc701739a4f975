"""Capture-format writer (SURVEY 8f "next" row 1): a standalone, non-ROS restatement of
ScannerNode::save_files (/root/reference/ros2_ws/src/system_manager/src/scanner_node.cpp:260-302) so
that synthetic and real sequences land in byte-identical `color/ depth/ poses/` trees.

    color/<label>_<n>.jpg   BGR8 via cv::imwrite (JPEG, default quality)
    depth/<label>_<n>.png   16-bit 1-channel PNG, millimetres: NaN -> 0, > 5.0 m -> 0, x1000,
                            cv::Mat::convertTo(CV_16U) = saturate_cast(round-half-even)
    poses/<label>_<n>.txt   4x4 camera(body)->map, std::fixed << setprecision(6)
"""
import os

import numpy as np

MAX_DEPTH_M = 5.0          # scanner_node.cpp:278


def depth_to_u16_mm(depth_m):
    """float32 metres -> uint16 millimetres exactly as scanner_node.cpp:277-281 does."""
    d = np.array(depth_m, dtype=np.float32, copy=True)
    d[np.isnan(d)] = 0.0                       # cv::patchNaNs(depth, 0)
    d[d > np.float32(MAX_DEPTH_M)] = 0.0       # depth.setTo(0, depth > 5.0)
    mm = d.astype(np.float64) * 1000.0         # convertTo(..., CV_16U, 1000.0): scale in double
    return np.clip(np.rint(mm), 0, 65535).astype(np.uint16)     # cvRound (half to even) + saturate


def pose_text(T):
    """4 lines x 4 numbers, fixed 6 decimals (scanner_node.cpp:294-298)."""
    T = np.asarray(T, np.float64).reshape(4, 4)
    return "".join(" ".join("%.6f" % v for v in T[r]) + "\n" for r in range(4))


def save_frame(base_dir, label, count, rgb_u8, depth_m_or_u16, pose_body_to_map):
    """Write one scan as <label>_<count>.{jpg,png,txt}; depth may be float metres or ready u16 mm."""
    import cv2
    for sub in ("color", "depth", "poses"):
        os.makedirs(os.path.join(base_dir, sub), exist_ok=True)
    depth = np.asarray(depth_m_or_u16)
    if depth.dtype != np.uint16:
        depth = depth_to_u16_mm(depth)
    stem = f"{label}_{count}"
    cv2.imwrite(os.path.join(base_dir, "color", stem + ".jpg"), np.ascontiguousarray(np.asarray(rgb_u8)[..., ::-1]))
    cv2.imwrite(os.path.join(base_dir, "depth", stem + ".png"), depth)
    with open(os.path.join(base_dir, "poses", stem + ".txt"), "w") as f:
        f.write(pose_text(pose_body_to_map))
    return stem
