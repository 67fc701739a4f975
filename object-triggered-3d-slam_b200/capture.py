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

# File-name patterns of the reference's capture tools: (color, depth, pose) templates over (label, count).
#   scanner       ScannerNode::save_files, scanner_node.cpp:268-299        <label>_<n>.{jpg,png,txt}
#   center_table  rgbd_capture_node_2.cpp:168-171 (manual SPACE capture)    center_table_color_0007.jpg ...
#   gt            rgbd_capture_node_gt.cpp:126-128 (ground-truth odometry)  gt_color_0007.png (PNG colour) ...
#   plain         _rgbd_capture_node.cpp:104-110 (older tool)               color_0007.png ...
PATTERNS = {
    "scanner": ("{label}_{n}.jpg", "{label}_{n}.png", "{label}_{n}.txt"),
    "center_table": ("center_table_color_{n:04d}.jpg", "center_table_depth_{n:04d}.png", "center_table_pose_{n:04d}.txt"),
    "gt": ("gt_color_{n:04d}.png", "gt_depth_{n:04d}.png", "gt_pose_{n:04d}.txt"),
    "plain": ("color_{n:04d}.png", "depth_{n:04d}.png", "pose_{n:04d}.txt"),
}


def pose_from_quaternion(qx, qy, qz, qw, tx, ty, tz):
    """4x4 pose from a TF / odometry message as the capture tools build it: tf2::Matrix3x3(q) (setRotation: s = 2 / |q|^2,
    products in tf2's order) + translation (rgbd_capture_node_2.cpp:199-221, rgbd_capture_node_gt.cpp:148-168)."""
    d = qx * qx + qy * qy + qz * qz + qw * qw
    s = 2.0 / d
    xs, ys, zs = qx * s, qy * s, qz * s
    wx, wy, wz = qw * xs, qw * ys, qw * zs
    xx, xy, xz = qx * xs, qx * ys, qx * zs
    yy, yz, zz = qy * ys, qy * zs, qz * zs
    return np.array([[1.0 - (yy + zz), xy - wz, xz + wy, tx],
                     [xy + wz, 1.0 - (xx + zz), yz - wx, ty],
                     [xz - wy, yz + wx, 1.0 - (xx + yy), tz],
                     [0.0, 0.0, 0.0, 1.0]])


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


def save_frame(base_dir, label, count, rgb_u8, depth_m_or_u16, pose_body_to_map, pattern="scanner"):
    """Write one scan in the layout of one of the reference's capture tools (PATTERNS; default: the scanner action
    server's <label>_<count>.{jpg,png,txt}); depth may be float metres or ready u16 mm.  Returns the three paths' stem
    for the scanner pattern, else the tuple of file names."""
    import cv2
    for sub in ("color", "depth", "poses"):
        os.makedirs(os.path.join(base_dir, sub), exist_ok=True)
    depth = np.asarray(depth_m_or_u16)
    if depth.dtype != np.uint16:
        depth = depth_to_u16_mm(depth)
    names = tuple(t.format(label=label, n=count) for t in PATTERNS[pattern])
    cv2.imwrite(os.path.join(base_dir, "color", names[0]), np.ascontiguousarray(np.asarray(rgb_u8)[..., ::-1]))
    cv2.imwrite(os.path.join(base_dir, "depth", names[1]), depth)
    with open(os.path.join(base_dir, "poses", names[2]), "w") as f:
        f.write(pose_text(pose_body_to_map))
    return f"{label}_{count}" if pattern == "scanner" else names
