"""o3d.visualization: headless no-ops (the reference opens a blocking GUI window,
fusion/hybrid_map.py:128-129; there is no display on a GPU box)."""
import os


def draw_geometries(geometries, window_name="Open3D", **kwargs):
    if os.environ.get("OTSLAM_HEADLESS", "1") != "0":
        print(f"[otslam_b200] draw_geometries('{window_name}'): headless, {len(list(geometries))} geometries not shown")
        return
    raise RuntimeError("no visualizer is available in otslam_b200 (set OTSLAM_HEADLESS=1)")
