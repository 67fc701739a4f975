"""Optional hook onto the REAL reference backend -- TEST INFRASTRUCTURE, not product code.

The reference scripts (/root/reference/3d_model/reconstruct_rgbd.py:79-113) do their arithmetic inside
the third-party `open3d` wheel.  It is not installable in the build container, so the oracle
(oracle.cpp) is a restatement and parity is "unpinned".  Wherever a box does have the wheel
(`import open3d` works), this module runs the reference's own call sequence on the same inputs so that
  * tests/test_open3d_crosscheck.py can pin the oracle against it, and
  * bench.py --impl reference / cpu_baseline can time it (kind "open3d <version>").
Nothing here is imported by the product package.
"""
import numpy as np


def available():
    try:
        import open3d  # noqa: F401
        return True
    except Exception:  # noqa: BLE001 -- a broken wheel counts as absent
        return False


def version():
    import open3d
    return str(getattr(open3d, "__version__", "unknown"))


def make_volume(voxel_length, sdf_trunc):
    """reconstruct_rgbd.py:79-83"""
    import open3d as o3d
    return o3d.pipelines.integration.ScalableTSDFVolume(
        voxel_length=float(voxel_length), sdf_trunc=float(sdf_trunc),
        color_type=o3d.pipelines.integration.TSDFVolumeColorType.RGB8)


def intrinsic(width, height, fx, fy, cx, cy):
    """reconstruct_rgbd.py:22-25"""
    import open3d as o3d
    return o3d.camera.PinholeCameraIntrinsic(int(width), int(height), float(fx), float(fy), float(cx), float(cy))


def integrate(volume, depth_u16, rgb_u8, intr, extrinsic, depth_scale=1000.0, depth_trunc=3.0):
    """One iteration of the reference loop body, reconstruct_rgbd.py:99-107."""
    import open3d as o3d
    color = o3d.geometry.Image(np.ascontiguousarray(rgb_u8, np.uint8))
    depth = o3d.geometry.Image(np.ascontiguousarray(depth_u16, np.uint16))
    rgbd = o3d.geometry.RGBDImage.create_from_color_and_depth(
        color, depth, depth_scale=float(depth_scale), depth_trunc=float(depth_trunc), convert_rgb_to_intensity=False)
    volume.integrate(rgbd, intr, np.ascontiguousarray(extrinsic, np.float64))
    return rgbd


def integrate_sequence(depth, rgb, whfxfycxcy, extrinsics, voxel_length, sdf_trunc, depth_trunc=3.0, frames=None):
    vol = make_volume(voxel_length, sdf_trunc)
    intr = intrinsic(*whfxfycxcy)
    for k in (range(len(depth)) if frames is None else frames):
        integrate(vol, depth[k], rgb[k], intr, extrinsics[k], 1000.0, depth_trunc)
    return vol


def near_surface_voxels(volume, voxel_length):
    """What the Python API exposes of the voxel state: extract_voxel_point_cloud() lists every voxel with
    weight != 0 and -0.98 <= tsdf < 0.98 as (centre, grey = (tsdf + 1) / 2).  Returns (global voxel
    index [n,3] int64 sorted lexicographically, tsdf [n] float32 in that order); (tsdf + 1) * 0.5 is exact
    in FP64 for a float32 tsdf, so the TSDF comes back bit for bit."""
    pc = volume.extract_voxel_point_cloud()
    pts = np.asarray(pc.points)
    grey = np.asarray(pc.colors)[:, 0] if len(pts) else np.zeros(0)
    idx = np.floor(pts / voxel_length).astype(np.int64)        # centre = (i + 0.5) * vl
    tsdf = (grey * 2.0 - 1.0).astype(np.float32)
    order = np.lexsort((idx[:, 2], idx[:, 1], idx[:, 0])) if len(idx) else np.zeros(0, np.int64)
    return idx[order], tsdf[order]


def mesh_arrays(volume):
    """reconstruct_rgbd.py:112-113"""
    m = volume.extract_triangle_mesh()
    m.compute_vertex_normals()
    return (np.asarray(m.vertices).copy(), np.asarray(m.vertex_colors).copy(), np.asarray(m.vertex_normals).copy(),
            np.asarray(m.triangles).copy())


def point_cloud_arrays(volume):
    pc = volume.extract_point_cloud()
    return np.asarray(pc.points).copy(), np.asarray(pc.colors).copy(), np.asarray(pc.normals).copy()


def cloud(points, colors=None):
    import open3d as o3d
    pc = o3d.geometry.PointCloud()
    pc.points = o3d.utility.Vector3dVector(np.ascontiguousarray(points, np.float64))
    if colors is not None:
        pc.colors = o3d.utility.Vector3dVector(np.ascontiguousarray(colors, np.float64))
    return pc


def time_frame_loop(depth, rgb, whfxfycxcy, extrinsics, voxel_length, sdf_trunc, budget_s, order):
    """Frames/s of volume.integrate (+ create_from_color_and_depth) over `order`, stopping after budget_s."""
    import time
    vol = make_volume(voxel_length, sdf_trunc)
    intr = intrinsic(*whfxfycxcy)
    done, t0 = 0, time.perf_counter()
    while True:
        k = order[done % len(order)]
        integrate(vol, depth[k], rgb[k], intr, extrinsics[k])
        done += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or done >= 64 * len(order):
            return done, dt
