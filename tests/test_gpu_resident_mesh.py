"""The drop-in's extract_triangle_mesh() leaves the mesh in HBM (lazy TriangleMesh): everything the
reference scripts do with it (len(mesh.vertices), compute_vertex_normals, sample_points_uniformly,
write_triangle_mesh -- reconstruct_rgbd.py:112-118, reconstruct_rgbd_filter.py:113-123) must give exactly
what the eager host path gives."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _volume(n_frames=6):
    import otslam_b200.o3d_compat as o3d
    from otslam_b200 import synth
    seq = synth.make_sequence("table", 300, subsample=(0, 300 // n_frames))
    d, c = seq.numpy()
    intr = o3d.camera.PinholeCameraIntrinsic(*synth.REF_INTRINSICS)
    vol = o3d.pipelines.integration.ScalableTSDFVolume(voxel_length=0.01, sdf_trunc=0.04,
                                                       color_type=o3d.pipelines.integration.TSDFVolumeColorType.RGB8)
    vol.integrate_sequence(d, c, intr, seq.extrinsic)
    return o3d, vol


def test_lazy_mesh_equals_eager_mesh(tmp_path):
    o3d, vol = _volume()
    ev, ec, en, ef, _ = vol._vol.extract_triangle_mesh(normals=True)          # eager: downloads everything
    mesh = vol.extract_triangle_mesh()
    assert mesh._res is not None                                              # still in HBM
    assert len(mesh.vertices) == len(ev) and len(mesh.triangles) == len(ef)   # no download for len()
    assert mesh._res is not None
    assert mesh.has_vertex_colors() and not mesh.has_vertex_normals()
    mesh.compute_vertex_normals()
    assert mesh.has_vertex_normals() and mesh._res is not None
    # sampling straight from HBM == sampling the downloaded mesh through the host operator
    hv, hc, hn, hf = vol._vol.mesh_download(len(ev), len(ef), normals=True)    # a copy; the mesh stays resident
    host = o3d.geometry.TriangleMesh()
    host.vertices, host.vertex_colors, host.vertex_normals, host.triangles = hv, hc, hn, hf
    a = mesh.sample_points_uniformly(20000, seed=4)
    b = host.sample_points_uniformly(20000, seed=4)
    assert mesh._res is not None
    for x, y in ((a.points, b.points), (a.colors, b.colors), (a.normals, b.normals)):
        assert (np.asarray(x) == np.asarray(y)).all()
    # reading the arrays downloads them once; values identical to the eager path
    assert (np.asarray(mesh.vertices) == ev).all() and mesh._res is None
    assert (np.asarray(mesh.triangles) == ef).all()
    assert (np.asarray(mesh.vertex_colors) == ec).all() and (np.asarray(mesh.vertex_normals) == hn).all()
    assert (hn == en).all()                    # normals are summed in triangle order: the same bits on every extraction
    p1, p2 = tmp_path / "lazy.ply", tmp_path / "host.ply"
    lazy2 = vol.extract_triangle_mesh()
    lazy2.compute_vertex_normals()
    assert o3d.io.write_triangle_mesh(str(p1), lazy2) and o3d.io.write_triangle_mesh(str(p2), host)
    assert p1.read_bytes() == p2.read_bytes()


def test_lazy_mesh_survives_reextract_and_reset():
    o3d, vol = _volume(4)
    m1 = vol.extract_triangle_mesh()
    n1 = len(m1.vertices)
    m2 = vol.extract_triangle_mesh()            # replaces the resident mesh: m1 must have been downloaded first
    assert m1._res is None and len(np.asarray(m1.vertices)) == n1
    vol.reset()                                 # frees it: m2 downloads
    assert m2._res is None and (np.asarray(m2.vertices) == np.asarray(m1.vertices)).all()
    m2.vertices = np.zeros((3, 3))              # assignment works on a materialised mesh
    assert len(m2.vertices) == 3
    e = vol.extract_triangle_mesh()
    assert len(e.vertices) == 0 and not e.has_triangles()
    with pytest.raises(RuntimeError):
        e.sample_points_uniformly(10)


def test_device_cloud_chain_equals_host_chain():
    """sample -> z mask -> voxel_down_sample -> remove_statistical_outlier on CUDA tensors (otslam_b200.cloud.DeviceCloud: the
    operators take device pointers, nothing crosses PCIe between them) == the same chain through host arrays, bit for bit."""
    from otslam_b200.cloud import DeviceCloud
    o3d, vol = _volume()
    mesh = vol.extract_triangle_mesh()
    mesh.compute_vertex_normals()
    host = mesh.sample_points_uniformly(60000, seed=9)
    pts, cols = np.asarray(host.points), np.asarray(host.colors)
    keep = pts[:, 2] >= 0.03
    hf = o3d.geometry.PointCloud()
    hf.points, hf.colors = pts[keep], cols[keep]
    hv = hf.voxel_down_sample(0.012)
    hs, hidx = hv.remove_statistical_outlier(20, 2.0)
    dc = DeviceCloud.sample_mesh(vol._vol, 60000, 9, colors=True, normals=True)
    assert dc.points.is_cuda and (dc.points.cpu().numpy() == pts).all() and (dc.normals.cpu().numpy() == np.asarray(host.normals)).all()
    df = dc.zfilter(0.03)
    assert len(df) == int(keep.sum()) and (df.points.cpu().numpy() == pts[keep]).all() and (df.colors.cpu().numpy() == cols[keep]).all()
    dv = df.voxel_down_sample(0.012)
    assert len(dv) == len(hv.points) and (dv.points.cpu().numpy() == hv.points).all() and (dv.colors.cpu().numpy() == hv.colors).all()
    ds, didx = dv.remove_statistical_outlier(20, 2.0)
    assert didx.is_cuda and didx.cpu().tolist() == hidx
    out = ds.to_pointcloud()
    assert (out.points == hs.points).all() and (out.colors == hs.colors).all() and 0 < len(out.points) < len(hv.points)
    # voxel_down_sample averages normals like colours: host compat and device chain agree, and a unit-normal
    # input gives means of length <= 1
    hn = o3d.geometry.PointCloud()
    hn.points, hn.colors, hn.normals = pts, cols, np.asarray(host.normals)
    hnv = hn.voxel_down_sample(0.012)
    dnv = dc.voxel_down_sample(0.012)
    assert hnv.has_normals() and (dnv.normals.cpu().numpy() == hnv.normals).all() and (dnv.points.cpu().numpy() == hnv.points).all()
    assert np.linalg.norm(hnv.normals, axis=1).max() <= 1 + 1e-12
