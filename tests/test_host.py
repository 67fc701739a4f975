"""Host-side logic that needs no GPU: the o3d_compat containers, PLY I/O layout,
dataset discovery of the drop-in scripts, capture-tree writer, and the per-frame error semantics."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

import otslam_b200.o3d_compat as o3d
from otslam_b200 import pipeline, synth


def test_vector_containers_copy_in_and_view_out():
    a = np.random.rand(5, 3)
    pc = o3d.geometry.PointCloud()
    pc.points = o3d.utility.Vector3dVector(a)
    a[0, 0] = 99.0
    assert pc.points[0, 0] != 99.0                       # copied in
    assert np.asarray(pc.points) is pc.points            # view out
    assert len(pc.points) == 5 and not pc.has_colors()
    with pytest.raises(RuntimeError):
        o3d.utility.Vector3dVector(np.zeros((3, 2)))


def test_pointcloud_paint_and_concat_rules():
    """colours survive += only if both sides have them."""
    a = o3d.geometry.PointCloud(); a.points = np.random.rand(4, 3); a.paint_uniform_color([0.2, 0.2, 0.2])
    b = o3d.geometry.PointCloud(); b.points = np.random.rand(3, 3); b.paint_uniform_color([1, 0, 0])
    c = a + b
    assert len(c.points) == 7 and c.has_colors() and (c.colors[:4] == 0.2).all() and (c.colors[4:] == [1, 0, 0]).all()
    assert len(a.points) == 4                            # + does not modify its operands
    d = o3d.geometry.PointCloud(); d.points = np.random.rand(2, 3)
    e = a + d
    assert len(e.points) == 6 and not e.has_colors()
    empty = o3d.geometry.PointCloud()
    empty += b
    assert empty.has_colors() and len(empty.points) == 3
    b.normals = np.random.rand(3, 3)
    f = a + b
    assert not f.has_normals()                            # the map has no normals -> dropped (hybrid_map.py:115)


def test_ply_layouts(tmp_path):
    """Open3D binary PLY: 27 B/point clouds need the GPU packer; normals+colours (51 B) and meshes
    (51 B/vertex + 13 B/face) are host-packed here."""
    pc = o3d.geometry.PointCloud()
    pc.points = np.random.rand(10, 3); pc.normals = np.random.rand(10, 3); pc.colors = np.random.rand(10, 3)
    p = str(tmp_path / "c.ply")
    assert o3d.io.write_point_cloud(p, pc)
    raw = open(p, "rb").read()
    hdr = raw[:raw.index(b"end_header\n") + 11]
    assert hdr.startswith(b"ply\nformat binary_little_endian 1.0\ncomment Created by Open3D\nelement vertex 10\nproperty double x")
    assert len(raw) - len(hdr) == 10 * 51
    q = o3d.io.read_point_cloud(p)
    assert (q.points == pc.points).all() and (q.normals == pc.normals).all()
    assert np.abs(q.colors - pc.colors).max() <= 0.5 / 255 + 1e-12
    m = o3d.geometry.TriangleMesh()
    m.vertices = np.random.rand(4, 3); m.vertex_normals = np.random.rand(4, 3); m.vertex_colors = np.random.rand(4, 3)
    m.triangles = np.array([[0, 1, 2], [1, 3, 2]])
    p2 = str(tmp_path / "m.ply")
    assert o3d.io.write_triangle_mesh(p2, m)
    raw = open(p2, "rb").read()
    hdr = raw[:raw.index(b"end_header\n") + 11]
    assert b"element face 2\nproperty list uchar uint vertex_indices\n" in hdr
    assert len(raw) - len(hdr) == 4 * 51 + 2 * 13
    r = o3d.io.read_triangle_mesh(p2)
    assert (r.vertices == m.vertices).all() and (r.triangles == m.triangles).all() and r.has_vertex_normals()
    as_cloud = o3d.io.read_point_cloud(p2)               # a mesh PLY read as a cloud yields its vertices
    assert len(as_cloud.points) == 4
    assert not o3d.io.write_point_cloud(str(tmp_path / "e.ply"), o3d.geometry.PointCloud())
    assert len(o3d.io.read_point_cloud(str(tmp_path / "missing.ply")).points) == 0


def test_read_ascii_ply(tmp_path):
    p = tmp_path / "a.ply"
    p.write_text("ply\nformat ascii 1.0\nelement vertex 2\nproperty float x\nproperty float y\nproperty float z\n"
                 "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n0 0 0 255 0 0\n1 2 3 0 255 0\n")
    pc = o3d.io.read_point_cloud(str(p))
    assert (pc.points == [[0, 0, 0], [1, 2, 3]]).all() and (pc.colors == [[1, 0, 0], [0, 1, 0]]).all()


def test_missing_image_is_empty_not_exception(tmp_path, capsys):
    img = o3d.io.read_image(str(tmp_path / "nope.png"))
    assert img.is_empty() and "Read image failed" in capsys.readouterr().out
    with pytest.raises(RuntimeError):
        o3d.geometry.RGBDImage.create_from_color_and_depth(img, img, convert_rgb_to_intensity=False)


def test_capture_tree_and_discovery(tmp_path, monkeypatch):
    intr = (64, 48, 56.0, 56.0, 32.5, 24.5)
    seq = synth.make_sequence("table", 12, intr=intr)
    base = str(tmp_path / "scan")
    synth.write_capture_tree(seq, base, label="Object_0")
    synth.write_capture_tree(seq, base, label="Object_1", start=1)
    assert sorted(os.listdir(os.path.join(base, "depth")))[0] == "Object_0_1.png"
    txt = open(os.path.join(base, "poses", "Object_0_1.txt")).read().split("\n")
    assert len(txt[0].split()) == 4 and txt[3].split() == ["0.000000", "0.000000", "0.000000", "1.000000"]
    monkeypatch.setenv("OTSLAM_BASE_DIR", base)
    sys.path.insert(0, os.path.join(ROOT, "3d_model"))
    for name in ("_common", "reconstruct_rgbd"):
        sys.modules.pop(name, None)
    mod = importlib.import_module("reconstruct_rgbd")
    try:
        assert mod.get_unique_object_names() == ["Object_0", "Object_1"]
        tr = mod.frame_triples("Object_0")
        # lexicographic order: _10 sorts before _2, identically for the three lists
        assert [os.path.basename(t[0]) for t in tr][:4] == ["Object_0_1.jpg", "Object_0_10.jpg", "Object_0_11.jpg", "Object_0_12.jpg"]
        assert all(os.path.basename(t[0])[:-4] == os.path.basename(t[1])[:-4] == os.path.basename(t[2])[:-4] for t in tr)
        assert os.path.isdir(os.path.join(base, "3d_reconst"))      # created at import like the reference
        assert (mod.T_fix == synth.T_FIX).all() and mod.intrinsics.width == 640
        assert (mod.VOXEL_LENGTH, mod.SDF_TRUNC) == (0.01, 0.04)    # reference defaults
    finally:
        sys.path.remove(os.path.join(ROOT, "3d_model"))
        for name in ("_common", "reconstruct_rgbd"):
            sys.modules.pop(name, None)


def test_load_frame_errors_follow_reference_semantics(tmp_path):
    intr = o3d.camera.PinholeCameraIntrinsic(64, 48, 56.0, 56.0, 32.5, 24.5)
    seq = synth.make_sequence("table", 3, intr=(64, 48, 56.0, 56.0, 32.5, 24.5))
    base = str(tmp_path / "scan")
    synth.write_capture_tree(seq, base)
    ok = [os.path.join(base, s, f"Object_0_1.{e}") for s, e in (("color", "jpg"), ("depth", "png"), ("poses", "txt"))]
    c, d, e = pipeline.load_frame(*ok, intr, synth.T_FIX)
    assert c.shape == (48, 64, 3) and d.dtype == np.uint16
    assert np.allclose(e, np.linalg.inv(np.loadtxt(ok[2]) @ synth.T_FIX))
    wrong = o3d.camera.PinholeCameraIntrinsic(640, 480, 1, 1, 1, 1)
    with pytest.raises(RuntimeError, match="Unsupported image format"):
        pipeline.load_frame(*ok, wrong, synth.T_FIX)
    with pytest.raises(Exception):
        pipeline.load_frame(ok[0], ok[1], os.path.join(base, "poses", "nope.txt"), intr, synth.T_FIX)

    class FakeVolume:
        def __init__(self):
            self.calls = []

        def integrate_sequence(self, d, c, intr_, e, s, t):
            self.calls.append(len(e))

    triples = [(ok[0], ok[1], ok[2], 1), (ok[0], os.path.join(base, "depth", "nope.png"), ok[2], 2), (ok[0], ok[1], ok[2], 3)]
    fv, errs = FakeVolume(), []
    n = pipeline.integrate_files(fv, triples, intr, synth.T_FIX, skip_errors=True, on_error=lambda l, e: errs.append(l))
    assert n == 2 and fv.calls == [2] and errs == [2]                  # reconstruct_rgbd_filter.py:108-109: skip the frame
    with pytest.raises(RuntimeError):                                   # reconstruct_rgbd.py: no try -> abort
        pipeline.integrate_files(FakeVolume(), triples, intr, synth.T_FIX, skip_errors=False)


def test_integrate_files_threaded_decode_keeps_file_order(tmp_path, monkeypatch):
    """Chunks are decoded by a thread pool one chunk ahead of the integration; the volume must still see
    every frame exactly once, in file order, with skipped frames dropped in place."""
    intr = o3d.camera.PinholeCameraIntrinsic(64, 48, 56.0, 56.0, 32.5, 24.5)
    seq = synth.make_sequence("table", 11, intr=(64, 48, 56.0, 56.0, 32.5, 24.5))
    base = str(tmp_path / "scan")
    synth.write_capture_tree(seq, base)
    monkeypatch.setattr(pipeline, "CHUNK_FRAMES", 4)
    monkeypatch.setenv("OTSLAM_DECODE_THREADS", "3")

    def triple(i, depth_ok=True):
        return (os.path.join(base, "color", f"Object_0_{i}.jpg"),
                os.path.join(base, "depth", f"Object_0_{i}.png" if depth_ok else "missing.png"),
                os.path.join(base, "poses", f"Object_0_{i}.txt"), i)

    triples = [triple(i, depth_ok=(i not in (3, 8))) for i in range(1, 12)]

    class Recorder:
        def __init__(self):
            self.ext, self.calls = [], []

        def integrate_sequence(self, d, c, intr_, e, s, t):
            assert d.shape[0] == c.shape[0] == e.shape[0]
            self.calls.append(len(e))
            self.ext += [x for x in e]

    rec, errs, prog = Recorder(), [], []
    n = pipeline.integrate_files(rec, triples, intr, synth.T_FIX, skip_errors=True, on_error=lambda l, e: errs.append(l),
                                 progress=lambda label, i, total: prog.append((label, i, total)))
    assert n == 9 and errs == [3, 8] and rec.calls == [3, 3, 3]
    want = [np.linalg.inv(np.loadtxt(triple(i)[2]) @ synth.T_FIX) for i in range(1, 12) if i not in (3, 8)]
    assert all((a == b).all() for a, b in zip(rec.ext, want))
    assert [p[0] for p in prog] == [i for i in range(1, 12) if i not in (3, 8)] and prog[-1][1:] == (11, 11)


def test_read_pose_equals_loadtxt(tmp_path):
    """The pose parser of the threaded loop must give np.loadtxt's bits (reconstruct_rgbd.py:90 uses loadtxt)."""
    from otslam_b200 import capture
    rng = np.random.default_rng(0)
    for k in range(50):
        T = rng.normal(size=(4, 4)) * 10.0 ** rng.integers(-3, 4)
        p = tmp_path / f"p{k}.txt"
        if k % 3 == 0:
            p.write_text(capture.pose_text(T))                          # what the capture nodes write (fixed, 6 decimals)
        elif k % 3 == 1:
            np.savetxt(p, T)                                            # scientific notation, 18 digits
        else:
            p.write_text("\n".join(" ".join(repr(float(x)) for x in r) for r in T) + "\n")
        a, b = pipeline.read_pose(str(p)), np.loadtxt(str(p))
        assert a.shape == (4, 4) and (a.view(np.int64) == b.view(np.int64)).all()
    c = tmp_path / "commented.txt"
    c.write_text("# pose\n" + capture.pose_text(np.eye(4)))                # anything unusual goes through loadtxt itself
    assert (pipeline.read_pose(str(c)) == np.eye(4)).all()
    with pytest.raises(Exception):
        pipeline.read_pose(str(tmp_path / "missing.txt"))


def test_read_poses_library_reader_equals_loadtxt(tmp_path):
    """otslam_read_pose_files (the GPU-decode loop's pose reader: host threads of the library, no interpreter time per file)
    gives np.loadtxt's bits for plain 16-number files and hands everything else back (status 1 / 2)."""
    from otslam_b200 import capture
    rng = np.random.default_rng(1)
    paths, want = [], []
    for k in range(120):
        T = rng.normal(size=(4, 4)) * 10.0 ** rng.integers(-8, 9)
        if k % 10 == 0:
            T.flat[rng.integers(0, 16)] = [0.0, -0.0, 1e-320, 1.7976931348623157e308, 5e-324][k // 10 % 5]
        p = tmp_path / f"p{k}.txt"
        if k % 4 == 0:
            p.write_text(capture.pose_text(T))
        elif k % 4 == 1:
            np.savetxt(p, T)
        elif k % 4 == 2:
            p.write_text("\n".join(" ".join(repr(float(x)) for x in r) for r in T) + "\n")
        else:
            p.write_text("\t".join(f"{x:+.17g}" for x in T.ravel()) + "\r\n")      # one line, tabs, explicit signs, CRLF
        paths.append(str(p)); want.append(np.loadtxt(str(p)).reshape(4, 4))
    poses, status = pipeline.read_poses(paths)
    assert (status == 0).all()
    assert (poses.view(np.int64) == np.stack(want).view(np.int64)).all()
    odd = {"comment": "# pose\n" + capture.pose_text(np.eye(4)), "commas": ",".join(["1.0"] * 16), "fifteen": " ".join(["1.0"] * 15),
           "seventeen": " ".join(["1.0"] * 17), "inf": " ".join(["inf"] * 16), "hex": " ".join(["0x1p3"] * 16),
           "glued": " ".join(["1.0x"] * 16), "underscore": " ".join(["1_0.0"] * 16), "empty": "", "exp": " ".join(["1e"] * 16),
           "nul": " ".join(["1.0"] * 15) + " 1.0\0 2.0"}
    op = []
    for name, txt in odd.items():
        q = tmp_path / f"{name}.txt"
        q.write_bytes(txt.encode())
        op.append(str(q))
    op.append(str(tmp_path / "missing.txt"))
    poses, status = pipeline.read_poses(op)
    assert list(status) == [1] * len(odd) + [2]
    assert pipeline.read_poses([])[1].shape == (0,)


def test_synth_matches_capture_contract():
    seq = synth.make_sequence("table", 300, subsample=(0, 150))
    d, c = seq.numpy()
    assert d.dtype == np.uint16 and d.shape == (2, 480, 640) and c.shape == (2, 480, 640, 3)
    assert d.max() <= 5000 and (d == 0).any()                          # > 5 m -> 0 (scanner_node.cpp:277-278)
    assert np.allclose(seq.pose_ros, np.round(seq.pose_ros, 6))
    T = seq.pose_ros[0] @ synth.T_FIX                                   # optical pose: z looks at the table
    fwd = T[:3, 2]
    to_target = np.array([0, 0, 0.5]) - T[:3, 3]
    assert np.dot(fwd, to_target / np.linalg.norm(to_target)) > 0.999
    # the analytic table top (z = 0.75) is visible at the image centre
    pts = T @ np.array([0, 0, d[0, 240, 320] / 1000.0, 1.0])
    assert abs(pts[2] - 0.75) < 0.01 or abs(pts[2]) < 0.01


def test_bench_reference_arm_runs_on_cpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--frames", "12", "--cpu-frames", "4"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["unit"] == "frames/s"


def test_capture_depth_conversion_matches_scanner_node():
    """scanner_node.cpp:277-281: NaN -> 0, > 5 m -> 0, x1000, round-half-even, saturate."""
    from otslam_b200 import capture
    d = np.array([[np.nan, 0.0, 0.0004, 0.0005, 0.0015, 1.2345, 4.9999, 5.0, 5.0001, 70.0]], np.float32)
    out = capture.depth_to_u16_mm(d)
    assert out.dtype == np.uint16
    exp = [0 if (np.isnan(x) or x > np.float32(5.0)) else int(np.rint(np.float64(x) * 1000.0)) for x in d[0]]
    assert out[0].tolist() == exp
    assert out[0, 0] == 0 and out[0, 6] == 5000 and out[0, 7] == 5000 and out[0, 8] == 0 and out[0, 9] == 0 and out[0, 4] == 2
    assert capture.pose_text(np.eye(4)).splitlines()[3] == "0.000000 0.000000 0.000000 1.000000"


def test_pcd_layout_and_round_trip(tmp_path):
    """north_star names .pcd beside .ply: PCD v0.7 in Open3D's layout (float32 xyz, rgb = float whose bits are r<<16|g<<8|b)."""
    rng = np.random.default_rng(3)
    pc = o3d.geometry.PointCloud()
    pc.points = rng.random((7, 3)); pc.colors = rng.random((7, 3))
    p = str(tmp_path / "c.pcd")
    assert o3d.io.write_point_cloud(p, pc)
    raw = open(p, "rb").read()
    hdr, body = raw[:raw.index(b"DATA binary\n") + 12], raw[raw.index(b"DATA binary\n") + 12:]
    assert hdr.decode().splitlines()[:4] == ["# .PCD v0.7 - Point Cloud Data file format", "VERSION 0.7", "FIELDS x y z rgb", "SIZE 4 4 4 4"]
    assert b"TYPE F F F F\nCOUNT 1 1 1 1\nWIDTH 7\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS 7\n" in hdr and len(body) == 7 * 16
    rec = np.frombuffer(body, np.dtype([("p", "<f4", 3), ("c", "<u4")]))
    assert (rec["p"] == pc.points.astype(np.float32)).all()
    cb = np.floor(pc.colors * 255.0 + 0.5).astype(np.uint32)
    assert (rec["c"] == ((cb[:, 0] << 16) | (cb[:, 1] << 8) | cb[:, 2])).all()
    q = o3d.io.read_point_cloud(p)
    assert (q.points == pc.points.astype(np.float32).astype(np.float64)).all() and np.abs(q.colors - pc.colors).max() <= 0.5 / 255 + 1e-12
    pc.normals = rng.random((7, 3))
    pa = str(tmp_path / "a.pcd")
    assert o3d.io.write_point_cloud(pa, pc, write_ascii=True)
    qa = o3d.io.read_point_cloud(pa)
    assert qa.has_normals() and np.abs(qa.points - pc.points).max() < 1e-6 and np.abs(qa.colors - pc.colors).max() <= 0.5 / 255 + 1e-12
    assert not o3d.io.write_point_cloud(str(tmp_path / "e.pcd"), o3d.geometry.PointCloud())


def test_capture_tool_name_patterns_and_tf_pose(tmp_path):
    """The manual capture tools' layouts (rgbd_capture_node_2.cpp:168-171, rgbd_capture_node_gt.cpp:126-128,
    _rgbd_capture_node.cpp:104-110) and their quaternion -> 4x4 pose."""
    from otslam_b200 import capture
    rgb = np.zeros((4, 6, 3), np.uint8); rgb[..., 0] = 200
    depth = np.full((4, 6), 1.5, np.float32)
    T = capture.pose_from_quaternion(0.0, 0.0, np.sin(0.25), np.cos(0.25), 1.0, 2.0, 0.5)      # yaw 0.5 rad
    assert np.allclose(T[:3, :3], [[np.cos(0.5), -np.sin(0.5), 0], [np.sin(0.5), np.cos(0.5), 0], [0, 0, 1]], atol=1e-15)
    assert np.allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-15) and T[:, 3].tolist() == [1.0, 2.0, 0.5, 1.0]
    want = {"center_table": ("center_table_color_0007.jpg", "center_table_depth_0007.png", "center_table_pose_0007.txt"),
            "gt": ("gt_color_0007.png", "gt_depth_0007.png", "gt_pose_0007.txt"),
            "plain": ("color_0007.png", "depth_0007.png", "pose_0007.txt")}
    import cv2
    for pat, names in want.items():
        base = str(tmp_path / pat)
        assert capture.save_frame(base, "ignored", 7, rgb, depth, T, pattern=pat) == names
        for sub, nm in zip(("color", "depth", "poses"), names):
            assert os.path.exists(os.path.join(base, sub, nm))
        d = cv2.imread(os.path.join(base, "depth", names[1]), cv2.IMREAD_UNCHANGED)
        assert d.dtype == np.uint16 and (d == 1500).all()
        assert np.allclose(np.loadtxt(os.path.join(base, "poses", names[2])), T, atol=5e-7)
    c = cv2.imread(os.path.join(str(tmp_path / "gt"), "color", "gt_color_0007.png"), cv2.IMREAD_UNCHANGED)
    assert (c[..., ::-1] == rgb).all()                     # PNG colour is lossless
    assert capture.save_frame(str(tmp_path / "s"), "Object_3", 12, rgb, depth, T) == "Object_3_12"


def test_interleave_plan_round_robin_runs():
    """pipeline.interleave_plan (config 3 arena ingest): every frame once, per-object order kept, 32-frame batches shared."""
    from otslam_b200 import pipeline
    for counts in ([150, 150, 150, 150], [5, 0, 40], [1], [33, 2, 2, 2, 2, 2, 2, 2], []):
        order = pipeline.interleave_plan(counts)
        assert sorted(order) == sorted((o, k) for o, c in enumerate(counts) for k in range(c))
        for o in range(len(counts)):
            assert [k for oo, k in order if oo == o] == list(range(counts[o]))
    first = pipeline.interleave_plan([150, 150, 150, 150])[:32]
    assert [o for o, _ in first] == [0] * 8 + [1] * 8 + [2] * 8 + [3] * 8


def test_raw_sidecar_round_trip_and_invalidation(tmp_path, monkeypatch):
    """OTSLAM_SIDECAR=1: the first decode of a capture triple stores the decoded arrays; later loads return exactly those
    arrays without decoding; touching a source file invalidates the side-car."""
    import cv2
    from otslam_b200 import capture, pipeline, synth
    rng = np.random.default_rng(1)
    H, W = 48, 64
    rgb = rng.integers(0, 255, (H, W, 3), dtype=np.uint8)
    depth = rng.integers(300, 3000, (H, W), dtype=np.uint16)
    base = str(tmp_path)
    capture.save_frame(base, "Object_0", 1, rgb, depth, np.eye(4))
    t = (os.path.join(base, "color", "Object_0_1.jpg"), os.path.join(base, "depth", "Object_0_1.png"),
         os.path.join(base, "poses", "Object_0_1.txt"), 1)
    intr = o3d.camera.PinholeCameraIntrinsic(W, H, 50.0, 50.0, 32.5, 24.5)
    d0, c0 = np.empty((H, W), np.uint16), np.empty((H, W, 3), np.uint8)
    monkeypatch.setenv("OTSLAM_SIDECAR", "0")
    e0, err = pipeline._decode_into(t, intr, synth.T_FIX, d0, c0)
    assert err is None and not os.path.exists(pipeline._sidecar_path(t[1]))
    monkeypatch.setenv("OTSLAM_SIDECAR", "1")
    d1, c1 = np.empty_like(d0), np.empty_like(c0)
    e1, err = pipeline._decode_into(t, intr, synth.T_FIX, d1, c1)          # decodes, writes the side-car
    assert err is None and os.path.exists(pipeline._sidecar_path(t[1])) and (d1 == d0).all() and (c1 == c0).all()
    calls = []
    real = cv2.imread
    monkeypatch.setattr(cv2, "imread", lambda *a, **k: calls.append(a) or real(*a, **k))
    d2, c2 = np.zeros_like(d0), np.zeros_like(c0)
    e2, err = pipeline._decode_into(t, intr, synth.T_FIX, d2, c2)          # served from the side-car
    assert err is None and not calls and (d2 == d0).all() and (c2 == c0).all() and (e2 == e0).all()
    os.utime(t[1], ns=(1, 1))                                              # the depth PNG "changed"
    d3, c3 = np.zeros_like(d0), np.zeros_like(c0)
    _, err = pipeline._decode_into(t, intr, synth.T_FIX, d3, c3)
    assert err is None and len(calls) == 2 and (d3 == d0).all()


def test_gpu_decode_file_loop_orchestration(tmp_path, monkeypatch):
    """pipeline._integrate_files_gpu with stand-in decoders (no GPU): DECODE_AHEAD chunks are prepared concurrently beside
    the one integrating, a decoder is never handed a new chunk before its previous chunk has been integrated, frames reach
    the volume exactly once in file order, frames the decoders pass on go through the stock decoder and are `put` back,
    and the print-and-skip / abort semantics are the host loop's."""
    import threading
    import time
    intr = o3d.camera.PinholeCameraIntrinsic(64, 48, 56.0, 56.0, 32.5, 24.5)
    seq = synth.make_sequence("table", 23, intr=(64, 48, 56.0, 56.0, 32.5, 24.5))
    base = str(tmp_path / "scan")
    synth.write_capture_tree(seq, base)
    monkeypatch.setattr(pipeline, "CHUNK_FRAMES", 4)
    monkeypatch.setattr(pipeline, "DECODE_AHEAD", 2)
    monkeypatch.setattr(pipeline, "gpu_decode_selfcheck", {"done": True, "failed": None})     # its own test below

    def triple(i, ok=True):
        return (os.path.join(base, "color", f"Object_0_{i}.jpg"),
                os.path.join(base, "depth", f"Object_0_{i}.png" if ok else "missing.png"),
                os.path.join(base, "poses", f"Object_0_{i}.txt"), i)

    passed_on = {2, 9, 14}                                       # status 1: valid files the GPU decoders do not cover
    missing = {6, 19}                                            # status 2 -> the stock path raises the reference's error
    triples = [triple(i, ok=(i not in missing)) for i in range(1, 24)]
    lock = threading.Lock()
    state = {"preparing": 0, "max_preparing": 0, "log": []}

    class FakeDecoder:
        def __init__(self, name):
            self.name, self.busy, self.loaded, self.puts = name, False, None, []

        def decode_files(self, cps, dps):
            with lock:
                assert not self.busy and self.loaded is None, "decoder reused before its chunk was integrated"
                self.busy = True
                state["preparing"] += 1
                state["max_preparing"] = max(state["max_preparing"], state["preparing"])
            time.sleep(0.05)                                     # the library call (GIL released)
            labels = [int(os.path.basename(p).split("_")[-1].split(".")[0]) for p in cps]
            cs = np.array([1 if l in passed_on else 0 for l in labels], np.int32)
            ds = np.array([2 if "missing" in d else 0 for d in dps], np.int32)
            with lock:
                self.busy, self.loaded = False, labels
                state["preparing"] -= 1
            return cs, ds

        def profile(self):
            return {"compressed_bytes": 0}

        def put(self, slot, depth, rgb):
            assert depth.shape == (48, 64) and rgb.shape == (48, 64, 3)
            self.puts.append(self.loaded[slot])

        def integrate(self, vol, slots, k, exts, depth_scale, depth_trunc):
            with lock:
                assert self.loaded is not None and not self.busy
                state["log"].append((self.name, [self.loaded[s] for s in slots], np.array(exts)))
                self.loaded = None
            time.sleep(0.01)

        def close(self):
            pass

    fakes = []

    def acquire(count, h, w, frames, device):
        assert (h, w, frames) == (48, 64, 4) and count == 3      # DECODE_AHEAD + 1
        fakes[:] = [FakeDecoder(i) for i in range(count)]
        pipeline._decoder_lock = pipeline._decoder_lock or threading.Lock()
        return list(fakes), ("fake", len(fakes))

    monkeypatch.setattr(pipeline, "_acquire_decoders", acquire)

    class Vol:
        class _vol:
            device = 0

    errs, prog = [], []
    n = pipeline._integrate_files_gpu(Vol, triples, intr, synth.T_FIX, 1000.0, 3.0, True, lambda l, i, t: prog.append(l),
                                      lambda l, e: errs.append((l, str(e))))
    pipeline._decoder_pool.pop(("fake", 3), None)
    good = [i for i in range(1, 24) if i not in missing]
    assert n == len(good) and [l for l, _ in errs] == sorted(missing) and prog == good
    assert all("Unsupported image format" in m for _, m in errs)
    assert [l for _, labels, _ in state["log"] for l in labels] == good                 # file order, each frame once
    assert [name for name, _, _ in state["log"]] == [0, 1, 2, 0, 1, 2]                  # six chunks round-robin over 3 decoders
    assert state["max_preparing"] == 2                                                   # two chunks in flight beside the integration
    assert sorted(l for f in fakes for l in f.puts) == sorted(passed_on)                 # stock-decoded frames went back into their slots
    want = [np.linalg.inv(np.loadtxt(triple(i)[2]) @ synth.T_FIX) for i in good]
    got = [e for _, _, exts in state["log"] for e in exts]
    assert all((a == b).all() for a, b in zip(got, want))
    assert len(pipeline.last_decode_profile) == 6 and sum(c["passed_on"] for c in pipeline.last_decode_profile) == 5
    # abort semantics: the first bad frame (6, in chunk 1) raises; chunk 0 has been integrated, nothing after it
    state["log"].clear()
    with pytest.raises(RuntimeError, match="Unsupported image format"):
        pipeline._integrate_files_gpu(Vol, triples, intr, synth.T_FIX, 1000.0, 3.0, False, None, None)
    pipeline._decoder_pool.pop(("fake", 3), None)
    assert [labels for _, labels, _ in state["log"]] == [[1, 2, 3, 4]]


def test_gpu_decode_selfcheck_guards_the_first_frame(tmp_path, monkeypatch):
    """Once per process the first GPU-decoded frame pair is compared with the stock decoders'.  Equal: the loop goes on and
    never checks again.  Different: GpuDecodeSelfCheckFailed before anything is integrated, and integrate_files re-runs the
    whole list on the host decoders (every frame integrated exactly once) and keeps GPU decoding off for the process."""
    import threading
    import cv2
    from otslam_b200.volume import TSDFVolume
    intr = o3d.camera.PinholeCameraIntrinsic(64, 48, 56.0, 56.0, 32.5, 24.5)
    seq = synth.make_sequence("table", 6, intr=(64, 48, 56.0, 56.0, 32.5, 24.5))
    base = str(tmp_path / "scan")
    synth.write_capture_tree(seq, base)
    triples = [(os.path.join(base, "color", f"Object_0_{i}.jpg"), os.path.join(base, "depth", f"Object_0_{i}.png"),
                os.path.join(base, "poses", f"Object_0_{i}.txt"), i) for i in range(1, 7)]
    monkeypatch.setattr(pipeline, "CHUNK_FRAMES", 4)
    log = []

    class FakeDecoder:
        wrong = False

        def decode_files(self, cps, dps):
            self.cps, self.dps = cps, dps
            return np.zeros(len(cps), np.int32), np.zeros(len(cps), np.int32)

        def fetch(self, first, count):
            d = cv2.imread(self.dps[first], cv2.IMREAD_UNCHANGED)[None].copy()
            c = cv2.imread(self.cps[first], cv2.IMREAD_UNCHANGED)[..., ::-1][None].copy()
            if FakeDecoder.wrong:
                c[0, 3, 5, 1] ^= 1                              # one bit of one colour sample
            return d, c

        def profile(self):
            return {"compressed_bytes": 0}

        def integrate(self, vol, slots, k, exts, s, t):
            log.append(("gpu", len(slots)))

        def close(self):
            pass

    def acquire(count, h, w, frames, device):
        pipeline._decoder_lock = pipeline._decoder_lock or threading.Lock()
        return [FakeDecoder() for _ in range(count)], ("fake2", count)

    monkeypatch.setattr(pipeline, "_acquire_decoders", acquire)

    class Vol:
        _vol = object.__new__(TSDFVolume)                       # passes the isinstance test; never touched by the fakes
        _vol.device = 0

        def integrate_sequence(self, d, c, i, e, s, t):
            log.append(("host", len(e)))

    monkeypatch.setenv("OTSLAM_GPU_DECODE", "1")
    monkeypatch.delenv("OTSLAM_SIDECAR", raising=False)
    state = {"done": False, "failed": None}
    monkeypatch.setattr(pipeline, "gpu_decode_selfcheck", state)
    assert pipeline.integrate_files(Vol(), triples, intr, synth.T_FIX) == 6
    assert log == [("gpu", 4), ("gpu", 2)] and state == {"done": True, "failed": None}
    # a decoder that gets one bit wrong
    log.clear()
    state.update(done=False, failed=None)
    FakeDecoder.wrong = True
    assert pipeline.integrate_files(Vol(), triples, intr, synth.T_FIX) == 6
    assert log == [("host", 4), ("host", 2)] and state["failed"] and "differs from the stock decoders" in state["failed"]
    log.clear()
    assert pipeline.integrate_files(Vol(), triples, intr, synth.T_FIX) == 6 and log == [("host", 4), ("host", 2)]   # stays off
    pipeline._decoder_pool.pop(("fake2", 2), None)
