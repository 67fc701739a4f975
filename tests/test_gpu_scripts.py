"""The four drop-in entry points, end to end on a generated capture tree (script-level
integration tests), compared with the same pipeline run through the oracle."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
from scipy.spatial import cKDTree

from conftest import ROOT, canon_mesh
from oracle import oracle

pytestmark = pytest.mark.gpu


def run_script(rel, env):
    e = dict(os.environ); e.update(env); e["OTSLAM_HEADLESS"] = "1"
    r = subprocess.run([sys.executable, os.path.join(ROOT, rel)], capture_output=True, text=True, env=e, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.fixture(scope="module")
def capture(tmp_path_factory):
    from otslam_b200 import synth
    base = str(tmp_path_factory.mktemp("scan"))
    seqs = {}
    for label, scene in (("Object_0", "table"), ("Object_1", "cone")):
        seq = synth.make_sequence(scene, 24, subsample=(0, 2))
        synth.write_capture_tree(seq, base, label=label)
        seqs[label] = seq
    return base, seqs


def oracle_from_files(base, label, n):
    """The reference loop restated on the files with the oracle (lexicographic order, JPEG decoded once)."""
    import glob
    import cv2
    from otslam_b200 import synth
    cf = sorted(glob.glob(os.path.join(base, "color", f"{label}_*.jpg")))
    v = oracle.Volume(0.01, 0.04)
    for f in cf:
        stem = os.path.basename(f)[:-4]
        col = cv2.imread(f, cv2.IMREAD_UNCHANGED)[..., ::-1]
        dep = cv2.imread(os.path.join(base, "depth", stem + ".png"), cv2.IMREAD_UNCHANGED)
        ext = np.linalg.inv(np.loadtxt(os.path.join(base, "poses", stem + ".txt")) @ synth.T_FIX)
        v.integrate(oracle.depth_convert(dep), np.ascontiguousarray(col), (565.6009, 565.6009, 320.5, 240.5), ext)
    return v


def test_reconstruct_rgbd_script(capture):
    import otslam_b200.o3d_compat as o3d
    base, seqs = capture
    out = run_script("3d_model/reconstruct_rgbd.py", {"OTSLAM_BASE_DIR": base})
    assert "Found 2 objects: ['Object_0', 'Object_1']" in out and "All reconstructions finished" in out
    for label in seqs:
        m = o3d.io.read_triangle_mesh(os.path.join(base, "3d_reconst", f"{label}.ply"))
        ov = oracle_from_files(base, label, len(seqs[label]))
        verts, cols, faces, ek = ov.extract_triangle_mesh()
        assert len(m.vertices) == len(verts) and len(m.triangles) == len(faces) and m.has_vertex_normals()
        # PLY vertex order is the GPU's; compare as sets through a KD-tree and exact sorted coordinates
        assert (np.sort(m.vertices.view([("x", "f8"), ("y", "f8"), ("z", "f8")]).ravel()) ==
                np.sort(np.ascontiguousarray(verts).view([("x", "f8"), ("y", "f8"), ("z", "f8")]).ravel())).all()
        dist, nn = cKDTree(verts).query(m.vertices)
        assert dist.max() == 0.0
        assert np.abs(m.vertex_colors - cols[nn]).max() <= 1.0 / 255            # colour bytes
        assert np.abs(np.linalg.norm(m.vertex_normals, axis=1) - 1).max() < 1e-9


def test_filter_script_and_optional_stage(capture):
    import otslam_b200.o3d_compat as o3d
    base, seqs = capture
    out = run_script("3d_model/reconstruct_rgbd_filter.py", {"OTSLAM_BASE_DIR": base, "OTSLAM_SAMPLE_SEED": "5"})
    assert "Points remaining" in out
    pc = o3d.io.read_point_cloud(os.path.join(base, "3d_reconst", "Object_0.ply"))
    ov = oracle_from_files(base, "Object_0", len(seqs["Object_0"]))
    verts, cols, faces, ek = ov.extract_triangle_mesh()
    assert 0 < len(pc.points) < 100000 and pc.has_colors() and not pc.has_normals()
    assert pc.points[:, 2].min() >= 0.03                                       # floor removed
    # sampled cloud lies on the oracle's mesh surface: chamfer to a dense oracle sampling <= 0.25 voxel-ish
    op, oc, _, _ = oracle.sample_uniform(verts, cols, None, faces, 400000, seed=1)
    op = op[op[:, 2] >= 0.03]
    assert cKDTree(op).query(pc.points)[0].mean() <= 0.25 * 0.01
    n_plain = len(pc.points)
    out = run_script("3d_model/reconstruct_rgbd_filter.py", {"OTSLAM_BASE_DIR": base, "OTSLAM_SAMPLE_SEED": "5",
                                                             "OTSLAM_POST_VOXEL": "0.02", "OTSLAM_POST_SOR": "20,2.0"})
    pc2 = o3d.io.read_point_cloud(os.path.join(base, "3d_reconst", "Object_0.ply"))
    assert 0 < len(pc2.points) < n_plain


def test_multi_ranges_script_and_missing_frames(capture):
    import otslam_b200.o3d_compat as o3d
    base, seqs = capture
    os.rename(os.path.join(base, "color", "Object_0_3.jpg"), os.path.join(base, "color", "hidden.jpg_"))
    os.rename(os.path.join(base, "depth", "Object_0_5.png"), os.path.join(base, "depth", "hidden.png_"))
    try:
        ranges = {"obj_a": [1, 6], "obj_b": [7, 12], "obj_none": [100, 101]}
        out = run_script("3d_model/multi_reconstruct_rgbd_filter.py", {"OTSLAM_BASE_DIR": base, "OTSLAM_OBJECT_RANGES": json.dumps(ranges)})
        assert "File missing Object_0_3.jpg, skipping" in out            # multi_reconstruct_rgbd_filter.py:78-80
        assert "Error on frame 5" in out                                 # unreadable depth -> skip (:102-103)
        assert "No frames were integrated for obj_none" in out           # :105-107
        for name in ("obj_a", "obj_b"):
            pc = o3d.io.read_point_cloud(os.path.join(base, "3d_reconst", f"{name}.ply"))
            assert len(pc.points) > 1000
        assert not os.path.exists(os.path.join(base, "3d_reconst", "obj_none.ply"))
    finally:
        os.rename(os.path.join(base, "color", "hidden.jpg_"), os.path.join(base, "color", "Object_0_3.jpg"))
        os.rename(os.path.join(base, "depth", "hidden.png_"), os.path.join(base, "depth", "Object_0_5.png"))


def test_hybrid_map_script(tmp_path):
    import cv2
    import otslam_b200.o3d_compat as o3d
    from otslam_b200 import synth
    mapdir, objdir = tmp_path / "2d_map", tmp_path / "objs"
    mapdir.mkdir(); objdir.mkdir()
    img = synth.occupancy_map(384, 300, 0.03, 7)
    cv2.imwrite(str(mapdir / "map_selective.pgm"), img)
    (mapdir / "map_selective.yaml").write_text("image: map_selective.pgm\nresolution: 0.05\norigin: [-9.6, -7.5, 0.0]\n"
                                                "negate: 0\noccupied_thresh: 0.65\nfree_thresh: 0.196\n")
    rng = np.random.default_rng(0)
    objs = []
    for i in range(3):
        pc = o3d.geometry.PointCloud(); pc.points = rng.normal(size=(5000 + i, 3)); pc.colors = rng.random((5000 + i, 3))
        o3d.io.write_point_cloud(str(objdir / f"object_{i}.ply"), pc)
        objs.append(pc.points)
    save = tmp_path / "out" / "hybrid.ply"
    out = run_script("fusion/hybrid_map.py", {"OTSLAM_MAP_BASE": str(mapdir), "OTSLAM_OBJ_DIR": str(objdir), "OTSLAM_HYBRID_SAVE": str(save)})
    assert "SUCCESS" in out
    got = o3d.io.read_point_cloud(str(save))
    mp = oracle.grid_to_points(img, 0.05, -9.6, -7.5, 100)
    exp_pts = np.concatenate([mp] + objs)
    assert (got.points == exp_pts).all()                                        # map first, then objects in sorted-file order
    exp_cols = np.concatenate([np.tile([51, 51, 51], (len(mp), 1)), np.tile([255, 0, 0], (len(exp_pts) - len(mp), 1))])
    assert (np.round(got.colors * 255).astype(int) == exp_cols).all()
    ref = oracle.pack_ply_cloud(exp_pts, exp_cols / 255.0)
    assert open(save, "rb").read().split(b"end_header\n", 1)[1] == ref.tobytes()   # byte-exact payload
    # map-only fallback (hybrid_map.py:106-109)
    out = run_script("fusion/hybrid_map.py", {"OTSLAM_MAP_BASE": str(mapdir), "OTSLAM_OBJ_DIR": str(tmp_path / "empty"), "OTSLAM_HYBRID_SAVE": str(save)})
    assert "Continuing with Map Only" in out
    assert len(o3d.io.read_point_cloud(str(save)).points) == len(mp)


def test_multi_objects_in_parallel_match_sequential(capture):
    """Config 3: the objects' frame loops through one multi-object arena (object id in the block key, one work list and one
    integration launch per batch) give exactly the PLYs of the reference-style sequential run."""
    import otslam_b200.o3d_compat as o3d
    base, seqs = capture
    ranges = {"par_a": [1, 4], "par_b": [5, 8], "par_c": [9, 12], "par_d": [2, 11]}
    env = {"OTSLAM_BASE_DIR": base, "OTSLAM_OBJECT_RANGES": json.dumps(ranges), "OTSLAM_SAMPLE_SEED": "3"}
    run_script("3d_model/multi_reconstruct_rgbd_filter.py", dict(env, OTSLAM_PARALLEL_OBJECTS="0"))
    seq_out = {k: open(os.path.join(base, "3d_reconst", f"{k}.ply"), "rb").read() for k in ranges}
    for k in ranges:
        os.remove(os.path.join(base, "3d_reconst", f"{k}.ply"))
    out = run_script("3d_model/multi_reconstruct_rgbd_filter.py", dict(env, OTSLAM_PARALLEL_OBJECTS="1"))
    assert "frames integrated" in out
    for k in ranges:
        assert open(os.path.join(base, "3d_reconst", f"{k}.ply"), "rb").read() == seq_out[k]


def test_reconstruct_rgbd_gt_script(tmp_path):
    """reconstruct_rgbd_gt.py on a tree written the way rgbd_capture_node_gt.cpp writes it (gt_color_0000.png,
    gt_depth_0000.png, gt_pose_0000.txt; poses are camera BODY -> map, the script's own T_fix = standard body -> optical)."""
    import otslam_b200.o3d_compat as o3d
    from otslam_b200 import capture, synth
    T_fix_gt = np.array([[0, 0, 1, 0], [-1, 0, 0, 0], [0, -1, 0, 0], [0, 0, 0, 1]], np.float64)
    base = str(tmp_path)
    poses = synth.trajectory("table", 24)[::4]
    seq = synth.make_sequence("table", 24, poses=poses)
    d, c = seq.numpy()
    pose_body = [np.round(T @ np.linalg.inv(T_fix_gt), 6) for T in poses]
    ext = []
    for k in range(len(poses)):
        capture.save_frame(base, "gt", k, c[k], d[k], pose_body[k], pattern="gt")
        ext.append(np.linalg.inv(np.loadtxt(os.path.join(base, "poses", f"gt_pose_{k:04d}.txt")) @ T_fix_gt))
    out = run_script("3d_model/reconstruct_rgbd_gt.py", {"OTSLAM_BASE_DIR": base})
    assert f"Found {len(poses)} frames" in out and "Success! Saved to" in out
    m = o3d.io.read_triangle_mesh(os.path.join(base, "3d_reconst", "object_reconst_gt.ply"))
    ov = oracle.Volume(0.01, 0.04)
    for k in range(len(poses)):                                     # colour PNGs are lossless: the oracle sees the same pixels
        ov.integrate(oracle.depth_convert(d[k]), c[k], seq.fxfycxcy, ext[k])
    verts, cols, faces, ek = ov.extract_triangle_mesh()
    assert len(verts) > 1000 and len(m.vertices) == len(verts) and len(m.triangles) == len(faces) and m.has_vertex_normals()
    assert cKDTree(verts).query(m.vertices)[0].max() == 0.0


def test_check_one_frame_script(tmp_path):
    """check_one_frame.py -- the reference's only call site of create_from_rgbd_image + voxel_down_sample
    (check_one_frame.py:27-28), on color_0000.png / depth_0000.png as _rgbd_capture_node.cpp names them."""
    import otslam_b200.o3d_compat as o3d
    from otslam_b200 import capture, synth
    base = str(tmp_path)
    seq = synth.make_sequence("chair_table", 40, subsample=(5, 40))
    d, c = seq.numpy()
    capture.save_frame(base, "x", 0, c[0], d[0], np.eye(4), pattern="plain")
    out = run_script("3d_model/check_one_frame.py", {"OTSLAM_BASE_DIR": base})
    assert "Displayed single-frame point cloud" in out
    got = o3d.io.read_point_cloud(os.path.join(base, "one_frame_cloud.ply"))
    pts, cols = oracle.backproject_rgbd(oracle.depth_convert(d[0], 1000.0, 5.0), c[0], seq.fxfycxcy)
    ep, ec, _, _ = oracle.voxel_down_sample(pts, cols, 0.01)
    assert len(got.points) == len(ep) > 1000
    o1, o2 = np.lexsort(got.points.T[::-1]), np.lexsort(ep.T[::-1])
    assert (got.points[o1] == ep[o2]).all()
    assert np.abs(got.colors[o1] - ec[o2]).max() <= 0.5 / 255 + 1e-9
