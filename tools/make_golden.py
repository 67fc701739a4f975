"""Generate tests/golden/small_sequence.npz: a tiny synthetic capture (inputs) plus digests of what
the ORACLE produces for it.  open3d (the reference's backend for this path) is not installable
here, so these are self-generated regression pins for the oracle and the CUDA path, not outputs
of the reference itself ("parity unpinned")."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from otslam_b200 import synth  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def lexorder(k):
    return np.lexsort(tuple(k[:, i] for i in range(k.shape[1] - 1, -1, -1)))


_c, _s = np.cos(0.3), np.sin(0.3)
POST_T = np.array([[_c, -_s, 0.0, 0.05], [_s, _c, 0.0, -0.1], [0.0, 0.0, 1.0, 0.2], [0.0, 0.0, 0.0, 1.0]])


def compute(depth, rgb, intr, extr, vl, trunc, post=True):
    v = oracle.Volume(vl, trunc)
    nupd = []
    for k in range(len(depth)):
        nupd.append(v.integrate(oracle.depth_convert(depth[k], 1000.0, 3.0), rgb[k], intr, extr[k]))
    keys, tsdf, w, col = v.export_blocks()
    verts, cols, faces, ek = v.extract_triangle_mesh()
    o = lexorder(ek)
    pts, pcols, pek = v.extract_point_cloud()
    po = lexorder(pek)
    if not post:      # HD fixture: frame loop + extraction only (k-NN over millions of points is not a seconds-scale oracle run)
        return {
            "n_blocks": int(len(keys)), "touched_updated": [[int(a), int(b)] for a, b in nupd],
            "keys": digest(keys), "weight": digest(w.astype(np.uint16)), "tsdf": digest(tsdf),
            "color_u8": digest(np.floor(col + 0.5).astype(np.uint8)),
            "mesh_nv": int(len(verts)), "mesh_nf": int(len(faces)), "mesh_ekeys": digest(ek[o]), "mesh_verts": digest(verts[o]),
            "pc_n": int(len(pts)), "pc_ekeys": digest(pek[po]), "pc_pts": digest(pts[po]),
        }
    # post stage on the canonically ordered extracted cloud (K10 / K11 / K15): order-dependent sums see the same order
    cp, cc = np.ascontiguousarray(pts[po]), np.ascontiguousarray(pcols[po])
    vp, vc, vk, vn = oracle.voxel_down_sample(cp, cc, 2.5 * vl)
    kept, _ = oracle.remove_statistical_outlier(cp, 20, 2.0)
    tp, _ = oracle.transform(cp, None, POST_T)
    return {
        "vds_n": int(len(vp)), "vds_pts": digest(vp), "vds_cols": digest(vc), "vds_counts": digest(vn.astype(np.int32)),
        "sor_n": int(len(kept)), "sor_idx": digest(np.asarray(kept, np.int64)), "xform_pts": digest(tp),
        "n_blocks": int(len(keys)), "touched_updated": [[int(a), int(b)] for a, b in nupd],
        "keys": digest(keys), "weight": digest(w.astype(np.uint16)), "tsdf": digest(tsdf),
        "color_u8": digest(np.floor(col + 0.5).astype(np.uint8)),
        "mesh_nv": int(len(verts)), "mesh_nf": int(len(faces)), "mesh_ekeys": digest(ek[o]), "mesh_verts": digest(verts[o]),
        "pc_n": int(len(pts)), "pc_ekeys": digest(pek[po]), "pc_pts": digest(pts[po]),
        "tsdf_samples": tsdf[::7, ::511].astype(np.float32).tolist()[:20],
    }


def main():
    intr = (160, 120, 565.6009 / 4, 565.6009 / 4, 80.5, 60.5)
    seq = synth.make_sequence("chair_table", 40, intr=intr, subsample=(0, 10))
    depth, rgb = seq.numpy()
    vl, trunc = 0.02, 0.08
    exp = compute(depth, rgb, seq.fxfycxcy, seq.extrinsic, vl, trunc)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "small_sequence.npz"), depth=depth, rgb=rgb,
                        extrinsic=seq.extrinsic, intr=np.array(seq.fxfycxcy), voxel=np.array([vl, trunc]),
                        expected=np.frombuffer(json.dumps(exp).encode(), np.uint8))
    print(json.dumps(exp)[:300])
    # config 4: 1280x720, K x2, 2 mm voxels / 8 mm truncation, the large-room scene; 8 frames spread over
    # the lawn-mower + orbit trajectory.  Inputs are committed because /root/reference-free boxes must not depend on
    # bit-identical re-rendering.
    seq = synth.make_sequence("room", 200, intr=synth.HD_INTRINSICS, subsample=(3, 25))
    depth, rgb = seq.numpy()
    vl, trunc = 0.002, 0.008
    exp = compute(depth, rgb, seq.fxfycxcy, seq.extrinsic, vl, trunc, post=False)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "hd_sequence.npz"), depth=depth, rgb=rgb,
                        extrinsic=seq.extrinsic, intr=np.array(seq.fxfycxcy), voxel=np.array([vl, trunc]),
                        expected=np.frombuffer(json.dumps(exp).encode(), np.uint8))
    print(json.dumps(exp)[:300])


if __name__ == "__main__":
    main()
