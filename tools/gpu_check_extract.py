"""GPU check: mesh / point extraction parity vs the oracle (dev tool)."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from otslam_b200 import synth, _lib
from otslam_b200.volume import TSDFVolume
from oracle import oracle

def canon(verts, cols, faces, ek):
    order = np.lexsort((ek[:, 3], ek[:, 2], ek[:, 1], ek[:, 0]))
    inv = np.empty_like(order); inv[order] = np.arange(len(order))
    f = inv[faces]
    # rotate each face so the smallest index is first, then sort rows
    r = np.argmin(f, axis=1)
    f = np.stack([np.take_along_axis(f, ((r + k) % 3)[:, None], 1)[:, 0] for k in range(3)], 1)
    f = f[np.lexsort((f[:, 2], f[:, 1], f[:, 0]))]
    return verts[order], cols[order], f, ek[order]

seq = synth.make_sequence("table", 300, subsample=(0, 50))
d, c = seq.numpy()
for vl in (0.01,):
    for slab in (None, (0, 2, 2, 0), (0, 2, 2, 1)):
        ov = oracle.Volume(vl, 4 * vl, slab=slab)
        for k in range(len(seq)):
            ov.integrate(oracle.depth_convert(d[k]), c[k], seq.fxfycxcy, seq.extrinsic[k])
        gv = TSDFVolume(vl, 4 * vl, slab=slab)
        gv.integrate_batch(d, c, seq.fxfycxcy, seq.extrinsic)
        print("slab", slab, "blocks", gv.num_blocks(), ov.num_blocks())
        t = time.time(); overts, ocols, ofaces, oek = ov.extract_triangle_mesh(); to = time.time() - t
        t = time.time(); gverts, gcols, gnrm, gfaces, gek = gv.extract_triangle_mesh(); tg = time.time() - t
        print("  mesh nv", len(gverts), len(overts), "nf", len(gfaces), len(ofaces), "t_oracle %.3f t_gpu %.3f" % (to, tg))
        if len(gverts) == len(overts) and len(gfaces) == len(ofaces):
            a = canon(overts, ocols, ofaces, oek); b = canon(gverts, gcols, gfaces, gek)
            print("  ekeys equal", bool((a[3] == b[3]).all()), "verts maxabs", np.abs(a[0] - b[0]).max(), "bitexact", bool((a[0] == b[0]).all()),
                  "cols maxabs", np.abs(a[1] - b[1]).max(), "faces equal", bool((a[2] == b[2]).all()))
            on = oracle.vertex_normals(gverts, gfaces)
            print("  normals maxabs", np.abs(on - gnrm).max())
        op, oc, oek2 = ov.extract_point_cloud()
        gp, gc, gek2 = gv.extract_point_cloud()
        print("  points", len(gp), len(op))
        if len(gp) == len(op):
            oo = np.lexsort((oek2[:, 3], oek2[:, 2], oek2[:, 1], oek2[:, 0])); go = np.lexsort((gek2[:, 3], gek2[:, 2], gek2[:, 1], gek2[:, 0]))
            print("  pts ekeys equal", bool((oek2[oo] == gek2[go]).all()), "pts maxabs", np.abs(op[oo] - gp[go]).max(), "cols maxabs", np.abs(oc[oo] - gc[go]).max())
        gv.close()
