"""First GPU check: integrate parity vs the oracle (dev tool; the pytest suite supersedes it)."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import torch
from otslam_b200 import synth, _lib
from otslam_b200.volume import TSDFVolume
from oracle import oracle

print("missing symbols:", _lib.MISSING)
seq = synth.make_sequence("table", 300, subsample=(0, 50))
d, c = seq.numpy()
for vl in (0.01, 0.005):
    ov = oracle.Volume(vl, 4 * vl)
    nupd = 0
    for k in range(len(seq)):
        nupd += ov.integrate(oracle.depth_convert(d[k]), c[k], seq.fxfycxcy, seq.extrinsic[k])[1]
    ok, ot, ow, oc = ov.export_blocks()
    for mode in ("frame", "batch_host", "batch_dev", "batch1"):
        gv = TSDFVolume(vl, 4 * vl)
        t = time.time()
        if mode == "frame":
            for k in range(len(seq)):
                gv.integrate_u16(d[k], c[k], seq.fxfycxcy, seq.extrinsic[k])
        elif mode == "batch_host":
            gv.integrate_batch(d, c, seq.fxfycxcy, seq.extrinsic)
        elif mode == "batch1":
            gv.set_batch(1)
            gv.integrate_batch(d, c, seq.fxfycxcy, seq.extrinsic)
        else:
            gv.integrate_batch(seq.depth.cuda(), seq.rgb.cuda(), seq.fxfycxcy, seq.extrinsic)
        dt = time.time() - t
        gk, gt, gw, gc = gv.export_blocks()
        st = gv.stats()
        same_keys = gk.shape == ok.shape and bool((gk == ok).all())
        print(vl, mode, "blocks", gk.shape[0], ok.shape[0], "keys_equal", same_keys, "stats", st, "oracle_nupd", nupd, "t=%.3f" % dt)
        if same_keys:
            print("   weight_equal", bool((gw == ow).all()), "tsdf_maxabs", float(np.abs(gt - ot)[ow > 0].max()),
                  "tsdf_bitexact", bool((gt == ot).all()), "color_maxabs", float(np.abs(gc - oc)[ow > 0].max()))
        gv.close()
print("launches", _lib.launch_count())
