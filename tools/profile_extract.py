"""dev: integrate the default bench workload once, then run the post stage kernels (extract mesh + normals, sample) a few times --
the command ncu is pointed at for profiles/extract_r02*.md.  Usage: python tools/profile_extract.py [frames]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from otslam_b200 import synth  # noqa: E402
from otslam_b200.volume import TSDFVolume  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
seq = synth.make_sequence("chair_table", n, device="cuda")
vol = TSDFVolume(0.005, 0.02)
vol.integrate_batch(seq.depth.contiguous(), seq.rgb.contiguous(), seq.fxfycxcy, seq.extrinsic)
for it in range(3):
    t0 = time.perf_counter()
    nv, nf = vol.extract_mesh_resident()
    torch.cuda.synchronize()
    print(f"extract {it}: {1e3 * (time.perf_counter() - t0):.3f} ms  nv {nv} nf {nf} blocks {vol.num_blocks()}")
pts, cols, _ = vol.mesh_sample(100000, 0)
print("sampled", len(pts))
