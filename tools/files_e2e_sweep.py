"""Frame loop from a capture tree on disk with the GPU decoders: frames/s against the number of chunks kept in preparation
(pipeline.DECODE_AHEAD) and the chunk size (pipeline.CHUNK_FRAMES).  One synthetic 1000-frame chair+table tree (640x480,
cv::imwrite JPEG + 16-bit PNG; every second depth frame with +-3 mm noise), volumes compared with the host-decoded loop.
Prints one JSON line.  Usage: python tools/files_e2e_sweep.py [frames]"""
import json
import os
import shutil
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import otslam_b200.o3d_compat as o3d
from otslam_b200 import capture, pipeline, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
seq = synth.make_sequence("chair_table", n)
dep, rgb = seq.numpy()
rng = np.random.default_rng(0)
base = tempfile.mkdtemp(prefix="otslam_sweep_")
try:
    for k in range(n):
        d = dep[k] if k % 2 == 0 else (dep[k].astype(np.int32) + rng.integers(-3, 4, dep[k].shape) * (dep[k] > 0)).clip(0, 65535).astype(np.uint16)
        capture.save_frame(base, "Object_0", k + 1, rgb[k], d, seq.pose_ros[k])
    triples = [(os.path.join(base, "color", f"Object_0_{j}.jpg"), os.path.join(base, "depth", f"Object_0_{j}.png"),
                os.path.join(base, "poses", f"Object_0_{j}.txt"), j) for j in range(1, n + 1)]
    size = sum(os.path.getsize(t[0]) + os.path.getsize(t[1]) for t in triples) / n
    intr = o3d.camera.PinholeCameraIntrinsic(*seq.intr)
    vol = o3d.pipelines.integration.ScalableTSDFVolume(voxel_length=0.005, sdf_trunc=0.02,
                                                       color_type=o3d.pipelines.integration.TSDFVolumeColorType.RGB8)

    def timed():
        best = 1e30
        for _ in range(3):                      # first pass warms the page cache, the decoders' buffers and the block pool
            vol.reset()
            t0 = time.perf_counter()
            m = pipeline.integrate_files(vol, triples, intr, synth.T_FIX)
            best = min(best, time.perf_counter() - t0)
        return m / best

    os.environ["OTSLAM_GPU_DECODE"] = "0"
    out = {"frames": n, "compressed_bytes_per_frame": int(size), "host_decode_frames_per_s": timed()}
    ref = vol._vol.stats()
    os.environ["OTSLAM_GPU_DECODE"] = "1"
    out["gpu_decode"] = []
    for chunk in (128, 256):
        for ahead in (1, 2, 3, 4, 6):
            pipeline.CHUNK_FRAMES, pipeline.DECODE_AHEAD = chunk, ahead
            fps = timed()
            out["gpu_decode"].append({"chunk": chunk, "ahead": ahead, "frames_per_s": fps, "identical_volume": vol._vol.stats() == ref})
            pipeline.release_decoders()
    print(json.dumps(out))
finally:
    shutil.rmtree(base, ignore_errors=True)
