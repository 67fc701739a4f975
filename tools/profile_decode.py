"""Workload for the ncu captures of the capture-file decoders (csrc/imgcodec.cu): one chunk of 256 frame pairs (640x480
chair+table frames as cv::imwrite stores them: quality-95 4:2:0 JPEG, 16-bit PNG; every second depth frame carries +-3 mm of
sensor-like noise so that the PNG stream is not all run-lengths) decoded twice through the C ABI.
    ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_decode.csv \
        python tools/profile_decode.py
    ncu --set full --clock-control none --import-source on -k regex:'png_inflate|jpeg_huff' -s 2 -c 2 \
        -o gpurun_out/prof_decode python tools/profile_decode.py
Without ncu it prints the decoder's wall / device times as one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cv2
import numpy as np

from otslam_b200 import synth
from otslam_b200.decoder import FrameDecoder

N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
seq = synth.make_sequence("chair_table", 64, subsample=(0, 2))            # 32 distinct frames, repeated
dep, rgb = seq.numpy()
rng = np.random.default_rng(0)
cfiles, dfiles = [], []
for k in range(len(dep)):
    d = dep[k] if k % 2 == 0 else (dep[k].astype(np.int32) + rng.integers(-3, 4, dep[k].shape) * (dep[k] > 0)).clip(0, 65535).astype(np.uint16)
    dfiles.append(cv2.imencode(".png", d)[1].tobytes())
    cfiles.append(cv2.imencode(".jpg", rgb[k][..., ::-1])[1].tobytes())
cf = [cfiles[i % len(cfiles)] for i in range(N)]
df = [dfiles[i % len(dfiles)] for i in range(N)]
dec = FrameDecoder(480, 640, N)
dec.decode_bytes(cf, df)                                                    # warm: buffers, context
t0 = time.perf_counter()
cs, ds = dec.decode_bytes(cf, df)
wall = time.perf_counter() - t0
prof = dec.profile()
print(json.dumps({"frames": N, "wall_ms": 1e3 * wall, "frames_per_s": N / wall, "ok": bool((cs == 0).all() and (ds == 0).all()),
                  "jpeg_bytes": int(np.mean([len(f) for f in cf])), "png_bytes": int(np.mean([len(f) for f in df])), **prof}))
