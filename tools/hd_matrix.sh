#!/bin/bash
# dev: HD (1280x720 / 2 mm) per-rank step time of an 8-rank slab run emulated on one GPU: slab thickness, z-split, rank
out=gpurun_out/hd_matrix.txt
: > $out
run() {
  python bench.py --hd --frames 400 --steps 3 --warmup 2 --no-post --no-cpu --no-e2e --hd-frames 0 "$@" > gpurun_out/emhd.json 2> gpurun_out/emhd.err || { echo "$* FAILED" >> $out; return; }
  python - "$*" >> $out <<'PY'
import json, sys
d = json.load(open("gpurun_out/emhd.json")); r = d["roofline"]
print(f"{sys.argv[1]:60s} {d['value']:8.0f} frames/s  step {d['ms_per_step']:7.3f} ms  K4 {r['launch_ms']*1e3:7.1f} us x {r['launches_per_step']:.0f}  share {r['kernel_share_of_step']:.3f}  blocks {d['n_blocks']}")
PY
}
run --emulate-world 1
for z in 1 2 4; do run --emulate-world 8 --zsplit $z; done
for t in 2 4; do for r in 0 5; do run --emulate-world 8 --emulate-rank $r --slab-thickness $t; done; done
run --emulate-world 8 --slab-axis 0 --slab-thickness 1
cat $out
