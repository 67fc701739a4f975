#!/bin/bash
# dev: K4 occupancy (via unused dynamic smem) vs step time for emulated rank 0 of W ranks
out=gpurun_out/pad_matrix.txt
: > $out
for w in ${W_LIST:-8 4 1}; do
 for z in ${ZS_LIST:-2 4}; do
  for pad in ${PAD_LIST:-0 24576 40960 61440}; do
    OTSLAM_K4_SMEM_PAD=$pad python bench.py --steps 5 --warmup 3 --no-post --no-cpu --no-e2e --hd-frames 0 --emulate-world $w --zsplit $z "$@" > gpurun_out/em.json 2> gpurun_out/em.err || { echo "w=$w z=$z pad=$pad FAILED" >> $out; continue; }
    python - "$w" "$z" "$pad" >> $out <<'PY'
import json, sys
d = json.load(open("gpurun_out/em.json"))
r = d["roofline"]
print(f"world {sys.argv[1]} zsplit {sys.argv[2]} pad {int(sys.argv[3])//1024:3d}K: {d['value']:9.0f} frames/s  step {d['ms_per_step']:7.3f} ms  K4 launch {r['launch_ms']*1e3:7.1f} us  share {r['kernel_share_of_step']:.3f}  pack {r['other_kernels_ms_per_step']['pack']:.3f} alloc {r['other_kernels_ms_per_step']['alloc']:.3f} ms/step")
PY
  done
 done
done
cat $out
