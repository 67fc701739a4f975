#!/bin/bash
# dev: per-rank step time of an N-rank slab run, emulated on ONE GPU (rank 0's slabs), for several z-split settings.
# usage: tools/emulate_matrix.sh [extra bench args]   -> gpurun_out/emulate_matrix.txt
out=gpurun_out/emulate_matrix.txt
: > $out
for w in 1 2 4 8; do
  for z in ${ZS_LIST:-0 2 4}; do
    if [ $w = 1 ] && [ $z = 8 ]; then continue; fi
    python bench.py --steps 5 --warmup 3 --no-post --no-cpu --no-e2e --hd-frames 0 --emulate-world $w --zsplit $z "$@" > gpurun_out/em.json 2> gpurun_out/em.err || { echo "w=$w z=$z FAILED" >> $out; tail -3 gpurun_out/em.err >> $out; continue; }
    python - "$w" "$z" >> $out <<'PY'
import json, sys
d = json.load(open("gpurun_out/em.json"))
r = d["roofline"]
print(f"world {sys.argv[1]} zsplit {sys.argv[2]}: {d['value']:9.0f} frames/s  step {d['ms_per_step']:7.3f} ms  K4 launch {r['launch_ms']*1e3:7.1f} us  share {r['kernel_share_of_step']:.3f}  pack {r['other_kernels_ms_per_step']['pack']:.3f} alloc {r['other_kernels_ms_per_step']['alloc']:.3f} ms/step  blocks {d['n_blocks']}")
PY
  done
done
cat $out
