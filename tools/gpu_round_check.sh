#!/bin/bash
# What a round-end check on the GPU box runs (through gpurun): the GPU test-suite, the smoke test and the default
# bench line.  Usage: /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/gpu_round_check.sh'
set -o pipefail
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python __graft_entry__.py --smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/bench_check.json") if l.startswith("{")][0])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "roofline.frac", round(d["roofline"]["frac"], 3),
      "cpu", round(d["cpu_baseline"]["value"], 1), "clocks", d["clocks"], "launches", d["gpu_launches"])
print("post stage wall ms", round(d["post_stage"]["total_wall_ms_config1_stage"], 2), "files_e2e", d["post_stage"].get("files_e2e"))
PY
