#!/bin/bash
# GPU check of the capture-file decoders (csrc/imgcodec.cu): their parity tests and the drop-in scripts on top of them, the
# default bench line, the chunk / look-ahead sweep of the file loop, the decoders' launch list and one ncu --set full capture
# of the inflate kernel.
# Usage: /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/decode_check.sh <tag>'
tag=${1:-dec}
timeout 200 python -m pytest tests/test_gpu_zz_decode.py tests/test_gpu_scripts.py -x -q --tb=short 2>&1 | tail -8
timeout 300 python bench.py > gpurun_out/${tag}_n1.json 2> gpurun_out/${tag}_n1.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${tag}_n1.json") if l.startswith("{")][0])
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), "clocks", d["clocks"])
    print(json.dumps(d["post_stage"]["files_e2e"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${tag}_n1.err").read()[-1500:])
PY
timeout 120 python tools/files_e2e_sweep.py > gpurun_out/${tag}_sweep.json 2> gpurun_out/${tag}_sweep.err; tail -1 gpurun_out/${tag}_sweep.json
timeout 60 python tools/profile_decode.py > gpurun_out/${tag}_decode_plain.log 2>&1; tail -1 gpurun_out/${tag}_decode_plain.log
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches_decode.csv python tools/profile_decode.py > gpurun_out/${tag}_ncul.log 2>&1
timeout 100 ncu --set full --clock-control none --import-source on -k regex:'png_inflate' -s 1 -c 1 -o gpurun_out/${tag}_prof_inflate python tools/profile_decode.py > gpurun_out/${tag}_ncuf.log 2>&1
ls gpurun_out/ | grep ${tag}
