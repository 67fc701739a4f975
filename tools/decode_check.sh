#!/bin/bash
# GPU check of the capture-file decoders (csrc/imgcodec.cu): the whole GPU suite (parity of the decoders, the drop-in scripts on
# top of them, everything else), the default bench line, the decoders' launch list and one ncu --set full capture of the two
# entropy-decoding kernels.
# Usage: /usr/local/graft/bin/gpurun --timeout 600 -- 'bash tools/decode_check.sh <tag>'
tag=${1:-dec}
timeout 300 python -m pytest tests -m gpu -x -q --tb=short 2>&1 | tail -12
timeout 60 python __graft_entry__.py --smoke 2>&1 | tail -1
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
timeout 60 python tools/profile_decode.py > gpurun_out/${tag}_decode_plain.log 2>&1; tail -1 gpurun_out/${tag}_decode_plain.log
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches_decode.csv python tools/profile_decode.py > gpurun_out/${tag}_ncul.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:'png_inflate|jpeg_huff' -s 2 -c 2 -o gpurun_out/${tag}_prof_decode python tools/profile_decode.py > gpurun_out/${tag}_ncuf.log 2>&1
ls -la gpurun_out/ | tail -5
