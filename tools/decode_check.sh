#!/bin/bash
# GPU check of the capture-file decoders (csrc/imgcodec.cu): parity tests, the drop-in scripts on top of them, a memcheck
# pass over the small cases and the bench's files_e2e block.
# Usage: /usr/local/graft/bin/gpurun --timeout 420 -- 'bash tools/decode_check.sh <tag>'
tag=${1:-dec}
timeout 240 python -m pytest tests/test_gpu_decode.py tests/test_gpu_scripts.py -x -q 2>&1 | tail -15
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_decode.py -x -q -k "other_sizes or status_codes" > gpurun_out/${tag}_memcheck.log 2>&1
echo "memcheck rc $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/${tag}_memcheck.log | head -8
timeout 240 python bench.py --no-cpu --no-e2e --steps 2 --warmup 1 --hd-frames 0 > gpurun_out/${tag}_files.json 2> gpurun_out/${tag}_files.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${tag}_files.json") if l.startswith("{")][0])
    print(round(d["value"]), json.dumps(d["post_stage"]["files_e2e"]))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${tag}_files.err").read()[-1500:])
PY
