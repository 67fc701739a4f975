#!/bin/bash
# quick GPU check of the drop-in scripts' path: script tests + the bench's files_e2e block
python -m pytest tests/test_gpu_scripts.py -x -q 2>&1 | tail -2
python bench.py --no-cpu --no-e2e --steps 2 --warmup 1 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(round(d['value']), d['post_stage']['files_e2e'])"
