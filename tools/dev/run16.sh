python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/t16.log
python bench.py --no-cpu --no-e2e --no-post --steps 3 > gpurun_out/bench16.json 2>&1
python bench.py --no-cpu --no-e2e --no-post --emulate-world 8 --steps 3 > gpurun_out/bench16_e8.json 2>&1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:alloc_kernel -s 4 -c 2 --csv --log-file gpurun_out/ncu16_alloc.csv python bench.py --no-cpu --no-e2e --no-post --emulate-world 8 --steps 1 --warmup 1 > gpurun_out/ncu16.log 2>&1
cat gpurun_out/t16.log; tail -3 gpurun_out/ncu16_alloc.csv
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench16*.json")):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); r=d["roofline"]; print(f, round(d["value"]), round(d["ms_per_step"],3), round(r["frac"],3), round(r["kernel_share_of_step"],3), r["other_kernels_ms_per_step"]["pack"], r["other_kernels_ms_per_step"]["alloc"])
PY
