python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/t14.log
python bench.py --no-post --cpu-frames 100 > gpurun_out/bench14.json 2> gpurun_out/bench14.err
python bench.py --no-cpu --no-e2e --no-post --emulate-world 8 > gpurun_out/bench14_e8.json 2>&1
python bench.py --no-cpu --no-e2e --no-post --emulate-world 2 > gpurun_out/bench14_e2.json 2>&1
python bench.py --no-cpu --no-e2e --no-post --hd --frames 200 > gpurun_out/bench14_hd.json 2>&1
python bench.py --no-cpu --no-e2e --no-post --hd --frames 200 --emulate-world 8 > gpurun_out/bench14_hd_e8.json 2>&1
cat gpurun_out/t14.log
python - <<PY
import json
for f in ["bench14.json","bench14_e8.json","bench14_e2.json","bench14_hd.json","bench14_hd_e8.json"]:
    for l in open("gpurun_out/"+f):
        if l.startswith("{"):
            d=json.loads(l); r=d["roofline"]; print(f, round(d["value"]), round(d["ms_per_step"],3), d["e2e"] and round(d["e2e"]["value"]), round(r["frac"],3), round(r["kernel_share_of_step"],3), r["other_kernels_ms_per_step"]["pack"], r["other_kernels_ms_per_step"]["alloc"], d["cpu_baseline"] and d["cpu_baseline"]["value"])
PY
