python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/t13.log
python bench.py > gpurun_out/bench13.json 2> gpurun_out/bench13.err
OTSLAM_NO_LPT=1 python bench.py --no-cpu --no-e2e --no-post > gpurun_out/bench13_nolpt.json 2>&1
python bench.py --no-cpu --no-e2e --no-post --emulate-world 8 > gpurun_out/bench13_e8.json 2>&1
OTSLAM_NO_LPT=1 python bench.py --no-cpu --no-e2e --no-post --emulate-world 8 > gpurun_out/bench13_e8_nolpt.json 2>&1
python bench.py --no-cpu --no-e2e --no-post --frames 300 --scene table > gpurun_out/bench13_c0.json 2>&1
cat gpurun_out/t13.log
python - <<PY
import json
for f in ["bench13.json","bench13_nolpt.json","bench13_e8.json","bench13_e8_nolpt.json","bench13_c0.json"]:
    for l in open("gpurun_out/"+f):
        if l.startswith("{"):
            d=json.loads(l); print(f, round(d["value"]), round(d["ms_per_step"],3), d["e2e"] and round(d["e2e"]["value"]), round(d["roofline"]["frac"],3), round(d["roofline"]["kernel_share_of_step"],3), d["cpu_baseline"] and d["cpu_baseline"]["value"])
PY
