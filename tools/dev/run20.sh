python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/t20.log
cat gpurun_out/t20.log
python tools/profile_filters.py > gpurun_out/filters_1M_c.json 2>&1; cat gpurun_out/filters_1M_c.json
python tools/profile_filters.py > gpurun_out/filters_1M_d.json 2>&1; cat gpurun_out/filters_1M_d.json
python bench.py --no-cpu --no-e2e --steps 3 > gpurun_out/bench20.json 2> gpurun_out/bench20.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench20.json") if l.startswith("{")][0])
print(round(d["value"]), d["ms_per_step"])
for k,v in d["post_stage"].items():
    print(k, v if not isinstance(v,dict) else {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_filters_r01e.csv python tools/profile_filters.py > gpurun_out/ncu20c.log 2>&1
