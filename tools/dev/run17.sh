python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/t17.log
python bench.py --no-cpu --no-e2e --steps 3 > gpurun_out/bench17.json 2> gpurun_out/bench17.err
cat gpurun_out/t17.log
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench17.json") if l.startswith("{")][0])
print(round(d["value"]), d["ms_per_step"])
for k,v in d["post_stage"].items():
    print(k, v if not isinstance(v,dict) else {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
PY
