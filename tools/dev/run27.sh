python -m pytest tests/test_gpu_ordered_sum.py tests/test_gpu_extract_cloud.py tests/test_gpu_resident_mesh.py -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu --no-e2e --steps 3 > gpurun_out/bench27.json 2> gpurun_out/bench27.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench27.json") if l.startswith("{")][0])
print(round(d["value"]), d["ms_per_step"])
for k,v in d["post_stage"].items():
    print(k, v if not isinstance(v,dict) else {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
PY
OUTLIER_FRAC=0 python tools/profile_filters.py | tail -1
