python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/t18.log
cat gpurun_out/t18.log
python bench.py > gpurun_out/bench18.json 2> gpurun_out/bench18.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench18.json") if l.startswith("{")][0])
print(round(d["value"]), d["ms_per_step"], d["e2e"]["value"], d["cpu_baseline"])
for k,v in d["post_stage"].items():
    print(k, v if not isinstance(v,dict) else {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu18a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:integrate_kernel -s 40 -c 1 -o gpurun_out/prof_integrate_r01d -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-post > gpurun_out/ncu18b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"knn_mean_dist_kernel|radix_scatter_kernel|radix_hist_kernel|voxel_mean_kernel|cell_key_kernel|zfilter_kernel|mc_classify_kernel|mc_vertices_kernel|sample_kernel|os_carry_kernel" -c 60 -o gpurun_out/prof_filters_r01d -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu18c.log 2>&1
ls -la gpurun_out/*.ncu-rep
