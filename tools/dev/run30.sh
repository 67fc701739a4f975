python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/bench30.json 2> gpurun_out/bench30.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench30.json') if l.startswith('{')][0])
print(round(d['value']), d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['clocks'], d['gpu_launches'], d['post_stage']['total_wall_ms_config1_stage'])
"
OTSLAM_PROFILE_ONCE=1 ncu --set full --clock-control none --import-source on -k regex:"knn_mean_dist_kernel" -c 1 -o gpurun_out/prof_knn_r01h -f python tools/profile_filters.py > gpurun_out/ncu30.log 2>&1
ls -la gpurun_out/prof_knn_r01h.ncu-rep
