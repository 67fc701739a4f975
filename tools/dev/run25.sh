python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for f in 0 0.01; do echo "outliers $f: $(OUTLIER_FRAC=$f python tools/profile_filters.py 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('sor device_ms', round(d['remove_statistical_outlier']['device_ms'],2), 'wall', round(d['remove_statistical_outlier']['wall_ms'],2),'vds', round(d['voxel_down_sample']['device_ms'],2), 'wall', round(d['voxel_down_sample']['wall_ms'],2))")"; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_filters_r01g.csv python tools/profile_filters.py > gpurun_out/ncu25.log 2>&1
grep knn gpurun_out/launches_filters_r01g.csv | tail -1
