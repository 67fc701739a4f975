for f in 0 0.001 0.01 0.05; do echo "outliers $f: $(OUTLIER_FRAC=$f python tools/profile_filters.py 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('sor device_ms', round(d['remove_statistical_outlier']['device_ms'],2), 'vds', round(d['voxel_down_sample']['device_ms'],2))")"; done
