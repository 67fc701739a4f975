for occ in 4 8 16 32 64; do for r in 1 2 3; do
echo "occ $occ rings $r: $(OTSLAM_KNN_DEBUG=1 OTSLAM_KNN_OCC=$occ OTSLAM_KNN_RINGS=$r python tools/profile_filters.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('knn:'): print(l.strip(), end=' | ')
    if l.startswith('{'): d=json.loads(l); print('sor device_ms', round(d['remove_statistical_outlier']['device_ms'],2), 'kept', d['remove_statistical_outlier']['out'])
")"
done; done
echo "rings 0: $(OTSLAM_KNN_RINGS=0 python tools/profile_filters.py 2>&1 | tail -1 | cut -c1-400)"
