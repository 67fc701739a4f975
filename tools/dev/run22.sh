python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t22.log
cat gpurun_out/t22.log
python tools/profile_filters.py > gpurun_out/filters_1M_e.json 2>&1; cat gpurun_out/filters_1M_e.json
python tools/profile_filters.py > gpurun_out/filters_1M_f.json 2>&1; cat gpurun_out/filters_1M_f.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_filters_r01f.csv python tools/profile_filters.py > gpurun_out/ncu22c.log 2>&1
grep knn gpurun_out/launches_filters_r01f.csv | tail -1
