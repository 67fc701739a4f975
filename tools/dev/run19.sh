python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/t19.log
cat gpurun_out/t19.log
python tools/profile_filters.py > gpurun_out/filters_1M.json 2>&1; cat gpurun_out/filters_1M.json
python tools/profile_filters.py > gpurun_out/filters_1M_b.json 2>&1; cat gpurun_out/filters_1M_b.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu19a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:integrate_kernel -s 40 -c 1 -o gpurun_out/prof_integrate_r01d -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-post > gpurun_out/ncu19b.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_filters_r01d.csv python tools/profile_filters.py > gpurun_out/ncu19c.log 2>&1
ncu --set full --clock-control none -k regex:"knn_mean_dist_kernel|radix_scatter_kernel|voxel_mean_kernel|cell_key_kernel|radix_hist_kernel" -c 14 -o gpurun_out/prof_filters_r01d -f python tools/profile_filters.py > gpurun_out/ncu19d.log 2>&1
ls -la gpurun_out/
