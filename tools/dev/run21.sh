python -m pytest tests/test_gpu_scripts.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/t21.log
cat gpurun_out/t21.log
ncu --set full --clock-control none --import-source on -k regex:knn_mean_dist_kernel -c 1 -o gpurun_out/prof_knn_r01e -f python tools/profile_filters.py > gpurun_out/ncu21.log 2>&1
ls -la gpurun_out
