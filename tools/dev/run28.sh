# final ncu captures of round 1 (each command exits 0 without ncu first: see run26/run27 logs)
ncu --set full --clock-control none -k regex:"knn_mean_dist_kernel|radix_scatter_kernel|radix_hist_kernel|voxel_mean_kernel|os_carry_kernel|os_quantise_kernel|cell_key_kernel|minmax_kernel" -c 16 -o gpurun_out/prof_filters_r01h -f python tools/profile_filters.py > gpurun_out/ncu28a.log 2>&1
ncu --set full --clock-control none -k regex:"alloc_kernel|pack_frames_kernel|order_list_kernel" -s 30 -c 3 -o gpurun_out/prof_pre_r01h -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-post > gpurun_out/ncu28b.log 2>&1
ncu --set full --clock-control none -k regex:"mc_classify_kernel|mc_vertices_kernel|mc_faces_kernel|vertex_normals_kernel|sample_kernel|tri_area_kernel" -c 6 -o gpurun_out/prof_extract_r01h -f python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu28c.log 2>&1
ls -la gpurun_out/*.ncu-rep
