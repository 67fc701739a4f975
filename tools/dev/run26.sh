K='integrate_kernel|alloc_kernel|pack_frames|order_list|mult_table|stats_kernel|mc_|scan_|pc_extract|face_normals|corner_fill|vertex_normals|tri_area|os_|ordered_accumulate|sample_kernel|zfilter|minmax|cell_key|radix_|seg_heads|voxel_mean|cell_hash|crowding|knn_|sor_select|clear_masks|scatter_base|rehash|export_kernel|merge_pack|grid_points|backproject|depth_convert|halo_'
python __graft_entry__.py --smoke 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$K" -c 2000 --csv --log-file gpurun_out/launches_r01e.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu26a.log 2>&1
grep -c integrate_kernel gpurun_out/launches_r01e.csv
python bench.py --steps 10 > gpurun_out/bench26.json 2> gpurun_out/bench26.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench26_ref.json 2> gpurun_out/bench26_ref.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench26.json') if l.startswith('{')][0])
print(round(d['value']), d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['clocks'], d['gpu_launches'])
d=json.loads([l for l in open('gpurun_out/bench26_ref.json') if l.startswith('{')][0]); print(d['value'], d['ms_per_step'])
"
