python -m pytest tests/test_gpu_extract_cloud.py -m gpu -x -q 2>&1 | tail -4
for ax in 0 3; do for r in 0 1 2 3 4 5 6 7; do
python bench.py --no-cpu --no-e2e --no-post --hd --frames 200 --steps 2 --warmup 1 --emulate-world 8 --emulate-rank $r --slab-axis $ax 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('hd axis $ax rank $r ms/step', round(d['ms_per_step'],3), 'blocks', d['config']['n_blocks'], 'nupd', round(d['roofline']['n_upd_per_frame']/1e6,2))"
done; done
for r in 0 3 5 7; do
python bench.py --no-cpu --no-e2e --no-post --steps 2 --warmup 1 --emulate-world 8 --emulate-rank $r --slab-axis 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('default axis 3 rank $r ms/step', round(d['ms_per_step'],3), 'blocks', d['config']['n_blocks'])"
done
