for fg in 0 1 2 4 8; do
OTSLAM_ALLOC_FG=$fg python bench.py --no-cpu --no-e2e --no-post --emulate-world 8 --steps 3 > gpurun_out/bench15_e8_fg$fg.json 2>&1
done
OTSLAM_ALLOC_FG=1 python bench.py --no-cpu --no-e2e --no-post --steps 3 > gpurun_out/bench15_fg1.json 2>&1
for fg in 0 1 8; do
OTSLAM_ALLOC_FG=$fg ncu --set full --clock-control none --import-source on -k regex:alloc_kernel -s 4 -c 1 -o gpurun_out/prof_alloc_fg$fg -f python bench.py --no-cpu --no-e2e --no-post --emulate-world 8 --steps 1 --warmup 1 > gpurun_out/ncu15_$fg.log 2>&1
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/bench15_*.json")):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); r=d["roofline"]; print(f, round(d["value"]), round(d["ms_per_step"],3), round(r["frac"],3), round(r["kernel_share_of_step"],3), r["other_kernels_ms_per_step"]["pack"], r["other_kernels_ms_per_step"]["alloc"])
PY
