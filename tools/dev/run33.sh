N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench33_n$N.json 2> gpurun_out/bench33_n$N.err
python - <<PY
import json
for f in ["bench33_n$N.json"]:
    for l in open("gpurun_out/"+f):
        if l.startswith("{"):
            d=json.loads(l); r=d["roofline"]; print(f, round(d["value"]), round(d["ms_per_step"],3), d["e2e"] and round(d["e2e"]["value"]), round(r["frac"],3), round(r["kernel_share_of_step"],3), d.get("halo_exchange_ms"), d.get("halo_planes_received_rank0"), d.get("extract_gather_ms"), d.get("gathered_points"), d["config"]["parallelism"])
PY
grep -i "error\|Traceback" gpurun_out/bench33_n$N.err | head -5
