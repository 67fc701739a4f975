N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench29_n$N.json 2> gpurun_out/bench29_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 2 --hd --frames 200 > gpurun_out/bench29_n${N}_hd.json 2> gpurun_out/bench29_n${N}_hd.err
python - <<PY
import json
for f in ["bench29_n$N.json","bench29_n${N}_hd.json"]:
    for l in open("gpurun_out/"+f):
        if l.startswith("{"):
            d=json.loads(l); r=d["roofline"]; print(f, round(d["value"]), round(d["ms_per_step"],3), d["e2e"] and round(d["e2e"]["value"]), round(r["frac"],3), round(r["kernel_share_of_step"],3), d.get("halo_exchange_ms"), d.get("extract_gather_ms"), d.get("gathered_points"))
PY
tail -3 gpurun_out/bench29_n$N.err
