"""dev: aggregate pinned host -> device copy bandwidth of the box with N ranks copying at the same time (torchrun).  The `e2e`
numbers of bench.py at N > 1 cannot exceed this: every frame has to cross PCIe once."""
import os
import time

import torch
import torch.distributed as dist

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
reps = 8
for _ in range(reps):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
t = torch.tensor([dt], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"ranks {world}: {reps * n * world / float(t.item()) / 1e9:.1f} GB/s aggregate pinned H2D ({reps * n / dt / 1e9:.1f} GB/s on rank 0)")
if world > 1:
    dist.destroy_process_group()
