"""Aggregate pinned host<->device bandwidth of the box with 1, 2, 4 ... ranks copying at the same time (VERDICT r1 item 1a:
the ceiling the end-to-end numbers have to be read against).  Run under torchrun, one rank per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_concurrent.py
Rank 0 prints one JSON line per (active ranks, direction)."""
import json, os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
MB = int(os.environ.get("PCIE_MB", "512"))
n = MB << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda")
d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
REPS = 6


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def allmax(x):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run(kind):
    for _ in range(REPS):
        if kind in ("d2h", "both"):
            with torch.cuda.stream(s1):
                h1.copy_(d1, non_blocking=True)
        if kind in ("h2d", "both"):
            with torch.cuda.stream(s2):
                d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize()


ks = [k for k in (1, 2, 4, 8) if k <= world]
for kind in ("d2h", "h2d", "both"):
    for k in ks:
        if rank < k:
            run(kind)   # warm
        barrier()
        t0 = time.perf_counter()
        if rank < k:
            run(kind)
        dt = time.perf_counter() - t0 if rank < k else 0.0
        barrier()
        dt = allmax(dt)
        if rank == 0:
            per_dir = k * REPS * n / dt / 1e9
            print(json.dumps({"active_ranks": k, "of": world, "kind": kind, "GBps_per_direction_aggregate": round(per_dir, 1),
                              "GBps_per_direction_per_rank": round(per_dir / k, 1), "MB_per_copy": MB}), flush=True)
if world > 1:
    dist.destroy_process_group()
