"""Aggregate pinned host -> device copy ceiling of the node: every rank copies a pinned buffer to its own GPU
with plain cudaMemcpyAsync (torch copy_, non_blocking) in a loop, all ranks at once; rank 0 prints the per-GPU and aggregate GB/s.
Run under torchrun with the N the bench will use:  python -m torch.distributed.run --nproc-per-node N scripts/h2d_probe.py [--bind]
`--bind` pins each rank to its GPU's NUMA node first (bench.bind_to_gpu_numa), which is what bench.py does.
Appends {"N": {...}} to profiles/r02_h2d_ceiling.json when --save is given."""
import json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
placement = None
if "--bind" in sys.argv:
    import bench
    placement = bench.bind_to_gpu_numa(local, world)
nbytes = 2 << 30
host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
host.fill_(1)
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
for _ in range(2):
    dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
reps = 8
t0 = time.perf_counter()
for _ in range(reps):
    dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
gbps = nbytes * reps / dt / 1e9
if world > 1:
    t = torch.tensor([gbps, dt], dtype=torch.float64, device="cuda")
    allt = torch.zeros(world, 2, dtype=torch.float64, device="cuda")
    dist.all_gather_into_tensor(allt, t)
    per = [float(v) for v in allt[:, 0].tolist()]; dmax = float(allt[:, 1].max())
else:
    per = [gbps]; dmax = dt
if rank == 0:
    out = dict(n_gpus=world, per_gpu_gbps_each=per, per_gpu_gbps=min(per), aggregate_gbps=nbytes * reps * world / dmax / 1e9, bytes_per_copy=nbytes, reps=reps,
               bound_to_numa="--bind" in sys.argv, placement_rank0=placement,
               how="pinned 2 GiB host buffer per rank -> its own GPU, cudaMemcpyAsync in a loop, all ranks concurrently, wall clock around the loop")
    print(json.dumps(out))
    if "--save" in sys.argv:
        p = ROOT / "profiles" / "r02_h2d_ceiling.json"
        d = json.loads(p.read_text()) if p.exists() else {}
        d[str(world)] = out
        p.write_text(json.dumps(d, indent=1))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
