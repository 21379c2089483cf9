"""Aggregate pinned-host -> device bandwidth with N ranks copying at once (torchrun), with and without binding each rank
to the CPUs next to its GPU: the ceiling of bench.py's e2e number under weak scaling.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe_multi.py
Prints one JSON line on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
results = {}
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

torch.cuda.set_device(local)
if world > 1:
    bench.init_distributed(torch, dist, local)
dev = torch.device("cuda", local)
n = 172 * 1000 * 1000


def measure(tag):
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbs = torch.tensor([n * 20 / (e0.elapsed_time(e1) / 1e3) / 1e9], dtype=torch.float64, device=dev)
    mn = gbs.clone()
    if world > 1:
        dist.all_reduce(gbs, op=dist.ReduceOp.SUM)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    results[tag] = {"aggregate_h2d_GBps": round(float(gbs), 1), "slowest_rank_GBps": round(float(mn), 1)}


measure("unbound")
aff = bench.bind_to_gpu_numa_node(local)
measure("bound_to_gpu_cpus")
if rank == 0:
    try:
        import subprocess
        topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
    except Exception:
        topo = ""
    print(json.dumps({"n_gpus": world, "copy_MB": 172, "cpu_affinity_rank0": aff, "results": results,
                      "topology_cpu_affinity_lines": [l for l in topo.splitlines() if l.startswith("GPU")][:8]}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
