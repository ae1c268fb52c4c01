"""dev/pcie_probe_ranks.py -- pinned H2D + D2H bandwidth with ALL ranks copying at the same time (one process per GPU,
torchrun): the host-side bound of bench.py's e2e figure at N GPUs (e2e scales 1.5x from 1 to 8 GPUs).
Prints per-rank and aggregate GB/s, alone (rank 0 only) and all together, plus each rank's CPU affinity / NUMA node."""
import os
import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
n = 1 << 27  # 512 MiB of float32 per direction per rank
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_a = torch.empty(n, dtype=torch.float32, device=dev)
d_b = torch.empty(n, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
gb = n * 4 / 1e9


def both():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def timed(active, reps=4):
    if active:
        both()
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    if active:
        for _ in range(reps):
            both()
    torch.cuda.synchronize()
    b.record()
    torch.cuda.synchronize()
    dist.barrier()
    return a.elapsed_time(b) / reps / 1e3


t_alone = timed(rank == 0)
t_all = timed(True)
try:
    aff = sorted(os.sched_getaffinity(0))
    aff_s = f"{aff[0]}-{aff[-1]} ({len(aff)} cpus)"
except Exception:
    aff_s = "?"
numa = "?"
try:
    bus = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
except Exception:
    bus = None
rows = [None] * world
dist.all_gather_object(rows, (rank, gb / t_all, aff_s))
if rank == 0:
    print(f"PCIE rank 0 alone: {gb / t_alone:.1f} GB/s each direction (H2D and D2H concurrently)")
    tot = 0.0
    for r, bw, a in rows:
        tot += bw
        print(f"PCIE all {world} ranks together: rank {r}: {bw:.1f} GB/s each direction, cpu affinity {a}")
    print(f"PCIE aggregate with {world} ranks: {tot:.1f} GB/s each direction = {tot / (gb / t_alone):.2f} x one rank alone")
dist.barrier()
dist.destroy_process_group()
