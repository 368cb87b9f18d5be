#!/usr/bin/env python
"""Development helper (multi-GPU box, torchrun): aggregate host->device copy rate when every rank copies from its own
pinned buffer at the same time, with and without binding the rank to the NUMA node of its GPU.  This is the ceiling of
the e2e figure of the pinned-caller entry at N GPUs."""
import os, time, glob
import torch, torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

def gpu_numa(idx):
    try:
        bus = torch.cuda.get_device_properties(idx).pci_bus_id
        dom = torch.cuda.get_device_properties(idx).pci_domain_id
        dev = torch.cuda.get_device_properties(idx).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        return int(open(path).read().strip())
    except Exception as e:
        return -1

def node_cpus(node):
    try:
        s = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        out = []
        for part in s.split(","):
            a, _, b = part.partition("-")
            out += list(range(int(a), int(b or a) + 1))
        return out
    except Exception:
        return []

def probe(tag):
    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); h.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(6): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    g = torch.tensor([6 * n / dt / 1e9], device="cuda", dtype=torch.float64)
    lst = [torch.zeros_like(g) for _ in range(world)]
    dist.all_gather(lst, g)
    if rank == 0:
        v = [float(x) for x in lst]
        print(f"{tag}: per-rank H2D GB/s {[round(x, 1) for x in v]} sum {sum(v):.1f}", flush=True)
    del h, d

nn = len(glob.glob("/sys/devices/system/node/node[0-9]*"))
node = gpu_numa(local)
if rank == 0:
    print(f"cpus {os.cpu_count()} numa nodes {nn}; affinity {len(os.sched_getaffinity(0))} cpus", flush=True)
allinfo = [None] * world
dist.all_gather_object(allinfo, (local, node))
if rank == 0: print("gpu -> numa node:", allinfo, flush=True)
probe("unbound")
cp = node_cpus(node) if node >= 0 else []
if cp:
    os.sched_setaffinity(0, cp)
    probe("bound to the GPU's NUMA node")
dist.destroy_process_group()
