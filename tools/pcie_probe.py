"""Times pinned host<->device copies of the e2e step's sizes (torch, CUDA events), with the allocating thread bound to
(a) wherever it happens to run, (b) the GPU's NUMA-local CPUs, (c) the CPUs of the other nodes."""
import glob
import os
import subprocess

import torch

dev = torch.device("cuda:0")
torch.cuda.init()
bdf = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", "0"], capture_output=True, text=True).stdout.strip().lower()
bdf = bdf[-12:] if len(bdf) > 12 else bdf
path = "/sys/bus/pci/devices/%s" % bdf
print("gpu", bdf, "numa_node", open(path + "/numa_node").read().strip() if os.path.exists(path) else "?", "local_cpulist",
      open(path + "/local_cpulist").read().strip() if os.path.exists(path) else "?")
print("nodes:", [(os.path.basename(n), open(n + "/cpulist").read().strip()) for n in sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))])
print("affinity now:", len(os.sched_getaffinity(0)), "cpus; running on", os.sched_getcpu() if hasattr(os, "sched_getcpu") else "?")


def parse(cl):
    out = set()
    for part in cl.split(","):
        if "-" in part:
            a, b = part.split("-")
            out |= set(range(int(a), int(b) + 1))
        elif part.strip():
            out.add(int(part))
    return out


def run(tag):
    for mb in (30.9, 12.9):
        n = int(mb * 1e6 / 8)
        h = torch.empty(n, dtype=torch.float64).pin_memory()
        h.zero_()
        d = torch.empty(n, dtype=torch.float64, device=dev)
        for direction in ("h2d", "d2h"):
            for _ in range(3):
                (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print("%-10s %s %6.2f MB  %.3f ms  %.1f GB/s" % (tag, direction, mb, ms, mb / ms))


all_cpus = os.sched_getaffinity(0)
run("default")
if os.path.exists(path):
    local = parse(open(path + "/local_cpulist").read().strip()) & all_cpus
    if local:
        os.sched_setaffinity(0, local)
        run("gpu-local")
    other = all_cpus - local
    if other:
        os.sched_setaffinity(0, other)
        run("remote")
