"""Facts about the GPU box that size the config-4 bench: host RAM / cores / NUMA, HBM actually allocatable,
how fast pinned host memory can be allocated, and the plain H2D rate from it."""
import os
import subprocess
import time

import torch


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=60).stdout.strip()
    except Exception as e:  # pragma: no cover
        return f"<{e}>"


print("nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
print(sh("free -g | head -3"))
print(sh("lscpu | egrep 'Model name|Socket|NUMA|Thread|Core'"))
print(sh("nvidia-smi --query-gpu=name,memory.total,memory.used,power.limit --format=csv"))
print(sh("nvidia-smi topo -m | head -20"))
print(sh("cat /sys/kernel/mm/transparent_hugepage/enabled; ulimit -l; cat /proc/meminfo | egrep 'HugePages_Total|Hugepagesize|MemAvailable'"))
free_b, total_b = torch.cuda.mem_get_info(0)
print("cuda mem free/total GB", free_b / 1e9, total_b / 1e9)
for gb in (4, 16):
    t0 = time.perf_counter()
    h = torch.empty(gb * (1 << 30), dtype=torch.uint8, pin_memory=True)
    t1 = time.perf_counter()
    d = torch.empty(gb * (1 << 30), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    t5 = time.perf_counter()
    print(f"pinned {gb} GiB: alloc {t1 - t0:.2f} s ({gb * 1.0737 / (t1 - t0):.1f} GB/s), H2D first {gb * 1.0737 / (t3 - t2):.1f} "
          f"GB/s, again {gb * 1.0737 / (t4 - t3):.1f} GB/s, D2H {gb * 1.0737 / (t5 - t4):.1f} GB/s")
    del h, d
    torch.cuda.empty_cache()
# largest single device allocation
lo = 0
for gb in (170, 175, 178, 180, 184, 188):
    try:
        x = torch.empty(gb * 10**9, dtype=torch.uint8, device="cuda")
        del x
        torch.cuda.empty_cache()
        lo = gb
    except Exception:
        break
print("largest device allocation that worked (GB, decimal):", lo)
