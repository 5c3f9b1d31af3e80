"""Host-side ceilings of this box: memcpy bandwidth by thread count, pinned H2D / D2H rates, single-thread pack rate.  python tools/host_probe.py"""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
import torch
b = load_package("binding")
L = b.load()
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    print(open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0])
except Exception:
    pass
for t in (1, 2, 4, 8, 12, 16):
    print("host memcpy, %2d threads: %.1f GB/s (read+write)" % (t, L.mm2b_measure_host_copy(t, 256 << 20)), flush=True)
n = 512 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4): fn()
    torch.cuda.synchronize()
    print("%s pinned: %.1f GB/s" % (name, 4 * n / (time.perf_counter() - t0) / 1e9), flush=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(4):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("h2d + d2h together: %.1f GB/s each direction" % (4 * n / dt / 1e9), flush=True)
