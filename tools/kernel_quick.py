"""Kernel-only timing of one preset (CUDA events): python tools/kernel_quick.py [preset] [n_reads] [real|model] [index|b]"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
import bench_workloads as BW
import torch
b = load_package("binding"); wl = load_package("workload")
L = b.load()
name = sys.argv[1] if len(sys.argv) > 1 else "map-ont"
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
source = sys.argv[3] if len(sys.argv) > 3 else "real"
index_out = (sys.argv[4] if len(sys.argv) > 4 else "index") == "index"
if source == "real":
    w = BW.real_seed_batch(name, n_reads, 1000)
    off, a = w["off"], w["a"]
    par = b.Params(*[int(x) for x in w["par"][:9]], float(w["par"][9]))
else:
    off, a = wl.preset_batch(name, n_reads, seed=1)
    par = b.Params()
db = b.DeviceBatch(par, off, a, index_out=index_out)
db.set_counting(True); db.run(); st = db.stats(); db.set_counting(False)
for i in range(3): db.run()
torch.cuda.synchronize()
ks = []
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(5): db.run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
for i in range(3):
    db.run(); ks.append(db.chain_kernel_ms())
print("%s/%s reads %d anchors %d | ms/batch %.3f K1 %.3f | reads/s %.3g anchors/s %.3g GCUPS %.1f | lanes/cell %.2f | heavy %d" % (name, source, len(off) - 1, len(a), ms, sum(ks) / 3, (len(off) - 1) / ms * 1e3, len(a) / ms * 1e3, st.cells_ref / ms / 1e6, st.cells_issued / max(st.cells_ref, 1), st.n_heavy_reads))
