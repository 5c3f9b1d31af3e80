import sys, time, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from __graft_entry__ import load_package
import torch
b = load_package("binding"); wl = load_package("workload")
L = b.load()
name = sys.argv[1] if len(sys.argv) > 1 else "map-ont"
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
off, a = wl.preset_batch(name, n_reads, seed=1)
db = b.DeviceBatch(b.Params(), off, a)
db.set_counting(True); db.run(); st = db.stats(); db.set_counting(False)
for i in range(3): db.run()
torch.cuda.synchronize()
ks = []
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(5): db.run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/5
for i in range(3):
    db.run(); ks.append(db.chain_kernel_ms())
print("%s reads %d anchors %d | ms/batch %.3f K1 %.3f | reads/s %.3g anchors/s %.3g GCUPS %.1f | lanes/cell %.2f" % (name, n_reads, len(a), ms, sum(ks)/3, n_reads/ms*1e3, len(a)/ms*1e3, st.cells_ref/ms/1e6, st.cells_issued/max(st.cells_ref,1)))
