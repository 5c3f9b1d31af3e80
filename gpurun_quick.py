import sys, time, numpy as np
sys.path.insert(0, '.')
from __graft_entry__ import load_package
import torch
b = load_package("binding"); wl = load_package("workload")
L = b.load()
print("devices", L.mm2b_cuda_device_count(), torch.cuda.get_device_name(0))
print("int32 peak Gops", L.mm2b_measure_int32_peak(0))
t0=time.time(); off,a = wl.synth_anchor_batch(20000, seed=1); print("gen", time.time()-t0, len(a))
db = b.DeviceBatch(b.Params(), off, a)
for i in range(3): db.run()
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(5): db.run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/5
st = db.stats()
print("ms/batch", ms, "reads/s", 20000/ms*1e3, "anchors/s", len(a)/ms*1e3, st.as_dict())
