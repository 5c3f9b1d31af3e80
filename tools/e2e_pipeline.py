import sys, time, os, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from __graft_entry__ import load_package
b = load_package("binding"); wl = load_package("workload")
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
off, a = wl.synth_anchor_batch(n_reads, seed=1)
n = len(a)
b.init(1)
h_a = b.PinnedArray(n, b.ANCHOR); h_a.array[:] = a
pin = {"u": b.PinnedArray(n, np.uint64), "b": b.PinnedArray(n, b.ANCHOR), "n_u": b.PinnedArray(n_reads, np.int32), "n_v": b.PinnedArray(n_reads, np.int32), "status": b.PinnedArray(n_reads, np.int32)}
out = {k: v.array for k, v in pin.items()}
for _ in range(3): res = b.chain_batch(b.Params(), off, h_a.array, out=out)
t0 = time.perf_counter()
for _ in range(5): res = b.chain_batch(b.Params(), off, h_a.array, out=out)
dt = (time.perf_counter() - t0) / 5
print("e2e ms", dt*1e3, "H2D GB/s equiv", 16*n/dt/1e9, res["stats"].as_dict(), file=sys.stderr)
