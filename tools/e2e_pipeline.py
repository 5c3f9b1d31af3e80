"""Host-buffer call timing: python tools/e2e_pipeline.py [n_reads] [default|index|round1]; MM2B_TRACE=1 prints the per-sub-batch timeline"""
import sys, time, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
import bench_workloads as BW
b = load_package("binding")
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
variant = sys.argv[2] if len(sys.argv) > 2 else "default"
w = BW.real_seed_batch("map-ont", n_reads, 1000)
off, a = w["off"], w["a"]
n = len(a)
b.init(1)
h_a = b.PinnedArray(n, b.ANCHOR); h_a.array[:] = a
pin = {"u": b.PinnedArray(n, np.uint64), "n_u": b.PinnedArray(n_reads, np.int32), "n_v": b.PinnedArray(n_reads, np.int32), "status": b.PinnedArray(n_reads, np.int32)}
mode, flags = {"default": ("default", 0), "index": ("index", 0), "round1": ("b", b.F_RAW_INPUT | b.F_DEVICE_GATHER)}[variant]
pin["bi" if mode == "index" else "b"] = b.PinnedArray(n, np.int32 if mode == "index" else b.ANCHOR)
out = {k: v.array for k, v in pin.items()}
for _ in range(3): res = b.chain_batch(b.Params(), off, h_a.array, out=out, mode=mode, flags=flags)
t0 = time.perf_counter()
for _ in range(5): res = b.chain_batch(b.Params(), off, h_a.array, out=out, mode=mode, flags=flags)
dt = (time.perf_counter() - t0) / 5
print(variant, "e2e ms %.3f" % (dt * 1e3), "anchors/s %.3g" % (n / dt), res["stats"].as_dict(), file=sys.stderr)
