"""Host-buffer call under different transfer settings (one process, re-initialising the library per setting):
   python tools/e2e_sweep.py [n_reads] [steps]
Prints ms per call for: pack in-flight limit 0 (raw only), 1, 2, 3, 99 (pack everything) x output {index, b (device gather), b (host gather)}."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
import bench_workloads as BW
b = load_package("binding")
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
w = BW.real_seed_batch("map-ont", n_reads, 1000)
off, a = w["off"], w["a"]
n = len(a)
b.load()
h_a = b.PinnedArray(n, b.ANCHOR); h_a.array[:] = a
pin = {"u": b.PinnedArray(n, np.uint64), "n_u": b.PinnedArray(n_reads, np.int32), "n_v": b.PinnedArray(n_reads, np.int32), "status": b.PinnedArray(n_reads, np.int32),
       "u_off": b.PinnedArray(n_reads + 1, np.int64), "b_off": b.PinnedArray(n_reads + 1, np.int64)}
h_b, h_bi = b.PinnedArray(n, b.ANCHOR), b.PinnedArray(n, np.int32)
settings = [dict(MM2B_PACK_INFLIGHT="0")]
settings += [dict(MM2B_PACK_INFLIGHT=str(k), MM2B_PACK_RING=str(r)) for r in (1, 0) for k in (1, 2, 99)]
settings += [dict(MM2B_PACK_INFLIGHT=str(k), MM2B_PACK_RING="1", MM2B_SUB_ANCHORS=str(4 << 20)) for k in (1, 2, 99)]
settings += [dict(MM2B_PACK_INFLIGHT=str(k), MM2B_PACK_RING="1", MM2B_PACK_CHUNK=str(c << 10)) for k in (2, 99) for c in (16, 32, 256)]
settings += [dict(MM2B_PACK_INFLIGHT="99", MM2B_PACK_RING="1", MM2B_HOST_THREADS=str(t)) for t in (6, 10)]
if os.environ.get("E2E_QUICK"): settings = [dict(MM2B_PACK_INFLIGHT="0"), dict(MM2B_PACK_INFLIGHT="99", MM2B_PACK_RING="1")]     # raw input, everything packed
for env in settings:
    for k, v in env.items():
        os.environ[k] = v
    b.init(1)
    row = []
    for name, mode, flags in (("index", "index", 0), ("b_dev", "b", b.F_DEVICE_GATHER)):
        out = {k: v.array for k, v in pin.items()}
        out["bi" if mode == "index" else "b"] = (h_bi if mode == "index" else h_b).array
        for _ in range(2):
            res = b.chain_batch(b.Params(), off, h_a.array, out=out, mode=mode, flags=flags)
        t0 = time.perf_counter()
        for _ in range(steps):
            res = b.chain_batch(b.Params(), off, h_a.array, out=out, mode=mode, flags=flags)
        ms = (time.perf_counter() - t0) * 1e3 / steps
        st = res["stats"]
        row.append("%s %.2f ms (packed %d raw %d, pack_ms %.0f)" % (name, ms, st.n_packed_subs, st.n_raw_subs, st.pack_ms))
    print(env, " | ".join(row), flush=True)
    b.shutdown()
    for k in env:
        os.environ.pop(k, None)
