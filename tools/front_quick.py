"""Seeding front end timing: python tools/front_quick.py [n_reads] [contexts per device] [map-ont|asm20] — one mm2b_map_batch call per step over
the bench's reads; prints wall ms per call and the CUDA-event time of every stage summed over sub-batches."""
import os, sys, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
if len(sys.argv) > 2: os.environ["MM2B_MAP_CTX"] = sys.argv[2]
from __graft_entry__ import load_package
import bench_workloads as BW
b = load_package("binding")
preset = sys.argv[3] if len(sys.argv) > 3 else "map-ont"
w = BW.real_seed_batch(preset, n_reads, 1000)
fi = BW.front_inputs(preset, n_reads, 1000)
b.init(1)
idx = b.Index(fi["index"])
pin = b.PinnedArray(len(fi["seq"]), np.uint8); pin.array[:] = fi["seq"]
par = b.Params(*[int(x) for x in w["par"][:9]], float(w["par"][9]))
for it in range(3):
    res = b.map_batch(idx, None, fi["mid_occ"], par, seq_off=fi["seq_off"], blob=pin.array, collect=False)
steps = 5
t0 = time.perf_counter()
for it in range(steps):
    res = b.map_batch(idx, None, fi["mid_occ"], par, seq_off=fi["seq_off"], blob=pin.array, collect=False)
ms = (time.perf_counter() - t0) * 1e3 / steps
st = res["stats"]
chk = b.map_batch(idx, None, fi["mid_occ"], par, seq_off=fi["seq_off"], blob=pin.array, collect=False)
print("%s: n_a sum %d (recorded %d), chains %d (recorded %d)" % (preset, int(chk["n_a"].sum()), int(w["off"][-1]), int(chk["n_u"].sum()), int(w["ref_n_u"].sum())))
print("reads %d bases %d | %.2f ms per call (%.3g reads/s, %.3g bases/s) | stage ms summed over %d sub-batches: sketch %.2f seed %.2f sort %.2f chain %.2f | minimizers %d anchors %d tie reads %d"
      % (n_reads, len(fi["seq"]), ms, n_reads / ms * 1e3, len(fi["seq"]) / ms * 1e3, st["n_segs"], st["sketch_ms"], st["seed_ms"], st["sort_ms"], st["chain_ms"], st["tot_mini"], st["tot_anchors"], st["n_tie_reads"]))
idx.close(); pin.free(); b.shutdown()
