"""Repeat one preset batch many times through both entry points and report any run whose output differs from the oracle."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
from oracle import oracle_py as O
b = load_package("binding"); wl = load_package("workload")
name = sys.argv[1] if len(sys.argv) > 1 else "ultralong"
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 600
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
off, a = wl.preset_batch(name, n_reads, seed=4242)
ref = O.replay(O.Params(), off, a, n_threads=16)
b.init(1)

def diff(res, tag):
    bad = []
    for r in range(n_reads):
        o, nu, nv = int(off[r]), int(ref["n_u"][r]), int(ref["n_v"][r])
        if int(res["n_u"][r]) != nu or int(res["n_v"][r]) != nv:
            bad.append((r, "counts", int(res["n_u"][r]), nu, int(res["n_v"][r]), nv)); continue
        uo, bo = int(res["u_off"][r]), int(res["b_off"][r])
        if not np.array_equal(res["u"][uo:uo + nu], ref["u"][o:o + nu]): bad.append((r, "u"))
        elif not np.array_equal(res["b"][bo:bo + nv], ref["b"][o:o + nv]):
            k = int(np.argmax(res["b"][bo:bo + nv] != ref["b"][o:o + nv])); bad.append((r, "b", k, nv))
    if bad: print(tag, "MISMATCH", len(bad), bad[:5], flush=True)
    return len(bad)

tot = 0
db = b.DeviceBatch(b.Params(), off, a)
for it in range(reps):
    db.run(); tot += diff(db.results(), "device run %d" % it)
for it in range(reps):
    tot += diff(b.chain_batch(b.Params(), off, a), "host run %d" % it)
print("done: %d mismatching reads over %d runs" % (tot, 2 * reps))
