import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from __graft_entry__ import load_package
import fuzz
from oracle import oracle_py as O, dumpio
b = load_package("binding")
b.init(1)
bad = 0
for seed, kw in ((1, {}), (2, dict(min_cnt=1, min_sc=5)), (3, dict(n_segs=2, max_dist_x=800, max_dist_y=600, bw=100)), (4, dict(max_iter=50, max_skip=3))):
    off, a = fuzz.mixed_batch(seed, n_reads=24, seg_ids=kw.get("n_segs", 1), scale=0.5)
    ref = O.replay(O.Params(**kw), off, a, n_threads=2)
    res = b.chain_batch(b.Params(**kw), off, a)
    bad += int(not np.array_equal(res["n_u"], ref["n_u"]))
rng = np.random.default_rng(5)
off, a = fuzz.batch([fuzz.dense_repeat(rng, 1500, width=4500, qwidth=4000), fuzz.many_chains(rng, 120, 4), fuzz.collinear(rng, 3000, 100)])
for kw in ({}, dict(min_cnt=1, min_sc=1)):
    ref = O.replay(O.Params(**kw), off, a, n_threads=2)
    res = b.chain_batch(b.Params(**kw), off, a)
    bad += int(not np.array_equal(res["n_u"], ref["n_u"]))
r = dumpio.read_dump("tests/golden/mt_map-ont.dump.gz")[0]
u, bb, _, _ = b.chain_read(b.Params(), r["a"])
bad += int(not np.array_equal(u, r["u"]))
b.shutdown()
print("sanitizer workload done, mismatches:", bad)
