"""GPU test: the range-checking twin of the library (kernels compiled with -DMM2B_DEBUG_CHECKS; every index into the per-read
scratch is checked on the device) runs adversarial and golden inputs with zero violations and bit-exact results.
compute-sanitizer is closed on this GPU pool, so this is the memory-safety check.  Runs in a subprocess because the
library path is fixed at first load."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

SCRIPT = r'''
import sys, numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
from __graft_entry__ import load_package
import fuzz
from oracle import oracle_py as O, dumpio
b = load_package("binding")
L = b.load()
assert L.mm2b_debug_flags() >> 31 == 1, "not the checking build"
b.init(1)
bad = 0
def check(kw, off, a):
    global bad
    ref = O.replay(O.Params(**kw), off, a, n_threads=8)
    for cnt in (False, True):
        b.set_counting(cnt)
        res = b.chain_batch(b.Params(**kw), off, a)
        ok = np.array_equal(res["n_u"], ref["n_u"]) and np.array_equal(res["n_v"], ref["n_v"].astype(np.int32))
        for r in range(len(off) - 1):
            o, nu, nv = int(off[r]), int(ref["n_u"][r]), int(ref["n_v"][r])
            ok = ok and np.array_equal(res["u"][res["u_off"][r]:res["u_off"][r] + nu], ref["u"][o:o + nu]) \
                    and np.array_equal(res["b"][res["b_off"][r]:res["b_off"][r] + nv], ref["b"][o:o + nv])
        bad += int(not ok)
    b.set_counting(False)
for seed, kw in ((1, {}), (2, dict(min_cnt=1, min_sc=5)), (3, dict(n_segs=2, max_dist_x=800, max_dist_y=600, bw=100)), (4, dict(max_iter=50, max_skip=3)),
                 (5, dict(is_cdna=1, max_dist_x=200000, max_dist_y=2000, bw=200000)), (6, dict(gap_scale=1.7, bw=2000)), (7, dict(min_cnt=0, min_sc=-3))):
    off, a = fuzz.mixed_batch(seed, n_reads=48, seg_ids=kw.get("n_segs", 1))
    check(kw, off, a)
rng = np.random.default_rng(5)
off, a = fuzz.batch([fuzz.dense_repeat(rng, 6000, width=4500, qwidth=4000), fuzz.many_chains(rng, 400, 4), fuzz.collinear(rng, 20000, 500),
                     fuzz.lattice(rng, 3000)] + [fuzz.dense_repeat(rng, n, width=max(4, n // 2), qwidth=max(4, n // 2)) for n in (1, 31, 32, 33, 255, 256, 257, 289)])
for kw in ({}, dict(min_cnt=1, min_sc=1), dict(max_iter=300, max_skip=2)):
    check(kw, off, a)
off, a = fuzz.edge_batch(11)
for kw in ({}, dict(max_dist_x=60000, max_dist_y=60000, bw=40000, max_skip=5), dict(max_skip=200)):
    check(kw, off, a)
for name in ("tandem_iter64", "syn_ccs", "sr_paired"):
    recs = dumpio.read_dump(%(root)r + "/tests/golden/" + name + ".dump.gz")
    groups = {}
    for r in recs: groups.setdefault(tuple(sorted(r["par"].as_dict().items())), []).append(r)
    for g in list(groups.values())[:6]:
        off, a = dumpio.to_batch(g)
        check(g[0]["par"].as_dict(), off, a)
flags = L.mm2b_debug_flags()
b.shutdown()
print("RESULT mismatches=%%d flags=0x%%x" %% (bad, flags & 0x7fffffff))
'''


def test_debug_build_sees_no_out_of_range_access(pkg):
    bld = pkg("build")
    bld.build_all()
    assert os.path.exists(bld.LIB_DBG)
    env = dict(os.environ, MM2B_LIB=bld.LIB_DBG)
    out = subprocess.run([sys.executable, "-c", SCRIPT % dict(root=ROOT)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("RESULT")][-1]
    assert line == "RESULT mismatches=0 flags=0x0", line
