"""GPU tests of the heavy-read kernel (one CTA per read, scans shared by its warps; csrc/chain_kernels.cu "Heavy reads").

By default only reads with long windows AND many window cells are routed to it (the tandem-repeat fixture qualifies).  With
MM2B_HEAVY_MIN_CELLS=0 every read with a mean window above 128 anchors goes that way, which turns the adversarial generators
into tests of the cooperative scan; MM2B_HEAVY=0 switches the kernel off.  The routing is fixed when the backend comes up,
so each setting runs in its own process.  Also through the range-checking build.
"""
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
b.init(1)
bad, heavy, reads = 0, 0, 0
def check(kw, off, a):
    global bad, heavy, reads
    ref = O.replay(O.Params(**kw), off, a, n_threads=8)
    res = b.chain_batch(b.Params(**kw), off, a)
    ok = np.array_equal(res["n_u"], ref["n_u"]) and np.array_equal(res["n_v"], ref["n_v"].astype(np.int32))
    for r in range(len(off) - 1):
        o, nu, nv = int(off[r]), int(ref["n_u"][r]), int(ref["n_v"][r])
        ok = ok and np.array_equal(res["u"][res["u_off"][r]:res["u_off"][r] + nu], ref["u"][o:o + nu]) \
                and np.array_equal(res["b"][res["b_off"][r]:res["b_off"][r] + nv], ref["b"][o:o + nv])
    bad += int(not ok)
    heavy += res["stats"].n_heavy_reads
    reads += len(off) - 1
rng = np.random.default_rng(5)
deep = fuzz.batch([fuzz.dense_repeat(rng, 6000, width=4500, qwidth=4000), fuzz.dense_repeat(rng, 3000, width=800, qwidth=6000),
                   fuzz.dense_repeat(rng, 9000, width=3000, qwidth=3000, span_jitter=True), fuzz.lattice(rng, 5000), fuzz.collinear(rng, 20000, 500),
                   fuzz.many_chains(rng, 400, 4)] + [fuzz.dense_repeat(rng, n, width=max(4, n // 2), qwidth=max(4, n // 2)) for n in (63, 64, 65, 129, 161, 257, 1025)])
for kw in ({}, dict(max_iter=300, max_skip=2), dict(min_cnt=1, min_sc=1), dict(max_skip=0), dict(max_skip=200), dict(max_iter=8000), dict(max_iter=129),
           dict(max_dist_x=60000, max_dist_y=60000, bw=40000, max_skip=5)):
    check(kw, *deep)
for seed, kw in ((1, {}), (2, dict(min_cnt=1, min_sc=5)), (3, dict(n_segs=2, max_dist_x=800, max_dist_y=600, bw=100)), (4, dict(max_iter=50, max_skip=3)),
                 (5, dict(gap_scale=1.7, bw=2000))):
    check(kw, *fuzz.mixed_batch(seed, n_reads=48, seg_ids=kw.get("n_segs", 1)))
check({}, *fuzz.edge_batch(11))
recs = dumpio.read_dump(%(root)r + "/tests/golden/tandem.dump.gz")
off, a = dumpio.to_batch(recs)
res = b.chain_batch(b.Params(**recs[0]["par"].as_dict()), off, a)
tandem_heavy = res["stats"].n_heavy_reads
for r, rec in enumerate(recs):
    bad += int(not (np.array_equal(res["u"][res["u_off"][r]:res["u_off"][r] + res["n_u"][r]], rec["u"])
                    and np.array_equal(res["b"][res["b_off"][r]:res["b_off"][r] + res["n_v"][r]], rec["b"])))
flags = b.load().mm2b_debug_flags()
b.shutdown()
print("RESULT mismatches=%%d heavy=%%d of %%d tandem_heavy=%%d flags=0x%%x" %% (bad, heavy, reads, tandem_heavy, flags & 0x7fffffff))
'''


def _run(env_extra, lib=None):
    env = dict(os.environ, **env_extra)
    if lib:
        env["MM2B_LIB"] = lib
    out = subprocess.run([sys.executable, "-c", SCRIPT % dict(root=ROOT)], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("RESULT")][-1]
    return dict(kv.split("=") for kv in line.split()[1:] if "=" in kv), line


def test_default_routing_takes_the_tandem_reads_only():
    r, line = _run({})
    assert r["mismatches"] == "0" and r["tandem_heavy"] == "3", line


def test_every_long_window_read_through_the_cooperative_scan():
    r, line = _run({"MM2B_HEAVY_MIN_CELLS": "0"})
    assert r["mismatches"] == "0" and int(r["heavy"]) >= 40 and r["tandem_heavy"] == "3", line


def test_switched_off():
    r, line = _run({"MM2B_HEAVY": "0"})
    assert r["mismatches"] == "0" and r["heavy"] == "0" and r["tandem_heavy"] == "0", line


def test_cooperative_scan_in_the_range_checking_build(pkg):
    bld = pkg("build")
    bld.build_all()
    r, line = _run({"MM2B_HEAVY_MIN_CELLS": "0"}, lib=bld.LIB_DBG)
    assert r["mismatches"] == "0" and int(r["heavy"]) >= 40 and r["flags"] == "0x0", line
