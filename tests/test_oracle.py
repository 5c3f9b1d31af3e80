"""CPU tests: pin the oracle (oracle/chain_oracle.c) against the reference.

1. against tests/golden/ — captures of the reference CLI's own mm_chain_dp calls (f/p/v per anchor,
   u[] and b[] per read), made by tests/golden/make_golden.py;
2. against the reference's compiled chain.c (oracle/_ref/libmm2ref.so) on seeded adversarial inputs —
   skipped on machines without the in-place reference build;
3. the restated unstable radix sort against the reference's radix_sort_128x on tie-heavy keys.
"""
import numpy as np
import pytest

import fuzz
from conftest import golden_names, load_golden


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_capture(oracle, name):
    recs = load_golden(name)
    assert recs, name
    for k, r in enumerate(recs):
        o = oracle.chain(r["par"], r["a"], want_fpv=True)
        assert np.array_equal(o["f"], r["f"]), (name, k, "f")
        assert np.array_equal(o["p"], r["p"]), (name, k, "p")
        assert np.array_equal(o["v"], r["v"]), (name, k, "v")
        assert np.array_equal(o["u"], r["u"]), (name, k, "u")
        assert np.array_equal(o["b"], r["b"]), (name, k, "b")
        assert (o["status"] == 2) == (not r["u_null"]), (name, k, "null-ness")


def test_golden_stats_match_survey(oracle):
    """MT-human vs MT-orang map-ont is one call, n=346, 8,970 cells, one chain of 342."""
    r = load_golden("mt_map-ont")[0]
    o = oracle.chain(r["par"], r["a"])
    assert len(r["a"]) == 346 and o["stats"].cells == 8970 and o["stats"].window_cells == 30829
    assert len(o["u"]) == 1 and int(o["u"][0]) & 0xffffffff == 342
    r = load_golden("mt_asm20")[0]
    o = oracle.chain(r["par"], r["a"])
    assert len(r["a"]) == 229 and o["stats"].cells == 5805


PARAM_SETS = [
    dict(),
    dict(min_cnt=1, min_sc=5),
    dict(max_iter=50, max_skip=3),
    dict(max_skip=0),
    dict(gap_scale=1.7, bw=2000),
    dict(is_cdna=1, max_dist_x=200000, max_dist_y=2000, bw=200000),
    dict(n_segs=2, max_dist_x=800, max_dist_y=600, bw=100, min_cnt=2, min_sc=25),
    dict(n_segs=3, is_cdna=1),
]


@pytest.mark.parametrize("pi", range(len(PARAM_SETS)))
def test_oracle_matches_compiled_reference_fuzz(oracle, pi):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libmm2ref.so not built (needs /root/reference)")
    kw = PARAM_SETS[pi]
    par = oracle.Params(**kw)
    off, a = fuzz.mixed_batch(100 + pi, n_reads=32, seg_ids=kw.get("n_segs", 1), scale=0.6)
    for r in range(len(off) - 1):
        ar = a[off[r]:off[r + 1]]
        o = oracle.chain(par, ar)
        ref = oracle.ref_chain(par, ar)
        assert np.array_equal(o["u"], ref["u"]), (kw, r, "u")
        assert np.array_equal(o["b"], ref["b"]), (kw, r, "b")
        assert (o["status"] != 2) == ref["u_null"], (kw, r)


@pytest.mark.parametrize("kw", [dict(), dict(max_dist_x=60000, max_dist_y=60000, bw=40000, max_skip=5), dict(min_cnt=1, min_sc=1, max_iter=40)])
def test_oracle_matches_compiled_reference_edge_shapes(oracle, kw):
    """x + max_dist_x carrying into the rid word, run boundaries inside blocks, chain links longer than 32 anchors."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libmm2ref.so not built (needs /root/reference)")
    par = oracle.Params(**kw)
    off, a = fuzz.edge_batch(11, scale=0.5)
    for r in range(len(off) - 1):
        ar = a[off[r]:off[r + 1]]
        o = oracle.chain(par, ar)
        ref = oracle.ref_chain(par, ar)
        assert np.array_equal(o["u"], ref["u"]), (kw, r, "u")
        assert np.array_equal(o["b"], ref["b"]), (kw, r, "b")


def test_replay_threads_and_reference_agree(oracle):
    par = oracle.Params()
    off, a = fuzz.mixed_batch(7, n_reads=48)
    r1 = oracle.replay(par, off, a, n_threads=1)
    r4 = oracle.replay(par, off, a, n_threads=4)
    for k in ("n_u", "n_v", "u", "b"):
        assert np.array_equal(r1[k], r4[k]), k
    assert r1["stats"].cells == r4["stats"].cells > 0
    if oracle.have_ref():
        rr = oracle.replay(par, off, a, n_threads=3, use_ref=True)
        for k in ("n_u", "n_v", "u", "b"):
            assert np.array_equal(r1[k], rr[k]), k


def test_sort_128x_restatement_is_the_reference_permutation(oracle):
    if not oracle.have_ref():
        pytest.skip("needs oracle/_ref/libmm2ref.so")
    import ctypes as C
    rng = np.random.default_rng(11)
    for n, nkeys in [(0, 1), (1, 1), (64, 5), (65, 3), (200, 7), (1000, 40), (5000, 300), (3000, 1), (70000, 1000)]:
        a = np.empty(n, oracle.ANCHOR)
        keys = rng.integers(0, 2 ** 63, max(nkeys, 1), dtype=np.uint64) >> np.uint64(int(rng.integers(0, 40)))
        a["x"] = keys[rng.integers(0, len(keys), n)]
        a["y"] = np.arange(n, dtype=np.uint64)
        mine, ref = a.copy(), a.copy()
        oracle.lib().mm2o_sort_128x(mine.ctypes.data_as(C.c_void_p), n)
        oracle.ref_lib().radix_sort_128x(ref.ctypes.data_as(C.c_void_p), C.c_void_p(ref.ctypes.data + 16 * n))
        assert np.array_equal(mine, ref), (n, nkeys)
        assert np.all(np.diff(mine["x"].astype(np.float64)) >= 0)
