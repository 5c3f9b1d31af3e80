"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI of
libmm2chain_b200.so; the oracle (pinned against the reference by tests/test_oracle.py) and the committed reference
captures in tests/golden/ are the checkers.  Bit-exact: integer scores, indices and anchor bytes must be identical.
"""
import numpy as np
import pytest

import fuzz
from conftest import golden_names, load_golden
from test_oracle import PARAM_SETS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def binding(pkg):
    b = pkg("binding")
    L = b.load()
    assert L.mm2b_cuda_device_count() > 0, "no CUDA device: the GPU tests must run on the B200 box"
    b.init(1)
    yield b
    b.shutdown()


def _groups(recs):
    """Group capture records by parameter set (the sr preset changes max_dist per read)."""
    out = {}
    for r in recs:
        out.setdefault(tuple(sorted(r["par"].as_dict().items())), []).append(r)
    return list(out.values())


def _compare_batch(res, recs_or_ref, off, name):
    for r in range(len(off) - 1):
        if isinstance(recs_or_ref, list):
            u_ref, b_ref = recs_or_ref[r]["u"], recs_or_ref[r]["b"]
            ok_ref = not recs_or_ref[r]["u_null"]
        else:
            o, nu, nv = int(off[r]), int(recs_or_ref["n_u"][r]), int(recs_or_ref["n_v"][r])
            u_ref, b_ref = recs_or_ref["u"][o:o + nu], recs_or_ref["b"][o:o + nv]
            ok_ref = None
        nu, nv = int(res["n_u"][r]), int(res["n_v"][r])
        assert nu == len(u_ref) and nv == len(b_ref), (name, r, nu, len(u_ref), nv, len(b_ref))
        uo, bo = int(res["u_off"][r]), int(res["b_off"][r])
        assert np.array_equal(res["u"][uo:uo + nu], u_ref), (name, r, "u")
        assert np.array_equal(res["b"][bo:bo + nv], b_ref), (name, r, "b")
        if ok_ref is not None:
            assert (int(res["status"][r]) == 2) == ok_ref, (name, r, "status")


@pytest.mark.parametrize("name", golden_names())
def test_device_batch_matches_reference_capture(binding, name):
    """f/p/v per anchor and u[]/b[] per read against what the reference CLI itself produced."""
    from oracle import dumpio
    for recs in _groups(load_golden(name)):
        off, a = dumpio.to_batch(recs)
        par = binding.Params(**recs[0]["par"].as_dict())
        db = binding.DeviceBatch(par, off, a, device=0, keep_fpv=True)
        db.run()
        f, p, v = db.fpv()
        res = db.results()
        for k, r in enumerate(recs):
            s, e = int(off[k]), int(off[k + 1])
            assert np.array_equal(f[s:e], r["f"]), (name, k, "f", int(np.argmax(f[s:e] != r["f"])))
            assert np.array_equal(p[s:e], r["p"]), (name, k, "p", int(np.argmax(p[s:e] != r["p"])))
            assert np.array_equal(v[s:e], r["v"]), (name, k, "v")
        _compare_batch(res, recs, off, name)
        db.close()


@pytest.mark.parametrize("name", ["mt_map-ont", "inv_map-ont", "syn_ont_n1m5", "sr_paired", "splice"])
def test_mm_chain_dp_dropin_matches_reference_capture(binding, name):
    """The per-read boundary itself: same 16 positional arguments, same NULL conventions as chain.c:29."""
    for r in load_golden(name):
        par = binding.Params(**r["par"].as_dict())
        u, b, u_null, b_null = binding.chain_read(par, r["a"])
        assert np.array_equal(u, r["u"]) and np.array_equal(b, r["b"])
        assert u_null == r["u_null"] and b_null == r["b_null"]
    u, b, u_null, b_null = binding.chain_read(binding.Params(), np.empty(0, binding.ANCHOR))
    assert u_null and b_null and len(u) == 0


@pytest.mark.parametrize("pi", range(len(PARAM_SETS)))
def test_host_batch_matches_oracle_fuzz(binding, oracle, pi):
    """Adversarial inputs (ties, dr==0 / dq==0, dense windows, many chains, empty and 1-anchor reads) through the host-buffer call."""
    kw = PARAM_SETS[pi]
    off, a = fuzz.mixed_batch(200 + pi, n_reads=64, seg_ids=kw.get("n_segs", 1))
    ref = oracle.replay(oracle.Params(**kw), off, a, n_threads=4)
    res = binding.chain_batch(binding.Params(**kw), off, a)
    _compare_batch(res, ref, off, kw)
    assert res["stats"].n_chains == int(ref["n_u"].sum()) and res["stats"].n_chained == int(ref["n_v"].sum())


def test_deep_lookback_and_long_reads(binding, oracle):
    """Windows far deeper than the 256-slot shared-memory ring, max_iter clamp, > 64 chains (radix tie order)."""
    rng = np.random.default_rng(5)
    reads = [fuzz.dense_repeat(rng, 6000, width=4500, qwidth=4000), fuzz.dense_repeat(rng, 3000, width=800, qwidth=6000),
             fuzz.many_chains(rng, 400, 4), fuzz.lattice(rng, 5000), fuzz.collinear(rng, 20000, 500)]
    off, a = fuzz.batch(reads)
    for kw in (dict(), dict(max_iter=300, max_skip=2), dict(min_cnt=1, min_sc=1)):
        ref = oracle.replay(oracle.Params(**kw), off, a, n_threads=8)
        res = binding.chain_batch(binding.Params(**kw), off, a)
        _compare_batch(res, ref, off, kw)


@pytest.mark.parametrize("kw", [dict(), dict(max_dist_x=60000, max_dist_y=60000, bw=40000, max_skip=5), dict(min_cnt=1, min_sc=1, max_iter=40),
                                dict(max_skip=200), dict(n_segs=2, max_dist_x=800, max_dist_y=600, bw=100, min_cnt=2, min_sc=25)])
def test_window_search_and_backtrack_edge_shapes(binding, oracle, kw):
    """x + max_dist_x carrying into the rid word (64-bit merge search), run boundaries inside blocks (low-word search),
    chain links longer than a 32-anchor block and interleaved tying chains (block-wise backtrack), quiet tails (max_skip=200)."""
    for seed in (11, 12):
        off, a = fuzz.edge_batch(seed)
        ref = oracle.replay(oracle.Params(**kw), off, a, n_threads=8)
        res = binding.chain_batch(binding.Params(**kw), off, a)
        _compare_batch(res, ref, off, kw)


def test_subbatching_and_order_independence(binding, oracle, monkeypatch):
    """Results must not depend on how the batch is cut into sub-batches or on read order."""
    off, a = fuzz.mixed_batch(77, n_reads=120, scale=0.5)
    par, opar = binding.Params(), oracle.Params()
    ref = oracle.replay(opar, off, a, n_threads=4)
    res = binding.chain_batch(par, off, a)
    _compare_batch(res, ref, off, "whole")
    perm = np.random.default_rng(1).permutation(len(off) - 1)
    reads = [a[off[r]:off[r + 1]] for r in perm]
    off2, a2 = fuzz.batch(reads)
    res2 = binding.chain_batch(par, off2, a2)
    for k, r in enumerate(perm):
        nu, nv = int(res["n_u"][r]), int(res["n_v"][r])
        assert int(res2["n_u"][k]) == nu and int(res2["n_v"][k]) == nv
        assert np.array_equal(res2["u"][res2["u_off"][k]:res2["u_off"][k] + nu], res["u"][res["u_off"][r]:res["u_off"][r] + nu])
        assert np.array_equal(res2["b"][res2["b_off"][k]:res2["b_off"][k] + nv], res["b"][res["b_off"][r]:res["b_off"][r] + nv])


def test_kernel_cell_tally_equals_oracle(binding, oracle, pkg):
    """GCUPS is quoted on reference-semantics cells; the kernel's own tally of them must equal the instrumented oracle's."""
    wl = pkg("workload")
    off, a = wl.synth_anchor_batch(300, seed=11)
    ref = oracle.replay(oracle.Params(), off, a, n_threads=4)
    res = binding.chain_batch(binding.Params(), off, a)
    assert res["stats"].cells_ref == 0          # the tally is a statistics option, off by default
    binding.set_counting(True)
    res = binding.chain_batch(binding.Params(), off, a)
    _compare_batch(res, ref, off, "synth")
    assert res["stats"].cells_ref == ref["stats"].cells and res["stats"].window_cells == ref["stats"].window_cells
    assert res["stats"].cells_issued >= res["stats"].cells_ref
    off, a = fuzz.mixed_batch(31, n_reads=48)
    for kw in (dict(), dict(max_skip=2, max_iter=40)):
        ref = oracle.replay(oracle.Params(**kw), off, a, n_threads=4)
        res = binding.chain_batch(binding.Params(**kw), off, a)
        assert res["stats"].cells_ref == ref["stats"].cells and res["stats"].window_cells == ref["stats"].window_cells, kw
        _compare_batch(res, ref, off, kw)
    binding.set_counting(False)


def test_edge_shapes(binding, oracle):
    """Read lengths around the 32-anchor block and 256-slot ring boundaries, windows that end exactly at the ring edge, extreme
    field values (q_span 255, reverse strand, rid 2^31-1, positions near 2^31), duplicate anchors, degenerate thresholds."""
    rng = np.random.default_rng(99)
    reads = []
    for n in (1, 2, 3, 31, 32, 33, 63, 64, 65, 95, 96, 97, 223, 224, 225, 255, 256, 257, 287, 288, 289, 511, 512, 513, 1023, 1025):
        reads.append(fuzz.dense_repeat(rng, n, width=max(4, n // 2), qwidth=max(4, n // 2)))        # everything in one window
        reads.append(fuzz.collinear(rng, n, 0, step=(1, 30)))
        reads.append(fuzz.lattice(rng, max(n, 4)))
    # windows whose start falls exactly on / just below the ring edge: uniform spacing s so that 5000 / s sweeps past 224..256
    for spacing in (19, 20, 21, 22, 23):
        k = np.arange(1500)
        a = np.empty(len(k), binding.ANCHOR)
        a["x"] = (np.uint64(1) << np.uint64(63)) | (np.uint64(2 ** 31 - 1) << np.uint64(32)) | (np.uint64(2 ** 31 - 40000) + (k * spacing).astype(np.uint64))
        a["y"] = (np.uint64(255) << np.uint64(32)) | (np.uint64(300) + (k * spacing).astype(np.uint64))
        reads.append(a)
    dup = fuzz.collinear(rng, 200, 20)
    reads.append(np.sort(np.concatenate([dup, dup, dup[:50]]), order="x", kind="stable"))          # duplicate anchors (dr == 0 and dq == 0)
    off, a = fuzz.batch(reads)
    for kw in (dict(), dict(max_skip=0), dict(max_skip=1, max_iter=1), dict(min_cnt=0, min_sc=-5), dict(min_cnt=1, min_sc=0, bw=0),
               dict(max_dist_x=0, max_dist_y=0), dict(max_iter=0), dict(bw=100000, max_dist_x=100000, max_dist_y=100000, max_iter=100000)):
        ref = oracle.replay(oracle.Params(**kw), off, a, n_threads=8)
        binding.set_counting(True)
        res = binding.chain_batch(binding.Params(**kw), off, a)
        binding.set_counting(False)
        _compare_batch(res, ref, off, kw)
        assert res["stats"].cells_ref == ref["stats"].cells, kw
        res = binding.chain_batch(binding.Params(**kw), off, a)          # the non-counting kernel variant
        _compare_batch(res, ref, off, kw)


def test_one_very_long_read(binding, oracle):
    """A single 300k-anchor read with a window that stays at the max_iter clamp (int32 indices, deep look-back throughout)."""
    rng = np.random.default_rng(3)
    off, a = fuzz.batch([fuzz.dense_repeat(rng, 300000, width=200000, qwidth=200000)])
    for kw in (dict(max_iter=400), dict(max_iter=5000, max_skip=3)):
        ref = oracle.replay(oracle.Params(**kw), off, a, n_threads=1)
        res = binding.chain_batch(binding.Params(**kw), off, a)
        _compare_batch(res, ref, off, kw)


def test_error_codes_and_empty_batches(binding):
    """The mm2b_* calls report misuse with return codes + mm2b_last_error() (they never exit); empty batches are fine."""
    import ctypes as C
    L = binding.load()
    par = binding.Params()
    off = np.array([0, 5, 9], np.int64)
    a = fuzz.collinear(np.random.default_rng(0), 9, 0)
    small_u = np.empty(3, np.uint64)
    with pytest.raises(binding.Mm2bError, match="u_cap and b_cap"):
        binding.chain_batch(par, off, a, out={"u": small_u})
    rc = L.mm2b_chain_batch(C.byref(par), 2, None, None, None, None, None, None, None, None, 0, None, 0, None)
    assert rc == -2 and b"NULL" in L.mm2b_last_error()
    res = binding.chain_batch(par, np.zeros(1, np.int64), np.empty(0, binding.ANCHOR))          # zero reads
    assert len(res["n_u"]) == 0 and res["u_off"][0] == 0
    res = binding.chain_batch(par, np.zeros(4, np.int64), np.empty(0, binding.ANCHOR))          # three empty reads
    assert list(res["n_u"]) == [0, 0, 0] and list(res["status"]) == [0, 0, 0]
    ws = L.mm2b_ws_create(0, 100, 10)
    assert ws and L.mm2b_ws_bytes(ws) > 0
    rc = L.mm2b_chain_batch_device(ws, C.byref(par), 11, 50, None, None, None, None, None, None, None, None, None, None)
    assert rc == -3 and b"capacity" in L.mm2b_last_error()
    L.mm2b_ws_destroy(ws)


def test_random_parameter_sweep(binding, oracle):
    """24 random parameter sets (degenerate values included) x 60 small random reads each, through both kernel variants."""
    rng = np.random.default_rng(2026)
    dom = dict(max_dist_x=[0, 50, 500, 5000], max_dist_y=[0, 50, 500, 5000], bw=[0, 10, 100, 500], max_skip=list(range(0, 31)),
               max_iter=[0, 1, 5, 50, 5000], min_cnt=[0, 1, 2, 3, 4], min_sc=[-5, 0, 10, 40], is_cdna=[0, 1], n_segs=[1, 2, 3], gap_scale=[0.5, 1.0, 1.7])
    for trial in range(24):
        kw = {k: (float if k == "gap_scale" else int)(v[rng.integers(0, len(v))]) for k, v in dom.items()}
        reads = []
        for _ in range(60):
            n = int(rng.integers(0, 120))
            style = int(rng.integers(0, 4))
            span = [1, 15, 19, 255][int(rng.integers(0, 4))]
            if style == 0:
                r = 1000 + np.cumsum(rng.integers(0, 40, n)); q = span + np.cumsum(rng.integers(0, 40, n))
            elif style == 1:
                r = rng.integers(0, 20000, n); q = rng.integers(span, 20000, n)
            elif style == 2:
                r = 500 + rng.integers(0, 6, n) * 17; q = span + rng.integers(0, 6, n) * 17
            else:
                r = 1000 + np.cumsum(rng.integers(1, 30, n)); q = span + np.cumsum(rng.integers(1, 30, n))
            rev = (rng.integers(0, 2, n) if style == 3 else np.zeros(n, np.int64)).astype(np.uint64)
            seg = rng.integers(0, kw["n_segs"], n).astype(np.uint64)
            a = np.empty(n, binding.ANCHOR)
            a["x"] = (rev << np.uint64(63)) | np.asarray(r, np.uint64)
            a["y"] = (seg << np.uint64(48)) | (np.uint64(span) << np.uint64(32)) | np.asarray(q, np.uint64)
            reads.append(a[np.argsort(a["x"], kind="stable")])
        off, a = fuzz.batch(reads)
        ref = oracle.replay(oracle.Params(**kw), off, a, n_threads=4)
        for counting in (False, True):
            binding.set_counting(counting)
            res = binding.chain_batch(binding.Params(**kw), off, a)
            _compare_batch(res, ref, off, kw)
            if counting:
                assert res["stats"].cells_ref == ref["stats"].cells, kw
        binding.set_counting(False)
