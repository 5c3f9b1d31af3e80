"""Long-read segmenting on the GPU (chain_kernels.cu: segment_kernel / chain_pieces_kernel): reads with usable x-gap cut points
(chain.c:192) are cut into pieces that are filled by warps of their own, the last piece to finish runs the per-read extraction.
Chimeric reads — several loci in one read — against the oracle, with thresholds lowered so that the path is taken; the same
batches with MM2B_SEG=0; reads that are one collinear chain must not be cut."""
import os

import numpy as np
import pytest

import fuzz

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def binding(pkg):
    b = pkg("binding")
    assert b.load().mm2b_cuda_device_count() > 0
    b.init(1)
    yield b
    b.shutdown()


def chimeric(rng, n_loci, per_locus, noise, span=15):
    """n_loci collinear clusters at distant reference positions / strands, consecutive in the query, plus scattered noise."""
    parts, q0 = [], 0
    for c in range(n_loci):
        a = fuzz.collinear(rng, per_locus + int(rng.integers(0, per_locus // 4 + 1)), 0, span=span, n_rid=2, genome=200_000_000)
        y = a["y"].copy()
        y += np.uint64(q0)                                   # this locus comes after the previous one in the query
        a = a.copy()
        a["y"] = y
        q0 = int((y & np.uint64(0xffffffff)).max()) + 200
        parts.append(a)
    if noise:
        parts.append(fuzz.collinear(rng, 1, noise, span=span, n_rid=2, genome=200_000_000))
    a = np.concatenate(parts)
    return a[np.argsort(a["x"], kind="stable")]


def _run(binding, oracle, reads, par_kw, env):
    off, a = fuzz.batch(reads)
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        db = binding.DeviceBatch(binding.Params(**par_kw), off, a, device=0)
        db.run()
        db.run()                                            # a second batch on the same workspace: control words are reset
        st, res = db.stats(), db.results()
        db.close()
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    ref = oracle.replay(oracle.Params(**par_kw), off, a, n_threads=8)
    assert np.array_equal(res["n_u"], ref["n_u"]) and np.array_equal(res["n_v"].astype(np.int64), ref["n_v"].astype(np.int64))
    for r in range(len(off) - 1):
        o, nu, nv = int(off[r]), int(ref["n_u"][r]), int(ref["n_v"][r])
        assert np.array_equal(res["u"][res["u_off"][r]:res["u_off"][r] + nu], ref["u"][o:o + nu]), r
        assert np.array_equal(res["b"][res["b_off"][r]:res["b_off"][r] + nv], ref["b"][o:o + nv]), r
    return st


LOW = {"MM2B_SEG_MIN_READ": "1200", "MM2B_SEG_MIN_PIECE": "250"}


@pytest.mark.parametrize("par_kw", [dict(), dict(min_cnt=1, min_sc=5), dict(max_skip=5, max_iter=64), dict(n_segs=2, gap_scale=1.7)])
def test_chimeric_reads_are_cut_and_chain_like_the_reference(binding, oracle, par_kw):
    rng = np.random.default_rng(17)
    reads = [chimeric(rng, int(rng.integers(2, 9)), int(rng.integers(300, 900)), int(rng.integers(0, 400))) for _ in range(40)]
    reads += [fuzz.collinear(rng, 2500, 300), fuzz.collinear(rng, 40, 10), chimeric(rng, 20, 300, 100)]     # one chain; short; more loci than pieces
    st = _run(binding, oracle, reads, par_kw, LOW)
    assert st.n_cut_reads >= 35, st.n_cut_reads
    st0 = _run(binding, oracle, reads, par_kw, dict(LOW, MM2B_SEG="0"))
    assert st0.n_cut_reads == 0


def test_collinear_reads_are_not_cut(binding, oracle):
    """One locus plus scattered seed hits: the isolated hits are cut points, but there is no second piece with work in it."""
    rng = np.random.default_rng(3)
    reads = [fuzz.collinear(rng, 3000, 1500, genome=200_000_000) for _ in range(12)]
    st = _run(binding, oracle, reads, dict(), {"MM2B_SEG_MIN_READ": "1200", "MM2B_SEG_MIN_PIECE": "600"})
    assert st.n_cut_reads == 0


def test_default_thresholds_on_long_chimeric_reads(binding, oracle):
    rng = np.random.default_rng(5)
    reads = [chimeric(rng, 6, 2600, 500) for _ in range(6)] + [fuzz.collinear(rng, 600, 200) for _ in range(50)]
    st = _run(binding, oracle, reads, dict(), {})
    assert st.n_cut_reads == 6
