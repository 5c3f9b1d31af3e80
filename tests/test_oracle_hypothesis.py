"""Property-based pinning of the oracle against the reference's compiled chain.c (oracle/_ref/libmm2ref.so): random small reads
with random chaining parameters, including degenerate ones.  Skipped where the in-place reference build is unavailable."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st


@st.composite
def reads(draw):
    n = draw(st.integers(0, 60))
    rng = np.random.default_rng(draw(st.integers(0, 2 ** 32 - 1)))
    style = draw(st.sampled_from(["cluster", "scatter", "dups", "two_strands", "carry"]))
    span = draw(st.sampled_from([1, 15, 19, 255]))
    if style == "cluster":
        r = 1000 + np.cumsum(rng.integers(0, 40, n)); q = span + np.cumsum(rng.integers(0, 40, n))
    elif style == "scatter":
        r = rng.integers(0, 20000, n); q = rng.integers(span, 20000, n)
    elif style == "dups":
        r = 500 + rng.integers(0, 6, n) * 17; q = span + rng.integers(0, 6, n) * 17
    elif style == "carry":      # positions just below 2^32 on rid 0, small ones on rid 1: x + max_dist_x carries into the rid word (chain.c:192)
        top = rng.random(n) < 0.6
        r = np.where(top, (1 << 32) - 1 - rng.integers(0, 6000, n), (1 << 32) + rng.integers(0, 6000, n)); q = span + rng.integers(0, 6000, n)
    else:
        r = 1000 + np.cumsum(rng.integers(1, 30, n)); q = span + np.cumsum(rng.integers(1, 30, n))
    rev = (rng.integers(0, 2, n) if style == "two_strands" else np.zeros(n, np.int64)).astype(np.uint64)
    seg = rng.integers(0, draw(st.integers(1, 3)), n).astype(np.uint64)
    a = np.empty(n, [("x", "<u8"), ("y", "<u8")])
    a["x"] = (rev << np.uint64(63)) | np.asarray(r, np.uint64)
    a["y"] = (seg << np.uint64(48)) | (np.uint64(span) << np.uint64(32)) | np.asarray(q, np.uint64)
    return a[np.argsort(a["x"], kind="stable")]


params = st.fixed_dictionaries(dict(
    max_dist_x=st.sampled_from([0, 50, 500, 5000]), max_dist_y=st.sampled_from([0, 50, 500, 5000]), bw=st.sampled_from([0, 10, 100, 500]),
    max_skip=st.integers(0, 30), max_iter=st.sampled_from([0, 1, 5, 50, 5000]), min_cnt=st.integers(0, 4), min_sc=st.sampled_from([-5, 0, 10, 40]),
    is_cdna=st.integers(0, 1), n_segs=st.integers(1, 3), gap_scale=st.sampled_from([0.5, 1.0, 1.7])))


@settings(max_examples=300, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(a=reads(), kw=params)
def test_oracle_equals_compiled_reference(oracle, a, kw):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libmm2ref.so not built (needs /root/reference)")
    par = oracle.Params(**kw)
    o = oracle.chain(par, a)
    ref = oracle.ref_chain(par, a)
    assert np.array_equal(o["u"], ref["u"]) and np.array_equal(o["b"], ref["b"])
    assert (o["status"] != 2) == ref["u_null"]
