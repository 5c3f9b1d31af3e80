"""GPU tests of the two PCIe diets of the host-buffer batch call (include/mm2chain_b200.h, ABI v3) and of concurrent calls.

  in   anchors packed on the host to 8-byte {x_lo, y_lo} + runs of the high words, restored by unpack_kernel in HBM
  out  chained anchors returned as int32 indices inside their read (bi[]), b[] gathered on the host or by the caller

Every variant must give byte-identical u[] / b[] to the oracle, and all variants must agree with each other.
"""
import threading

import numpy as np
import pytest

import fuzz

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def binding(pkg):
    b = pkg("binding")
    assert b.load().mm2b_cuda_device_count() > 0
    import os
    os.environ["MM2B_SUB_ANCHORS"] = "150000"       # several sub-batches per call, several packing chunks per sub-batch
    os.environ["MM2B_PACK_CHUNK"] = "20000"
    os.environ["MM2B_PACK_INFLIGHT"] = "99"         # pack every sub-batch (the default packs only what the helper threads keep up with)
    b.init(1)
    yield b
    b.shutdown()
    os.environ.pop("MM2B_SUB_ANCHORS", None)
    os.environ.pop("MM2B_PACK_CHUNK", None)
    os.environ.pop("MM2B_PACK_INFLIGHT", None)


def _per_read(res, off, key):
    out = []
    for r in range(len(off) - 1):
        o = int(res["u_off" if key == "u" else "b_off"][r])
        n = int(res["n_u" if key == "u" else "n_v"][r])
        out.append(res[key][o:o + n])
    return out


def _check_against(res, ref, off, with_b=True):
    assert np.array_equal(res["n_u"], ref["n_u"]) and np.array_equal(res["n_v"], ref["n_v"].astype(np.int32))
    for r in range(len(off) - 1):
        o, nu, nv = int(off[r]), int(ref["n_u"][r]), int(ref["n_v"][r])
        assert np.array_equal(res["u"][res["u_off"][r]:res["u_off"][r] + nu], ref["u"][o:o + nu]), ("u", r)
        if with_b:
            assert np.array_equal(res["b"][res["b_off"][r]:res["b_off"][r] + nv], ref["b"][o:o + nv]), ("b", r)


def test_unpack_kernel_restores_every_anchor(binding, pkg):
    wl = pkg("workload")
    off, a = wl.synth_anchor_batch(300, seed=2)
    rng = np.random.default_rng(1)
    a = a.copy()
    # sprinkle other rids / strands / flag bits so that both run lists have many entries, some one anchor long
    m = rng.random(len(a)) < 0.01
    a["x"][m] ^= np.uint64(5) << np.uint64(32)
    m = rng.random(len(a)) < 0.003
    a["y"][m] |= np.uint64(1) << np.uint64(41)
    db = binding.DeviceBatch(binding.Params(), off, a, device=0)
    back = db.unpack_into_place(a)
    assert np.array_equal(back["x"], a["x"]) and np.array_equal(back["y"], a["y"])
    for n in (1, 2, 255, 256, 257, 513):                # block-boundary shapes of the kernel (256 anchors per block)
        back = db.unpack_into_place(a[:n])
        assert np.array_equal(back, a[:n]), n
    db.close()


@pytest.mark.parametrize("seed", [1, 2])
def test_all_transfer_variants_agree_with_the_oracle(binding, oracle, pkg, seed):
    wl = pkg("workload")
    off, a = wl.synth_anchor_batch(900, seed=seed)
    assert len(a) > 3 * 150000
    par = binding.Params()
    ref = oracle.replay(oracle.Params(), off, a, n_threads=8)
    res = binding.chain_batch(par, off, a, mode="b", flags=binding.F_HOST_GATHER)     # packed in, indices out, b[] gathered on the host
    _check_against(res, ref, off)
    assert res["stats"].n_packed_subs >= 3 and res["stats"].n_raw_subs == 0
    assert res["stats"].h2d_bytes < 9 * len(a) + 64 * len(off) and res["stats"].d2h_bytes < 4 * int(ref["n_v"].sum()) + 8 * int(ref["n_u"].sum()) + 64 * len(off)
    for mode, flags in (("b", binding.F_RAW_INPUT), ("b", binding.F_HOST_GATHER), ("b", binding.F_RAW_INPUT | binding.F_HOST_GATHER), ("both", 0)):
        r2 = binding.chain_batch(par, off, a, mode=mode, flags=flags)
        _check_against(r2, ref, off)
        if flags & binding.F_RAW_INPUT:
            assert r2["stats"].n_packed_subs == 0 and r2["stats"].h2d_bytes >= 16 * len(a)
    ri = binding.chain_batch(par, off, a, mode="index")
    _check_against(ri, ref, off, with_b=False)
    ri["b"] = binding.gather_b(off, a, ri)                                    # what a caller that still holds a[] does
    _check_against(ri, ref, off)
    rb = binding.chain_batch(par, off, a, mode="both")
    for r in range(len(off) - 1):                                            # indices and anchors describe the same chains
        bo, nv = int(rb["b_off"][r]), int(rb["n_v"][r])
        assert np.array_equal(a[int(off[r]) + rb["bi"][bo:bo + nv]], rb["b"][bo:bo + nv])


def test_bounded_packing_mixes_packed_and_raw_subbatches(binding, pkg, oracle):
    """MM2B_PACK_INFLIGHT=1: at most one sub-batch is being packed at a time, the others go over raw; results do not change."""
    import os
    b = binding
    b.shutdown()
    saved = {k: os.environ.pop(k, None) for k in ("MM2B_PACK_INFLIGHT", "MM2B_PACK_CHUNK")}
    os.environ["MM2B_SUB_ANCHORS"] = "60000"
    os.environ["MM2B_PACK_INFLIGHT"] = "1"
    try:
        b.init(1)
        off, a = pkg("workload").synth_anchor_batch(1200, seed=9)
        ref = oracle.replay(oracle.Params(), off, a, n_threads=8)
        res = b.chain_batch(b.Params(), off, a, mode="both")
        _check_against(res, ref, off)
        st = res["stats"]
        assert st.n_packed_subs + st.n_raw_subs >= 10 and st.n_packed_subs >= 1
        res = b.chain_batch(b.Params(), off, a)          # mm2b_chain_batch itself sends the input as it is
        _check_against(res, ref, off)
        assert res["stats"].n_packed_subs == 0
    finally:
        b.shutdown()
        os.environ.pop("MM2B_PACK_INFLIGHT", None)
        for k, v in saved.items():
            if v is not None:
                os.environ[k] = v
        os.environ["MM2B_SUB_ANCHORS"] = "150000"
        b.init(1)


def test_high_words_too_varied_fall_back_to_raw_per_subbatch(binding, oracle):
    """Segment ids / q_span changing on every anchor (paired reads, homopolymer-compressed index): the run lists overflow their budget
    and those sub-batches must go over as plain mm128_t — with the same results."""
    rng = np.random.default_rng(5)
    reads = [fuzz.dense_repeat(rng, 6000, span_jitter=True, seg_ids=2) for _ in range(60)]
    off = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    a = np.concatenate(reads)
    par = dict(n_segs=2, max_iter=200)
    ref = oracle.replay(oracle.Params(**par), off, a, n_threads=8)
    res = binding.chain_batch(binding.Params(**par), off, a, mode="both")
    assert res["stats"].n_raw_subs >= 2 and res["stats"].n_packed_subs == 0
    _check_against(res, ref, off)


def test_many_concurrent_small_calls(binding, oracle):
    """Calls from many host threads share the device workers (a worker serves several calls at once); every call's Job lives on
    its caller's stack, so a worker touching it after the caller has left would show up here as a crash or a wrong result."""
    off, a = fuzz.mixed_batch(21, n_reads=40, scale=0.4)
    ref = oracle.replay(oracle.Params(), off, a, n_threads=4)
    errs = []

    def caller(k):
        try:
            for it in range(25):
                lo = (k + it) % 20
                sub_off = off[lo:lo + 21] - off[lo]
                sub_a = a[off[lo]:off[lo + 20]]
                res = binding.chain_batch(binding.Params(), sub_off, sub_a, mode="index" if it & 1 else "default")
                assert np.array_equal(res["n_u"], ref["n_u"][lo:lo + 20]), (k, it)
                for r in range(20):
                    o, nu = int(off[lo + r]), int(ref["n_u"][lo + r])
                    assert np.array_equal(res["u"][res["u_off"][r]:res["u_off"][r] + nu], ref["u"][o:o + nu]), (k, it, r)
        except Exception as e:      # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=caller, args=(k,)) for k in range(12)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs[:3]
