"""GPU test of the host backend's sub-batch pipeline and device sharding: many small sub-batches (MM2B_SUB_ANCHORS) pulled by
one worker thread per visible device, three slots in flight each, results gathered in input order.  With one GPU this
exercises the pipeline; with several (gpurun --gpus N) it also exercises read sharding across devices."""
import os

import numpy as np
import pytest

import fuzz

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def binding(pkg):
    b = pkg("binding")
    L = b.load()
    assert L.mm2b_cuda_device_count() > 0
    os.environ["MM2B_SUB_ANCHORS"] = "20000"
    b.init()                       # all visible devices
    yield b
    b.shutdown()
    os.environ.pop("MM2B_SUB_ANCHORS", None)


def test_many_subbatches_all_devices(binding, oracle, pkg):
    wl = pkg("workload")
    n_dev = binding.load().mm2b_num_devices()
    assert n_dev >= 1
    off, a = wl.synth_anchor_batch(1500, seed=5)
    assert len(a) > 30 * 20000          # dozens of sub-batches
    ref = oracle.replay(oracle.Params(), off, a, n_threads=8)
    binding.set_counting(True)
    for _ in range(2):                  # second call reuses grown slots
        res = binding.chain_batch(binding.Params(), off, a)
        assert np.array_equal(res["n_u"], ref["n_u"]) and np.array_equal(res["n_v"], ref["n_v"].astype(np.int32))
        for r in range(len(off) - 1):
            o, nu, nv = int(off[r]), int(ref["n_u"][r]), int(ref["n_v"][r])
            assert np.array_equal(res["u"][res["u_off"][r]:res["u_off"][r] + nu], ref["u"][o:o + nu]), r
            assert np.array_equal(res["b"][res["b_off"][r]:res["b_off"][r] + nv], ref["b"][o:o + nv]), r
        assert res["stats"].cells_ref == ref["stats"].cells
    binding.set_counting(False)


def test_concurrent_callers(binding, oracle):
    """mm2b_chain_batch and mm_chain_dp are thread-safe: several host threads at once (ctypes releases the GIL)."""
    import threading
    off, a = fuzz.mixed_batch(9, n_reads=60, scale=0.5)
    ref = oracle.replay(oracle.Params(), off, a, n_threads=4)
    errs = []

    def batch_caller():
        try:
            res = binding.chain_batch(binding.Params(), off, a)
            assert np.array_equal(res["n_u"], ref["n_u"])
            assert res["stats"].n_chained == int(ref["n_v"].sum())
        except Exception as e:      # noqa: BLE001
            errs.append(repr(e))

    def read_caller(lo, hi):
        try:
            for r in range(lo, hi):
                u, b, _, _ = binding.chain_read(binding.Params(), a[off[r]:off[r + 1]])
                o, nu, nv = int(off[r]), int(ref["n_u"][r]), int(ref["n_v"][r])
                assert np.array_equal(u, ref["u"][o:o + nu]) and np.array_equal(b, ref["b"][o:o + nv]), r
        except Exception as e:      # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=batch_caller) for _ in range(3)] + \
         [threading.Thread(target=read_caller, args=(k * 15, k * 15 + 15)) for k in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs


def test_batcher_aggregates_many_mm_chain_dp_callers(binding, oracle, pkg):
    """64 host threads call the per-read drop-in at once (an oversubscribed `-t`): the cross-thread batcher must hand every
    caller its own read's result, bit-exact, whatever flight it ended up in."""
    import threading
    wl = pkg("workload")
    off, a = wl.synth_anchor_batch(640, seed=21)
    ref = oracle.replay(oracle.Params(), off, a, n_threads=8)
    errs, n_threads = [], 64

    def caller(t):
        try:
            for r in range(t, len(off) - 1, n_threads):
                u, b, u_null, b_null = binding.chain_read(binding.Params(), a[off[r]:off[r + 1]])
                o, nu, nv = int(off[r]), int(ref["n_u"][r]), int(ref["n_v"][r])
                assert np.array_equal(u, ref["u"][o:o + nu]) and np.array_equal(b, ref["b"][o:o + nv]), r
        except Exception as e:      # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=caller, args=(t,)) for t in range(n_threads)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs[:3]


def test_map_batch_over_all_devices(binding, oracle):
    """The seeding front end shards sub-batches of reads over every bound device (an index replica on each) and concurrent callers
    share the devices: results equal chaining the seeding oracle's anchors, whatever device a read landed on."""
    import threading
    import seedgen
    from oracle import seed_py
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "seed_golden.npz"))
    flat = dict(k=int(g["k"]), w=int(g["w"]), keys=g["keys"], vals=g["vals"], pos=g["pos"])
    blob, off = bytes(g["seq"]), g["seq_off"]
    seqs = [blob[off[i]:off[i + 1]] for i in range(len(off) - 1)] * 3
    oi = seed_py.Index(flat)
    par = oracle.Params()
    ref = []
    for q in seqs[:len(seqs) // 3]:
        mv = seed_py.sketch(q, flat["w"], flat["k"])
        a, rep, mp = oi.seed(mv, len(q), int(g["mid_occ"]))
        ref.append((oracle.chain(par, a), rep, len(a)))
    ref = ref * 3
    gi = binding.Index(flat)
    os.environ["MM2B_MAP_SUB_BYTES"] = "20000"
    errs = []

    def caller():
        try:
            res = binding.map_batch(gi, seqs, int(g["mid_occ"]), binding.Params(**par.as_dict()))
            assert res["stats"]["n_segs"] >= 8
            for i, (rc, rep, n_a) in enumerate(ref):
                assert int(res["n_a"][i]) == n_a and int(res["rep_len"][i]) == rep and int(res["status"][i]) == rc["status"], i
                assert np.array_equal(res["u"][i], rc["u"]) and np.array_equal(res["b"][i], rc["b"]), i
        except Exception as e:      # noqa: BLE001
            errs.append(repr(e)[:300])

    try:
        th = [threading.Thread(target=caller) for _ in range(3)]
        for t in th:
            t.start()
        for t in th:
            t.join()
    finally:
        os.environ.pop("MM2B_MAP_SUB_BYTES", None)
        gi.close()
    assert not errs, errs
