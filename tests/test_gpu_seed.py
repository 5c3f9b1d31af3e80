"""GPU parity tests of the seeding front end (include/mm2seed_b200.h; run on the B200 box: pytest -m gpu), all through the C ABI.
Checkers: the committed fixture of the reference's own seeding (tests/golden/seed_golden.npz), the reference-side tool
oracle/_ref/mm2-seed-ref (the reference's collect_minimizers / collect_seed_hits compiled from its sources; prebuilt, travels with
the repository) on generated reads, and the oracles (oracle/seed_oracle.c, oracle/chain_oracle.c).  Bit-exact: minimizers, sorted
anchors including the order of equal keys, rep_len, mini_pos, and the chains that come out of the chaining kernels behind them."""
import os

import numpy as np
import pytest

import seedgen
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def binding(pkg):
    b = pkg("binding")
    assert b.load().mm2b_cuda_device_count() > 0, "no CUDA device: the GPU tests must run on the B200 box"
    b.init(1)
    yield b
    b.shutdown()


@pytest.fixture(scope="module")
def seed_oracle(oracle):
    from oracle import seed_py
    return seed_py


@pytest.fixture(scope="module")
def golden():
    g = np.load(os.path.join(GOLDEN, "seed_golden.npz"))
    flat = dict(k=int(g["k"]), w=int(g["w"]), keys=g["keys"], vals=g["vals"], pos=g["pos"])
    blob, off = bytes(g["seq"]), g["seq_off"]
    return flat, [blob[off[i]:off[i + 1]] for i in range(len(off) - 1)], int(g["mid_occ"])


@pytest.fixture(scope="module")
def generated(tmp_path_factory, seed_oracle):
    if not seed_oracle.have_tool():
        pytest.skip("oracle/_ref/mm2-seed-ref was not built (needs /root/reference at build time)")
    td = tmp_path_factory.mktemp("seed")
    chroms = seedgen.make_reference()
    reads = seedgen.make_reads(chroms)
    ref, q = str(td / "ref.fa"), str(td / "reads.fa")
    seedgen.write_fasta(ref, chroms), seedgen.write_fasta(q, reads)
    out = {}
    for preset, occ in (("map-ont", None), ("asm20", None), ("map-ont", 300)):
        sf, xf = str(td / ("s_%s_%s.bin" % (preset, occ))), str(td / ("i_%s_%s.bin" % (preset, occ)))
        seed_oracle.run_tool(preset, ref, q, sf, xf, occ)
        out[(preset, occ)] = (seed_oracle.read_seeds(sf), seed_oracle.read_index(xf))
    return [s.tobytes() for _, s in reads], out


def test_index_lookup_is_mm_idx_get(binding, seed_oracle, golden):
    flat, seqs, _ = golden
    gi, oi = binding.Index(flat), seed_oracle.Index(flat)
    try:
        present = flat["keys"][::7] >> np.uint64(1)
        rng = np.random.default_rng(3)
        absent = rng.integers(0, 1 << (2 * flat["k"]), 5000, dtype=np.uint64)
        q = np.concatenate([present, absent])
        n_occ, val = gi.lookup(q)
        for m, n, v in zip(q[:3000].tolist() + q[-2000:].tolist(), n_occ[:3000].tolist() + n_occ[-2000:].tolist(), val[:3000].tolist() + val[-2000:].tolist()):
            n_ref, v_ref = oi.get(m)
            assert n == n_ref, (m, n, n_ref)
            if n == 1:
                assert v == v_ref                      # the position itself
            elif n > 1:
                assert int(flat["pos"][v >> 32]) == v_ref and (v & 0xffffffff) == n
        assert int((n_occ[:len(present)] > 0).sum()) == len(present)
    finally:
        gi.close()


def _compare_seeds(dbg, recs, tag):
    n_tie = 0
    for i, r in enumerate(recs):
        m0, m1 = int(dbg["mini_off"][i]), int(dbg["mini_off"][i + 1])
        assert m1 - m0 == len(r["mv"]) and np.array_equal(dbg["mini"][m0:m1], r["mv"]), "%s read %d: minimizers differ from mm_sketch" % (tag, i)
        a0, a1 = int(dbg["a_off"][i]), int(dbg["a_off"][i + 1])
        assert a1 - a0 == len(r["a"]), "%s read %d: %d anchors, reference %d" % (tag, i, a1 - a0, len(r["a"]))
        assert np.array_equal(dbg["a"][a0:a1], r["a"]), "%s read %d: anchors differ from collect_seed_hits" % (tag, i)
        assert int(dbg["rep_len"][i]) == r["rep_len"], "%s read %d: rep_len" % (tag, i)
        nmp = int(dbg["n_mini_pos"][i])
        assert nmp == len(r["mini_pos"]) and np.array_equal(dbg["mini_pos"][m0:m0 + nmp].astype(np.uint64), r["mini_pos"] & np.uint64(0xffffffff)), "%s read %d: mini_pos" % (tag, i)
        n_tie += int(len(r["a"]) > 1 and bool(np.any(r["a"]["x"][1:] == r["a"]["x"][:-1])))
    return n_tie


def test_seeds_match_committed_reference_fixture(binding, seed_oracle, golden):
    flat, seqs, mid_occ = golden
    gi, oi = binding.Index(flat), seed_oracle.Index(flat)
    try:
        dbg = binding.seed_debug(gi, seqs, mid_occ)
        recs = []
        for q in seqs:                                   # the oracle is pinned to the fixture's hashes by tests/test_seed_oracle.py
            mv = seed_oracle.sketch(q, flat["w"], flat["k"])
            a, rep, mp = oi.seed(mv, len(q), mid_occ)
            recs.append(dict(mv=mv, a=a, rep_len=rep, mini_pos=mp))
        n_tie = _compare_seeds(dbg, recs, "fixture")
        assert n_tie >= 5 and dbg["n_tie_reads"] == n_tie
    finally:
        gi.close()


def test_seeds_match_the_reference_tool(binding, generated):
    seqs, recs = generated
    for key, (rec, flat) in recs.items():
        gi = binding.Index(flat)
        try:
            dbg = binding.seed_debug(gi, seqs, rec["mid_occ"])
            n_tie = _compare_seeds(dbg, rec["reads"], str(key))
            assert dbg["n_tie_reads"] == n_tie and n_tie >= 5
        finally:
            gi.close()


def test_map_batch_chains_equal_chaining_the_reference_anchors(binding, oracle, generated):
    """Sequences in, chains out: u[] / b[] must be what the chaining oracle gives on the reference's own anchors; sub-batching
    (three contexts, many small sub-batches) must not change anything."""
    seqs, recs = generated
    rec, flat = recs[("map-ont", None)]
    gi = binding.Index(flat)
    par = oracle.Params()
    try:
        for sub_bytes in (None, "40000"):
            if sub_bytes:
                os.environ["MM2B_MAP_SUB_BYTES"] = sub_bytes
            try:
                res = binding.map_batch(gi, seqs, rec["mid_occ"], binding.Params(**par.as_dict()))
            finally:
                os.environ.pop("MM2B_MAP_SUB_BYTES", None)
            if sub_bytes:
                assert res["stats"]["n_segs"] > 20
            for i, r in enumerate(rec["reads"]):
                assert int(res["n_a"][i]) == len(r["a"]) and int(res["n_mini"][i]) == len(r["mv"]) and int(res["rep_len"][i]) == r["rep_len"]
                assert np.array_equal(res["mini_pos"][i].astype(np.uint64), r["mini_pos"] & np.uint64(0xffffffff))
                ref = oracle.chain(par, r["a"])
                assert int(res["status"][i]) == ref["status"], (i, int(res["status"][i]), ref["status"])
                assert np.array_equal(res["u"][i], ref["u"]) and np.array_equal(res["b"][i], ref["b"]), "read %d: chains differ" % i
    finally:
        gi.close()


def test_minimizer_buffer_too_small_is_run_again(pkg, seed_oracle, golden):
    """The sketch writes into a buffer sized for the usual minimizer density; when a batch has more (low-complexity sequence can push
    a minimizer at every position) the kernel only counts and the host runs it again with the size it learned.  MM2B_TEST_SMALL_MV=1
    forces that path (read once per process: run in a child)."""
    import subprocess
    import sys
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from __graft_entry__ import load_package; b = load_package('binding'); b.init(1)\n"
            "g = np.load(%r); flat = dict(k=int(g['k']), w=int(g['w']), keys=g['keys'], vals=g['vals'], pos=g['pos'])\n"
            "blob, off = bytes(g['seq']), g['seq_off']; seqs = [blob[off[i]:off[i + 1]] for i in range(len(off) - 1)]\n"
            "gi = b.Index(flat); d = b.seed_debug(gi, seqs, int(g['mid_occ']))\n"
            "import hashlib; print(' '.join(hashlib.sha1(d['mini'][int(d['mini_off'][i]):int(d['mini_off'][i + 1])].tobytes()).hexdigest() for i in range(len(seqs))))\n"
            "print(' '.join(hashlib.sha1(d['a'][int(d['a_off'][i]):int(d['a_off'][i + 1])].tobytes()).hexdigest() for i in range(len(seqs))))\n"
            "gi.close(); b.shutdown()\n"
            % (os.path.dirname(GOLDEN.rstrip('/')).rsplit('/tests', 1)[0], os.path.join(GOLDEN, 'seed_golden.npz')))
    out = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=dict(os.environ, MM2B_TEST_SMALL_MV="1"), timeout=300)
    assert out.returncode == 0, out.stderr.decode()[-2000:]
    lines = out.stdout.decode().strip().splitlines()
    g = np.load(os.path.join(GOLDEN, "seed_golden.npz"))
    assert lines[-2].split() == [str(x) for x in g["mv_sha"]] and lines[-1].split() == [str(x) for x in g["a_sha"]]


def test_unsupported_configurations_are_refused(binding, golden):
    L = binding.load()
    assert L.mm2b_map_supported(15, 10, 0, 1, 0, 0) == 1
    for args in ((16, 10, 0, 1, 0, 0), (15, 10, 1, 1, 0, 0), (15, 10, 0, 2, 0, 0), (15, 10, 0, 1, 0x400000, 0), (15, 10, 0, 1, 0x001, 0), (15, 100, 0, 1, 0, 0), (15, 10, 0, 1, 0, 20)):
        assert L.mm2b_map_supported(*args) == 0
    flat, seqs, mid_occ = golden
    bad = dict(flat, k=16)
    gi = binding.Index(bad)
    try:
        with pytest.raises(binding.Mm2bError):
            binding.map_batch(gi, seqs[:2], mid_occ)
    finally:
        gi.close()


def test_empty_and_unseedable_reads(binding, golden):
    """Reads without a single minimizer (empty, shorter than k, all N) in a batch of their own and mixed with real ones."""
    flat, seqs, mid_occ = golden
    gi = binding.Index(flat)
    try:
        res = binding.map_batch(gi, [b"", b"NNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNN", b"ACGTACG", b""], mid_occ)
        assert res["n_a"].tolist() == [0, 0, 0, 0] and res["n_u"].tolist() == [0, 0, 0, 0] and res["status"].tolist() == [0, 0, 0, 0]
        mixed = [b"", seqs[0], b"NNNN", seqs[1], b""]
        res = binding.map_batch(gi, mixed, mid_occ)
        alone = binding.map_batch(gi, [seqs[0], seqs[1]], mid_occ)
        for i, j in ((1, 0), (3, 1)):
            assert int(res["n_a"][i]) == int(alone["n_a"][j]) and np.array_equal(res["u"][i], alone["u"][j]) and np.array_equal(res["b"][i], alone["b"][j])
            assert np.array_equal(res["mini_pos"][i], alone["mini_pos"][j])
        assert res["n_a"][[0, 2, 4]].tolist() == [0, 0, 0]
    finally:
        gi.close()


def test_exit_without_shutdown_does_not_hang(golden):
    """A host that exits without mm2b_shutdown (main.c returns early on several error paths) must still exit promptly."""
    import subprocess
    import sys
    from conftest import ROOT
    code = ("import sys; sys.path.insert(0, %r); from __graft_entry__ import load_package; b = load_package('binding'); b.init(1)\n"
            "import numpy as np; off = np.array([0, 3], np.int64); a = np.zeros(3, b.ANCHOR); a['x'] = [10, 20, 30]; a['y'] = [(15 << 32) | 10, (15 << 32) | 20, (15 << 32) | 30]\n"
            "r = b.chain_batch(b.Params(), off, a); print('ok', int(r['n_u'][0]))\n" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=120)
    assert b"ok" in out.stdout, out.stderr.decode()[-1500:]


def test_pinned_pool_hands_blocks_back(binding):
    """mm2b_host_alloc / _free go through a pool (pinning costs ~0.5 ms per MB): a freed block is handed out again for a request of
    about its size, mm2b_host_reserve fills the pool ahead of time, mm2b_host_pool_trim empties it."""
    import ctypes as C
    L = binding.load()
    L.mm2b_host_reserve.restype, L.mm2b_host_reserve.argtypes = None, [C.c_size_t, C.c_int]
    L.mm2b_host_pool_trim.restype, L.mm2b_host_pool_trim.argtypes = None, []
    L.mm2b_host_pool_trim()
    p = L.mm2b_host_alloc(8 << 20)
    assert p
    L.mm2b_host_free(p)
    q = L.mm2b_host_alloc(6 << 20)                   # fits the pooled 8 MB block (at most twice the size asked for)
    assert q == p
    r = L.mm2b_host_alloc(1 << 20)                   # too small a request for an 8 MB block, and the pool is empty anyway
    assert r and r != p
    L.mm2b_host_free(q), L.mm2b_host_free(r)
    L.mm2b_host_reserve(4 << 20, 3)
    got = [L.mm2b_host_alloc(4 << 20) for _ in range(3)]
    assert len(set(got)) == 3 and all(got)
    for g in got:
        L.mm2b_host_free(g)
    L.mm2b_host_pool_trim()
