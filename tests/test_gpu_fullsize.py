"""GPU parity at BASELINE.json's full sizes: every read of a 100k-read map-ont batch (59 M anchors), of a CCS/asm20 batch and
of an ultra-long batch is compared with the oracle (which chains the same batch on all host cores in about a second), plus
size-independent properties of the outputs (chains are collinear sub-sequences of the input, sorted by reference position,
scores and counts consistent, idempotent under re-chaining)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def binding(pkg):
    b = pkg("binding")
    assert b.load().mm2b_cuda_device_count() > 0
    b.init(1)
    yield b
    b.shutdown()


def _compare_all(res, ref, off):
    n_u, n_v = ref["n_u"].astype(np.int64), ref["n_v"].astype(np.int64)
    assert np.array_equal(res["n_u"], ref["n_u"]) and np.array_equal(res["n_v"].astype(np.int64), n_v)
    # gather both sides into dense per-read-ordered arrays and compare in one shot
    def gather(vals, starts, counts):
        idx = np.repeat(starts - np.concatenate([[0], np.cumsum(counts)[:-1]]), counts) + np.arange(int(counts.sum()))
        return vals[idx]
    assert np.array_equal(gather(res["u"], res["u_off"][:-1], n_u), gather(ref["u"], off[:-1], n_u))
    gb, rb = gather(res["b"], res["b_off"][:-1], n_v), gather(ref["b"], off[:-1], n_v)
    assert np.array_equal(gb["x"], rb["x"]) and np.array_equal(gb["y"], rb["y"])
    return gb


@pytest.mark.parametrize("name,n_reads", [("map-ont", 100000), ("asm20", 12000), ("ultralong", 600)])
def test_full_size_batches_bit_exact(binding, oracle, pkg, name, n_reads):
    wl = pkg("workload")
    off, a = wl.preset_batch(name, n_reads, seed=4242)
    ref = oracle.replay(oracle.Params(), off, a, n_threads=16)
    res = binding.chain_batch(binding.Params(), off, a)
    b = _compare_all(res, ref, off)
    binding.set_counting(True)                  # the statistics variant of the kernel must give the same chains and the oracle's cell count
    res_c = binding.chain_batch(binding.Params(), off, a)
    binding.set_counting(False)
    _compare_all(res_c, ref, off)
    assert res_c["stats"].cells_ref == ref["stats"].cells
    # properties that hold at any size
    n_v = res["n_v"].astype(np.int64)
    n_u = res["n_u"].astype(np.int64)         # u[] is packed per sub-batch with gaps between sub-batches: sum exactly n_u entries per read
    starts = np.repeat(res["u_off"][:-1] - np.concatenate([[0], np.cumsum(n_u)[:-1]]), n_u) + np.arange(int(n_u.sum()))
    lens = (res["u"][starts] & np.uint64(0xffffffff)).astype(np.int64)
    cnt_from_u = np.bincount(np.repeat(np.arange(len(n_u)), n_u), weights=lens, minlength=len(n_u)).astype(np.int64)
    assert np.array_equal(cnt_from_u, n_v)                               # sum of chain lengths == anchors returned
    assert int(n_v.sum()) <= len(a) and np.all(n_v <= np.diff(off))
    # every returned anchor is an input anchor of the same read
    key = lambda arr: arr["x"].astype(np.uint64) * np.uint64(1000003) ^ arr["y"]
    rid_b = np.repeat(np.arange(len(n_v)), n_v)
    rid_a = np.repeat(np.arange(len(n_v)), np.diff(off))
    sa = set(zip(rid_a[:200000].tolist(), key(a[:200000]).tolist()))
    m = rid_b <= rid_a[199999]
    assert all(t in sa for t in zip(rid_b[m][:50000].tolist(), key(b[m][:50000]).tolist()))


def test_rechaining_own_output_is_idempotent(binding, oracle, pkg):
    """Chaining the chained anchors of single-chain reads again returns the same chain (a fixed point of mm_chain_dp)."""
    wl = pkg("workload")
    off, a = wl.preset_batch("map-ont", 3000, seed=7)
    r1 = binding.chain_batch(binding.Params(), off, a)
    single = np.nonzero(r1["n_u"] == 1)[0]
    assert len(single) > 2000
    reads = [r1["b"][r1["b_off"][r]:r1["b_off"][r] + r1["n_v"][r]] for r in single]
    off2 = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([len(x) for x in reads], out=off2[1:])
    a2 = np.concatenate(reads)
    r2 = binding.chain_batch(binding.Params(), off2, a2)
    assert np.all(r2["n_u"] == 1) and np.array_equal(r2["n_v"], r1["n_v"][single])
    for k, r in enumerate(single):
        assert r2["u"][r2["u_off"][k]] == r1["u"][r1["u_off"][r]]
        assert np.array_equal(r2["b"][r2["b_off"][k]:r2["b_off"][k] + r2["n_v"][k]], reads[k])


def test_real_seeding_replay_against_reference_outputs(binding, pkg, tmp_path):
    """The reference itself as the checker, at scale: simulate ONT reads at the sequence level against a 100 Mbp reference, let
    the reference CLI (built in place, software chaining) seed and chain them while the capture shim records every mm_chain_dp
    call, then replay the captured anchors through the GPU batch call and compare with the reference's OWN u[] / b[]."""
    import os
    import subprocess
    from conftest import ROOT
    from oracle import dumpio
    cli = os.path.join(ROOT, "oracle", "_ref", "minimap2-sw")
    if not os.path.exists(cli):
        pytest.skip("oracle/_ref/minimap2-sw was not built (needs /root/reference at build time)")
    seqsim = pkg("seqsim")
    ref = seqsim.gen_reference(100_000_000, seed=1)
    seqsim.write_fasta(str(tmp_path / "ref.fa"), [("chr1", ref)])
    seqsim.write_fasta(str(tmp_path / "q.fa"), seqsim.gen_reads(ref, 6000, 10000, 0.10, seed=23))
    dump = str(tmp_path / "d.bin")
    subprocess.run([cli, "-t", "16", "-x", "map-ont", str(tmp_path / "ref.fa"), str(tmp_path / "q.fa")], env=dict(os.environ, MM2_DUMP=dump),
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
    recs = dumpio.read_dump(dump)
    assert len(recs) >= 5900
    off, a = dumpio.to_batch(recs)
    res = binding.chain_batch(binding.Params(**recs[0]["par"].as_dict()), off, a)
    bad = 0
    for r, rec in enumerate(recs):
        nu, nv = int(res["n_u"][r]), int(res["n_v"][r])
        ok = nu == len(rec["u"]) and nv == len(rec["b"]) and (int(res["status"][r]) == 2) == (not rec["u_null"]) \
            and np.array_equal(res["u"][res["u_off"][r]:res["u_off"][r] + nu], rec["u"]) \
            and np.array_equal(res["b"][res["b_off"][r]:res["b_off"][r] + nv], rec["b"])
        bad += int(not ok)
    assert bad == 0, "%d of %d reads differ from the reference's own output" % (bad, len(recs))
