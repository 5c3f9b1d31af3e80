"""CPU tests of the seeding oracle (oracle/seed_oracle.c: mm_sketch, mm_idx_get, collect_matches, collect_seed_hits restated):
  * pinned against the reference itself — oracle/_ref/mm2-seed-ref runs the reference's own collect_minimizers / collect_seed_hits
    (map.c:64-78, :215-247) on generated reads and records minimizers, sorted anchors, rep_len and mini_pos;
  * pinned against a committed fixture of those records (tests/golden/seed_golden.npz, written by tests/golden/make_seed_golden.py);
  * the position-parallel formulation the sketch kernel uses, modelled in Python, against the sequential restatement."""
import hashlib
import os

import numpy as np
import pytest

import seedgen
from conftest import GOLDEN


@pytest.fixture(scope="module")
def seed_oracle(oracle):
    from oracle import seed_py
    return seed_py


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
    return reads, out


def test_oracle_is_the_reference_on_generated_reads(seed_oracle, generated):
    reads, recs = generated
    for key, (rec, flat) in recs.items():
        ix = seed_oracle.Index(flat)
        n_tie = 0
        for (name, q), r in zip(reads, rec["reads"]):
            mv = seed_oracle.sketch(q.tobytes(), rec["w"], rec["k"])
            assert np.array_equal(mv, r["mv"]), "%s %s: minimizers differ from mm_sketch" % (key, name)
            a, rep, mp = ix.seed(mv, len(q), rec["mid_occ"])
            assert np.array_equal(a, r["a"]), "%s %s: anchors differ from collect_seed_hits" % (key, name)
            assert rep == r["rep_len"] and np.array_equal(mp, r["mini_pos"]), "%s %s: rep_len / mini_pos differ" % (key, name)
            n_tie += int(len(a) > 1 and bool(np.any(a["x"][1:] == a["x"][:-1])))
        assert n_tie >= 5          # the unstable order of equal keys is exercised


def test_oracle_against_committed_fixture(seed_oracle):
    g = np.load(os.path.join(GOLDEN, "seed_golden.npz"))
    flat = dict(k=int(g["k"]), w=int(g["w"]), keys=g["keys"], vals=g["vals"], pos=g["pos"])
    ix = seed_oracle.Index(flat)
    seqs = bytes(g["seq"])
    off = g["seq_off"]
    for i in range(len(off) - 1):
        q = seqs[off[i]:off[i + 1]]
        mv = seed_oracle.sketch(q, flat["w"], flat["k"])
        a, rep, mp = ix.seed(mv, len(q), int(g["mid_occ"]))
        assert hashlib.sha1(mv.tobytes()).hexdigest() == str(g["mv_sha"][i]), "read %d: minimizers" % i
        assert hashlib.sha1(a.tobytes()).hexdigest() == str(g["a_sha"][i]), "read %d: anchors" % i
        assert rep == int(g["rep_len"][i]) and len(mp) == int(g["n_mini_pos"][i])


# ---- the sketch kernel's formulation -------------------------------------------------------------------------------------------
NONE = (1 << 64) - 1
CODE = {**{c: 0 for c in b"Aa\x00"}, **{c: 1 for c in b"Cc\x01"}, **{c: 2 for c in b"Gg\x02"}, **{c: 3 for c in b"TtUu\x03"}}


def _hash64(key, mask):
    key = (~key + (key << 21)) & mask
    key ^= key >> 24
    key = (key + (key << 3) + (key << 8)) & mask
    key ^= key >> 14
    key = (key + (key << 2) + (key << 4)) & mask
    key ^= key >> 28
    return (key + (key << 31)) & mask


def model_sketch(seq, w, k):
    """What sketch_kernel (csrc/seed_kernels.cu) computes: every position decides from X[t-w .. t] and the run of valid bases what
    mm_sketch pushes at its step; the minimum of the ring is always the newest minimal entry of the last w positions."""
    L = len(seq)
    c = [CODE.get(ch, 4) for ch in seq]
    mask = (1 << 2 * k) - 1
    run = [0] * L
    for i in range(L):
        run[i] = 0 if c[i] == 4 else (run[i - 1] if i else 0) + 1
    X, Z = {}, {}

    def x_at(j):
        if j < 0 or j >= L:
            return NONE
        if j not in X:
            x, z = NONE, 0
            if run[j] >= k:
                f = fw = 0
                for m in range(k):
                    f |= c[j - k + 1 + m] << (2 * m)                # oldest base lowest: the kernel's packed field
                    fw |= c[j - k + 1 + m] << (2 * (k - 1 - m))
                rv = f ^ mask
                z = 0 if fw < rv else 1
                x = _hash64(rv if z else fw, mask) << 8 | k
            X[j], Z[j] = x, z
        return X[j]

    out = []
    for t in range(L):
        l = run[t]
        xm, jm = NONE, t - w
        for j in range(t - w, t):
            if x_at(j) <= xm:
                xm, jm = x_at(j), j
        after, em = jm, []
        if l == w + k - 1 and xm != NONE:
            em += [j for j in range(t - w + 1, t) if x_at(j) == xm and j != jm]
        if x_at(t) <= xm:
            if l >= w + k and xm != NONE:
                em.append(jm)
            after = t
        elif jm == t - w:
            if l >= w + k - 1 and xm != NONE:
                em.append(jm)
            xn, jn = NONE, t - w + 1
            for j in range(t - w + 1, t + 1):
                if x_at(j) <= xn:
                    xn, jn = x_at(j), j
            if l >= w + k - 1 and xn != NONE:
                em += [j for j in range(t - w + 1, t + 1) if x_at(j) == xn and j != jn]
            after = jn
        if t == L - 1 and x_at(after) != NONE:
            em.append(after)
        out += [(X[j], j << 1 | Z[j]) for j in em]
    return out


def test_position_parallel_sketch_equals_the_sequential_one(seed_oracle):
    rng = np.random.default_rng(1)
    cases = []
    for it in range(120):
        L = int(rng.integers(1, 300))
        kind = it % 5
        if kind == 0:
            s = bytes(rng.choice(list(b"ACGT"), L))
        elif kind == 1:
            s = bytes(rng.choice(list(b"ACGTN"), L, p=[.23, .23, .23, .23, .08]))
        elif kind == 2:
            s = bytes(rng.choice(list(b"AC"), L))
        elif kind == 3:
            s = bytes(rng.choice(list(b"AAAAAAAAAAAAAAAAAAAAN"), L))
        else:
            u = bytes(rng.choice(list(b"ACGT"), int(rng.integers(1, 12))))
            s = (u * (L // len(u) + 1))[:L]
        cases.append(s)
    for s in cases:
        for w, k in ((10, 15), (5, 15), (10, 19), (19, 19), (1, 15), (3, 5), (64, 27)):
            ref = seed_oracle.sketch(s, w, k)
            mod = model_sketch(s, w, k)
            assert len(ref) == len(mod) and all((int(a["x"]), int(a["y"])) == m for a, m in zip(ref, mod)), (w, k, s)
