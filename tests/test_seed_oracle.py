"""CPU tests of the seeding oracle (oracle/seed_oracle.c: mm_sketch, mm_idx_get, collect_matches, collect_seed_hits restated):
  * pinned against the reference itself — oracle/_ref/mm2-seed-ref runs the reference's own collect_minimizers / collect_seed_hits
    (map.c:64-78, :215-247) on generated reads and records minimizers, sorted anchors, rep_len and mini_pos;
  * pinned against a committed fixture of those records (tests/golden/seed_golden.npz, written by tests/golden/make_seed_golden.py);
  * the position-parallel formulations the two sketch kernels use (one position per thread; eight per thread), modelled in Python,
    against the sequential restatement."""
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


def _brev64(x):
    return int("{:064b}".format(x)[::-1], 2)


def model_sketch8(seq, w, k, TILE=1920, P=8, HALO_THREADS=16):
    """What sketch8_kernel (csrc/seed_kernels.cu) computes, thread by thread: 8 consecutive positions per thread, the k-mer and its
    reverse complement rolled from the k - 1 bases in front (taken from the neighbours' packed words), the nine windows around the
    thread's positions from suffix minima over the w hashes in front of its first position combined with prefix minima over its
    own (newest minimum + 'hash occurs twice'), tiles of 1,920 positions after a halo of 128."""
    L = len(seq)
    mask = (1 << 2 * k) - 1
    M64 = (1 << 64) - 1
    threads = TILE // P + HALO_THREADS
    span, halo = threads * P, HALO_THREADS * P
    out = []
    for t0 in range(0, L, TILE):
        pb = t0 - halo
        raw = [seq[pb + i] if 0 <= pb + i < L else ord("N") for i in range(span)]
        pk16, bad8s = [0] * (threads + 4), [0] * threads
        for tid in range(threads):
            two = bad = 0
            for j in range(P):
                c = CODE.get(raw[tid * P + j], 4)
                two |= (c & 3) << 2 * j
                bad |= (c >> 2) << j
            pk16[4 + tid], bad8s[tid] = two, bad
        K, Z, RUN = [NONE] * span, [0] * span, [0] * span
        last_bad = -1                                       # block-wide max scan in the kernel
        for tid in range(threads):
            i0 = tid * P
            run = 128 if last_bad < 0 else min(128, i0 - 1 - last_bad)
            prev = 0
            for e in range(4):
                prev |= pk16[tid + e] << (16 * e)            # the 32 bases in front of i0, oldest lowest
            f = prev >> (64 - 2 * (k - 1)) if k > 1 else 0
            rv = ((f ^ (mask >> 2)) << 2) & M64
            x = _brev64(f)
            x = ((x >> 1) & 0x5555555555555555) | ((x & 0x5555555555555555) << 1)
            fw = x >> (64 - 2 * (k - 1)) if k > 1 else 0
            for j in range(P):
                c = (pk16[4 + tid] >> 2 * j) & 3
                fw = ((fw << 2) | c) & mask
                rv = (rv >> 2) | ((3 ^ c) << 2 * (k - 1))
                run = 0 if (bad8s[tid] >> j) & 1 else min(128, run + 1)
                RUN[i0 + j] = run
                if run >= k:
                    z = 0 if fw < rv else 1
                    K[i0 + j], Z[i0 + j] = _hash64(rv if z else fw, mask) << 8 | k, z
            if bad8s[tid]:
                last_bad = i0 + bad8s[tid].bit_length() - 1
        for tid in range(HALO_THREADS, threads):
            i0 = tid * P
            p0 = pb + i0
            xk = K[i0:i0 + P]
            suf = [NONE, i0 - 1, False]
            win = [None] * (P + 1)

            def older(i):
                xo = K[i0 - w + i]
                if xo < suf[0]:
                    suf[0], suf[1], suf[2] = xo, i0 - w + i, False
                elif xo == suf[0]:
                    suf[2] = True
            for i in range(w - 1, P, -1):
                older(i)
            for i in range(P, -1, -1):
                if i < w:
                    older(i)
                win[i] = tuple(suf)
            pre = [xk[0], i0, False]
            for j in range(1, P + 1):
                m = list(pre)
                if j < w:
                    o = win[j]
                    if o[0] < pre[0]:
                        m = list(o)
                    elif o[0] == pre[0]:
                        m[2] = True
                win[j] = tuple(m)
                if j < P and xk[j] <= pre[0]:
                    pre = [xk[j], i0 + j, xk[j] == pre[0]]
            for j in range(P):
                i, t = i0 + j, p0 + j
                if t >= L:
                    break
                l, (xm, jm, ties), xt = RUN[i], win[j], xk[j]
                em, after = [], jm
                if l == w + k - 1 and xm != NONE and ties:
                    em += [q for q in range(i - w + 1, i) if K[q] == xm and q != jm]
                if xt <= xm:
                    if l >= w + k and xm != NONE:
                        em.append(jm)
                    after = i
                elif jm == i - w:
                    if l >= w + k - 1 and xm != NONE:
                        em.append(jm)
                    xn, jn, tn = win[j + 1]
                    if l >= w + k - 1 and xn != NONE and tn:
                        em += [q for q in range(i - w + 1, i + 1) if K[q] == xn and q != jn]
                    after = jn
                if t == L - 1 and K[after] != NONE:
                    em.append(after)
                out += [(K[q], (pb + q) << 1 | Z[q]) for q in em]
    return out


def test_eight_positions_per_thread_sketch_equals_the_sequential_one(seed_oracle):
    rng = np.random.default_rng(2)
    cases = []
    for it, L in enumerate([1, 7, 8, 9, 24, 25, 100, 127, 128, 129, 300, 1919, 1920, 1921, 2047, 2048, 3839, 3841, 4100]):
        kind = it % 5
        if kind == 0:
            s = bytes(rng.choice(list(b"ACGT"), L))
        elif kind == 1:
            s = bytes(rng.choice(list(b"ACGTN"), L, p=[.24, .24, .24, .24, .04]))
        elif kind == 2:
            s = bytes(rng.choice(list(b"AC"), L))
        elif kind == 3:
            s = bytes(rng.choice(list(b"AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAAN"), L))
        else:
            u = bytes(rng.choice(list(b"ACGT"), int(rng.integers(1, 12))))
            s = (u * (L // len(u) + 1))[:L]
        cases.append(s)
    for s in cases:
        for w, k in ((10, 15), (10, 19), (19, 19), (11, 21), (8, 3), (64, 27)):
            ref = seed_oracle.sketch(s, w, k)
            mod = model_sketch8(s, w, k)
            assert len(ref) == len(mod) and all((int(a["x"]), int(a["y"])) == m for a, m in zip(ref, mod)), (w, k, len(s))
