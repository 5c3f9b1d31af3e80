"""Deterministic inputs for the seeding tests: a small reference with everything that makes seeding interesting (a tandem array,
runs of N, homopolymer / dinucleotide / trinucleotide stretches, a segmental duplication on another chromosome) and reads drawn
from it at 0 / 1 / 10 % error with N runs, lower case, plus hand-made edge cases around k, w + k and the 512-position tile of the
sketch kernel."""
import numpy as np

B = np.frombuffer(b"ACGT", np.uint8)


def _rnd(rng, n):
    return B[rng.integers(0, 4, n)]


def _mutate(rng, s, e):
    u = rng.random(len(s))
    sub = u < e / 3
    dele = (u >= e / 3) & (u < 2 * e / 3)
    ins = (u >= 2 * e / 3) & (u < e)
    s = s.copy()
    idx = np.searchsorted(B, s[sub])
    s[sub] = B[(idx + rng.integers(1, 4, int(sub.sum()))) % 4]
    keep = ~dele
    reps = np.where(ins, 2, 1)[keep]
    out = np.repeat(s[keep], reps)
    pos = np.cumsum(reps) - 1
    ins_at = pos[ins[keep]]
    out[ins_at] = _rnd(rng, len(ins_at))
    return out


def _rc(s):
    m = np.full(256, ord("N"), np.uint8)
    for a, b in zip(b"ACGTacgt", b"TGCAtgca"):
        m[a] = b
    return m[s][::-1]


def make_reference(seed=7, scale=1.0):
    rng = np.random.default_rng(seed)
    n1 = int(600000 * scale)
    c1 = _rnd(rng, n1)
    unit = _rnd(rng, 300)
    t0 = n1 // 6
    for c in range(40):
        u = unit.copy()
        m = rng.random(300) < 0.02
        u[m] = _rnd(rng, int(m.sum()))
        c1[t0 + c * 300:t0 + (c + 1) * 300] = u
    q = n1 // 3
    c1[q:q + 700] = ord("N")
    c1[q + 5000:q + 5400] = ord("A")
    c1[q + 9000:q + 9600] = np.tile(np.frombuffer(b"AC", np.uint8), 300)
    c1[q + 12000:q + 12900] = np.tile(np.frombuffer(b"ACG", np.uint8), 300)
    chroms = [("chr1", c1), ("chr2", _rnd(rng, int(250000 * scale)))]
    chroms.append(("chr3", _mutate(rng, c1[t0 - 50000 if t0 > 50000 else 0:t0 + 50000], 0.01)))
    return chroms


def make_reads(chroms, n_reads=300, seed=11, mean_len=6000):
    rng = np.random.default_rng(seed)
    reads = []
    for i in range(n_reads):
        _, s = chroms[int(rng.integers(0, len(chroms)))]
        L = int(max(60, rng.gamma(4, mean_len / 4)))
        st = int(rng.integers(0, max(1, len(s) - L)))
        q = s[st:st + L].copy()
        if rng.integers(0, 2):
            q = _rc(q)
        q = _mutate(rng, q, [0.0, 0.01, 0.1][i % 3])
        if i % 7 == 0 and len(q) > 50:
            p = int(rng.integers(0, len(q) - 40))
            q[p:p + int(rng.integers(1, 40))] = ord("N")
        if i % 11 == 0:
            q = np.frombuffer(q.tobytes().lower(), np.uint8).copy()
        if i % 13 == 0:
            q[::int(rng.integers(2, 50))] = ord("n")
        reads.append(("r%d" % i, q))
    c1 = chroms[0][1]
    t0 = len(c1) // 6
    N = lambda n: np.full(n, ord("N"), np.uint8)
    reads += [("short_lt_k", _rnd(rng, 10)), ("len_k", _rnd(rng, 15)), ("len_wk", _rnd(rng, 24)), ("len_wk1", _rnd(rng, 25)),
              ("allA", np.full(500, ord("A"), np.uint8)), ("allN", N(300)), ("AC", np.tile(np.frombuffer(b"AC", np.uint8), 400)),
              ("tandem_cross", c1[max(0, t0 - 5000):t0 + 15000].copy()),
              ("tile_edges", np.concatenate([_rnd(rng, 505), N(3), _rnd(rng, 520), N(1), _rnd(rng, 1200)])),
              ("N_at_tile", np.concatenate([_rnd(rng, 511), N(1), _rnd(rng, 512), N(30), _rnd(rng, 600)])),
              ("one_base", _rnd(rng, 1))]
    return reads


def write_fasta(path, recs):
    with open(path, "wb") as fh:
        for name, s in recs:
            fh.write(b">" + name.encode() + b"\n" + (s.tobytes() if hasattr(s, "tobytes") else bytes(s)) + b"\n")
