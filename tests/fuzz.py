"""Seeded adversarial anchor generators shared by the CPU (oracle vs reference) and GPU (CUDA vs oracle) tests.

Anchors follow the packing of map.c:232-241: x = rev<<63 | rid<<32 | ref_pos, y = seg<<48 | flags(40-43) | q_span<<32 | q_pos,
sorted ascending by x (ties in arbitrary but fixed order, like the unstable sort at map.c:245 leaves them).
"""
import numpy as np

ANCHOR = np.dtype([("x", "<u8"), ("y", "<u8")])


def _pack(rev, rid, rpos, qpos, span, seg=0, flags=0):
    x = (rev.astype(np.uint64) << np.uint64(63)) | (rid.astype(np.uint64) << np.uint64(32)) | rpos.astype(np.uint64)
    y = (np.asarray(seg).astype(np.uint64) << np.uint64(48)) | (np.asarray(flags).astype(np.uint64) << np.uint64(40)) | \
        (np.asarray(span).astype(np.uint64) << np.uint64(32)) | qpos.astype(np.uint64)
    a = np.empty(len(x), ANCHOR)
    a["x"], a["y"] = x, y
    return a[np.argsort(a["x"], kind="stable")]


def collinear(rng, n_true, n_noise, span=15, qlen=None, indel=0.08, step=(8, 40), n_rid=3, genome=5_000_000, seg_ids=1, tie_rate=0.0):
    """One dense collinear cluster (a mapped read) plus uniformly scattered noise anchors."""
    steps = rng.integers(step[0], step[1], n_true)
    q = np.cumsum(steps) + span
    drift = np.cumsum(np.where(rng.random(n_true) < indel, rng.integers(-6, 7, n_true), 0))
    r0 = int(rng.integers(10_000, genome - 10_000 - int(q[-1]) if n_true else genome))
    r = r0 + q + drift
    if tie_rate > 0:   # duplicate some reference / query positions to provoke dr==0, dq==0 and equal-score ties
        m = rng.random(n_true) < tie_rate
        r[m] = np.roll(r, 1)[m]
        m = rng.random(n_true) < tie_rate
        q[m] = np.roll(q, 1)[m]
    qlen = int(q[-1]) + span + 1 if n_true else 1000
    rev = np.full(n_true, int(rng.integers(0, 2)))
    rid = np.full(n_true, int(rng.integers(0, n_rid)))
    nq = rng.integers(span, qlen, n_noise)
    nr = rng.integers(span, genome, n_noise)
    rev = np.concatenate([rev, rng.integers(0, 2, n_noise)])
    rid = np.concatenate([rid, rng.integers(0, n_rid, n_noise)])
    n = n_true + n_noise
    seg = rng.integers(0, seg_ids, n) if seg_ids > 1 else np.zeros(n, np.int64)
    spans = np.full(n, span)
    return _pack(rev, rid, np.concatenate([np.abs(r), nr]), np.concatenate([q, nq]), spans, seg=seg)


def dense_repeat(rng, n, span=15, width=3000, qwidth=3000, seg_ids=1, span_jitter=False):
    """Everything inside one max_dist_x window: long inner loops, many ties, many chains (tandem-repeat like)."""
    rpos = 100_000 + np.sort(rng.integers(0, width, n))
    qpos = span + rng.integers(0, qwidth, n)
    spans = rng.integers(5, 256, n) if span_jitter else np.full(n, span)
    seg = rng.integers(0, seg_ids, n) if seg_ids > 1 else np.zeros(n, np.int64)
    flags = rng.integers(0, 16, n)
    return _pack(np.zeros(n, np.int64), np.zeros(n, np.int64), rpos, qpos, spans, seg=seg, flags=flags)


def lattice(rng, n, period=37, span=15):
    """Periodic structure: many equal-score candidates, exercising the strict-> tie rule and duplicate chain keys."""
    k = np.arange(n)
    rpos = 50_000 + (k // 4) * period
    qpos = span + (k % 4) * period * 3 + (k // 4) * period
    return _pack(np.zeros(n, np.int64), np.zeros(n, np.int64), rpos, qpos, np.full(n, span))


def many_chains(rng, n_chains, per_chain, span=15, dup_start=True):
    """Many short chains; with dup_start several chains begin at the same reference x (tie keys in the final sort)."""
    parts = []
    for c in range(n_chains):
        r0 = 1_000_000 + (c // 3 if dup_start else c) * 20_000
        q0 = span + c * (per_chain * 20 + 6000)
        k = np.arange(per_chain)
        parts.append((r0 + k * 18, q0 + k * 18))
    r = np.concatenate([p[0] for p in parts])
    q = np.concatenate([p[1] for p in parts])
    n = len(r)
    return _pack(np.zeros(n, np.int64), np.zeros(n, np.int64), r, q, np.full(n, span))


def batch(reads):
    off = np.zeros(len(reads) + 1, np.int64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    a = np.concatenate(reads) if len(reads) else np.empty(0, ANCHOR)
    return off, a


def mixed_batch(seed, n_reads=40, seg_ids=1, scale=1.0):
    rng = np.random.default_rng(seed)
    reads = []
    for i in range(n_reads):
        kind = i % 8
        if kind == 0:
            reads.append(collinear(rng, int(rng.integers(1, 900 * scale + 2)), int(rng.integers(0, 300 * scale + 1)), seg_ids=seg_ids))
        elif kind == 1:
            reads.append(collinear(rng, int(rng.integers(5, 400 * scale + 6)), int(rng.integers(0, 50)), tie_rate=0.15, seg_ids=seg_ids))
        elif kind == 2:
            reads.append(dense_repeat(rng, int(rng.integers(2, 700 * scale + 3)), seg_ids=seg_ids))
        elif kind == 3:
            reads.append(lattice(rng, int(rng.integers(4, 500 * scale + 5))))
        elif kind == 4:
            reads.append(many_chains(rng, int(rng.integers(1, 40 * scale + 2)), int(rng.integers(2, 9))))
        elif kind == 5:
            reads.append(np.empty(0, ANCHOR) if i % 16 == 5 else collinear(rng, 1, int(rng.integers(0, 3))))
        elif kind == 6:
            reads.append(dense_repeat(rng, int(rng.integers(2, 300 * scale + 3)), width=200, qwidth=200, span_jitter=True, seg_ids=seg_ids))
        else:
            reads.append(collinear(rng, int(rng.integers(30, 1500 * scale + 31)), 0, indel=0.3, step=(1, 200)))
    return batch(reads)

def high_positions(rng, n, n_rid=3, span=15, spread=30_000):
    """Reference positions just below 2^32, plus small positions on the next rid: `x + max_dist_x` carries into the
    rid/strand word (chain.c:192 does the addition on the whole 64-bit x), so the window search cannot work on low words."""
    top = (1 << 32) - 1
    rid = rng.integers(0, n_rid, n)
    hi = rng.random(n) < 0.6
    rpos = np.where(hi, top - rng.integers(0, spread, n), rng.integers(0, spread, n))
    qpos = span + rng.integers(0, spread, n)
    return _pack(rng.integers(0, 2, n), rid, rpos, qpos, np.full(n, span))


def many_runs(rng, n, n_rid=60, span=15):
    """Short runs of equal (strand, rid): several run boundaries inside every block of 32 anchors, each run a small cluster."""
    rid = rng.integers(0, n_rid, n)
    rev = rng.integers(0, 2, n)
    r0 = 200_000 + (rid * 7919 + rev * 104729) % 50_000
    k = rng.integers(0, 12, n)
    return _pack(rev, rid, r0 + k * 21 + rng.integers(0, 3, n), span + k * 21 + rng.integers(0, 3, n), np.full(n, span))


def interleaved(rng, n_diag, per, span=15):
    """n_diag chains on diagonals far apart in the query but interleaved on the reference: every chain link skips
    n_diag - 1 anchors (links longer than a 32-anchor block when n_diag > 32), and the chains tie in score."""
    k = np.repeat(np.arange(per), n_diag)
    d = np.tile(np.arange(n_diag), per)
    rpos = 300_000 + k * (n_diag + 5) + d
    qpos = span + d * 6000 + k * (n_diag + 5)
    return _pack(np.zeros(len(k), np.int64), np.zeros(len(k), np.int64), rpos, qpos, np.full(len(k), span))


def edge_batch(seed, scale=1.0):
    """The shapes the window search and the block-wise backtrack have to get right; see tests/test_gpu_parity.py."""
    rng = np.random.default_rng(seed)
    s = lambda v: max(2, int(v * scale))
    reads = [high_positions(rng, s(900)), high_positions(rng, s(300), n_rid=1, spread=6000), many_runs(rng, s(1500)),
             many_runs(rng, s(200), n_rid=5), interleaved(rng, 40, s(30)), interleaved(rng, 3, s(200)), interleaved(rng, 70, s(12)),
             np.concatenate([dense_repeat(rng, s(700), width=4500, qwidth=4000), high_positions(rng, s(200))]),
             collinear(rng, s(2500), s(400), n_rid=40), collinear(rng, s(800), s(800), n_rid=200, genome=400_000)]
    reads[7] = reads[7][np.argsort(reads[7]["x"], kind="stable")]
    return batch(reads)
