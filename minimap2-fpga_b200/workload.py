"""Synthetic chaining workloads at the ANCHOR level (the input of mm_chain_dp).

synth_anchor_batch() draws, per read, one collinear cluster (the read's true locus, minimizers every ~5.5 bp of which
~1/5 survive 10 % error, with indel drift between reference and query coordinates) plus uniformly scattered random
seed hits, and packs them exactly as collect_seed_hits does (map.c:232-241): x = rev<<63 | rid<<32 | ref_pos,
y = q_span<<32 | q_pos, sorted by x.  It is used by __graft_entry__.smoke() and by quick tests; bench.py's workloads are the
anchors the reference's own CLI seeds for simulated reads (bench_workloads.py), this model only with `--source model`.
"""
import numpy as np

ANCHOR = np.dtype([("x", "<u8"), ("y", "<u8")])


def synth_read(rng, read_len, k=15, keep=0.2, density=2.0 / 11.0, err_indel=0.067, noise_rate=0.023, genome=100_000_000):
    n_min = max(1, int(read_len * density))
    qpos = np.sort(rng.choice(np.arange(k, max(read_len, k + n_min + 1)), size=n_min, replace=False))
    kept = qpos[rng.random(n_min) < keep]
    gaps = np.diff(np.concatenate([[0], kept]))
    drift = np.cumsum(rng.binomial(gaps, err_indel / 2) - rng.binomial(gaps, err_indel / 2))
    start = int(rng.integers(0, genome - 2 * read_len - 10))
    rpos_t = start + kept + drift
    rev_t = np.full(len(kept), int(rng.integers(0, 2)), np.uint64)
    n_noise = rng.poisson(noise_rate * read_len)
    q_n = rng.integers(k, max(read_len, k + 1), n_noise)
    r_n = rng.integers(k, genome, n_noise)
    rev_n = rng.integers(0, 2, n_noise).astype(np.uint64)
    q = np.concatenate([kept, q_n]).astype(np.uint64)
    r = np.concatenate([np.abs(rpos_t), r_n]).astype(np.uint64)
    rev = np.concatenate([rev_t, rev_n])
    a = np.empty(len(q), ANCHOR)
    a["x"] = (rev << np.uint64(63)) | r
    a["y"] = (np.uint64(k) << np.uint64(32)) | q
    return a[np.argsort(a["x"], kind="stable")]


def synth_anchor_batch(n_reads, seed=0, mean_len=10000, min_len=500, **kw):
    rng = np.random.default_rng(seed)
    reads = []
    for _ in range(n_reads):
        L = max(min_len, int(rng.gamma(4.0, mean_len / 4.0)))
        reads.append(synth_read(rng, L, **kw))
    off = np.zeros(n_reads + 1, np.int64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    return off, (np.concatenate(reads) if reads else np.empty(0, ANCHOR))


# BASELINE.json configs at the anchor level.  Model constants are calibrated against real minimap2 seeding of simulated reads
# (SURVEY.md 8d probe: ONT 588 anchors/read, 15.9 cells/anchor, 61 % chained; CCS 2,248 anchors/read, 26.8 cells/anchor,
# 99.9 % chained; ultra-long 7.5 k anchors/read); tests/test_workload_model.py re-checks the statistics with the oracle.
PRESETS = {
    # name: (mean read length, generator keyword arguments, chaining preset)
    "map-ont": (10000, dict(k=15, keep=0.2, err_indel=0.067, noise_rate=0.023), "map-ont"),
    "asm20": (15000, dict(k=19, keep=0.826, err_indel=0.0067, noise_rate=0.0002), "asm20"),
    "ultralong": (120000, dict(k=15, keep=0.2, err_indel=0.067, noise_rate=0.023), "map-ont"),
}


def preset_batch(name, n_reads, seed=0):
    mean_len, kw, _ = PRESETS[name]
    return synth_anchor_batch(n_reads, seed=seed, mean_len=mean_len, **kw)
