"""Synthetic reference / read simulator for the BASELINE.json configs (sequence level, numpy).

Shape: reference = i.i.d. uniform ACGT from default_rng(seed); reads have
length max(500, Gamma(k=4, theta=mean/4)), uniform start, random strand, per-base error `err` split
1/3 substitution, 1/3 deletion, 1/3 insertion.  Used to make the tests/golden fixtures (through the
reference CLI) and small end-to-end inputs; bench.py's large workloads come from the C++ seed
generator in workload/ which follows the same model.
"""
import numpy as np

_ALPHA = np.frombuffer(b"ACGT", dtype=np.uint8)


def gen_reference(length, seed=1):
    return np.random.default_rng(seed).integers(0, 4, length, dtype=np.uint8)


def mutate(seq, err, rng):
    """Apply sub/del/ins errors at total rate `err` to a uint8 {0..3} array."""
    n = len(seq)
    u = rng.random(n)
    out = seq.copy()
    sub = u < err / 3
    out[sub] = (out[sub] + rng.integers(1, 4, int(sub.sum()), dtype=np.uint8)) & 3
    dele = (u >= err / 3) & (u < 2 * err / 3)
    ins = (u >= 2 * err / 3) & (u < err)
    reps = np.ones(n, np.int64)
    reps[dele] = 0
    reps[ins] = 2
    res = np.repeat(out, reps)
    # second copy of every inserted position becomes a random base
    ends = np.cumsum(reps)[ins] - 1
    res[ends] = rng.integers(0, 4, len(ends), dtype=np.uint8)
    return res


def revcomp(seq):
    return (3 - seq)[::-1]


def gen_reads(ref, n_reads, mean_len, err, seed, min_len=500):
    rng = np.random.default_rng(seed + 1000)
    reads = []
    for i in range(n_reads):
        L = max(min_len, int(rng.gamma(4.0, mean_len / 4.0)))
        L = min(L, len(ref))
        start = int(rng.integers(0, len(ref) - L + 1))
        strand = int(rng.integers(0, 2))
        s = ref[start:start + L]
        if strand:
            s = revcomp(s)
        reads.append(("r%d_%d_%d_%d" % (i, start, start + L, strand), mutate(s, err, rng)))
    return reads


def to_ascii(seq):
    return _ALPHA[seq].tobytes()


def write_fasta(path, records, width=0):
    """records: iterable of (name, uint8 array)."""
    with open(path, "wb") as fh:
        for name, seq in records:
            fh.write(b">" + name.encode() + b"\n")
            fh.write(to_ascii(seq))
            fh.write(b"\n")
