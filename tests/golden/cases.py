"""Inputs of the golden cases: FASTA files (the reference's bundled test data + seeded synthetic sequences) and the CLI
arguments of each case.  Shared by make_golden.py (which runs the REFERENCE CLI on them) and tests/test_gpu_cli.py (which
runs the same reference CLI linked against the B200 backend and expects byte-identical PAF)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

TEST = os.path.join(ROOT, "oracle", "_ref", "test")      # copies of /root/reference/test/*.fa made by oracle/Makefile


def tandem_reference(rng, length, unit_len, copies, at, div=0.02):
    ref = rng.integers(0, 4, length, dtype=np.uint8)
    unit = rng.integers(0, 4, unit_len, dtype=np.uint8)
    for c in range(copies):
        u = unit.copy()
        m = rng.random(unit_len) < div
        u[m] = (u[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) & 3
        ref[at + c * unit_len: at + (c + 1) * unit_len] = u
    return ref


def build_cases(td):
    """Write every input FASTA under `td` and return [(name, cli_args)] in a fixed order."""
    seqsim = load_package("seqsim")
    cases = []
    # --- config 0: the reference's bundled inputs -------------------------------------------------
    for preset in ("map-ont", "asm20"):
        cases.append(("mt_" + preset, ["-x", preset, TEST + "/MT-human.fa", TEST + "/MT-orang.fa"]))
        cases.append(("inv_" + preset, ["-x", preset, TEST + "/t-inv.fa", TEST + "/q-inv.fa"]))
    cases.append(("inv_n1m5", ["-x", "map-ont", "-n", "1", "-m", "5", TEST + "/t-inv.fa", TEST + "/q-inv.fa"]))
    # --- config 1/2 shape, small: synthetic ONT / CCS reads vs a random reference ---------------------
    ref = seqsim.gen_reference(4_000_000, seed=1)
    seqsim.write_fasta(td + "/ref.fa", [("chrS", ref)])
    seqsim.write_fasta(td + "/ont.fa", seqsim.gen_reads(ref, 24, 10000, 0.10, seed=7))
    seqsim.write_fasta(td + "/ccs.fa", seqsim.gen_reads(ref, 6, 15000, 0.01, seed=8))
    cases.append(("syn_ont", ["-x", "map-ont", td + "/ref.fa", td + "/ont.fa"]))
    cases.append(("syn_ccs", ["-x", "asm20", td + "/ref.fa", td + "/ccs.fa"]))
    cases.append(("syn_ont_n1m5", ["-x", "map-ont", "-n", "1", "-m", "5", td + "/ref.fa", td + "/ont.fa"]))
    # --- config 3 shape, small: tandem repeats (dense windows, max_iter clamp, max_skip, many chains) ---
    rng = np.random.default_rng(5)
    tref = tandem_reference(rng, 300_000, 400, 60, 100_000)
    seqsim.write_fasta(td + "/tref.fa", [("chrT", tref)])
    reads = []
    for i, (s, e) in enumerate([(90_000, 140_000), (100_000, 124_000), (95_000, 112_000)]):
        seq = tref[s:e]
        if i == 1:
            seq = seqsim.revcomp(seq)
        reads.append(("t%d" % i, seqsim.mutate(seq, 0.06, rng)))
    seqsim.write_fasta(td + "/tq.fa", reads)
    cases.append(("tandem", ["-x", "map-ont", "-f", "100000", td + "/tref.fa", td + "/tq.fa"]))
    cases.append(("tandem_iter64", ["-x", "map-ont", "-f", "100000", "--max-chain-iter", "64", "--max-chain-skip", "5", td + "/tref.fa", td + "/tq.fa"]))
    # --- other presets through the same function: paired short reads (n_segs=2), splice (is_cdna), gap scale ---
    sref = ref[:400_000]
    seqsim.write_fasta(td + "/sref.fa", [("chrS", sref)])
    r1, r2 = [], []
    for i in range(60):
        st = int(rng.integers(0, len(sref) - 600))
        frag = sref[st:st + int(rng.integers(300, 550))]
        r1.append(("p%d" % i, seqsim.mutate(frag[:150], 0.01, rng)))
        r2.append(("p%d" % i, seqsim.mutate(seqsim.revcomp(frag)[:150], 0.01, rng)))
    seqsim.write_fasta(td + "/r1.fa", r1)
    seqsim.write_fasta(td + "/r2.fa", r2)
    cases.append(("sr_paired", ["-x", "sr", td + "/sref.fa", td + "/r1.fa", td + "/r2.fa"]))
    sp = []
    for i in range(12):
        st = int(rng.integers(0, len(sref) - 40_000))
        exons, pos = [], st
        for _ in range(int(rng.integers(3, 8))):
            el = int(rng.integers(80, 400))
            exons.append(sref[pos:pos + el])
            pos += el + int(rng.integers(200, 4000))
        seq = np.concatenate(exons)
        if i & 1:
            seq = seqsim.revcomp(seq)
        sp.append(("s%d" % i, seqsim.mutate(seq, 0.03, rng)))
    seqsim.write_fasta(td + "/sp.fa", sp)
    cases.append(("splice", ["-x", "splice", td + "/sref.fa", td + "/sp.fa"]))
    cases.append(("syn_ont_gapscale", ["-x", "map-ont", "--chain-gap-scale", "1.7", "-r", "2000", td + "/ref.fa", td + "/ont.fa"]))
    seqsim.write_fasta(td + "/ovl.fa", seqsim.gen_reads(ref[:50_000], 10, 8000, 0.08, seed=9))
    cases.append(("ava", ["-x", "ava-ont", td + "/ovl.fa", td + "/ovl.fa"]))
    return cases
