#!/usr/bin/env python
"""Regenerate tests/golden/ from the REFERENCE itself (run in the build container only).

Every fixture is a capture of real mm_chain_dp calls made by the reference CLI built in place from
/root/reference (oracle/_ref/minimap2-sw, software chaining, `make -C oracle ref`), recorded by
oracle/dump_shim.c: inputs a[], the reference's f/p/v after the DP fill, and its final u[]/b[].
PAF md5s of the same runs go to paf_md5.json.  Usage:  python tests/golden/make_golden.py
"""
import gzip
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

seqsim = load_package("seqsim")
GOLD = os.path.join(ROOT, "tests", "golden")
CLI = os.path.join(ROOT, "oracle", "_ref", "minimap2-sw")
TEST = os.path.join(ROOT, "oracle", "_ref", "test")


def run(name, args, md5s):
    with tempfile.TemporaryDirectory() as td:
        dump = os.path.join(td, "d.bin")
        env = dict(os.environ, MM2_DUMP=dump)
        paf = subprocess.run([CLI, "-t", "1"] + args, env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
        md5s[name] = dict(args=[os.path.basename(a) if os.path.isabs(a) else a for a in args],
                          md5=hashlib.md5(paf).hexdigest(), lines=paf.count(b"\n"))
        with open(dump, "rb") as fi, gzip.GzipFile(os.path.join(GOLD, name + ".dump.gz"), "wb", 9, mtime=0) as fo:
            shutil.copyfileobj(fi, fo)
    print(name, md5s[name]["md5"], md5s[name]["lines"], os.path.getsize(os.path.join(GOLD, name + ".dump.gz")))


def tandem_reference(rng, length, unit_len, copies, at, div=0.02):
    ref = rng.integers(0, 4, length, dtype=np.uint8)
    unit = rng.integers(0, 4, unit_len, dtype=np.uint8)
    for c in range(copies):
        u = unit.copy()
        m = rng.random(unit_len) < div
        u[m] = (u[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) & 3
        ref[at + c * unit_len: at + (c + 1) * unit_len] = u
    return ref


def main():
    md5s = {}
    td = tempfile.mkdtemp()
    # --- config 0: the reference's bundled inputs -------------------------------------------------
    for preset in ("map-ont", "asm20"):
        run("mt_" + preset, ["-x", preset, TEST + "/MT-human.fa", TEST + "/MT-orang.fa"], md5s)
        run("inv_" + preset, ["-x", preset, TEST + "/t-inv.fa", TEST + "/q-inv.fa"], md5s)
    run("inv_n1m5", ["-x", "map-ont", "-n", "1", "-m", "5", TEST + "/t-inv.fa", TEST + "/q-inv.fa"], md5s)
    # --- config 1/2 shape, small: synthetic ONT / CCS reads vs a random reference ---------------------
    ref = seqsim.gen_reference(4_000_000, seed=1)
    seqsim.write_fasta(td + "/ref.fa", [("chrS", ref)])
    seqsim.write_fasta(td + "/ont.fa", seqsim.gen_reads(ref, 24, 10000, 0.10, seed=7))
    seqsim.write_fasta(td + "/ccs.fa", seqsim.gen_reads(ref, 6, 15000, 0.01, seed=8))
    run("syn_ont", ["-x", "map-ont", td + "/ref.fa", td + "/ont.fa"], md5s)
    run("syn_ccs", ["-x", "asm20", td + "/ref.fa", td + "/ccs.fa"], md5s)
    run("syn_ont_n1m5", ["-x", "map-ont", "-n", "1", "-m", "5", td + "/ref.fa", td + "/ont.fa"], md5s)
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
    run("tandem", ["-x", "map-ont", "-f", "100000", td + "/tref.fa", td + "/tq.fa"], md5s)
    run("tandem_iter64", ["-x", "map-ont", "-f", "100000", "--max-chain-iter", "64", "--max-chain-skip", "5", td + "/tref.fa", td + "/tq.fa"], md5s)
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
    run("sr_paired", ["-x", "sr", td + "/sref.fa", td + "/r1.fa", td + "/r2.fa"], md5s)
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
    run("splice", ["-x", "splice", td + "/sref.fa", td + "/sp.fa"], md5s)
    run("syn_ont_gapscale", ["-x", "map-ont", "--chain-gap-scale", "1.7", "-r", "2000", td + "/ref.fa", td + "/ont.fa"], md5s)
    seqsim.write_fasta(td + "/ovl.fa", seqsim.gen_reads(ref[:50_000], 10, 8000, 0.08, seed=9))
    run("ava", ["-x", "ava-ont", td + "/ovl.fa", td + "/ovl.fa"], md5s)
    shutil.rmtree(td)
    with open(os.path.join(GOLD, "paf_md5.json"), "w") as fh:
        json.dump(md5s, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
