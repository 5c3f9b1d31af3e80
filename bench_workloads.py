"""Bench inputs: real seeding instead of an anchor model.

The chaining hot path starts where the reference's seeding ends (mm_map_frag, map.c:291-316), so the most faithful input for
its benchmark is what the reference hands to mm_chain_dp for the reads BASELINE.json describes: a numpy-simulated reference and
read set (minimap2-fpga_b200/seqsim.py) is written to FASTA, the reference's own CLI — built in place
from /root/reference by oracle/Makefile, software chaining — sketches, seeds and chains it ONCE while the capture shim
(oracle/dump_shim.c, MM2_DUMP) records every mm_chain_dp call, and the recorded anchors become the CSR batch the bench replays.
The recorded outputs of the reference (n_u, u[], a per-read hash of b[]) ride along so that bench.py can check the GPU's results
against the reference's own on the full workload.  Nothing here is timed, and nothing here is on the product path: it is the
data generator.  Results are cached on disk (MM2B_CACHE_DIR, default /tmp/mm2b_cache) so that the reference arm, the B200 arm and
the scaling sweep on one box generate each input once.

Presets (BASELINE.json configs[1..3]):
  map-ont     10 kb mean, 10 % error, -x map-ont           100 Mbp i.i.d. reference
  asm20       15 kb mean,  1 % error, -x asm20             same reference
  ultralong  120 kb mean, 10 % error, -x map-ont           same reference (7.5k anchors/read: no repeats to seed from)
  tandem     100 kb reads across 24 kb tandem arrays, 6 % error, -x map-ont -f 100000: >50k anchors/read, windows at the max_iter clamp
"""
import fcntl
import hashlib
import json
import os
import struct
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

ANCHOR = np.dtype([("x", "<u8"), ("y", "<u8")])
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "minimap2-sw")
CACHE = os.environ.get("MM2B_CACHE_DIR", "/tmp/mm2b_cache")
GENOME = 100_000_000
_HDR = struct.Struct("<II9ifqii")

PRESETS = {
    # name: (mean read length, error rate, CLI arguments, default reads per GPU)
    "map-ont": (10000, 0.10, ["-x", "map-ont"], 100000),
    "asm20": (15000, 0.01, ["-x", "asm20"], 50000),
    "ultralong": (120000, 0.10, ["-x", "map-ont"], 2000),
    "tandem": (100000, 0.06, ["-x", "map-ont", "-f", "100000"], 96),
}
DESCRIPTION = {
    "map-ont": "map-ont (BASELINE configs[1]): %d synthetic ONT reads per GPU (10 kb mean, ~10%% error) vs 100 Mbp random reference",
    "asm20": "asm20 (BASELINE configs[2]): %d synthetic CCS-like reads per GPU (15 kb mean, ~1%% error) vs 100 Mbp random reference",
    "ultralong": "ultra-long map-ont (BASELINE configs[3]): %d synthetic ONT reads per GPU (120 kb mean, ~10%% error) vs 100 Mbp random reference",
    "tandem": "ultra-long map-ont across tandem repeats (BASELINE configs[3]): %d reads of 100 kb over 24 kb arrays (400 bp unit x 60), -f 100000: "
              ">50k anchors per read, windows at the max_iter clamp",
}


def available():
    return os.path.exists(REF_CLI)


class _Lock:
    def __init__(self, path):
        self.path = path

    def __enter__(self):
        os.makedirs(os.path.dirname(self.path), exist_ok=True)
        self.fh = open(self.path, "w")
        fcntl.flock(self.fh, fcntl.LOCK_EX)

    def __exit__(self, *exc):
        fcntl.flock(self.fh, fcntl.LOCK_UN)
        self.fh.close()


def _tandem_reference(seqsim):
    """10 Mbp random reference with 24 kb tandem arrays (400 bp unit, 60 copies, 2 % divergence between copies) every 100 kb."""
    rng = np.random.default_rng(5)
    ref = rng.integers(0, 4, 10_000_000, dtype=np.uint8)
    starts = list(range(60_000, len(ref) - 200_000, 100_000))
    for at in starts:
        unit = rng.integers(0, 4, 400, dtype=np.uint8)
        for c in range(60):
            u = unit.copy()
            m = rng.random(400) < 0.02
            u[m] = (u[m] + rng.integers(1, 4, int(m.sum()), dtype=np.uint8)) & 3
            ref[at + c * 400: at + (c + 1) * 400] = u
    return ref, starts


def _reference_files(name, seqsim, threads):
    """FASTA + .mmi of the preset's reference, built once per cache directory."""
    kind = "tandem" if name == "tandem" else "random100m"
    preset = PRESETS[name][2][1]
    fa, mmi = os.path.join(CACHE, kind + ".fa"), os.path.join(CACHE, "%s.%s.mmi" % (kind, preset))
    with _Lock(os.path.join(CACHE, kind + ".lock")):
        if not os.path.exists(fa):
            ref = _tandem_reference(seqsim)[0] if kind == "tandem" else seqsim.gen_reference(GENOME, seed=1)
            seqsim.write_fasta(fa + ".tmp", [("chr1", ref)])
            os.replace(fa + ".tmp", fa)
        if not os.path.exists(mmi):
            subprocess.run([REF_CLI, "-x", preset, "-t", str(threads), "-d", mmi + ".tmp", fa], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
            os.replace(mmi + ".tmp", mmi)
    return fa, mmi


def _gen_chunk(args):
    name, n, mean_len, err, seed = args
    seqsim = load_package("seqsim")
    if name == "tandem":
        ref, starts = _tandem_reference(seqsim)
        rng = np.random.default_rng(seed)
        reads = []
        for i in range(n):
            at = starts[int(rng.integers(0, len(starts)))]
            s = at - int(rng.integers(20_000, 56_000))          # the read covers the whole 24 kb array and 76 kb of unique flanks
            seq = ref[s:s + mean_len]
            if rng.integers(0, 2):
                seq = seqsim.revcomp(seq)
            reads.append(("t%d_%d" % (seed, i), seqsim.mutate(seq, err, rng)))
        return [(nm, seqsim.to_ascii(sq)) for nm, sq in reads]
    ref = np.fromfile(os.path.join(CACHE, "random100m.u8"), dtype=np.uint8)
    return [(nm, seqsim.to_ascii(sq)) for nm, sq in seqsim.gen_reads(ref, n, mean_len, err, seed=seed)]


def _simulate_reads(name, n_reads, seed, path, procs):
    from multiprocessing import get_context
    seqsim = load_package("seqsim")
    mean_len, err = PRESETS[name][0], PRESETS[name][1]
    if name != "tandem":
        raw = os.path.join(CACHE, "random100m.u8")
        with _Lock(os.path.join(CACHE, "random100m.lock")):
            if not os.path.exists(raw):
                seqsim.gen_reference(GENOME, seed=1).tofile(raw + ".tmp")
                os.replace(raw + ".tmp", raw)
    per = max(1, min(4000, (n_reads + procs - 1) // procs))
    jobs = [(name, min(per, n_reads - s), mean_len, err, seed * 100003 + k) for k, s in enumerate(range(0, n_reads, per))]
    with open(path, "wb") as fh:
        if procs > 1 and len(jobs) > 1:
            with get_context("fork").Pool(procs) as pool:
                for k, chunk in enumerate(pool.imap(_gen_chunk, jobs)):
                    for nm, seq in chunk:
                        fh.write(b">c%d_" % k + nm.encode() + b"\n" + seq + b"\n")
        else:
            for k, job in enumerate(jobs):
                for nm, seq in _gen_chunk(job):
                    fh.write(b">c%d_" % k + nm.encode() + b"\n" + seq + b"\n")


def _parse_dump(path):
    """Recorded mm_chain_dp calls -> CSR batch + the reference's own results."""
    buf = np.fromfile(path, dtype=np.uint8)
    mv = memoryview(buf)
    pos, n_tot = 0, len(buf)
    a_parts, u_parts, n_u, n_v, bh, par0 = [], [], [], [], [], None
    while pos < n_tot:
        (magic, flags, mdx, mdy, bw, skip, it, cnt, sc, cdna, segs, gs, n, nu, nv) = _HDR.unpack_from(mv, pos)
        assert magic == 0x4443324D, "bad dump record at %d" % pos
        par = (mdx, mdy, bw, skip, it, cnt, sc, cdna, segs, gs)
        par0 = par0 or par
        pos += 64
        a = np.frombuffer(mv, ANCHOR, n, pos)
        pos += 16 * n
        if flags & 1:
            pos += 12 * n
        u = np.frombuffer(mv, "<u8", nu, pos)
        pos += 8 * nu
        b = np.frombuffer(mv, "<u8", 2 * nv, pos)
        pos += 16 * nv
        if par != par0:
            continue                      # (a preset that re-chains with other arguments; none of the bench presets does)
        a_parts.append(a), u_parts.append(u), n_u.append(nu), n_v.append(nv)
        bh.append(int(np.bitwise_xor.reduce(b * np.uint64(0x9E3779B97F4A7C15) + np.arange(1, 2 * nv + 1, dtype=np.uint64))) if nv else 0)
    off = np.zeros(len(a_parts) + 1, np.int64)
    np.cumsum([len(x) for x in a_parts], out=off[1:])
    return dict(off=off, a=np.concatenate(a_parts) if a_parts else np.empty(0, ANCHOR), ref_n_u=np.asarray(n_u, np.int32), ref_n_v=np.asarray(n_v, np.int32),
                ref_u=np.concatenate(u_parts) if u_parts else np.empty(0, np.uint64), ref_b_hash=np.asarray(bh, np.uint64), par=par0)


def b_hash(b_words):
    """Per-read hash of b[] used by _parse_dump (b as a flat uint64 view: x0, y0, x1, y1, ...)."""
    n = len(b_words)
    return int(np.bitwise_xor.reduce(b_words * np.uint64(0x9E3779B97F4A7C15) + np.arange(1, n + 1, dtype=np.uint64))) if n else 0


def real_seed_batch(name, n_reads, seed, threads=None, verbose=False):
    """Return dict(off, a, ref_n_u, ref_n_v, ref_u, ref_b_hash, par, meta) for `n_reads` reads of preset `name`."""
    if not available():
        raise RuntimeError("bench_workloads: %s is missing (oracle/Makefile builds it where /root/reference exists; the built binary travels with the repo)" % REF_CLI)
    threads = threads or os.cpu_count() or 1
    key = "%s_r%d_s%d" % (name, n_reads, seed)
    d = os.path.join(CACHE, key)
    t0 = time.time()
    with _Lock(os.path.join(CACHE, key + ".lock")):
        if not os.path.exists(os.path.join(d, "meta.json")):
            os.makedirs(d, exist_ok=True)
            seqsim = load_package("seqsim")
            fa, mmi = _reference_files(name, seqsim, threads)
            q = os.path.join(d, "reads.fa")
            _simulate_reads(name, n_reads, seed, q, procs=max(1, min(threads, 16)))
            t_sim = time.time() - t0
            dump = os.path.join(d, "dump.bin")
            env = dict(os.environ, MM2_DUMP=dump, MM2_DUMP_NO_FPV="1")
            cmd = [REF_CLI] + PRESETS[name][2] + ["-t", str(threads), mmi, q]
            p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, check=True)
            paf_md5, paf_lines = hashlib.md5(p.stdout).hexdigest(), p.stdout.count(b"\n")
            t_map = time.time() - t0 - t_sim
            w = _parse_dump(dump)
            for k in ("off", "a", "ref_n_u", "ref_n_v", "ref_u", "ref_b_hash"):
                np.save(os.path.join(d, k + ".npy"), w[k])
            meta = dict(preset=name, reads_requested=n_reads, calls_recorded=len(w["off"]) - 1, anchors=int(w["off"][-1]), seed=seed, par=list(w["par"]),
                        cli=" ".join(["minimap2-sw"] + PRESETS[name][2]), paf_md5=paf_md5, paf_lines=paf_lines, simulate_s=round(t_sim, 1), seed_and_chain_s=round(t_map, 1))
            os.remove(dump)
            if name not in ("map-ont", "asm20"):    # these keep their reads: the seeding front end maps the sequences themselves
                os.remove(q)
            json.dump(meta, open(os.path.join(d, "meta.json"), "w"))
    meta = json.load(open(os.path.join(d, "meta.json")))
    out = {k: np.load(os.path.join(d, k + ".npy")) for k in ("off", "a", "ref_n_u", "ref_n_v", "ref_u", "ref_b_hash")}
    out["par"], out["meta"] = tuple(meta["par"]), meta
    meta["load_s"] = round(time.time() - t0, 1)
    if verbose:
        print("[bench_workloads] %s: %s" % (key, meta), file=sys.stderr, flush=True)
    return out


SEED_TOOL = os.path.join(ROOT, "oracle", "_ref", "mm2-seed-ref")


def front_inputs(name, n_reads, seed, threads=None):
    """Inputs of the seeding front end for the same reads real_seed_batch(name, n_reads, seed) recorded: dict(seq_off, seq (uint8), index
    (flat arrays as the product's mm2b_index_flatten writes them, via oracle/_ref/mm2-seed-ref), mid_occ) — or None when the reads were
    not kept or the tool is missing."""
    threads = threads or os.cpu_count() or 1
    d = os.path.join(CACHE, "%s_r%d_s%d" % (name, n_reads, seed))
    q = os.path.join(d, "reads.fa")
    if not (os.path.exists(q) and os.path.exists(SEED_TOOL)):
        return None
    seqsim = load_package("seqsim")
    fa, mmi = _reference_files(name, seqsim, threads)
    flat = mmi + ".flat"
    with _Lock(mmi + ".flat.lock"):
        if not os.path.exists(flat):
            empty = os.path.join(CACHE, "empty.fa")
            open(empty, "w").close()
            subprocess.run([SEED_TOOL, PRESETS[name][2][1], mmi, empty, flat + ".seeds", flat + ".tmp"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            os.replace(flat + ".tmp", flat)
    from oracle import seed_py
    idx = seed_py.read_index(flat)
    buf = np.fromfile(q, dtype=np.uint8)
    nl = np.flatnonzero(buf == 10)                    # two lines per read: >name, sequence
    starts, ends = nl[0::2] + 1, nl[1::2]
    lens = (ends - starts).astype(np.int64)
    seq_off = np.zeros(len(lens) + 1, np.int64)
    np.cumsum(lens, out=seq_off[1:])
    seq = np.empty(int(seq_off[-1]), np.uint8)
    for i in range(len(lens)):                        # (slice copies: the name lines and newlines stay behind)
        seq[seq_off[i]:seq_off[i + 1]] = buf[starts[i]:ends[i]]
    return dict(seq_off=seq_off, seq=seq, index=idx, mid_occ=idx["mid_occ"])


if __name__ == "__main__":
    nm = sys.argv[1] if len(sys.argv) > 1 else "map-ont"
    w = real_seed_batch(nm, int(sys.argv[2]) if len(sys.argv) > 2 else PRESETS[nm][3], 1000, verbose=True)
    n = np.diff(w["off"])
    print("reads %d anchors %d (mean %.0f, max %d) chains %d chained %.3f" % (len(n), w["off"][-1], n.mean(), n.max(), int(w["ref_n_u"].sum()), w["ref_n_v"].sum() / max(1, w["off"][-1])))
