#!/usr/bin/env python
"""One-time calibration of the anchor-level workload model (minimap2-fpga_b200/workload.py) against REAL minimap2 seeding.

Simulates reads at the sequence level (seqsim.py, BASELINE.json shape: 100 Mbp random reference, ONT 10 kb / 10 % error, CCS
15 kb / 1 % error, ultra-long 120 kb), maps them with the reference CLI built in place (oracle/_ref/minimap2-sw, -x map-ont /
asm20), captures every mm_chain_dp call, and records the statistics the chaining kernel's cost depends on.  The result,
workload_calibration.json, is what tests/test_workload_model.py holds the anchor model to.  Build container only (~5 min).
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402
from oracle import dumpio, oracle_py as O  # noqa: E402

seqsim = load_package("seqsim")
CLI = os.path.join(ROOT, "oracle", "_ref", "minimap2-sw")


def stats_of(recs):
    n = np.array([len(r["a"]) for r in recs])
    off, a = dumpio.to_batch(recs)
    r = O.replay(recs[0]["par"], off, a, n_threads=8)
    st = r["stats"]
    return dict(reads=len(recs), anchors_per_read=float(n.mean()), max_anchors=int(n.max()), cells_per_anchor=st.cells / len(a),
                window_cells_per_anchor=st.window_cells / len(a), chained_fraction=st.n_chained / len(a), chains_per_read=st.n_chains / len(recs))


def main():
    out = {}
    with tempfile.TemporaryDirectory() as td:
        ref = seqsim.gen_reference(100_000_000, seed=1)
        seqsim.write_fasta(td + "/ref.fa", [("chr1", ref)])
        for name, preset, n_reads, mean_len, err in (("map-ont", "map-ont", 400, 10000, 0.10), ("asm20", "asm20", 150, 15000, 0.01),
                                                     ("ultralong", "map-ont", 24, 120000, 0.10)):
            seqsim.write_fasta(td + "/q.fa", seqsim.gen_reads(ref, n_reads, mean_len, err, seed=11))
            dump = td + "/d.bin"
            subprocess.run([CLI, "-t", "8", "-x", preset, td + "/ref.fa", td + "/q.fa"], env=dict(os.environ, MM2_DUMP=dump),
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
            out[name] = stats_of(dumpio.read_dump(dump))
            print(name, out[name], flush=True)
    with open(os.path.join(HERE, "workload_calibration.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
