"""Writes tests/golden/seed_golden.npz: what the REFERENCE's own seeding (oracle/_ref/mm2-seed-ref: collect_minimizers + collect_seed_hits
compiled from /root/reference) produces for a small generated read set — per read the sha1 of its minimizers and of its sorted
anchors, rep_len and the number of mini_pos entries — together with the inputs (sequences, flat index).  Run once where
/root/reference exists:  python tests/golden/make_seed_golden.py"""
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import seedgen  # noqa: E402
from oracle import seed_py as S  # noqa: E402

chroms = seedgen.make_reference(seed=21, scale=0.12)
reads = seedgen.make_reads(chroms, n_reads=60, seed=22, mean_len=2500)
with tempfile.TemporaryDirectory() as td:
    ref, q, sf, xf = (os.path.join(td, n) for n in ("ref.fa", "reads.fa", "s.bin", "i.bin"))
    seedgen.write_fasta(ref, chroms), seedgen.write_fasta(q, reads)
    S.run_tool("map-ont", ref, q, sf, xf, 20)
    rec, flat = S.read_seeds(sf), S.read_index(xf)
off = np.zeros(len(reads) + 1, np.int64)
np.cumsum([len(s) for _, s in reads], out=off[1:])
np.savez_compressed(os.path.join(HERE, "seed_golden.npz"), k=rec["k"], w=rec["w"], mid_occ=rec["mid_occ"], keys=flat["keys"], vals=flat["vals"], pos=flat["pos"],
                    seq=np.frombuffer(b"".join(s.tobytes() for _, s in reads), np.uint8), seq_off=off,
                    mv_sha=np.array([hashlib.sha1(r["mv"].tobytes()).hexdigest() for r in rec["reads"]]),
                    a_sha=np.array([hashlib.sha1(r["a"].tobytes()).hexdigest() for r in rec["reads"]]),
                    rep_len=np.array([r["rep_len"] for r in rec["reads"]], np.int32), n_mini_pos=np.array([len(r["mini_pos"]) for r in rec["reads"]], np.int32),
                    n_a=np.array([len(r["a"]) for r in rec["reads"]], np.int64))
print("reads", len(reads), "anchors", sum(len(r["a"]) for r in rec["reads"]), "reads with equal keys",
      sum(int(len(r["a"]) > 1 and bool(np.any(r["a"]["x"][1:] == r["a"]["x"][:-1]))) for r in rec["reads"]))
