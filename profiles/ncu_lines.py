#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: executed warp-instructions and stall samples.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv ; python ncu_lines.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
# first kernel instance only
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
end = next((i for i in range(hi + 1, len(rows)) if rows[i] and rows[i][0] == "File Path"), len(rows))
hdr = rows[hi]
ci, smp = hdr.index("Instructions Executed"), hdr.index("# Samples")
agg = {}
for r in rows[hi + 1:end]:
    if len(r) <= ci or r[2] != "-":      # keep the per-line summary rows (Address == "-")
        continue
    try:
        agg[int(r[0])] = (int(r[ci]), int(r[smp]), r[1])
    except ValueError:
        pass
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("total warp-instructions %d, samples %d" % (tot, ts))
for ln, (n, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.2f%% inst %5.2f%% smp  L%-4d %s" % (100.0 * n / tot, 100.0 * s / max(ts, 1), ln, src.strip()[:120]))
