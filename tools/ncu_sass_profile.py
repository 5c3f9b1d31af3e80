"""Per-SASS-instruction execution counts of one kernel from an ncu report, grouped by source line and as a linear listing.
   python tools/ncu_sass_profile.py <report.ncu-rep> [n_anchors] > listing.txt
Columns of the listing: source line, address, executed warp-instructions (per anchor if n_anchors given), stall samples, SASS."""
import csv, io, subprocess, sys
rep = sys.argv[1]
n_anchors = float(sys.argv[2]) if len(sys.argv) > 2 else None
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] in ("Address", "Line No") and "Instructions Executed" in r)
hdr = rows[hi]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot = 0
out = []
for r in rows[hi + 1:]:
    if len(r) <= iex or not r[iex]:
        continue
    try:
        ex = int(r[iex])
    except ValueError:
        continue
    tot += ex
    out.append((r[ia], ex, int(r[ismp] or 0), r[isrc]))
print("# total executed warp-instructions (all captured launches): %d" % tot)
scale = 1.0
if n_anchors:
    # the source page sums the captured launches
    import re
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    n_launch = max(1, len(list(csv.reader(io.StringIO(raw)))) - 2)
    scale = 1.0 / (n_anchors * n_launch)
    print("# launches %d, per-anchor scale applied; total per anchor %.2f" % (n_launch, tot * scale))
for a, ex, smp, s in out:
    print("%s\t%10.4f\t%6d\t%s" % (a, ex * scale, smp, s))
