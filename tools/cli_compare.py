"""Reference CLI end to end: software chaining vs the per-read drop-in vs the phase-split caller (python tools/cli_compare.py [n_reads] [preset])."""
import os, sys, time, subprocess, hashlib, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
import bench_workloads as BW
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
preset = sys.argv[2] if len(sys.argv) > 2 else "map-ont"
td = tempfile.mkdtemp()
t0 = time.time()
threads = os.cpu_count() or 8
fa, mmi = BW._reference_files(preset, load_package("seqsim"), threads)
q = td + "/q.fa"
BW._simulate_reads(preset, n_reads, 11, q, procs=min(threads, 16))
print("inputs ready in %.1f s (%d reads, %s)" % (time.time() - t0, n_reads, preset), flush=True)
R = "oracle/_ref/"
def run(exe, t, env=None):
    t0 = time.time()
    p = subprocess.run([R + exe] + BW.PRESETS[preset][2] + ["-t", str(t), mmi, q], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=dict(os.environ, **(env or {})))
    dt = time.time() - t0
    err = p.stderr.decode().splitlines()
    tr = [l for l in err if "batcher" in l][-1:] + [l for l in err if "front end:" in l or "init:" in l] + [l for l in err if "mapped" in l or "loaded/built" in l or "Real time" in l]
    return dt, hashlib.md5(p.stdout).hexdigest(), p.stdout.count(b"\n"), tr
runs = [("minimap2-sw", threads, None), ("minimap2-b200-batch", threads, None), ("minimap2-b200-batch", threads, {"MM2B_TRACE": "1"}), ("minimap2-b200-batch", threads, {"MM2B_FRONT": "0"}),
        ("minimap2-b200", 256, None), ("minimap2-sw", threads, None)]
for exe, t, env in runs:
    dt, md5, lines, tr = run(exe, t, env)
    print("%-20s -t %-3d %-18s wall %.2f s  %d PAF lines  md5 %s" % (exe, t, env or "", dt, lines, md5[:8]), flush=True)
    for l in tr: print("      " + l)
