import os, sys, time, subprocess, hashlib, tempfile
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from __graft_entry__ import load_package
seqsim = load_package("seqsim")
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
td = tempfile.mkdtemp()
t0 = time.time()
ref = seqsim.gen_reference(100_000_000, seed=1)
seqsim.write_fasta(td + "/ref.fa", [("chr1", ref)])
seqsim.write_fasta(td + "/q.fa", seqsim.gen_reads(ref, n_reads, 10000, 0.10, seed=11))
print("inputs written in %.1f s" % (time.time() - t0), flush=True)
# index once so that both runs only map
subprocess.run(["oracle/_ref/minimap2-sw", "-x", "map-ont", "-d", td + "/ref.mmi", td + "/ref.fa"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
def run(exe, t, env=None):
    t0 = time.time()
    p = subprocess.run([exe, "-x", "map-ont", "-t", str(t), td + "/ref.mmi", td + "/q.fa"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=dict(os.environ, **(env or {})))
    dt = time.time() - t0
    err = p.stderr.decode().splitlines()
    tr = [l for l in err if "batcher" in l or "chaining calls" in l][-2:] + [l for l in err if "mapped" in l or "loaded/built" in l or "Real time" in l]
    return dt, hashlib.md5(p.stdout).hexdigest(), p.stdout.count(b"\n"), tr
runs = [("oracle/_ref/minimap2-sw", 16, None), ("oracle/_ref/minimap2-b200", 16, {"MM2B_TRACE": "1"}), ("oracle/_ref/minimap2-b200", 128, {"MM2B_TRACE": "1"}),
        ("oracle/_ref/minimap2-b200", 512, {"MM2B_TRACE": "1"})]
if len(sys.argv) > 2 and sys.argv[2] == "fiber":     # the reference CLI on the fiber-based kt_for (host/fiber_for.cpp): -t = reads in flight
    runs = [("oracle/_ref/minimap2-sw", 16, None), ("oracle/_ref/minimap2-fiber-b200", 512, {"MM2B_TRACE": "1"}), ("oracle/_ref/minimap2-fiber-b200", 2048, {"MM2B_TRACE": "1"})]
for exe, t, env in runs:
    dt, md5, lines, tr = run(exe, t, env)
    print("%-16s -t %-3d wall %.2f s  %d PAF lines  md5 %s" % (exe.split("/")[-1], t, dt, lines, md5[:8]), flush=True)
    for l in tr: print("      " + l)
