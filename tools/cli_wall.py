"""Where the wall time of the phase-split CLI goes: python tools/cli_wall.py [n_reads] — minimap2-sw vs minimap2-b200-batch (default, MM2B_RESERVE=0), three runs each,
wall from the shell and `Real time` from minimap2 itself, with the library's own start-up / shutdown trace."""
import os, sys, time, subprocess, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
import bench_workloads as BW
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
threads = os.cpu_count() or 8
fa, mmi = BW._reference_files("map-ont", load_package("seqsim"), threads)
q = tempfile.mkdtemp() + "/q.fa"
BW._simulate_reads("map-ont", n_reads, 11, q, procs=min(threads, 16))
for exe, env in (("minimap2-sw", {}), ("minimap2-b200-batch", {"MM2B_TRACE": "1"}), ("minimap2-b200-batch", {"MM2B_TRACE": "1", "MM2B_RESERVE": "0"}), ("minimap2-sw", {})):
    for rep in range(3):
        t0 = time.time()
        p = subprocess.run(["oracle/_ref/" + exe, "-x", "map-ont", "-t", str(threads), mmi, q], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=dict(os.environ, **env))
        dt = time.time() - t0
        err = p.stderr.decode().splitlines()
        real = [l for l in err if "Real time" in l]
        extra = [l.split("] ")[-1] for l in err if "shutdown:" in l or "init: contexts" in l or "loaded/built" in l]
        print("%-20s %-22s wall %.2f s | %s | %s" % (exe, env.get("MM2B_RESERVE", ""), dt, real[0].split(";")[0] if real else "", " ; ".join(extra)), flush=True)
