import sys, time, threading, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from __graft_entry__ import load_package
b = load_package("binding"); wl = load_package("workload")
off, a = wl.synth_anchor_batch(800, seed=5)
t0 = time.time(); b.init(1); print("init %.3f s" % (time.time() - t0), flush=True)
reads = [np.ascontiguousarray(a[off[r]:off[r+1]]) for r in range(800)]
par = b.Params()
t0 = time.time(); b.chain_read(par, reads[0]); print("first call %.3f s" % (time.time() - t0), flush=True)
t0 = time.time()
for r in range(200): b.chain_read(par, reads[r])
dt = time.time() - t0
print("single thread: %.3f ms per call (mean n=%d)" % (dt / 200 * 1e3, np.mean([len(x) for x in reads[:200]])), flush=True)
for nt in (4, 16, 64):
    def work(t):
        for r in range(t, 800, nt): b.chain_read(par, reads[r])
    th = [threading.Thread(target=work, args=(t,)) for t in range(nt)]
    t0 = time.time()
    for t in th: t.start()
    for t in th: t.join()
    dt = time.time() - t0
    print("%d threads: %.3f ms per read, %.0f reads/s" % (nt, dt / 800 * 1e3, 800 / dt), flush=True)
b.shutdown()
