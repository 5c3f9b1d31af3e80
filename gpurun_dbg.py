import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from __graft_entry__ import load_package
import fuzz
from oracle import oracle_py as O
b = load_package("binding")
kw = dict(gap_scale=1.7, bw=2000)
off, a = fuzz.mixed_batch(204, n_reads=64)
for kw in (dict(gap_scale=1.7, bw=2000), dict(bw=2000), dict(gap_scale=1.7)):
    db = b.DeviceBatch(b.Params(**kw), off, a, keep_fpv=True); db.run(); f,p,v = db.fpv(); res = db.results()
    nbad = 0
    for r in range(len(off)-1):
        ar = a[off[r]:off[r+1]]
        o = O.chain(O.Params(**kw), ar, want_fpv=True)
        s,e = off[r], off[r+1]
        bad = np.nonzero((f[s:e]!=o['f'])|(p[s:e]!=o['p'])|(v[s:e]!=o['v']))[0]
        if len(bad):
            nbad += 1
            i = bad[0]
            if nbad <= 2:
                print(kw, 'read', r, 'n', e-s, 'first bad i', i, 'gpu f,p,v', f[s+i], p[s+i], v[s+i], 'ref', o['f'][i], o['p'][i], o['v'][i], 'nbad', len(bad))
                pj, pr = p[s+i], o['p'][i]
                for j in (pj, pr):
                    if j >= 0: print('   cand j', j, 'x', ar['x'][j]&0xffffffff, 'y', ar['y'][j]&0xffffffff, 'f', o['f'][j], ' i: x', ar['x'][i]&0xffffffff,'y', ar['y'][i]&0xffffffff)
    print(kw, 'reads with fpv mismatch', nbad)
