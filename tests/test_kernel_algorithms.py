"""CPU checks of the warp-parallel reformulations the CUDA kernels rely on (csrc/chain_kernels.cu), each against the plain
sequential statement of the reference loop it replaces (chain.c:197-235, chain.c:374-391, chain.c:192).  Small pure-Python
models on seeded random data: they pin the *arguments* the kernels' exactness rests on, so that a GPU is only needed to check the
CUDA transcription (tests/test_gpu_*.py), not the mathematics.

  1. one chunk of 32 cells: records from a running strict-'>' max, hits, and the saturating n_skip counter in closed form
     (Lindley recursion on vote masks), including the break lane;
  2. in-chunk visit stamps as a one-hot OR over the lanes;
  3. a scan folded chunk by chunk from state-independent summaries (the heavy-read kernel's rounds);
  4. the greedy backtrack taken 32 anchors per step by pointer jumping inside aligned blocks;
  5. the window start as a lower bound found by halving steps / by merging batches of 32 candidates.
"""
import numpy as np
import pytest

INT_MIN = -(1 << 31)


# ---------------------------------------------------------------------------------------------------------------
# sequential statements (the reference's loops, cell by cell)
# ---------------------------------------------------------------------------------------------------------------
def seq_chunk(sc, hit, max_f, max_j, n_skip, max_skip, jt):
    """chain.c:226-233 over one chunk: sc[l] is INT_MIN for `continue`d cells; hit[l] = t[j] == i.  Returns the new state,
    whether the loop broke and at which lane."""
    for lane in range(32):
        if sc[lane] == INT_MIN:
            continue
        if sc[lane] > max_f:
            max_f, max_j = sc[lane], jt - lane
            if n_skip > 0:
                n_skip -= 1
        elif hit[lane]:
            n_skip += 1
            if n_skip > max_skip:
                return max_f, max_j, n_skip, True, lane
    return max_f, max_j, n_skip, False, 32


def popc(m):
    return bin(m & 0xFFFFFFFF).count("1")


def lowest_lane(m):
    return (m & -m).bit_length() - 1


# ---------------------------------------------------------------------------------------------------------------
# 1. one chunk resolved from vote masks (scan_predecessors cases (i)-(iii) / resolve_records)
# ---------------------------------------------------------------------------------------------------------------
def vote_chunk(sc, hit, max_f, max_j, n_skip, max_skip, jt):
    valid = [s != INT_MIN for s in sc]
    hitv = sum(1 << l for l in range(32) if valid[l] and hit[l])
    cand = sum(1 << l for l in range(32) if sc[l] > max_f)
    if cand == 0:                                           # (i) no record: the counter only goes up
        x0 = n_skip
        n_skip += popc(hitv)
        if n_skip > max_skip:
            brk = next(l for l in range(32) if hitv >> l & 1 and x0 + popc(hitv & ((2 << l) - 1)) == max_skip + 1)
            return max_f, max_j, n_skip, True, brk
        return max_f, max_j, n_skip, False, 32
    top = max(sc)
    last = sc.index(top)                                    # first occurrence = nearest j
    recmask = 1 << last
    if cand & (recmask - 1) == 0:                           # (ii) one record
        hm = hitv & ~recmask
        h1, h2 = hm & (recmask - 1), hm & ~(recmask - 1)
        x1 = n_skip + popc(h1)
        x2 = x1 - 1 if x1 > 0 else 0
        x3 = x2 + popc(h2)
        early = x1 > max_skip
        broke = early or x3 > max_skip
        if not early:
            max_f, max_j = top, jt - last
        brk = 32
        if broke:
            run = h1 if early else h2
            k = max(1, max_skip + 1 - (n_skip if early else x2))
            brk = next(l for l in range(32) if run >> l & 1 and popc(run & ((2 << l) - 1)) == k)
        return max_f, max_j, x3, broke, brk
    r = lowest_lane(cand)                                   # (iii) several records: walk the prefix maxima
    while r != last:
        recmask |= 1 << r
        above = sum(1 << l for l in range(32) if sc[l] > sc[r]) & (0xFFFFFFFE << r)
        r = lowest_lane(above)
    hm = hitv & ~recmask
    take, broke, brk = recmask, False, 32
    if hm == 0:
        n_skip = max(0, n_skip - popc(recmask))
    else:
        floor_all, done, corr = 0, 0, [0] * 32
        rm = recmask
        while rm:
            below = (rm - 1) & ~rm
            done += 1
            s_r = n_skip + popc(hm & below) - done
            floor_all = min(floor_all, s_r)
            for l in range(32):
                if not below >> l & 1:
                    corr[l] = min(corr[l], s_r)
            rm &= rm - 1
        over = 0
        for l in range(32):
            le = (2 << l) - 1
            x = n_skip + popc(hm & le) - popc(recmask & le) - corr[l]
            if hm >> l & 1 and x > max_skip:
                over |= 1 << l
        if over:
            broke, brk = True, lowest_lane(over)
            take = recmask & ((1 << brk) - 1)
        else:
            n_skip = n_skip + popc(hm) - done - floor_all
    if take:
        l2 = take.bit_length() - 1
        max_f, max_j = sc[l2], jt - l2
    return max_f, max_j, n_skip, broke, brk


def random_chunk(rng, dense):
    n_valid = rng.integers(0, 33)
    sc = [INT_MIN] * 32
    for l in rng.choice(32, n_valid, replace=False):
        sc[l] = int(rng.integers(-5, 12)) if dense else int(rng.integers(-50, 400))
    hit = [bool(rng.random() < (0.8 if dense else 0.3)) for _ in range(32)]
    return sc, hit


@pytest.mark.parametrize("dense", [True, False])
def test_chunk_from_vote_masks_equals_the_sequential_loop(dense):
    rng = np.random.default_rng(3 + dense)
    for _ in range(6000):
        sc, hit = random_chunk(rng, dense)
        max_f, n_skip, max_skip = int(rng.integers(-6, 12)), int(rng.integers(0, 30)), int(rng.integers(0, 30))
        n_skip = min(n_skip, max_skip)                      # the counter never exceeds max_skip between chunks
        ref = seq_chunk(sc, hit, max_f, -1, n_skip, max_skip, 1000)
        got = vote_chunk(sc, hit, max_f, -1, n_skip, max_skip, 1000)
        if ref[3]:
            assert got[3] and got[:2] == ref[:2] and got[4] == ref[4], (sc, hit, max_f, n_skip, max_skip, ref, got)
        else:
            assert got == ref, (sc, hit, max_f, n_skip, max_skip, ref, got)


# ---------------------------------------------------------------------------------------------------------------
# 2. stamps inside a chunk as a one-hot OR (and memory stamps only for later chunks)
# ---------------------------------------------------------------------------------------------------------------
def test_one_hot_stamps_equal_t_array_stamps():
    rng = np.random.default_rng(5)
    for _ in range(300):
        i = int(rng.integers(40, 400))
        st = int(rng.integers(0, i))
        p = [int(rng.integers(-1, j)) if j > 0 else -1 for j in range(i)]       # p[j] < j
        valid = [bool(rng.random() < 0.6) for _ in range(i)]
        t = [-1] * i                                                             # reference: t[p[j]] = i after visiting valid cell j
        ref_hit = {}
        for j in range(i - 1, st - 1, -1):
            if not valid[j]:
                continue
            ref_hit[j] = t[j] == i
            if p[j] >= 0:
                t[p[j]] = i
        mem = [-1] * i                                                           # kernel: per chunk, one-hot OR + stamps of earlier chunks
        jt = i - 1
        while jt >= st:
            hot = 0
            for lane in range(32):
                j = jt - lane
                if j >= st and valid[j] and 0 <= jt - p[j] < 32:
                    hot |= 1 << (jt - p[j])
            for lane in range(32):
                j = jt - lane
                if j >= st and valid[j]:
                    assert (bool(hot >> lane & 1) or mem[j] == i) == ref_hit[j]
            for lane in range(32):                                               # the scan moves on: stamps for later chunks only
                j = jt - lane
                if j >= st and valid[j] and st <= p[j] < jt - 31:
                    mem[p[j]] = i
            jt -= 32


# ---------------------------------------------------------------------------------------------------------------
# 3. a scan folded from state-independent chunk summaries (chain_heavy_kernel / coop_scan)
# ---------------------------------------------------------------------------------------------------------------
def fold_scan(chunks, q_span, max_skip, i, warps):
    max_f, max_j, n_skip = q_span, -1, 0
    for g0 in range(0, len(chunks), warps):
        rnd = chunks[g0:g0 + warps]
        summ = [(max(sc), sum(1 for l in range(32) if sc[l] != INT_MIN and hit[l])) for sc, hit in rnd]
        k0 = 0
        while k0 < len(rnd):
            kr = next((k for k in range(k0, len(rnd)) if summ[k][0] > max_f), 32)
            acc, kb = n_skip, 32
            for k in range(k0, len(rnd)):
                acc += summ[k][1]
                if acc > max_skip:
                    kb = k
                    break
            if kb < kr:
                return max_f, max_j
            if kr == 32:
                n_skip += sum(s[1] for s in summ[k0:])
                break
            n_skip += sum(s[1] for s in summ[k0:kr])
            sc, hit = rnd[kr]
            max_f, max_j, n_skip, broke, _ = vote_chunk(sc, hit, max_f, max_j, n_skip, max_skip, i - 1 - 32 * (g0 + kr))
            if broke:
                return max_f, max_j
            k0 = kr + 1
    return max_f, max_j


@pytest.mark.parametrize("warps", [1, 8, 16])
def test_scan_folded_from_chunk_summaries_equals_the_sequential_scan(warps):
    rng = np.random.default_rng(7 + warps)
    for _ in range(1500):
        n_chunks = int(rng.integers(1, 40))
        dense = bool(rng.random() < 0.5)
        chunks = [random_chunk(rng, dense) for _ in range(n_chunks)]
        if rng.random() < 0.5:                              # far chunks rarely beat the running max: the common shape
            chunks = chunks[:2] + [([s if s == INT_MIN else s - 400 for s in sc], hit) for sc, hit in chunks[2:]]
        q_span, max_skip, i = 15, int(rng.integers(0, 40)), 5000
        max_f, max_j, n_skip = q_span, -1, 0
        for c, (sc, hit) in enumerate(chunks):
            max_f, max_j, n_skip, broke, _ = seq_chunk(sc, hit, max_f, max_j, n_skip, max_skip, i - 1 - 32 * c)
            if broke:
                break
        assert fold_scan(chunks, q_span, max_skip, i, warps) == (max_f, max_j)


# ---------------------------------------------------------------------------------------------------------------
# 4. backtrack 32 anchors per step (extract_chains)
# ---------------------------------------------------------------------------------------------------------------
def seq_backtrack(p, ends):
    used, out = [False] * len(p), []
    for e in ends:
        j, path = e, []
        while True:                                          # do-while of chain.c:379-383
            path.append(j)
            used[j] = True
            j = p[j]
            if j < 0 or used[j]:
                break
        out.append((path, j))
    return out


def block_backtrack(p, ends):
    n = len(p)
    w = list(p)                                              # used-mark folded into the link word: -3 - p
    out = []
    for e in ends:
        cur, first, path = e, True, []
        while True:
            base, el = cur & ~31, cur & 31
            ww = [w[base + l] if base + l < n else -1 for l in range(32)]
            used = [x <= -2 for x in ww]
            pp = [-3 - x if x <= -2 else x for x in ww]
            if not first and used[el]:
                stop = cur
                break
            nxt = [pp[l] - base if pp[l] >= base and not used[pp[l] - base] else -1 for l in range(32)]
            mask = [1 << l for l in range(32)]
            while any(x >= 0 for x in nxt):                  # pointer jumping, all lanes at once
                m2 = [mask[nxt[l]] if nxt[l] >= 0 else 0 for l in range(32)]
                n2 = [nxt[nxt[l]] if nxt[l] >= 0 else -1 for l in range(32)]
                mask = [mask[l] | m2[l] for l in range(32)]
                nxt = n2
            stretch = mask[el]
            for l in range(31, -1, -1):                      # walk order = descending index
                if stretch >> l & 1:
                    path.append(base + l)
                    if not used[l]:
                        w[base + l] = -3 - ww[l]
            out_link = pp[lowest_lane(stretch)]
            first = False
            if out_link < 0 or out_link >= base:
                stop = out_link
                break
            cur = out_link
        out.append((path, stop))
    return out


def test_blockwise_backtrack_equals_the_serial_walk():
    rng = np.random.default_rng(11)
    for _ in range(400):
        n = int(rng.integers(1, 400))
        reach = int(rng.choice([1, 3, 40, 400]))
        p = [int(rng.integers(max(-1, j - reach), j)) if j > 0 and rng.random() < 0.9 else -1 for j in range(n)]
        ends = [int(e) for e in rng.choice(n, min(n, int(rng.integers(1, 12))), replace=True)]   # repeats: two ends sharing a peak
        assert block_backtrack(p, ends) == seq_backtrack(p, ends)


# ---------------------------------------------------------------------------------------------------------------
# 5. window start: halving-step lower bound, and the merge over batches of 32 candidates (dp_fill)
# ---------------------------------------------------------------------------------------------------------------
def test_window_start_searches_equal_the_linear_scan():
    rng = np.random.default_rng(13)
    for _ in range(300):
        n = int(rng.integers(1, 300))
        x = np.sort(rng.integers(0, int(rng.choice([200, 5000, 100000])), n)).astype(np.int64)
        win, max_iter = int(rng.choice([0, 50, 1000])), int(rng.choice([5, 64, 5000]))
        st, ref = 0, []
        for i in range(n):                                   # chain.c:192-193
            while st < i and x[i] > x[st] + win:
                st += 1
            if i - st > max_iter:
                st = i - max_iter
            ref.append(st)
        st_carry = 0
        for base in range(0, n, 32):
            ks = range(base, min(n, base + 32))
            got = {}
            for k in ks:                                     # halving steps from just before the range
                lo, hi = st_carry, k
                pos, span = lo - 1, max(1, max(kk - st_carry for kk in ks))
                step = 1 << (span.bit_length() - 1)
                while step:
                    q = pos + step
                    if q < hi and x[k] > x[q] + win:
                        pos = q
                    step >>= 1
                got[k] = pos + 1
            s0, pos_m, more = st_carry, {k: st_carry - 1 for k in ks}, {k: st_carry < k for k in ks}
            while any(more.values()):                        # merge: 32 candidate starts per pass
                v = [x[s0 + l] + win if s0 + l < n else 0 for l in range(32)]
                for k in ks:
                    c = -1
                    for r in range(6):
                        q = c + (16 >> r if r < 5 else 1)
                        if s0 + q < k and x[k] > v[q]:
                            c = q
                    if more[k]:
                        pos_m[k] = s0 + c
                        more[k] = c == 31 and s0 + 32 < k
                s0 += 32
            for k in ks:
                st_k = got[k]
                assert pos_m[k] + 1 == st_k
                if k - st_k > max_iter:
                    st_k = k - max_iter
                assert st_k == ref[k], (k, st_k, ref[k])
            st_carry = ref[ks[-1]]                          # the (clamped) start of the block's last anchor carries over
