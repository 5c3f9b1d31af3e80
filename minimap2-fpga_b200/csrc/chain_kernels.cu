// Hand-written sm_100a kernels for minimap2's chaining hot path (mm_chain_dp, /root/reference/chain.c:29-423).
//
// Design (B200-first, not a translation of the FPGA shift-register kernel or of the CPU loop):
//   * one WARP per read, persistent CTAs (2 warps, 16 CTAs/SM = 32 warps/SM) pulling reads longest-first from a
//     global work counter, so 148 SMs x 32 warps chain 4736 reads concurrently;
//   * anchors stream in 32 at a time with one coalesced 16-byte load per lane; each lane finds its own anchor's window
//     start (halving-step lower bound over the shared-memory ring, or a register merge against coalesced batches of
//     candidate starts when the window reaches below the ring), anchors with an empty window are finished in parallel,
//     and the most recent 256 anchors' {x_lo, y_lo, f, p | v, t} live in a per-warp shared-memory ring (6 KB); deeper
//     look-back (rare) reads L2;
//   * the inner loop over predecessors j = i-1 .. st is evaluated 32 lanes at a time; the order-dependent parts of
//     the reference loop (strict '>' running max, t[] stamps, the n_skip counter and its break, chain.c:226-233)
//     are recovered exactly from warp votes: REDUX.MAX + ballots for the records, a one-hot REDUX.OR for the stamps
//     that land inside the chunk (memory stamps only when the scan moves to another chunk), and a closed-form
//     (Lindley) evaluation of the n_skip counter on the two vote masks;
//   * chain ends / peaks, the descending sort, the priority backtrack (32 anchors per step by pointer jumping inside
//     aligned blocks of p[]) and the final order by reference position (including the reference's unstable radix-sort
//     tie order) run in the same warp right after the fill while f/p/v are still in L1/L2; the warp then reserves its share
//     of the packed output from two atomic cursors and writes u[] and b[] (or the 4-byte indices of b[]'s anchors) itself.
//   * the few reads with very long windows (tandem repeats) get a CTA of 16 warps and a ring that holds a whole window;
//     their long scans are shared by the warps (chain_heavy_kernel, "Heavy reads" below).
// No tensor cores (nothing here is a contraction) and no collective (reads are independent).
#include "chain_kernels.cuh"
#include <limits.h>

namespace mm2b {

namespace {

constexpr unsigned FULL = 0xffffffffu;
#ifndef MM2B_RING
#define MM2B_RING 256
#endif
constexpr int LIGHT_RING = MM2B_RING;    // ring slots per warp in the warp-per-read kernel (power of two)
constexpr int RING_ARRAYS = 6;          // x_lo, y_lo, f, p, v, t
// 2 warps per CTA, 16 CTAs per SM (32 warps/SM, 64 registers): small CTAs retire as soon as their reads are done, so when
// several sub-batch kernels share the GPU (the host-buffer pipeline) SM slots are handed on at 2-read granularity.
#ifndef MM2B_WARPS_PER_CTA
#define MM2B_WARPS_PER_CTA 2
#endif
constexpr int WARPS_PER_CTA = MM2B_WARPS_PER_CTA;
#ifndef MM2B_WARPS_PER_SM
#define MM2B_WARPS_PER_SM 32
#endif
constexpr int CTAS_PER_SM = MM2B_WARPS_PER_SM / WARPS_PER_CTA;
constexpr int32_t MARK_SUCC = 0x7ffffffe;   // "has a successor" (chain.c:351); DP stamps are anchor indices < 2^31-2
constexpr int SEG_SHIFT = 48;               // MM_SEED_SEG_SHIFT, mmpriv.h:22

// Debug build (-DMM2B_DEBUG_CHECKS, libmm2chain_b200_dbg.so): every index into the per-read scratch is range-checked and
// violations are OR-ed into a device flag word (mm2b_debug_flags()).  compute-sanitizer is closed on this GPU pool, so this is
// the memory-safety net the tests use (tests/test_gpu_debug_build.py).  The release build compiles the checks away.
#ifdef MM2B_DEBUG_CHECKS
__device__ unsigned g_dbg_flags = 0;
#define MM2B_CHK(cond, code) do { if (!(cond)) atomicOr(&g_dbg_flags, (code)); } while (0)
#else
#define MM2B_CHK(cond, code) do { } while (0)
#endif
#define MM2B_SORT_CHK(cond, code) MM2B_CHK(cond, code)
}  // namespace
}  // namespace mm2b
#include "sort_replay.cuh"      // W16, insertion_by_x, flag_sort_by_x_lane0: the reference's radix_sort_128x replayed exactly
namespace mm2b {
namespace {

// W16 (sort_replay.cuh) here: (first-anchor x, start-in-PATH << 32 | chain index)

struct ReadCtx {
	const ulonglong2 *A;
	int32_t *F, *P, *V, *T;
	uint64_t *X, *U;
	int n;
};

__device__ __forceinline__ unsigned lanemask_lt(int lane) { return (1u << lane) - 1u; }
__device__ __forceinline__ unsigned bits_below(int k) { return k >= 32 ? FULL : ((1u << k) - 1u); }   // lanes [0,k)

// x86 cvttss2si semantics for (int)float, which is what the reference's `(int)(dd * avg)` compiles to
__device__ __forceinline__ int f2i_x86(float f)
{
	return (f >= -2147483648.f && f < 2147483648.f) ? __float2int_rz(f) : INT_MIN;
}
__device__ __forceinline__ int d2i_x86(double d)
{
	return (d >= -2147483648.0 && d < 2147483648.0) ? __double2int_rz(d) : INT_MIN;
}

// Lowest set bit of a non-zero mask as a lane index.
__device__ __forceinline__ int lowest_lane(unsigned m) { return __popc((m - 1u) & ~m); }

// Per-warp shared-memory ring: the most recent RING anchors, one slot each.
//   slotA[s] = {x_lo, y_lo, f, p}   one 16-byte LDS fetches a predecessor
//   slotB[s] = {v, t}               v = peak score on the path (chain.c:237); t = visit stamp (chain.c:229,233)
struct Ring {
	int4 *a;
	int2 *b;
	int32_t *coop;          // mailbox of the heavy-read kernel (CTA-cooperative scans); unused by the warp-per-read kernel
};

struct DpConst {
	int max_dist_x, max_dist_y, bw, max_skip, max_iter, max_dq_same;
	bool cap_dr, is_cdna;
	float avg;
	double gap_scale;
};

// One anchor's scan over its predecessors j = i-1 .. st in 32-lane chunks, nearest first (chain.c:197-235).
// DEEP=false: the whole window [st, i) is resident in the ring, so every access is shared memory.
// DEEP=true : the window reaches below the ring (CCS reads, repeats); chunks that still lie inside the ring — the nearest
//             ones, where the loop usually ends — take the same shared-memory path, only deeper chunks read L1/L2.
//
// The order-dependent parts of the reference loop are recovered exactly from warp votes:
//  * records (chain.c:226, strict '>' running max): the first lane above max_f, then the first later lane above that, ...
//    — chunks hold 0-2 records in practice, so a warp-uniform loop beats a 5-step shuffle prefix-max;
//  * stamps (chain.c:233): every visited cell stamps t[p[j]] = i and a later cell is a "hit" if it finds its own t[j] == i.
//    A stamp only lands on an index smaller than its writer's, so within a chunk the stamps are a one-hot OR over the lanes
//    (REDUX.OR, no memory); stamps for cells of later chunks go to the ring (or to t[] in HBM below the ring) only when the
//    scan does move on, and are read back there.  Stamps from lanes past the break carry a value (i) that is never compared
//    again; stamps below st are never read;
//  * n_skip (chain.c:228-231) is a Lindley recursion x_t = max(x_{t-1} + d_t, 0), d = +1 (hit), -1 (record), 0 (other),
//    with closed form x_t = S_t - min(0, min_{s<=t} S_s), S_t = n_skip + sum d.  S only drops at records, so the running
//    minimum is a minimum over the record lanes: scalar work on the two vote masks, no shuffles;
//  * the loop breaks at the first hit lane with x > max_skip; the new (max_f, max_j) is the last record before it.
// Scores of one chunk: the 32 predecessors j = jt - lane (chain.c:199-220) and what the bookkeeping needs of them.  Nothing in
// here depends on the state of the scan (max_f, n_skip).
struct Chunk {
	int n_act;              // cells of the reference loop covered by this chunk
	int j, s;               // this lane's predecessor and its ring slot
	int32_t sc;             // score through j, INT_MIN when the reference `continue`s (or the lane is past the window)
	int32_t pj, tj_deep;    // p[j]; the memory stamp of j when it was fetched with the cell (deep chunks)
	bool valid, in_ring;
};

template <int RING, bool GENERAL, bool DEEP>
__device__ __forceinline__ void score_chunk(Chunk &k, const DpConst &c, const ReadCtx &rc, const Ring &ring, int lane, int i, int st, int ring_lo, int jt,
                                            int32_t xi, int32_t qi, int32_t q_span, int32_t sidi)
{
	k.n_act = jt - st + 1 < 32 ? jt - st + 1 : 32;
	const bool act = lane < k.n_act;
	k.j = jt - lane;
	k.s = k.j & (RING - 1);
	const int j = k.j, s = k.s;
	int32_t xj, yj, fj, pj, sidj = sidi;
	k.tj_deep = -1;
	k.in_ring = !DEEP || jt - k.n_act + 1 >= ring_lo;          // warp-uniform: the whole chunk is resident in the ring
	if (k.in_ring) {
		const int4 q = ring.a[s];                              // lanes past the window read a stale slot; masked by `act`
		xj = q.x, yj = q.y, fj = q.z, pj = q.w;
		if (GENERAL) {
			sidj = act ? (int32_t)(__ldg(&rc.A[j].y) >> SEG_SHIFT & 0xff) : sidi;
			__syncwarp();
		}
	} else {
		// (never the first chunk of a scan: that one is always resident.)  The cell's memory stamp is fetched together with
		// the cell: the stamps of earlier chunks are in place by now (end of their loop body) and the ones from inside this
		// chunk come through the one-hot OR, so a deep chunk costs one round trip to L2, not two.
		xj = 0, yj = 0, fj = 0, pj = -1;
		if (act) {
			if (j >= ring_lo) {
				const int4 q = ring.a[s];
				xj = q.x, yj = q.y, fj = q.z, pj = q.w;
				k.tj_deep = ring.b[s].y;
				if (GENERAL) sidj = (int32_t)(__ldg(&rc.A[j].y) >> SEG_SHIFT & 0xff);
			} else {                                           // deep look-back: L1/L2
				MM2B_CHK(j >= 0 && j < i, 0x1);
				const ulonglong2 t = __ldg(rc.A + j);
				xj = (int32_t)t.x, yj = (int32_t)t.y, fj = rc.F[j], pj = rc.P[j];
				k.tj_deep = rc.T[j];
				sidj = (int32_t)(t.y >> SEG_SHIFT & 0xff);
			}
		}
		__syncwarp();
	}
	k.pj = pj;
	// inside the window 0 <= dr <= max_dist_x, so the low words give dr exactly (chain.c:199)
	const int32_t dr = (int32_t)((uint32_t)xi - (uint32_t)xj);
	const int32_t dq = (int32_t)((uint32_t)qi - (uint32_t)yj);                // chain.c:200
	const int32_t diff = dr - dq;
	const int32_t dd = diff < 0 ? -diff : diff;                              // chain.c:204
	bool valid;
	int32_t sc;
	if (!GENERAL) {
		valid = act && dr != 0 && (uint32_t)(dq - 1) < (uint32_t)c.max_dq_same && dd <= c.bw;   // chain.c:202-205
		const int32_t md = dq < dr ? dq : dr;
		sc = md < q_span ? md : q_span;                                       // chain.c:207-208
		const float fdd = __int2float_rn(dd);                                 // exact: dd <= bw < 2^24 on this path
		const int c_lin = __float2int_rz(__fmul_rn(fdd, c.avg));              // chain.c:218
		int lg = (__float_as_int(fdd) >> 23) - 127;                           // ilog2_32(dd) read off the float exponent (chain.c:209) ...
		lg = lg < 0 ? 0 : lg;                                                 // ... and 0 for dd == 0
		sc = sc - (c_lin + (lg >> 1)) + fj;
	} else {
		const bool same = sidi == sidj;
		valid = act && !((same && dr == 0) || dq <= 0)                        // chain.c:202
		            && !((same && dq > c.max_dist_y) || dq > c.max_dist_x)     // chain.c:203
		            && !(same && dd > c.bw)                                    // chain.c:205
		            && !(c.cap_dr && same && dr > c.max_dist_y);               // chain.c:206
		const int32_t md = dq < dr ? dq : dr;
		sc = md > q_span ? q_span : md;
		const int lg = dd ? 31 - __clz(dd) : 0;
		int gap = 0;
		if (c.is_cdna || !same) {                                             // chain.c:211-217
			const int c_lin = f2i_x86(__fmul_rn(__int2float_rn(dd), c.avg));
			if (!same && dr == 0) ++sc;
			else if (dr > dq || !same) gap = c_lin < lg ? c_lin : lg;
			else gap = c_lin + (lg >> 1);
		} else gap = f2i_x86(__fmul_rn(__int2float_rn(dd), c.avg)) + (lg >> 1);
		sc -= d2i_x86(__dadd_rn(__dmul_rn((double)gap, c.gap_scale), .499));  // chain.c:219
		sc += fj;
	}
	k.sc = valid ? sc : INT_MIN;
	k.valid = valid;
	if (GENERAL) __syncwarp();                            // the cost switch above branches per lane
}

// The scan goes on to the next chunk: stamps (chain.c:233) for the cells of later chunks go to memory now.
template <int RING, bool DEEP>
__device__ __forceinline__ void stamp_later_chunks(const Chunk &k, const ReadCtx &rc, const Ring &ring, int i, int st, int ring_lo, int jt)
{
	const int32_t pj = k.pj;
	if (k.valid && pj >= st && pj < jt - 31) {
		MM2B_CHK(pj < k.j && pj > i - RING - 32 - 5000000, 0x2);
		MM2B_CHK(DEEP || pj >= ring_lo, 0x4);
		if (!DEEP || pj >= ring_lo) ring.b[pj & (RING - 1)].y = i;
		else rc.T[pj] = i;
	}
	__syncwarp();
}

// The order-dependent part for one chunk: hits, records, the n_skip counter, the break (chain.c:226-233).  Returns whether the
// reference's loop ends inside this chunk.  FIRST = the chunk next to anchor i: no stamp of this scan can be in memory yet, and
// max_f / max_j / n_skip still hold their initial values (q_span, -1, 0), which the compiler folds.
// Every case ends in its own `return` (no common tail): once inlined, "the loop ends here" is a branch, never a flag in a register.
template <int RING, bool DEEP, bool COUNT, bool FIRST>
__device__ __forceinline__ bool resolve_chunk(const Chunk &k, const DpConst &c, const ReadCtx &rc, const Ring &ring, int lane, int i, int st, int ring_lo, int jt,
                                              int32_t &max_f, int32_t &max_j, int &n_skip, unsigned &n_cells)
{
	// hits (chain.c:229,233): cell j was stamped by an earlier-visited valid cell whose predecessor it is.  A stamp from
	// inside this chunk is a bit in a one-hot OR over the lanes (lane of the target = jt - p[j]); memory stamps are only
	// written when the scan moves on to another chunk (stamp_later_chunks), so the usual single-chunk scan has no
	// store -> load round trip at all.  `hitv` still includes the record lanes; they are masked out below.
	unsigned hot;
	asm("shl.b32 %0, 1, %1;" : "=r"(hot) : "r"(jt - k.pj));      // PTX shl clamps: 0 for targets beyond this chunk (distance >= 32)
	hot = __reduce_or_sync(FULL, k.valid ? hot : 0u);
	unsigned hitv;
	if (FIRST) hitv = __ballot_sync(FULL, k.valid) & hot;
	else {
		const int32_t tj = k.in_ring ? ring.b[k.s].y : k.tj_deep;
		hitv = __ballot_sync(FULL, k.valid && (tj == i || (hot >> lane & 1u)));
	}
	// records (chain.c:226, strict '>'), n_skip, whether the loop breaks in this chunk, and the last record before the
	// break = the new running max.  The largest score of the chunk (first occurrence) is the LAST record.
	// (the branch is on a vote result, not on `top <= max_f`: ptxas must see a warp-uniform condition — convergence note below)
	const unsigned cand = __ballot_sync(FULL, k.sc > max_f);
#define MM2B_TALLY(broke_, brk_) if (COUNT) n_cells += (broke_) ? (brk_) + 1 : k.n_act
	if (cand == 0) {                                      // (i) no record: the counter only goes up
		const int x0 = n_skip;
		n_skip += __popc(hitv);
		if (n_skip > c.max_skip) {
			if (COUNT) {
				const unsigned le = lanemask_lt(lane) | (1u << lane);
				MM2B_TALLY(true, lowest_lane(__ballot_sync(FULL, ((hitv >> lane) & 1u) && x0 + __popc(hitv & le) == c.max_skip + 1)));
			}
			return true;
		}
		MM2B_TALLY(false, 0);
		stamp_later_chunks<RING, DEEP>(k, rc, ring, i, st, ring_lo, jt);
		return false;
	}
	const int32_t top = __reduce_max_sync(FULL, k.sc);
	// the nearest lane holding the maximum = the lane with the largest j among them: a second REDUX instead of vote + bit scan
	const int32_t top_j = __reduce_max_sync(FULL, k.sc == top ? k.j : INT_MIN);
	const int last = jt - top_j;
	unsigned recmask = 1u << last;
	const unsigned before = recmask - 1u;
	if ((cand & before) == 0) {                           // (ii) one record: a run of hits before it and one after it
		const unsigned h1 = hitv & before;
		const int x1 = n_skip + __popc(h1);                                    // counter when the record is reached
		if (x1 > c.max_skip) {                                                 // the loop ends before it reaches the record
			if (COUNT) {
				const unsigned le = lanemask_lt(lane) | (1u << lane);
				int kk = c.max_skip + 1 - n_skip;
				kk = kk < 1 ? 1 : kk;
				MM2B_TALLY(true, lowest_lane(__ballot_sync(FULL, ((h1 >> lane) & 1u) && __popc(h1 & le) == kk)));
			}
			return true;
		}
		max_f = top, max_j = top_j;
		const unsigned h2 = hitv & ~(before | recmask);
		const int x2 = max(x1 - 1, 0);                                         // ... after the record
		n_skip = x2 + __popc(h2);                                              // ... at the end of the chunk
		if (n_skip > c.max_skip) {
			if (COUNT) {
				const unsigned le = lanemask_lt(lane) | (1u << lane);
				int kk = c.max_skip + 1 - x2;
				kk = kk < 1 ? 1 : kk;
				MM2B_TALLY(true, lowest_lane(__ballot_sync(FULL, ((h2 >> lane) & 1u) && __popc(h2 & le) == kk)));
			}
			return true;
		}
		MM2B_TALLY(false, 0);
		stamp_later_chunks<RING, DEEP>(k, rc, ring, i, st, ring_lo, jt);
		return false;
	}
	// (iii) several records (rare)
	for (int r = lowest_lane(cand); r != last;) {
		recmask |= 1u << r;
		const int32_t t = __shfl_sync(FULL, k.sc, r);
		r = lowest_lane(__ballot_sync(FULL, k.sc > t) & (0xfffffffeu << r));
	}
	const unsigned hm = hitv & ~recmask;
	bool broke = false;
	int brk = 32;
	unsigned take = recmask;                              // records visited before the break
	if (hm == 0) {                                        // only decrements: saturating subtraction
		n_skip -= __popc(recmask);
		n_skip = n_skip > 0 ? n_skip : 0;
	} else {
		int corr = 0, floor_all = 0, done = 0;            // corr: min(0, min S over the records at or before this lane)
		for (unsigned rm = recmask; rm; rm &= rm - 1) {
			const unsigned below = (rm - 1u) & ~rm;       // lanes before this record
			const int S_r = n_skip + __popc(hm & below) - (++done);
			floor_all = S_r < floor_all ? S_r : floor_all;
			if ((below >> lane & 1u) == 0 && S_r < corr) corr = S_r;
		}
		const unsigned le = lanemask_lt(lane) | (1u << lane);
		const int x = n_skip + __popc(hm & le) - __popc(recmask & le) - corr;
		const unsigned over = __ballot_sync(FULL, ((hm >> lane) & 1u) && x > c.max_skip);
		if (over) broke = true, brk = lowest_lane(over), take = recmask & bits_below(brk);
		else n_skip = n_skip + __popc(hm) - done - floor_all;
	}
	// records are strictly increasing and ties went to the nearest j, so the last one taken is the new running max
	if (take) {
		const int l2 = 31 - __clz(take);
		max_f = __shfl_sync(FULL, k.sc, l2);
		max_j = jt - l2;
	}
	MM2B_TALLY(broke, brk);
	if (broke) return true;
	stamp_later_chunks<RING, DEEP>(k, rc, ring, i, st, ring_lo, jt);
	return false;
#undef MM2B_TALLY
}

template <int RING, bool GENERAL, bool DEEP, bool COUNT>
__device__ __forceinline__ void scan_predecessors(const DpConst &c, const ReadCtx &rc, const Ring &ring, int lane, int i, int st, int ring_lo,
                                                  int32_t xi, int32_t qi, int32_t q_span, int32_t sidi,
                                                  int32_t &max_f, int32_t &max_j, unsigned &n_chunks, unsigned &n_cells)
{
	int jt = i - 1;                                          // the caller only comes here with a non-empty window (st < i)
	Chunk k;
	// Quiet scan: n_skip grows by at most one per visited cell, so a window of at most max_skip cells cannot trigger the break of
	// chain.c:230.  Then stamps, hits and n_skip cannot change the outcome and only the running maximum matters (strict '>',
	// nearest first = the largest j among the lanes holding a chunk's maximum).  Short windows are a large part of a noisy read.
	if (i - st <= c.max_skip) {
#pragma unroll 1
		do {
			if (COUNT) ++n_chunks;
			score_chunk<RING, GENERAL, DEEP>(k, c, rc, ring, lane, i, st, ring_lo, jt, xi, qi, q_span, sidi);
			const int32_t best = __reduce_max_sync(FULL, k.sc);
			const int32_t best_j = __reduce_max_sync(FULL, k.sc == best ? k.j : INT_MIN);   // unconditional: keeps the warp converged by construction
			if (best > max_f) max_f = best, max_j = best_j;
			if (COUNT) n_cells += k.n_act;
		} while ((jt -= 32) >= st);
		return;
	}
	int n_skip = 0;
	if (COUNT) ++n_chunks;
	score_chunk<RING, GENERAL, false>(k, c, rc, ring, lane, i, st, ring_lo, jt, xi, qi, q_span, sidi);      // the nearest chunk is always resident
	if (resolve_chunk<RING, DEEP, COUNT, true>(k, c, rc, ring, lane, i, st, ring_lo, jt, max_f, max_j, n_skip, n_cells)) return;
#pragma unroll 1
	while ((jt -= 32) >= st) {
		if (COUNT) ++n_chunks;
		score_chunk<RING, GENERAL, DEEP>(k, c, rc, ring, lane, i, st, ring_lo, jt, xi, qi, q_span, sidi);
		if (resolve_chunk<RING, DEEP, COUNT, false>(k, c, rc, ring, lane, i, st, ring_lo, jt, max_f, max_j, n_skip, n_cells)) return;
	}
}

// ---------------------------------------------------------------------------------------------------------------
// Heavy reads: scans shared by the warps of a CTA
// ---------------------------------------------------------------------------------------------------------------
// A read in tandem repeats has windows at the max_iter clamp (thousands of anchors) and scans dozens of chunks per anchor; on
// one warp that is a serial chain of ~100 instructions per chunk.  The heavy-read kernel gives such a read a CTA of HEAVY_WARPS
// warps and a ring that holds the whole window.  Warp 0 runs the read exactly like the warp-per-read kernel; for a long scan
// it posts the anchor in a mailbox and all warps take the window's chunks round-robin, two per warp and round (32 chunks):
//   A  every warp scores its chunks and writes ALL their stamps (chain.c:233) to the ring                        -> barrier
//   B  every warp reads its cells' stamps and reduces each chunk to a summary that does not depend on the scan's state:
//      the chunk's maximum score and its number of stamped valid cells (a stamp only comes from a nearer cell, and the
//      nearer chunks stamped in A of this round or earlier)                                                        -> barrier
//   C  every warp folds the round's summaries in chunk order with the same replicated state (max_f, max_j, n_skip): a chunk
//      whose maximum does not beat max_f has no record, so it is n_skip += hits and one compare (chain.c:229-231); a chunk
//      with a record is resolved in detail by its owner (which still holds the lanes' scores) and published       -> barrier
// Stamps and summaries of chunks past the break are computed in vain and never looked at, like the lanes past the break
// inside a chunk.
constexpr int HEAVY_RING = 8192;           // covers max_iter = 5000 entirely: no look-back below the ring in this kernel
constexpr int HEAVY_WARPS = 16;
constexpr int COOP_MIN_CELLS = 128;        // windows longer than this are scanned cooperatively
constexpr int COOP_CPW = 2;                // chunks per warp and round of a cooperative scan
constexpr int COOP_ROUND = HEAVY_WARPS * COOP_CPW;     // = 32: one summary per lane in the fold
static_assert(COOP_ROUND == 32 && COOP_CPW == 2, "the fold keeps chunk k's summary in lane k");
constexpr int COOP_WORDS = 224;            // mailbox: [0,8) job, [16,80) chunk summaries, [80,208) detailed results
enum { JOB_SCAN = 1, JOB_EXIT = 2 };
enum { BAR_JOB = 1, BAR_STAMPS = 2, BAR_SUMM = 3, BAR_DETAIL = 4 };

__device__ __forceinline__ void cta_bar(int id) { asm volatile("bar.sync %0, %1;" :: "r"(id), "n"(HEAVY_WARPS * 32) : "memory"); }

// One chunk whose maximum beats the running max, i.e. with at least one record: cases (ii) and (iii) of scan_predecessors
// (kept as a separate copy so that the warp-per-read kernel's hot loop stays exactly as tuned).  Returns whether the loop breaks.
__device__ __forceinline__ bool resolve_records(int32_t sc, unsigned hitv, int lane, int jt, int max_skip, int32_t &max_f, int32_t &max_j, int &n_skip)
{
	const unsigned cand = __ballot_sync(FULL, sc > max_f);
	const int32_t top = __reduce_max_sync(FULL, sc);
	const int32_t top_j = __reduce_max_sync(FULL, sc == top ? jt - lane : INT_MIN);
	const int last = jt - top_j;
	unsigned recmask = 1u << last;
	if ((cand & (recmask - 1u)) == 0) {                       // one record
		const unsigned hitmask = hitv & ~recmask;
		const unsigned h1 = hitmask & (recmask - 1u), h2 = hitmask & ~(recmask - 1u);
		const int x1 = n_skip + __popc(h1);
		const int x2 = x1 > 0 ? x1 - 1 : 0;
		const int x3 = x2 + __popc(h2);
		const bool early = x1 > max_skip;
		if (!early) max_f = top, max_j = top_j;
		n_skip = x3;
		return early || x3 > max_skip;
	}
	for (int r = lowest_lane(cand); r != last;) {             // several records
		recmask |= 1u << r;
		const int32_t t = __shfl_sync(FULL, sc, r);
		r = lowest_lane(__ballot_sync(FULL, sc > t) & (0xfffffffeu << r));
	}
	const unsigned hm = hitv & ~recmask;
	bool broke = false;
	unsigned take = recmask;
	if (hm == 0) {
		n_skip -= __popc(recmask);
		n_skip = n_skip > 0 ? n_skip : 0;
	} else {
		int corr = 0, floor_all = 0, done = 0;
		for (unsigned rm = recmask; rm; rm &= rm - 1) {
			const unsigned below = (rm - 1u) & ~rm;
			const int S_r = n_skip + __popc(hm & below) - (++done);
			floor_all = S_r < floor_all ? S_r : floor_all;
			if ((below >> lane & 1u) == 0 && S_r < corr) corr = S_r;
		}
		const unsigned le = lanemask_lt(lane) | (1u << lane);
		const int x = n_skip + __popc(hm & le) - __popc(recmask & le) - corr;
		const unsigned over = __ballot_sync(FULL, ((hm >> lane) & 1u) && x > max_skip);
		if (over) broke = true, take = recmask & bits_below(lowest_lane(over));
		else n_skip = n_skip + __popc(hm) - done - floor_all;
	}
	if (take) {
		const int l2 = 31 - __clz(take);
		max_f = __shfl_sync(FULL, sc, l2);
		max_j = jt - l2;
	}
	return broke;
}

// The scan of anchor i over [st, i) by all warps of the CTA (same-segment, non-cDNA cost only); w = this warp's index.
// A round covers COOP_ROUND = 32 chunks (1,024 cells), two per warp: the scans this kernel exists for visit ~900 cells before the
// max_skip break, so they finish in one round — one barrier per phase instead of two rounds' worth — and lane k of every warp
// holds chunk k's summary in the fold.
template <int RING>
__device__ void coop_scan(const DpConst &c, const Ring &ring, int lane, int w, int i, int st, int32_t xi, int32_t qi, int32_t q_span,
                          int32_t &max_f_out, int32_t &max_j_out)
{
	int32_t *summ = ring.coop + 16, *detail = ring.coop + 16 + 2 * COOP_ROUND;
	int32_t max_f = q_span, max_j = -1;
	int n_skip = 0;
	const int n_chunks = (i - st + 31) >> 5;
	bool broke = false;
	for (int g0 = 0; g0 < n_chunks && !broke; g0 += COOP_ROUND) {
		// A: this warp's chunks of the round (chunk w and chunk w + HEAVY_WARPS)
		int32_t sc[COOP_CPW];
		bool valid[COOP_CPW];
		int jt[COOP_CPW], sl[COOP_CPW];
#pragma unroll
		for (int q = 0; q < COOP_CPW; ++q) {
			jt[q] = i - 1 - 32 * (g0 + w + q * HEAVY_WARPS);
			const int left = jt[q] - st + 1;                          // <= 0 for a warp without a chunk in the last round
			const bool act = lane < left;
			const int j = jt[q] - lane;
			sl[q] = j & (RING - 1);
			sc[q] = INT_MIN, valid[q] = false;
			if (left > 0) {                                           // (warp-uniform)
				const int4 r = ring.a[sl[q]];
				const int32_t pj = r.w;
				const int32_t dr = (int32_t)((uint32_t)xi - (uint32_t)r.x);
				const int32_t dq = (int32_t)((uint32_t)qi - (uint32_t)r.y);
				const int32_t diff = dr - dq;
				const int32_t dd = diff < 0 ? -diff : diff;
				valid[q] = act && dr != 0 && (uint32_t)(dq - 1) < (uint32_t)c.max_dq_same && dd <= c.bw;    // chain.c:202-205
				const int32_t md = dq < dr ? dq : dr;
				int32_t v = md < q_span ? md : q_span;
				const float fdd = __int2float_rn(dd);
				const int c_lin = __float2int_rz(__fmul_rn(fdd, c.avg));
				int lg = (__float_as_int(fdd) >> 23) - 127;
				lg = lg < 0 ? 0 : lg;
				sc[q] = valid[q] ? v - (c_lin + (lg >> 1)) + r.z : INT_MIN;
				MM2B_CHK(!valid[q] || (j >= st && j < i && pj < j), 0x100);
				if (valid[q] && pj >= st) ring.b[pj & (RING - 1)].y = i;
			}
			__syncwarp();
		}
		cta_bar(BAR_STAMPS);
		// B: state-independent summaries
		unsigned hitv[COOP_CPW];
#pragma unroll
		for (int q = 0; q < COOP_CPW; ++q) {
			const int32_t tj = ring.b[sl[q]].y;
			hitv[q] = __ballot_sync(FULL, valid[q] && tj == i);
			const int32_t best = __reduce_max_sync(FULL, sc[q]);
			if (lane == 0) summ[2 * (w + q * HEAVY_WARPS)] = best, summ[2 * (w + q * HEAVY_WARPS) + 1] = __popc(hitv[q]);
		}
		__syncwarp();
		cta_bar(BAR_SUMM);
		// C: fold in chunk order (every warp, same values).  Lane k holds chunk k's summary; votes find the first chunk with a
		// record and, from a prefix sum of the hit counts, the first chunk in which the counter passes max_skip.
		const int2 mine = *(const int2*)&summ[2 * lane];
		__syncwarp();
		const int in_round = n_chunks - g0 < COOP_ROUND ? n_chunks - g0 : COOP_ROUND;
		for (int k0 = 0; k0 < in_round;) {                        // chunks [k0, in_round) are still to be folded
			const bool todo = lane >= k0 && lane < in_round;
			const unsigned recs = __ballot_sync(FULL, todo && mine.x > max_f);
			int h = todo ? mine.y : 0;                            // inclusive prefix sum of the hits over the chunks still to fold
#pragma unroll
			for (int d = 1; d < COOP_ROUND; d <<= 1) {
				const int o = __shfl_up_sync(FULL, h, d);
				if (lane >= d) h += o;
			}
			const unsigned over = __ballot_sync(FULL, todo && n_skip + h > c.max_skip);
			const int kr = recs ? lowest_lane(recs) : 32, kb = over ? lowest_lane(over) : 32;
			if (kb < kr) { broke = true; break; }                 // the counter passes max_skip before any record (chain.c:229-231)
			if (kr == 32) { n_skip += __shfl_sync(FULL, h, in_round - 1); break; }
			if (kr > k0) n_skip += __shfl_sync(FULL, h, kr - 1);  // up to the chunk with a record; its owner resolves that one in detail
			if (w == (kr & (HEAVY_WARPS - 1))) {
				const int q = kr / HEAVY_WARPS;
				int32_t f2 = max_f, j2 = max_j;
				int s2 = n_skip;
				const bool b2 = resolve_records(q ? sc[COOP_CPW - 1] : sc[0], q ? hitv[COOP_CPW - 1] : hitv[0], lane, q ? jt[COOP_CPW - 1] : jt[0], c.max_skip, f2, j2, s2);
				if (lane == 0) *(int4*)&detail[4 * kr] = make_int4(f2, j2, s2, b2 ? 1 : 0);
				__syncwarp();
			}
			cta_bar(BAR_DETAIL);
			int4 d4 = make_int4(0, 0, 0, 0);
			if (lane == 0) d4 = *(const int4*)&detail[4 * kr];
			__syncwarp();
			max_f = __shfl_sync(FULL, d4.x, 0), max_j = __shfl_sync(FULL, d4.y, 0), n_skip = __shfl_sync(FULL, d4.z, 0);
			broke = __shfl_sync(FULL, d4.w, 0) != 0;
			if (broke) break;
			k0 = kr + 1;
		}
	}
	max_f_out = max_f, max_j_out = max_j;
}

// The sequential step for one block of 32 anchors: the anchors flagged in `todo` (non-empty window), in index order.
template <int RING, bool GENERAL, bool DEEP, bool COUNT, bool COOP>
__device__ __forceinline__ void chain_block(const DpConst &c, const ReadCtx &rc, const Ring &ring, int lane, int base, int ring_lo,
                                            unsigned todo, int32_t seg, int st_k, unsigned &n_chunks, unsigned &n_cells)
{
	if (todo == 0) return;
	const int4 *slot0 = ring.a + (base & (RING - 1));      // a block is 32-aligned, so its slots do not wrap inside the ring
	do {
		const int ii = lowest_lane(todo);
		todo &= todo - 1;
		const int i = base + ii;
		const int4 *slot = slot0 + ii;
		const int4 me = *slot;                    // broadcast read: this anchor's own slot still holds its defaults
		// The window start must come through a shuffle, not through shared memory: it bounds the chunk loop, and ptxas only
		// treats the loop as warp-uniform (no BRA.DIV guards on the collectives inside) when the bound is a shuffle/vote result.
		const int st = __shfl_sync(FULL, st_k, ii);
		const int32_t q_span = me.z;
		const int32_t sidi = GENERAL ? __shfl_sync(FULL, seg, ii) : 0;
		int32_t max_f = q_span, max_j = -1;
		if (COOP && !GENERAL && i - st > COOP_MIN_CELLS) {     // a long scan in the heavy-read kernel: all warps of the CTA share it
			if (lane == 0) {
				ring.coop[0] = JOB_SCAN, ring.coop[1] = i, ring.coop[2] = st, ring.coop[3] = me.x, ring.coop[4] = me.y, ring.coop[5] = q_span;
				ring.coop[6] = __float_as_int(c.avg);
			}
			__syncwarp();
			cta_bar(BAR_JOB);
			coop_scan<RING>(c, ring, lane, 0, i, st, me.x, me.y, q_span, max_f, max_j);
		} else
		scan_predecessors<RING, GENERAL, DEEP, COUNT>(c, rc, ring, lane, i, st, ring_lo, me.x, me.y, q_span, sidi, max_f, max_j, n_chunks, n_cells);
		// f[i], p[i] (chain.c:236): one lane publishes them to the anchor's slot; untouched if no predecessor won.
		// v[i] is not needed by the scan at all; it is filled in per block afterwards (dp_fill).
		MM2B_CHK(max_j < i && max_j >= -1 && (max_j < 0 || max_j >= st) && (DEEP || max_j < 0 || max_j >= ring_lo), 0x8);
		if (lane == 0) *(int2*)&slot->z = make_int2(max_f, max_j);      // (when nothing won these are the defaults the slot already holds)
		__syncwarp();
	} while (todo);
}

// ---------------------------------------------------------------------------------------------------------------
// DP fill: f[], p[], v[] for one read (chain.c:184-238).  GENERAL=false is the map-ont / asm20 shape
// (one segment id, !is_cdna, gap_scale == 1) with a pure-integer + one float-multiply cost; GENERAL=true carries
// the full cost switch of chain.c:211-219 (cross-segment, cDNA, gap_scale in double).
//
// Anchors are taken 32 at a time.  For a block, every lane publishes {x_lo, y_lo, f = q_span, p = -1 | v = q_span, t = -1} to
// the ring and finds its own anchor's window start st (chain.c:192-193) as a lower bound —
// st_i = max(lower_bound{s : x_s + max_dist_x >= x_i}, i - max_iter) since the input is sorted by x.  Anchors whose window is empty
// (isolated seed hits: ~40 % of a noisy ONT read) are final at that point; only the others take the sequential step.
// ---------------------------------------------------------------------------------------------------------------
// RANGED: fill only the anchors [k_begin, k_end) of the read, where k_begin is a cut point — a[k_begin].x > a[k_begin - 1].x + max_dist_x,
// so no anchor from k_begin on can reach one before it (chain.c:192) and the piece is an independent DP (long-read segmenting).
template <int RING, bool GENERAL, bool COUNT, bool COOP, bool RANGED = false>
__device__ void dp_fill(const mm2b_params_t &par, const ReadCtx &rc, float avg, bool lo_safe, int32_t *smem, int32_t *coop, int lane,
                        unsigned long long &n_chunks64, unsigned long long &n_cells64, unsigned long long &n_window64, int k_begin = 0, int k_end = 0)
{
	Ring ring;
	ring.a = (int4*)smem, ring.b = (int2*)(smem + 4 * RING), ring.coop = coop;
	const ulonglong2 *A = rc.A;
	const int n = rc.n;
	DpConst c;
	c.max_dist_x = par.max_dist_x, c.max_dist_y = par.max_dist_y, c.bw = par.bw, c.max_skip = par.max_skip, c.max_iter = par.max_iter;
	c.max_dq_same = par.max_dist_x < par.max_dist_y ? par.max_dist_x : par.max_dist_y;   // chain.c:203, same segment
	c.cap_dr = par.n_segs > 1 && !par.is_cdna;                                           // chain.c:206
	c.is_cdna = par.is_cdna != 0, c.avg = avg, c.gap_scale = (double)par.gap_scale;
	const uint64_t win = (uint64_t)(int64_t)par.max_dist_x;
	unsigned n_chunks = 0, n_cells = 0;
	const int kb = RANGED ? k_begin : 0, ke = RANGED ? k_end : n;
	int st_carry = kb;          // window start of the previous block's last anchor (window starts never move backwards)
	int run_carry = kb;         // first anchor of the run of equal high words (strand, rid) the previous block ended in
	uint32_t hi_carry = 0;      // ... and that high word

	for (int base = kb & ~31; base < ke; base += 32) {
		const int k = base + lane;
		const bool in = k < ke && (!RANGED || k >= kb);
		uint64_t x = 0, y = 0;
		if (in) {
			const ulonglong2 t = __ldg(A + k);
			x = t.x, y = t.y;
		}
		__syncwarp();
		const int32_t seg = (int32_t)(y >> SEG_SHIFT & 0xff);
		if (in) {                                          // publish first: the window search below probes the ring
			const int s = k & (RING - 1);
			const int32_t q_span = (int32_t)(y >> 32 & 0xff);
			ring.a[s] = make_int4((int32_t)x, (int32_t)y, q_span, -1);
			ring.b[s] = make_int2(q_span, -1);
		}
		__syncwarp();
		const int ring_lo = base + 32 - RING;      // anchors with index >= ring_lo are resident in the ring
		// chain.c:192: first s in [st_carry, k] with !(x > a[s].x + max_dist_x) — a 64-bit compare, strand/rid are in the high word.
		// Anchors are sorted, so the anchors sharing k's high word are a run [run_k, k]; when low word + max_dist_x cannot carry
		// (`lo_safe`, checked over the whole read) every anchor before the run is out of reach and inside the run the compare is
		// on the low words, which the ring holds for the most recent RING anchors: the search then runs on shared memory
		// (probes below the ring read the low word from L1/L2).  Without `lo_safe` it is the plain 64-bit search over L1/L2.
		const uint32_t xh = (uint32_t)(x >> 32);
		uint32_t xh_prev = __shfl_up_sync(FULL, xh, 1);
		if (lane == 0) xh_prev = hi_carry;
		const unsigned chg = __ballot_sync(FULL, in && (k == kb || xh != xh_prev)) & (lanemask_lt(lane) | (1u << lane));
		const int run_k = chg ? base + (31 - __clz(chg)) : run_carry;
		// Lower bound by halving steps (no data-dependent branch, the same trip count for all lanes): pos = last index known to
		// be out of reach, starting just before the range.
		const int lo = lo_safe && run_k > st_carry ? run_k : st_carry, hi = in ? k : lo;
		int pos = lo - 1;
		const int span = __reduce_max_sync(FULL, hi - lo);
		if (lo_safe && __all_sync(FULL, lo >= ring_lo)) {
			const uint32_t xl = (uint32_t)x - (uint32_t)win;              // x_s + win < x_k on the low words; cannot wrap when it matters:
			const bool any_far = (uint32_t)x >= (uint32_t)win;           // x_k < win means every anchor of the run is within reach
			for (int step = 1 << (31 - __clz(span | 1)); step; step >>= 1) {
				const int q = pos + step;
				if (q < hi && any_far && (uint32_t)ring.a[q & (RING - 1)].x < xl) pos = q;
			}
		} else {
			// The range reaches below the ring (long windows: CCS reads, repeats) or the low words are not enough.  Window starts
			// move forward by about one block per block, so instead of ~10 dependent probes into L2 the warp merges: it loads
			// the next 32 candidate starts with one coalesced read and every lane counts, by a binary search over the lanes'
			// registers, how many of them are out of its reach; lanes that exhaust the batch go on to the next one.
			int s0 = __reduce_min_sync(FULL, in ? lo : INT_MAX);
			bool more = in && lo < k;
			while (__any_sync(FULL, more)) {
				const int sidx = s0 + lane;
				uint64_t v = 0;
				if (sidx < n) v = __ldg(&A[sidx].x) + win;
				__syncwarp();
				int c = -1;                                   // last out-of-reach entry of the batch
#pragma unroll
				for (int r = 0; r < 6; ++r) {                 // steps 16, 8, 4, 2, 1 and once more 1: c can reach 31
					const int q = c + (r < 5 ? 16 >> r : 1);
					const uint64_t vq = __shfl_sync(FULL, v, q);
					if (s0 + q < k && x > vq) c = q;
				}
				if (more) {
					pos = s0 + c;
					more = c == 31 && s0 + 32 < k;
				}
				__syncwarp();
				s0 += 32;
			}
		}
		__syncwarp();
		int st_k = pos + 1;
		if (k - st_k > c.max_iter) st_k = k - c.max_iter;                                // chain.c:193
		if (COUNT && in) n_window64 += (unsigned)(k - st_k);             // per-lane partial sums of the window sizes (chain.c:192-193)
		unsigned todo = __ballot_sync(FULL, in && st_k < k);
		const bool deep_block = __any_sync(FULL, in && st_k < ring_lo);    // some window in this block reaches below the ring
		{
			const int last_lane = (ke - base < 32 ? ke - base : 32) - 1;
			st_carry = __shfl_sync(FULL, st_k, last_lane);
			run_carry = __shfl_sync(FULL, run_k, last_lane);
			hi_carry = __shfl_sync(FULL, xh, last_lane);
		}
		__syncwarp();

		if (!deep_block) chain_block<RING, GENERAL, false, COUNT, COOP>(c, rc, ring, lane, base, ring_lo, todo, seg, st_k, n_chunks, n_cells);
		else chain_block<RING, GENERAL, true, COUNT, false>(c, rc, ring, lane, base, ring_lo, todo, seg, st_k, n_chunks, n_cells);
		// Block epilogue.  v[i] = max(f[i], v[p[i]]) (chain.c:237) is the maximum of f along the chain behind i.  Inside the block
		// it is resolved by pointer jumping over the lanes (5 rounds cover 32 anchors); a chain that leaves the block picks up the
		// finished v of an earlier block from the ring (or from HBM when it reaches below the ring).  Then one coalesced write
		// of the block's f/p/v (needed by deep look-back and by the backtrack).
		{
			const int s = k & (RING - 1);
			int2 fp = make_int2(0, -1);
			if (in) fp = *(const int2*)&ring.a[s].z;
			__syncwarp();
			int32_t v = fp.x, pp = fp.y;
#pragma unroll
			for (int r = 0; r < 5; ++r) {
				const bool inb = pp >= base;
				const int src = inb ? pp - base : lane;
				const int32_t vs = __shfl_sync(FULL, v, src), ps = __shfl_sync(FULL, pp, src);
				if (inb) v = v > vs ? v : vs, pp = ps;
			}
			if (in) {
				if (pp >= 0) {
					const int32_t v_prev = pp >= ring_lo ? ring.b[pp & (RING - 1)].x : rc.V[pp];
					v = v > v_prev ? v : v_prev;
				}
				ring.b[s].x = v;
				rc.F[k] = fp.x, rc.P[k] = fp.y, rc.V[k] = v;
				// chain.c:349-351, folded in: t[] was zeroed before the fill and every later writer of t[p] — this mark or a
				// deep-look-back stamp (an anchor index >= 1) — means "p has a successor", so chain ends are the anchors with t == 0
				if (fp.y >= 0) rc.T[fp.y] = MARK_SUCC;
			}
		}
		__syncwarp();            // (see the note on convergence at the top of the kernel)
		n_chunks64 += n_chunks, n_cells64 += n_cells;
		n_chunks = 0, n_cells = 0;
	}
	__syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-level sorts used after the fill
// ---------------------------------------------------------------------------------------------------------------

// Ascending LSD radix sort of n 64-bit keys by one warp; 8-bit digits, only the bytes that differ between keys.
// `hist` is 256 ints of shared memory.  Result ends in `keys` (tmp is the ping-pong buffer).
__device__ void warp_radix_sort_u64(uint64_t *keys, uint64_t *tmp, int n, int *hist, int lane)
{
	uint64_t diff = 0;
	const uint64_t k0 = keys[0];
	for (int k = lane; k < n; k += 32) diff |= keys[k] ^ k0;
	__syncwarp();
#pragma unroll
	for (int d = 16; d; d >>= 1) diff |= __shfl_xor_sync(FULL, diff, d);
	diff = __shfl_sync(FULL, diff, 0);          // same value in every lane already; the broadcast makes the `continue` below provably uniform
	uint64_t *src = keys, *dst = tmp;
	for (int shift = 0; shift < 64; shift += 8) {
		if (((diff >> shift) & 0xff) == 0) continue;
		for (int b = lane; b < 256; b += 32) hist[b] = 0;
		__syncwarp();
		for (int base = 0; base < n; base += 32) {     // histogram without atomics (shared-memory atomics compile to retry loops
			const int k = base + lane;                   // that cost ptxas its convergence proof for the whole kernel): the first
			const int dig = k < n ? (int)(src[k] >> shift & 0xff) : 256;      // lane of each digit group adds the group size
			const unsigned peers = __match_any_sync(FULL, dig);
			if (k < n && (peers & lanemask_lt(lane)) == 0) hist[dig] += __popc(peers);
			__syncwarp();
		}
		{	// exclusive prefix over 256 bins: 8 consecutive bins per lane
			int loc[8], sum = 0;
#pragma unroll
			for (int q = 0; q < 8; ++q) loc[q] = hist[lane * 8 + q], sum += loc[q];
			int incl = sum;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const int o = __shfl_up_sync(FULL, incl, d);
				if (lane >= d) incl += o;
			}
			int run = incl - sum;
#pragma unroll
			for (int q = 0; q < 8; ++q) hist[lane * 8 + q] = run, run += loc[q];
		}
		__syncwarp();
		for (int base = 0; base < n; base += 32) {     // stable scatter, 32 keys at a time in input order
			const int k = base + lane;
			const bool act = k < n;
			const uint64_t key = act ? src[k] : 0;
			const int dig = act ? (int)(key >> shift & 0xff) : 256;
			const unsigned peers = __match_any_sync(FULL, dig);
			const int rank = __popc(peers & lanemask_lt(lane));
			int pos = 0;
			if (act) pos = hist[dig] + rank;
			__syncwarp();
			if (act) {
				dst[pos] = key;
				if (rank == 0) hist[dig] += __popc(peers);
			}
			__syncwarp();
		}
		uint64_t *t = src; src = dst; dst = t;
	}
	if (src != keys) for (int k = lane; k < n; k += 32) keys[k] = src[k];
	__syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// Everything after the fill for one read: chain.c:348-422
// ---------------------------------------------------------------------------------------------------------------
// Output (chain.c:412-420's copies, done once): the read's warp reserves its share of the packed output arrays from two atomic
// cursors (uo / bo are returned as the read's offsets) and writes u[] and either the chained anchors themselves (b, 16 B each)
// or their indices inside the read (bi, 4 B each: the caller still holds a[]).  The layout is packed but not in read order.
struct OutArgs {
	unsigned long long *cursor;     // [0] entries of u handed out, [1] entries of b / bi handed out
	uint64_t *u;
	ulonglong2 *b;                  // or nullptr
	int32_t *bi;                    // or nullptr
};

__device__ void extract_chains(const mm2b_params_t &par, const ReadCtx &rc, const OutArgs &out, int32_t *smem, int lane, int &n_u_out, int &n_v_out, int &status,
                               unsigned long long &uo_out, unsigned long long &bo_out)
{
	const int n = rc.n;
	const int32_t *F = rc.F;
	int32_t *P = rc.P, *V = rc.V, *T = rc.T;
	uint64_t *U = rc.U;

	// chain.c:352-367: chain ends whose path peak reaches min_sc, each walked back to that peak
	int n_u = 0;
	for (int base = 0; base < n; base += 32) {
		const int k = base + lane;
		bool is_end = false;
		uint64_t key = 0;
		if (k < n && T[k] == 0 && V[k] >= par.min_sc) {          // no successor (marks: dp_fill's block epilogue)
			is_end = true;
			int j = k;
			while (j >= 0 && F[j] < V[j]) j = P[j];
			if (j < 0) j = k;
			key = (uint64_t)(uint32_t)F[j] << 32 | (uint32_t)j;
		}
		__syncwarp();
		const unsigned m = __ballot_sync(FULL, is_end);
		MM2B_CHK(n_u + __popc(m) <= n, 0x20);
		if (is_end) U[n_u + __popc(m & lanemask_lt(lane))] = key;
		__syncwarp();
		n_u += __popc(m);
	}
	__syncwarp();
	if (n_u == 0) { n_u_out = 0, n_v_out = 0, status = MM2B_READ_NO_CHAIN; return; }   // chain.c:355-358
	status = MM2B_READ_OK;

	// chain.c:368-372: descending order (a total order on values, so any correct sort gives the reference's result)
	if (n_u > 1) {
		if (n_u <= 32) {
			const uint64_t key = lane < n_u ? U[lane] : 0;
			__syncwarp();
			int rank = 0;
			for (int t = 0; t < n_u; ++t) {     // two ends can share a peak => equal keys: break ties by position
				const uint64_t kt = __shfl_sync(FULL, key, t);
				rank += kt > key || (kt == key && t < lane);
			}
			__syncwarp();
			if (lane < n_u) U[rank] = key;
		} else {
			warp_radix_sort_u64(U, rc.X, n_u, smem, lane);
			for (int k = lane; k < n_u / 2; k += 32) {
				const uint64_t t = U[k];
				U[k] = U[n_u - 1 - k], U[n_u - 1 - k] = t;
			}
		}
		__syncwarp();
	}

	// chain.c:374-391: greedy backtrack in score order.  Chains are taken one after the other (a chain stops where a better
	// one already passed; used-marks persist even when a candidate is dropped), but the walk along one chain is done 32 anchors
	// at a time: the warp loads the aligned block of p[] that holds the next anchor of the walk, resolves the stretch of the
	// path inside the block by pointer jumping over the lanes (links only point to smaller indices), appends it to PATH and
	// follows the link that leaves the block.  Chains of neighbouring anchors — the bulk of all chained anchors — cost a few
	// warp instructions per anchor instead of one dependent load each.
	// The used-mark of an anchor (t[j] = 1 at chain.c:381) is folded into its p[] word: a used anchor stores -3 - p (<= -2).
	// p[] has no reader after the backtrack.
	int32_t *PATH = V;      // v[] is dead from here on, exactly like the reference reuses it (chain.c:380)
	int n_v = 0, n_kept = 0;
	for (int i = 0; i < n_u; ++i) {
		const uint64_t key = U[i];
		const int n_v0 = n_v;
		int cur = __shfl_sync(FULL, (int32_t)key, 0);        // next anchor of the walk (through a shuffle: see the convergence note)
		bool first = true;                                   // do-while of chain.c:379-383: the first anchor is taken unconditionally
		int stop;                                            // the j the reference's loop ends with: -1 = past the head, else the first used anchor
		for (;;) {
			const int base = cur & ~31, e = cur & 31;
			const int k = base + lane;
			const int32_t w = k < n ? P[k] : -1;
			const bool used = w <= -2;
			const int32_t pp = used ? -3 - w : w;            // the anchor's predecessor, mark removed
			const unsigned usedmask = __ballot_sync(FULL, used);
			if (!first && (usedmask >> e & 1u)) { stop = cur; break; }                 // already on a better chain
			int nxt = (pp >= base && !(usedmask >> ((pp - base) & 31) & 1u)) ? pp - base : -1;   // link followed inside the block
			unsigned mask = 1u << lane;
			while (__any_sync(FULL, nxt >= 0)) {             // <= 5 rounds
				const int src = nxt >= 0 ? nxt : lane;
				const unsigned m2 = __shfl_sync(FULL, mask, src);
				const int n2 = __shfl_sync(FULL, nxt, src);
				if (nxt >= 0) mask |= m2, nxt = n2;
			}
			const unsigned path = __shfl_sync(FULL, mask, e);                        // this block's stretch of the walk
			const int32_t out = __shfl_sync(FULL, pp, lowest_lane(path));              // the link that leaves it
			MM2B_CHK(n_v + __popc(path) <= n && cur >= 0 && cur < n, 0x40);
			if (path >> lane & 1u) {
				PATH[n_v + __popc(path >> 1 >> lane)] = k;                           // walk order = descending index
				if (!used) P[k] = -3 - w;
			}
			__syncwarp();
			n_v += __popc(path);
			first = false;
			if (out < 0 || out >= base) { stop = out; break; }                       // past the head, or an in-block link that was not followed: used
			cur = out;
		}
		const int len = n_v - n_v0;
		const int32_t f_stop = __shfl_sync(FULL, stop >= 0 ? F[stop] : 0, 0);
		const int32_t sc = (int32_t)(key >> 32) - f_stop;                            // chain.c:385-388
		const bool keep = len >= par.min_cnt && (stop < 0 || sc >= par.min_sc);
		// (stored whether kept or not — slot n_kept <= i is already consumed — so that the branch stays lane-only; see the convergence note)
		if (lane == 0) U[n_kept] = (uint64_t)(stop < 0 ? (uint32_t)(key >> 32) : (uint32_t)sc) << 32 | (uint64_t)(uint32_t)len;
		__syncwarp();
		n_kept += keep;
		if (!keep) n_v = n_v0;
	}
	__syncwarp();
	n_u = n_kept;
	n_u_out = n_u, n_v_out = n_v;
	if (n_u == 0) return;

	// chain.c:405-411: w[i] = (x of the chain's first anchor, start-in-PATH << 32 | i); F, P and X are dead, W aliases them
	W16 *W = (W16*)rc.F;
	{
		int carry = 0;
		for (int base = 0; base < n_u; base += 32) {
			const int i = base + lane;
			const int len = i < n_u ? (int32_t)U[i] : 0;
			int incl = len;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const int o = __shfl_up_sync(FULL, incl, d);
				if (lane >= d) incl += o;
			}
			const int k0 = carry + incl - len;
			if (i < n_u) {
				W16 w;
				w.x = rc.A[PATH[k0 + len - 1]].x;
				w.y = (uint64_t)(uint32_t)k0 << 32 | (uint32_t)i;
				W[i] = w;
			}
			__syncwarp();
			carry += __shfl_sync(FULL, incl, 31);
		}
	}
	__syncwarp();
	if (n_u > 1) {
		if (n_u <= 64) {        // stable (insertion sort in the reference): rank by (x, original index)
			W16 w0 = {0, 0}, w1 = {0, 0};
			if (lane < n_u) w0 = W[lane];
			if (lane + 32 < n_u) w1 = W[lane + 32];
			__syncwarp();
			int r0 = 0, r1 = 0;
			for (int t = 0; t < n_u; ++t) {
				const uint64_t xt = W[t].x;
				r0 += xt < w0.x || (xt == w0.x && t < lane);
				r1 += xt < w1.x || (xt == w1.x && t < lane + 32);
			}
			__syncwarp();
			if (lane < n_u) W[r0] = w0;
			if (lane + 32 < n_u) W[r1] = w1;
		} else if (lane == 0) {
			// worklist of pending (begin, count, shift) ranges: T is dead once the chain ends have been read off it
			int3 *work = (int3*)(rc.T);
			const int cap = (int)((int64_t)n * 4 / (int64_t)sizeof(int3));
			flag_sort_by_x_lane0(W, n_u, smem, work, cap);
		}
		__syncwarp();
	}

	// chain.c:412-420: final u[] and b[] (or the indices of b[]'s anchors), written straight into the packed output
	unsigned long long uo = 0, bo = 0;
	if (lane == 0) uo = atomicAdd(&out.cursor[0], (unsigned long long)n_u), bo = atomicAdd(&out.cursor[1], (unsigned long long)n_v);
	__syncwarp();
	uo = __shfl_sync(FULL, uo, 0), bo = __shfl_sync(FULL, bo, 0);
	uo_out = uo, bo_out = bo;
	uint64_t *u_dst = out.u + uo;
	int pos = 0;
	for (int i = 0; i < n_u; ++i) {
		const W16 w = W[i];
		const int src = (int32_t)w.y, k0 = (int32_t)(w.y >> 32);
		const uint64_t uu = U[src];
		const int len = (int32_t)uu;
		if (lane == 0) u_dst[i] = uu;
		MM2B_CHK(src >= 0 && src < n_u && k0 >= 0 && len >= 1 && k0 + len <= n_v && pos + len <= n_v, 0x80);
		if (out.bi) {
			int32_t *dst = out.bi + bo + pos;
			for (int j = lane; j < len; j += 32) dst[j] = PATH[k0 + (len - 1 - j)];
		} else {
			ulonglong2 *dst = out.b + bo + pos;
			for (int j = lane; j < len; j += 32) dst[j] = __ldg(rc.A + PATH[k0 + (len - 1 - j)]);
		}
		__syncwarp();
		pos += len;
	}
	__syncwarp();
}

// ---------------------------------------------------------------------------------------------------------------
// K1: persistent warp-per-read kernel
// ---------------------------------------------------------------------------------------------------------------
// Convergence note: every lane-dependent `if` / lane-strided loop in this kernel is followed by an explicit __syncwarp()
// before the next loop back-edge or warp collective.  Without it ptxas cannot prove that the warp is converged anywhere inside
// the persistent loop and guards EVERY shuffle/vote/redux with a `BRA.DIV` + `WARPSYNC.COLLECTIVE` slow path (2 extra
// instructions per collective: ~16 of the ~150 per anchor).  Found by bisecting SASS; one missing __syncwarp() after the
// `if (lane == 0) ...` at the end of a read was enough to poison the whole kernel.
// mm_chain_dp for read r by one warp: prologue (chain.c:46-49), fill, extraction.  RING_N / COOP select the kernel flavour.
template <int RING_N, bool COUNT, bool COOP>
__device__ __forceinline__ void chain_one_read(const BatchArgs &args, int64_t r, int32_t *ring, int32_t *coop, int lane,
                                               unsigned long long &n_chunks, unsigned long long &n_general, unsigned long long &n_cells,
                                               unsigned long long &n_window)
{
	const int64_t o = args.off[r];
	const int64_t n64 = args.off[r + 1] - o;
	if (n64 <= 0) {                                                               // chain.c:38-41
		if (lane == 0) args.n_u[r] = 0, args.n_v[r] = 0, args.status[r] = MM2B_READ_EMPTY, args.u_off[r] = 0, args.b_off[r] = 0;
		__syncwarp();
		return;
	}
	ReadCtx rc;
	rc.n = (int)n64;
	rc.A = (const ulonglong2*)(args.a + o);
	uint8_t *s = args.scratch + (size_t)o * SCRATCH_BYTES_PER_ANCHOR;
	const size_t n = (size_t)n64;
	rc.F = (int32_t*)s, rc.P = (int32_t*)(s + 4 * n), rc.X = (uint64_t*)(s + 8 * n);
	rc.V = (int32_t*)(s + 16 * n), rc.T = (int32_t*)(s + 20 * n), rc.U = (uint64_t*)(s + 24 * n);

	// chain.c:46-49: zero t[], sum the 8-bit q_span fields; also find out whether every anchor carries the same segment id
	uint64_t sum = 0;
	uint32_t seg_diff = 0;
	const uint32_t seg0 = (uint32_t)(__ldg(&rc.A[0].y) >> SEG_SHIFT & 0xff);
	uint32_t max_xl = 0;                     // largest low word of x (reference position): see the window search in dp_fill
	for (int k = lane; k < rc.n; k += 32) {
		const ulonglong2 t = __ldg(rc.A + k);
		sum += t.y >> 32 & 0xff;
		seg_diff |= (uint32_t)(t.y >> SEG_SHIFT & 0xff) ^ seg0;
		max_xl = max_xl > (uint32_t)t.x ? max_xl : (uint32_t)t.x;
		rc.T[k] = 0;
	}
	__syncwarp();
	const bool lo_safe = args.par.max_dist_x >= 0 && (uint64_t)__reduce_max_sync(FULL, max_xl) + (uint64_t)args.par.max_dist_x < (1ull << 32);
#pragma unroll
	for (int d = 16; d; d >>= 1) {
		sum += __shfl_xor_sync(FULL, sum, d);
		seg_diff |= __shfl_xor_sync(FULL, seg_diff, d);
	}
	// `.01 * (float)sum_qspan / n` — double arithmetic on a float-rounded sum, rounded once more to float
	const float avg = __double2float_rn(__ddiv_rn(__dmul_rn(.01, (double)__ull2float_rn(sum)), (double)n64));
	// (through a vote so that ptxas sees a warp-uniform branch: see the convergence note above)
	const bool general = __any_sync(FULL, seg_diff != 0) || args.par.is_cdna || args.par.gap_scale != 1.0f || args.par.bw >= (1 << 24) || args.par.bw < 0
	                  || args.par.n_segs > 1 || args.par.max_dist_x <= 0 || args.par.max_dist_y <= 0;
	__syncwarp();
	if (general) {
		++n_general;
		dp_fill<RING_N, true, COUNT, false>(args.par, rc, avg, lo_safe, ring, coop, lane, n_chunks, n_cells, n_window);
	} else {
		dp_fill<RING_N, false, COUNT, COOP>(args.par, rc, avg, lo_safe, ring, coop, lane, n_chunks, n_cells, n_window);
	}
	if (args.dbg_fpv) {      // test hook (MM2B_KEEP_FPV=1): keep f/p/v as they are at chain.c:238, before the extraction reuses v
		int32_t *d = args.dbg_fpv + o;
		for (int k = lane; k < rc.n; k += 32) d[k] = rc.F[k], d[args.n_anchors + k] = rc.P[k], d[2 * args.n_anchors + k] = rc.V[k];
		__syncwarp();
	}
	int n_u = 0, n_v = 0, status = MM2B_READ_OK;
	unsigned long long uo = 0, bo = 0;
	OutArgs out;
	out.cursor = args.out_cursor, out.u = args.u, out.b = (ulonglong2*)args.b, out.bi = args.bi;
	extract_chains(args.par, rc, out, ring, lane, n_u, n_v, status, uo, bo);
	if (lane == 0) args.n_u[r] = n_u, args.n_v[r] = n_v, args.status[r] = status, args.u_off[r] = (int64_t)uo, args.b_off[r] = (int64_t)bo;
	__syncwarp();
}

template <bool COUNT>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, CTAS_PER_SM)
chain_reads_kernel(const BatchArgs args)
{
	__shared__ __align__(16) int32_t smem_ring[WARPS_PER_CTA][RING_ARRAYS * LIGHT_RING];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int32_t *ring = smem_ring[warp];
	unsigned long long n_chunks = 0, n_general = 0, n_cells = 0, n_window = 0;

	for (;;) {
		int64_t slot = 0;
		if (lane == 0) slot = atomicAdd(args.work_counter, 1);
		__syncwarp();
		slot = __shfl_sync(FULL, slot, 0);
		if (slot >= args.n_reads) break;
		const int64_t r = args.order ? args.order[slot] : slot;
		if (args.heavy_flag) {                                                        // reads taken by the heavy-read kernel
			int hv = 0;
			if (lane == 0) hv = args.heavy_flag[r];
			__syncwarp();
			if (__shfl_sync(FULL, hv, 0)) continue;
		}
		chain_one_read<LIGHT_RING, COUNT, false>(args, r, ring, nullptr, lane, n_chunks, n_general, n_cells, n_window);
	}
	if (COUNT) {
#pragma unroll
		for (int d = 16; d; d >>= 1) n_window += __shfl_xor_sync(FULL, n_window, d);
	}
	if (lane == 0) {
		if (n_window) atomicAdd(&args.counters[3], n_window);
		if (n_chunks) atomicAdd(&args.counters[0], n_chunks);
		if (n_general) atomicAdd(&args.counters[1], n_general);
		if (n_cells) atomicAdd(&args.counters[2], n_cells);
	}
}

// ---------------------------------------------------------------------------------------------------------------
// Long-read segmenting.  If a[k].x > a[k-1].x + max_dist_x then no anchor from k on can have a predecessor before k (chain.c:192: the
// window start st moves up to k and never back), so the DP over [k, ...) does not depend on anything before k: k is a cut point
// and the pieces between cut points are independent fills.  A read whose cut points give at least two pieces of seg_min_piece
// anchors (reads spanning several loci: chimeric reads, repeats with copies further apart than max_dist_x) is cut here; its pieces
// are filled concurrently by different warps and the warp that finishes the read's last piece runs the extraction, which is per
// read (chain.c:348-422).  A read that is one collinear chain has no usable cut point: most of a simulated read's cut points
// isolate single seed hits, and those are settled 32 at a time by the fill anyway.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) segment_kernel(const BatchArgs args)
{
	__shared__ int32_t starts[SEG_MAX_PIECES + 1];
	__shared__ unsigned long long sum_s;
	__shared__ unsigned seg_diff_s, max_xl_s;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint64_t win = (uint64_t)(int64_t)args.par.max_dist_x;
	// One CTA per long read.  The processing order (longest reads first, 8 length buckets per octave) says where the long reads are:
	// slots are taken until the reads are clearly shorter than seg_min_read.
	for (int64_t slot = blockIdx.x; slot < args.n_reads; slot += gridDim.x) {
		const int64_t r = args.order ? args.order[slot] : slot;
		const int64_t o = args.off[r], n64 = args.off[r + 1] - o;
		if (args.order && n64 * 10 < (int64_t)args.seg_min_read * 9) break;           // (block-uniform) the rest of the order is shorter still
		if (n64 < args.seg_min_read || n64 >= (1ll << 31) || args.heavy_flag[r]) continue;
		const int n = (int)n64, n_blocks = (n + 31) >> 5;
		const ulonglong2 *A = (const ulonglong2*)(args.a + o);
		uint8_t *scr = args.scratch + (size_t)o * SCRATCH_BYTES_PER_ANCHOR;
		int32_t *T = (int32_t*)(scr + 20 * (size_t)n);
		uint32_t *cutw = (uint32_t*)(scr + 24 * (size_t)n);                 // the read's U region: one word of cut flags per 32 anchors
		if (threadIdx.x == 0) sum_s = 0, seg_diff_s = 0, max_xl_s = 0;
		__syncthreads();
		uint64_t sum = 0;
		uint32_t seg_diff = 0, max_xl = 0;
		const uint32_t seg0 = (uint32_t)(__ldg(&A[0].y) >> SEG_SHIFT & 0xff);
		// Pass 1, all warps, no dependency between blocks of 32 anchors: which anchors are cut points, and the sums the read's
		// prologue needs (chain.c:46-49); t[] is zeroed here for the pieces' fills.
#pragma unroll 2
		for (int blk = warp; blk < n_blocks; blk += 8) {
			const int k = (blk << 5) + lane;
			uint64_t x = 0, xp = 0;
			if (k < n) {
				const ulonglong2 t = __ldg(A + k);
				x = t.x;
				sum += t.y >> 32 & 0xff;
				seg_diff |= (uint32_t)(t.y >> SEG_SHIFT & 0xff) ^ seg0;
				max_xl = max_xl > (uint32_t)t.x ? max_xl : (uint32_t)t.x;
				T[k] = 0;
				if (k > 0) xp = __ldg(&A[k - 1].x);
			}
			// (no cut where x + max_dist_x wraps: fewer cuts are always correct)
			const unsigned cuts = __ballot_sync(FULL, k < n && k > 0 && xp + win >= xp && x > xp + win);
			if (lane == 0) cutw[blk] = cuts;
		}
#pragma unroll
		for (int d = 16; d; d >>= 1) {
			sum += __shfl_xor_sync(FULL, sum, d);
			seg_diff |= __shfl_xor_sync(FULL, seg_diff, d);
			const uint32_t m = __shfl_xor_sync(FULL, max_xl, d);
			max_xl = max_xl > m ? max_xl : m;
		}
		if (lane == 0) atomicAdd(&sum_s, (unsigned long long)sum), atomicOr(&seg_diff_s, seg_diff), atomicMax(&max_xl_s, max_xl);
		__syncthreads();                                                    // cutw[] and the sums are in place (block-scope visibility)
		if (warp != 0) continue;
		sum = sum_s, seg_diff = seg_diff_s, max_xl = max_xl_s;
		// Pass 2: a piece must hold seg_min_piece anchors that HAVE a predecessor in reach (the sequential work of the fill; an anchor
		// that is a cut point is an isolated seed hit as far as what lies before it goes), and starts at the first cut point where
		// the running count since the previous start has got there.  Lane = block of 32 anchors, 32 blocks per step; a piece spans at
		// least seg_min_piece >= 1024 anchors, so a step holds at most one start.
		int n_p = 1, work = 0, work_at_start = 0;
		if (lane == 0) starts[0] = 0;
		__syncwarp();
		for (int b0 = 0; b0 < n_blocks; b0 += 32) {
			const int blk = b0 + lane;
			unsigned cuts = 0, valid = 0;
			if (blk < n_blocks) {
				cuts = cutw[blk];
				const int left = n - (blk << 5);
				valid = left >= 32 ? FULL : ((1u << left) - 1u);
				if (blk == 0) valid &= ~1u;                                 // anchor 0 has nothing before it
			}
			const unsigned linked = valid & ~cuts;
			int incl = __popc(linked);
			const int mine = incl;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const int v = __shfl_up_sync(FULL, incl, d);
				if (lane >= d) incl += v;
			}
			const int w_before = work + incl - mine;                        // linked anchors before this block
			// first cut point of this block with at least `need` linked anchors of the block in front of it
			const int need = work_at_start + args.seg_min_piece - w_before;
			unsigned cand = cuts;
			if (need > 0) {
				if (need > mine) cand = 0;
				else {
					const unsigned p = __fns(linked, 0, need);                 // position of the need-th linked anchor
					cand = cuts & ~((2u << p) - 1u);                         // cut points after it
				}
			}
			const unsigned has = __ballot_sync(FULL, cand != 0);
			if (has && n_p < SEG_MAX_PIECES) {
				const int bl = lowest_lane(has);
				const int cl = lowest_lane(__shfl_sync(FULL, cand, bl));
				const unsigned lk = __shfl_sync(FULL, linked, bl);
				const int wb = __shfl_sync(FULL, w_before, bl);
				if (lane == 0) starts[n_p] = ((b0 + bl) << 5) + cl;
				++n_p, work_at_start = wb + __popc(lk & lanemask_lt(cl));
			}
			work += __shfl_sync(FULL, incl, 31);
		}
		if (work - work_at_start < args.seg_min_piece && n_p > 1) --n_p;     // a tail with little work stays with the piece before it
		if (n_p < 2) continue;
		const bool lo_safe = args.par.max_dist_x >= 0 && (uint64_t)max_xl + (uint64_t)args.par.max_dist_x < (1ull << 32);
		const bool general = seg_diff != 0 || args.par.is_cdna || args.par.gap_scale != 1.0f || args.par.bw >= (1 << 24) || args.par.bw < 0
		                  || args.par.n_segs > 1 || args.par.max_dist_x <= 0 || args.par.max_dist_y <= 0;
		int idx = 0, item0 = 0;
		if (lane == 0) {
			idx = atomicAdd(&args.seg_ctl[0], 1);
			if (idx < args.seg_cap) item0 = atomicAdd(&args.seg_ctl[1], n_p);
		}
		idx = __shfl_sync(FULL, idx, 0), item0 = __shfl_sync(FULL, item0, 0);
		if (idx >= args.seg_cap) continue;                                   // (more long reads than room: the rest stay whole)
		__syncwarp();
		SegRead *sr = args.seg_reads + idx;
		if (lane == 0) {
			sr->read = (int32_t)r, sr->n_pieces = n_p, sr->flags = (lo_safe ? 1 : 0) | (general ? 2 : 0);
			sr->avg = __double2float_rn(__ddiv_rn(__dmul_rn(.01, (double)__ull2float_rn(sum)), (double)n64));     // chain.c:49
			sr->start[n_p] = n;
			args.seg_done[idx] = 0;
			args.heavy_flag[r] = 2;
		}
		if (lane < n_p) sr->start[lane] = starts[lane], args.seg_items[item0 + lane] = idx << 8 | lane;
		__syncwarp();
	}
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32, CTAS_PER_SM)
chain_pieces_kernel(const BatchArgs args)
{
	__shared__ __align__(16) int32_t smem_ring[WARPS_PER_CTA][RING_ARRAYS * LIGHT_RING];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	int32_t *ring = smem_ring[warp];
	unsigned long long n_chunks = 0, n_general = 0, n_cells = 0, n_window = 0;
	int n_items = 0;
	if (lane == 0) n_items = args.seg_ctl[1];
	n_items = __shfl_sync(FULL, n_items, 0);
	for (;;) {
		int it = 0;
		if (lane == 0) it = atomicAdd(&args.seg_ctl[2], 1);
		__syncwarp();
		it = __shfl_sync(FULL, it, 0);
		if (it >= n_items) break;
		const int item = args.seg_items[it], idx = item >> 8, piece = item & 0xff;
		const SegRead *sr = args.seg_reads + idx;
		const int64_t r = sr->read, o = args.off[r];
		ReadCtx rc;
		rc.n = (int)(args.off[r + 1] - o);
		rc.A = (const ulonglong2*)(args.a + o);
		uint8_t *s = args.scratch + (size_t)o * SCRATCH_BYTES_PER_ANCHOR;
		const size_t n = (size_t)rc.n;
		rc.F = (int32_t*)s, rc.P = (int32_t*)(s + 4 * n), rc.X = (uint64_t*)(s + 8 * n);
		rc.V = (int32_t*)(s + 16 * n), rc.T = (int32_t*)(s + 20 * n), rc.U = (uint64_t*)(s + 24 * n);
		const int flags = __shfl_sync(FULL, sr->flags, 0);                   // (through a shuffle: warp-uniform by construction)
		const int k0 = sr->start[piece], k1 = sr->start[piece + 1];
		if (flags & 2) {
			++n_general;
			dp_fill<LIGHT_RING, true, false, false, true>(args.par, rc, sr->avg, (flags & 1) != 0, ring, nullptr, lane, n_chunks, n_cells, n_window, k0, k1);
		} else dp_fill<LIGHT_RING, false, false, false, true>(args.par, rc, sr->avg, (flags & 1) != 0, ring, nullptr, lane, n_chunks, n_cells, n_window, k0, k1);
		__threadfence();                                                    // this piece's f / p / v / t before the count below
		int done = 0;
		if (lane == 0) done = atomicAdd(&args.seg_done[idx], 1);
		__syncwarp();
		done = __shfl_sync(FULL, done, 0);
		if (done != sr->n_pieces - 1) continue;
		__threadfence();                                                    // the read's last piece: every other piece is in place
		int n_u = 0, n_v = 0, status = MM2B_READ_OK;
		unsigned long long uo = 0, bo = 0;
		OutArgs out;
		out.cursor = args.out_cursor, out.u = args.u, out.b = (ulonglong2*)args.b, out.bi = args.bi;
		extract_chains(args.par, rc, out, ring, lane, n_u, n_v, status, uo, bo);
		if (lane == 0) args.n_u[r] = n_u, args.n_v[r] = n_v, args.status[r] = status, args.u_off[r] = (int64_t)uo, args.b_off[r] = (int64_t)bo;
		__syncwarp();
	}
	if (lane == 0 && n_general) atomicAdd(&args.counters[1], n_general);
	if (blockIdx.x == 0 && threadIdx.x == 0) args.counters[5] = (unsigned long long)(args.seg_ctl[0] < args.seg_cap ? args.seg_ctl[0] : args.seg_cap);
}

// ---------------------------------------------------------------------------------------------------------------
// K1h: heavy reads, one CTA per read (see "Heavy reads" above).  Warp 0 is the read's warp; the others serve its long scans.
// ---------------------------------------------------------------------------------------------------------------
constexpr int HEAVY_SMEM_BYTES = (RING_ARRAYS * HEAVY_RING + COOP_WORDS) * 4;

__global__ void __launch_bounds__(HEAVY_WARPS * 32, 1)
chain_heavy_kernel(const BatchArgs args)
{
	extern __shared__ __align__(16) int32_t heavy_smem[];
	int32_t *ring = heavy_smem, *coop = heavy_smem + RING_ARRAYS * HEAVY_RING;
	const int lane = threadIdx.x & 31;
	const int warp = __shfl_sync(FULL, (int)(threadIdx.x >> 5), 0);      // through a shuffle: warp-uniform by construction (convergence note)
	if (warp != 0) {
		DpConst c;
		c.max_dist_x = args.par.max_dist_x, c.max_dist_y = args.par.max_dist_y, c.bw = args.par.bw, c.max_skip = args.par.max_skip, c.max_iter = args.par.max_iter;
		c.max_dq_same = args.par.max_dist_x < args.par.max_dist_y ? args.par.max_dist_x : args.par.max_dist_y;
		c.cap_dr = false, c.is_cdna = false, c.avg = 0.f, c.gap_scale = 1.0;
		Ring rg;
		rg.a = (int4*)ring, rg.b = (int2*)(ring + 4 * HEAVY_RING), rg.coop = coop;
		for (;;) {
			cta_bar(BAR_JOB);
			int v = 0;
			if (lane < 8) v = coop[lane];
			__syncwarp();
			if (__shfl_sync(FULL, v, 0) == JOB_EXIT) break;
			c.avg = __int_as_float(__shfl_sync(FULL, v, 6));
			int32_t f, jj;
			coop_scan<HEAVY_RING>(c, rg, lane, warp, __shfl_sync(FULL, v, 1), __shfl_sync(FULL, v, 2), __shfl_sync(FULL, v, 3), __shfl_sync(FULL, v, 4),
			                      __shfl_sync(FULL, v, 5), f, jj);
		}
		return;
	}
	unsigned long long n_chunks = 0, n_general = 0, n_cells = 0, n_window = 0;
	int n_heavy = 0;
	if (lane == 0) n_heavy = *args.heavy_count;
	__syncwarp();
	n_heavy = __shfl_sync(FULL, n_heavy, 0);
	n_heavy = n_heavy < args.heavy_cap ? n_heavy : args.heavy_cap;            // (the counter also counted the reads turned away)
	for (;;) {
		int slot = 0;
		if (lane == 0) slot = atomicAdd(args.heavy_counter, 1);
		__syncwarp();
		slot = __shfl_sync(FULL, slot, 0);
		if (slot >= n_heavy) break;
		chain_one_read<HEAVY_RING, false, true>(args, args.heavy_list[slot], ring, coop, lane, n_chunks, n_general, n_cells, n_window);
	}
	if (lane == 0) {
		coop[0] = JOB_EXIT;
		if (n_general) atomicAdd(&args.counters[1], n_general);
		if (blockIdx.x == 0) args.counters[4] = (unsigned long long)n_heavy;
	}
	__syncwarp();
	cta_bar(BAR_JOB);
}

// Which reads go to the heavy-read kernel: an estimate of the read's window cells (sum over anchors of min(i - st, max_iter),
// chain.c:192-193) from 32 sampled anchors.  One warp per read.
constexpr int HEAVY_MIN_ANCHORS = 64;

__global__ void __launch_bounds__(256) classify_heavy_kernel(const BatchArgs args)
{
	const int lane = threadIdx.x & 31;
	const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
	for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < args.n_reads; r += n_warps) {
		const int64_t o = args.off[r], n = args.off[r + 1] - o;
		bool heavy = false;
		// (a window has at most max_iter cells, so a read with fewer than heavy_min_cells / max_iter anchors cannot qualify: no sampling)
		if (n >= HEAVY_MIN_ANCHORS && n * (int64_t)args.par.max_iter >= args.heavy_min_cells) {
			const ulonglong2 *A = (const ulonglong2*)(args.a + o);
			const uint64_t win = (uint64_t)(int64_t)args.par.max_dist_x;
			const int64_t k = (int64_t)(lane + 1) * n / 33;
			const uint64_t x = __ldg(&A[k].x);
			int64_t lo = 0, hi = k;
			while (lo < hi) {
				const int64_t mid = (lo + hi) >> 1;
				if (x > __ldg(&A[mid].x) + win) lo = mid + 1;
				else hi = mid;
			}
			__syncwarp();
			long long w = k - lo;
			w = w > args.par.max_iter ? args.par.max_iter : w;
#pragma unroll
			for (int d = 16; d; d >>= 1) w += __shfl_xor_sync(FULL, w, d);
			const long long mean_window = w / 32;
			heavy = mean_window > COOP_MIN_CELLS && mean_window * n >= args.heavy_min_cells;
		}
		if (lane == 0) {
			if (heavy) {                                          // beyond one wave of CTAs the warp-per-read kernel has the higher throughput
				const int idx = atomicAdd(args.heavy_count, 1);
				if (idx < args.heavy_cap) args.heavy_list[idx] = (int32_t)r;
				else heavy = false;
			}
			args.heavy_flag[r] = heavy ? 1 : 0;
		}
		__syncwarp();
	}
}

// ---------------------------------------------------------------------------------------------------------------
// K0: processing order, longest reads first (LPT) — 256 logarithmic length buckets, order inside a bucket is free
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int length_bucket(int64_t n)
{
	const float l = __log2f((float)(n + 1));
	int b = (int)(l * 8.f);
	b = b > 255 ? 255 : b;
	return 255 - b;
}
__global__ void order_hist_kernel(int64_t n_reads, const int64_t *off, int *hist)
{
	const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r < n_reads) atomicAdd(&hist[length_bucket(off[r + 1] - off[r])], 1);
}
__global__ void order_scan_kernel(int *hist)        // 256 threads, one block: exclusive prefix in place
{
	__shared__ int s[256];
	const int t = threadIdx.x;
	s[t] = hist[t];
	__syncthreads();
	for (int d = 1; d < 256; d <<= 1) {
		const int v = t >= d ? s[t - d] : 0;
		__syncthreads();
		s[t] += v;
		__syncthreads();
	}
	hist[t] = s[t] - hist[t];
}
__global__ void order_scatter_kernel(int64_t n_reads, const int64_t *off, int *cursor, int32_t *order)
{
	const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r < n_reads) order[atomicAdd(&cursor[length_bucket(off[r + 1] - off[r])], 1)] = (int32_t)r;
}

// ---------------------------------------------------------------------------------------------------------------
// K-unpack: anchors arrive over PCIe as 8 bytes (low words of x and y) plus run-length lists of the high words — in the sorted
// input the strand/rid word of x changes a few times per read and the flags/q_span/segment word of y hardly ever (it is the
// k-mer length unless the index is homopolymer-compressed).  This kernel restores the 16-byte mm128_t in HBM (24 B of HBM
// traffic per anchor, ~0.3 % of the chaining kernel's time) so that everything downstream sees the reference's layout.
// runs: {first anchor of the run, high word}, sorted by first anchor, run 0 starts at anchor 0.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int last_run_at_or_before(const uint2 *runs, int lo, int hi, uint32_t i)   // largest r in [lo, hi] with runs[r].x <= i
{
	while (lo < hi) {
		const int mid = (lo + hi + 1) >> 1;
		if (__ldg(&runs[mid].x) <= i) lo = mid;
		else hi = mid - 1;
	}
	return lo;
}

__global__ void __launch_bounds__(256) unpack_kernel(int64_t n, const uint2 *lo, const uint2 *xruns, int n_xruns, const uint2 *yruns, int n_yruns, ulonglong2 *a)
{
	__shared__ int range[4];
	for (int64_t base = (int64_t)blockIdx.x * 256; base < n; base += (int64_t)gridDim.x * 256) {
		const uint32_t first = (uint32_t)base, last = (uint32_t)(base + 255 < n - 1 ? base + 255 : n - 1);
		if (threadIdx.x < 4) {          // the runs this block of 256 anchors can touch
			const uint2 *runs = threadIdx.x < 2 ? xruns : yruns;
			const int nr = threadIdx.x < 2 ? n_xruns : n_yruns;
			range[threadIdx.x] = last_run_at_or_before(runs, 0, nr - 1, (threadIdx.x & 1) ? last : first);
		}
		__syncthreads();
		const int64_t i = base + threadIdx.x;
		if (i < n) {
			const uint2 w = __ldg(lo + i);
			const uint32_t xh = __ldg(&xruns[last_run_at_or_before(xruns, range[0], range[1], (uint32_t)i)].y);
			const uint32_t yh = __ldg(&yruns[last_run_at_or_before(yruns, range[2], range[3], (uint32_t)i)].y);
			a[i] = make_ulonglong2((uint64_t)xh << 32 | w.x, (uint64_t)yh << 32 | w.y);
		}
		__syncthreads();
	}
}

// ---------------------------------------------------------------------------------------------------------------
// INT32 issue-rate micro-benchmark (roofline denominator): eight independent integer chains per thread, all SMs.  Each source
// statement `a = max(a + k, b ^ i)` compiles to TWO SASS instructions on sm_100a — LOP3.LUT (the xor) and VIADDMNMX (add and
// max fused) — so the benchmark counts 2 lane-instructions per statement (SASS excerpt: profiles/int32_peak_sass.txt).  The
// three source-level operations (add, xor, max) per statement are what SURVEY.md 8d's "30 ops per cell" counts, hence the
// op-counted peak is 1.5 x the instruction peak; bench.py reports both and says which one each fraction uses.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) int32_peak_kernel(int iters, int *out)
{
	int a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
	const int k = blockIdx.x | 1;
	for (int i = 0; i < iters; ++i) {
#pragma unroll
		for (int u = 0; u < 8; ++u) {
			a0 = max(a0 + k, a1 ^ i); a1 = min(a1 + k, a2 ^ i); a2 = max(a2 + k, a3 ^ i); a3 = min(a3 + k, a4 ^ i);
			a4 = max(a4 + k, a5 ^ i); a5 = min(a5 + k, a6 ^ i); a6 = max(a6 + k, a7 ^ i); a7 = min(a7 + k, a0 ^ i);
		}
	}
	const int r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
	if (r == 0x7fffffff) out[0] = r;
}

}  // anonymous namespace

// ---------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------
int launch_order(int64_t n_reads, const int64_t *off, int32_t *order, int *bucket_scratch, cudaStream_t stream)
{
	if (n_reads <= 0) return 0;
	cudaMemsetAsync(bucket_scratch, 0, 256 * sizeof(int), stream);
	const int grid = (int)((n_reads + 255) / 256);
	order_hist_kernel<<<grid, 256, 0, stream>>>(n_reads, off, bucket_scratch);
	order_scan_kernel<<<1, 256, 0, stream>>>(bucket_scratch);
	order_scatter_kernel<<<grid, 256, 0, stream>>>(n_reads, off, bucket_scratch, order);
	return 3;
}

int launch_chain(const BatchArgs &args, int n_sms, cudaStream_t stream)
{
	if (args.n_reads <= 0) return 0;
	cudaMemsetAsync(args.work_counter, 0, sizeof(int), stream);
	cudaMemsetAsync(args.counters, 0, 6 * sizeof(unsigned long long), stream);
	cudaMemsetAsync(args.out_cursor, 0, 2 * sizeof(unsigned long long), stream);
	int64_t ctas = (args.n_reads + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
	const int64_t resident = (int64_t)n_sms * CTAS_PER_SM;
	if (ctas > resident) ctas = resident;
	int launches = 1;
	if (args.heavy_list) {          // classify, then the heavy reads on CTAs of their own (same stream: a heavy read outlasts everything else anyway)
		static bool attr_set[64];
		int dev = 0;
		cudaGetDevice(&dev);
		if (dev < 64 && !attr_set[dev]) {
			cudaFuncSetAttribute(chain_heavy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HEAVY_SMEM_BYTES);
			attr_set[dev] = true;
		}
		cudaMemsetAsync(args.heavy_count, 0, 2 * sizeof(int), stream);          // heavy_count and heavy_counter are adjacent
		const int64_t warps_needed = args.n_reads, blocks = (warps_needed + 7) / 8;
		classify_heavy_kernel<<<(int)(blocks > 4 * n_sms ? 4 * n_sms : blocks), 256, 0, stream>>>(args);
		chain_heavy_kernel<<<n_sms, HEAVY_WARPS * 32, HEAVY_SMEM_BYTES, stream>>>(args);
		launches += 2;
	} else if (args.heavy_flag) cudaMemsetAsync(args.heavy_flag, 0, (size_t)args.n_reads, stream);
	if (args.heavy_flag && args.seg_reads) {            // cut the long reads that can be cut, fill their pieces side by side
		cudaMemsetAsync(args.seg_ctl, 0, 4 * sizeof(int), stream);
		segment_kernel<<<(int)(args.n_reads > 8 * n_sms ? 8 * n_sms : args.n_reads), 256, 0, stream>>>(args);
		chain_pieces_kernel<<<n_sms * 4, WARPS_PER_CTA * 32, 0, stream>>>(args);
		launches += 2;
	}
	if (args.count_cells) chain_reads_kernel<true><<<(int)ctas, WARPS_PER_CTA * 32, 0, stream>>>(args);
	else chain_reads_kernel<false><<<(int)ctas, WARPS_PER_CTA * 32, 0, stream>>>(args);
	return launches;
}

int heavy_ring_slots() { return HEAVY_RING; }
int heavy_min_window() { return COOP_MIN_CELLS; }

int launch_unpack(int64_t n_anchors, const uint2 *lo, const uint2 *xruns, int n_xruns, const uint2 *yruns, int n_yruns, mm2b_anchor_t *a,
                  int n_sms, cudaStream_t stream)
{
	if (n_anchors <= 0) return 0;
	int64_t blocks = (n_anchors + 255) / 256;
	const int64_t cap = (int64_t)n_sms * 8;
	if (blocks > cap) blocks = cap;
	unpack_kernel<<<(int)blocks, 256, 0, stream>>>(n_anchors, lo, xruns, n_xruns, yruns, n_yruns, (ulonglong2*)a);
	return 1;
}

unsigned debug_flags()
{
#ifdef MM2B_DEBUG_CHECKS
	unsigned f = 0;
	cudaDeviceSynchronize();
	cudaMemcpyFromSymbol(&f, g_dbg_flags, sizeof(f));
	return f | 0x80000000u;         // top bit: this IS the checking build
#else
	return 0;
#endif
}

double measure_int32_peak(int device)
{
	int prev = 0, n_sms = 0;
	cudaGetDevice(&prev);
	if (cudaSetDevice(device) != cudaSuccess) return -1.0;
	cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, device);
	int *out = nullptr;
	cudaMalloc(&out, sizeof(int));
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0), cudaEventCreate(&e1);
	const int iters = 4096, grid = n_sms * 8;
	int32_peak_kernel<<<grid, 256>>>(64, out);      // warm-up
	double best = 0;
	for (int rep = 0; rep < 5; ++rep) {
		cudaEventRecord(e0);
		int32_peak_kernel<<<grid, 256>>>(iters, out);
		cudaEventRecord(e1);
		cudaEventSynchronize(e1);
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		// per inner statement: LOP3 + VIADDMNMX = 2 lane-instructions (see the comment at int32_peak_kernel)
		const double ops = (double)grid * 256.0 * iters * 8.0 * 8.0 * 2.0;
		const double rate = ops / (ms * 1e-3) / 1e9;
		if (rate > best) best = rate;
	}
	cudaEventDestroy(e0), cudaEventDestroy(e1);
	cudaFree(out);
	cudaSetDevice(prev);
	return best;
}

}  // namespace mm2b
