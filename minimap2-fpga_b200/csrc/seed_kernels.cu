// Hand-written sm_100a kernels for minimap2's seeding front end — SURVEY.md 8(f) next-4 (mm_sketch, /root/reference/sketch.c:77-143)
// and next-1 (collect_matches / collect_seed_hits over mm_idx_get, map.c:90-247, index.c:81-98).  Anchors are born in HBM, sorted
// there the way the reference's radix_sort_128x leaves them (map.c:245), and go straight into the chaining kernels.
//
// Design (B200-first; the reference is a per-read sequential scan with a w-entry ring buffer and a khash probe per minimizer):
//   * sketch: one thread per base position, one CTA per tile of 512 positions.  The CTA packs the tile's bases to 2 bits with
//     warp ballots / REDUX.OR, every thread cuts its k-mer and the reverse complement out of two 64-bit words (funnel shift,
//     BREV), hashes it, and the w + 1 hashes around it in shared memory tell it everything mm_sketch's state machine does at its
//     position: the ring buffer's minimum is always the NEWEST minimal entry of the last w positions, so "what is pushed at
//     step t" is a function of X[t-w .. t] and of the distance to the last ambiguous base.  One pass: a tile counts what it pushes,
//     gets its place in the output from a decoupled look-back over the tiles in front of it, and writes the minimizers in the
//     reference's order, duplicates of its first-window rule included.
//   * index: one open-addressing table in HBM (linear probing, load <= 0.5) instead of 2^b khash buckets; one probe thread per
//     minimizer, all minimizers of a sub-batch at once (the probes are dependent HBM accesses: parallelism is what hides them).
//   * matches: one warp per read walks its minimizers 32 at a time: occurrence filter, repeat length (a merge of intervals the
//     reference does sequentially, here a "previous flagged lane" vote), tandem flags, mini_pos and the anchor offsets.
//   * anchors: expanded by the read's warp, sorted by a warp-level LSD radix sort over the key bytes that differ.  Sorted order is
//     unique except among equal keys, whose order in the reference is whatever its in-place MSD radix sort leaves; reads that
//     hold equal keys (a minimizer repeated in the query hitting one reference position) are expanded again and put through an
//     exact replay of that sort (sort_replay.cuh), so a[] is byte-identical to the reference's in every case.
// No tensor cores (nothing here is a contraction) and no collective (reads are independent).
#include "seed_kernels.cuh"
#include "sort_replay.cuh"
#include <limits.h>
#include <stdlib.h>
#include <type_traits>

namespace mm2b {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int HALO = 96;                                    // positions loaded in front of a tile: >= w + k, a multiple of 32
constexpr int TILE_SPAN = SKETCH_TILE + HALO;               // 608 = 19 warps of positions
constexpr uint64_t EMPTY = ~0ull;
constexpr uint64_t SEED_TANDEM = 1ull << 42;                // MM_SEED_TANDEM, mmpriv.h:19

__device__ __forceinline__ unsigned lanemask_lt(int lane) { return (1u << lane) - 1u; }

// seq_nt4_table (sketch.c:9-26): A/a 0, C/c 1, G/g 2, T/t/U/u 3, the byte values 0..3 themselves, everything else 4
__device__ __forceinline__ int nt4(unsigned ch)
{
	if (ch < 4) return (int)ch;
	const unsigned u = ch & 0xdfu;                          // only clears bit 5: 65 comes from 'A' or 'a' alone, and so on
	return u == 65 ? 0 : u == 67 ? 1 : u == 71 ? 2 : (u == 84 || u == 85) ? 3 : 4;
}

// hash64 of sketch.c:28-38
__device__ __forceinline__ uint64_t hash64(uint64_t key, uint64_t mask)
{
	key = (~key + (key << 21)) & mask;
	key = key ^ key >> 24;
	key = ((key + (key << 3)) + (key << 8)) & mask;
	key = key ^ key >> 14;
	key = ((key + (key << 2)) + (key << 4)) & mask;
	key = key ^ key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}

__device__ __forceinline__ uint64_t table_slot(uint64_t minimizer, int log2cap)
{
	return (minimizer * 0x9E3779B97F4A7C15ull) >> (64 - log2cap);
}

// ---------------------------------------------------------------------------------------------------------------
// Sketch
// ---------------------------------------------------------------------------------------------------------------
// mm_sketch's state after processing position t (sketch.c:89-139; non-HPC, odd k so no k-mer equals its reverse complement):
//   buf      the infos of positions t-w+1 .. t        info = (hash << 8 | k, t << 1 | strand) or "invalid" (all ones)
//   min      the NEWEST entry of buf with the smallest x — `<=` when a new info arrives (sketch.c:121), `>=` in the rescan
//            that runs oldest to newest (sketch.c:126-129); an invalid entry is just the largest possible x
//   l        valid bases since the last ambiguous one (or since the start of the read)
// so with X[j] the x of position j (invalid before the read and wherever l < k) and m = newest minimal of X[t-w .. t-1], step t
// pushes, in this order:
//   S  l == w+k-1 and m valid:  every j in (t-w, t) with X[j] == X[m], j != m                       (sketch.c:115-120, first window)
//   A  X[t] <= X[m]:            m, if l >= w+k and m valid                                          (sketch.c:121-123)
//   B  else if m == t-w:        m, if l >= w+k-1 and m valid; then with n = newest minimal of X[t-w+1 .. t], if l >= w+k-1 and
//                               n valid: every j in (t-w, t] with X[j] == X[n], j != n              (sketch.c:124-137)
//   E  at the last position:    the minimum after the step, if valid                                (sketch.c:141-142)
// A position can be pushed more than once (S and B at the same step): the output is a sequence, not a set, and is reproduced as such.
struct Emit {
	unsigned long long mask_s, mask_b;      // bit d <-> position t - w + 1 + d
	int first_m, last_e;                    // position pushed by A / B's first push (or -1); by E (or -1)
};

// hash64 on 32-bit words, for 2k <= 32: every step of sketch.c:28-38 ends in `& mask`, so carrying only the low 32 bits is exact
__device__ __forceinline__ uint32_t hash32(uint32_t key, uint32_t mask)
{
	key = (~key + (key << 21)) & mask;
	key = key ^ key >> 24;
	key = ((key + (key << 3)) + (key << 8)) & mask;
	key = key ^ key >> 14;
	key = ((key + (key << 2)) + (key << 4)) & mask;
	key = key ^ key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}

// One pass: a tile counts what it pushes, learns how many minimizers the tiles before it pushed from a decoupled look-back over
// their published counts (tiles are handed out by an atomic ticket so that every predecessor is already running), and writes.
// tile_state[t] = flag << 62 | count: flag 1 = the tile's own count, 2 = the inclusive prefix up to and including it.
// If mv is too small for the batch the writes are dropped but the counting goes on: the host sees the total and runs again.
// K32: 2k < 32, the hash fits 32 bits with room for "invalid" (all ones): the window scans compare 32-bit words.  W: the window size
// when it is one of the presets' (the scan is unrolled), 0 for any other w.
template <bool K32, int W>
__global__ void __launch_bounds__(SKETCH_TILE) sketch_kernel(const SeedArgs s)
{
	typedef typename std::conditional<K32, uint32_t, uint64_t>::type key_t;
	constexpr key_t NOKEY = (key_t)~(key_t)0;
	__shared__ uint64_t X[TILE_SPAN];
	__shared__ uint32_t H32[K32 ? TILE_SPAN : 1];           // K32: the hashes alone, all ones where X is invalid
	__shared__ uint64_t pk[TILE_SPAN / 32 + 1];             // 2-bit bases, 32 per word, earlier base in the lower bits
	__shared__ uint32_t badw[TILE_SPAN / 32 + 1];           // one bit per position: ambiguous base or outside the read
	__shared__ uint8_t zs[TILE_SPAN];                       // strand of the position's k-mer
	__shared__ int16_t jm_s[SKETCH_TILE];                   // per position: newest minimal entry of the window in front of it ...
	__shared__ uint8_t ties_s[SKETCH_TILE];                 // ... and whether its hash occurs more than once in that window
	__shared__ int warp_sum[SKETCH_TILE / 32];
	__shared__ int tile_s;
	__shared__ long long excl_s;

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) {
		pk[TILE_SPAN / 32] = 0, badw[TILE_SPAN / 32] = 0;
		tile_s = atomicAdd(s.tile_ticket, 1);
	}
	__syncthreads();
	const int tile = tile_s;
	const int r = s.tile_read[tile];                        // which read this tile belongs to (filled by the host with the offsets)
	const int64_t so = s.seq_off[r];
	const int L = (int)(s.seq_off[r + 1] - so);
	const int t0 = (tile - s.tile_off[r]) * SKETCH_TILE;
	const int pb = t0 - HALO;                               // position of shared-memory index 0
	const int k = s.k, w = W ? W : s.w;
	const uint64_t mask = (1ull << 2 * k) - 1;
	const key_t *K = K32 ? (const key_t*)H32 : (const key_t*)X;         // what the window scans compare

	// 1. bases -> 2-bit words and the bitmap of ambiguous positions (warps 0..18 of positions; the CTA has 16 warps)
	for (int base = warp * 32; base < TILE_SPAN; base += SKETCH_TILE) {
		const int pos = pb + base + lane;
		int c = 4;
		if (pos >= 0 && pos < L) c = nt4(s.seq[so + pos]);
		const unsigned bad = __ballot_sync(FULL, c == 4);
		const unsigned two = (unsigned)(c & 3) << 2 * (lane & 15);
		const unsigned lo = __reduce_or_sync(FULL, lane < 16 ? two : 0u), hi = __reduce_or_sync(FULL, lane >= 16 ? two : 0u);
		if (lane == 0) badw[base >> 5] = bad, pk[base >> 5] = (uint64_t)hi << 32 | lo;
	}
	__syncthreads();

	// valid bases up to and including shared index i (0 when i itself is ambiguous); 128 when the run reaches past what is loaded —
	// every threshold below is <= w + k <= 92, and a thread only asks about indices that have that much in front of them
	auto run_len = [&](int i) -> int {
		int wi = i >> 5;
		const int b = i & 31;
		unsigned m = badw[wi] & (0xffffffffu >> (31 - b));
		if (m) return b - (31 - __clz(m));
		int l = b + 1;
		for (int back = 0; back < 3; ++back) {
			if (--wi < 0) return 128;
			m = badw[wi];
			if (m) return l + __clz(m);
			l += 32;
		}
		return 128;
	};

	// 2. X for the positions t0 - w .. t0 + TILE - 1
	for (int i = HALO - w + tid; i < TILE_SPAN; i += SKETCH_TILE) {
		uint64_t x = EMPTY;
		uint32_t h32 = ~0u;
		int z = 0;
		const int p0 = i - k + 1;
		const uint64_t bad_bits = ((uint64_t)badw[(p0 >> 5) + 1] << 32 | badw[p0 >> 5]) >> (p0 & 31);     // positions p0 .. p0 + 31
		if ((bad_bits & ((1ull << k) - 1)) == 0) {                          // k valid bases end here: l >= k (sketch.c:113)
			const int wi = p0 >> 5, sh = (p0 & 31) * 2;
			uint64_t f = pk[wi] >> sh;
			if (sh) f |= pk[wi + 1] << (64 - sh);
			f &= mask;                                              // oldest base in the lowest bits ...
			const uint64_t rv = f ^ mask;                           // ... which is how the reverse k-mer is kept (sketch.c:109): complement only
			uint64_t fw = __brevll(f);                              // forward k-mer: oldest base on top (sketch.c:108) = the 2-bit groups reversed
			fw = ((fw >> 1) & 0x5555555555555555ull) | ((fw & 0x5555555555555555ull) << 1);
			fw >>= 64 - 2 * k;
			z = fw < rv ? 0 : 1;                                    // sketch.c:111 (fw != rv for odd k)
			const uint64_t km = z ? rv : fw;
			const uint64_t h = K32 ? (uint64_t)hash32((uint32_t)km, (uint32_t)mask) : hash64(km, mask);
			x = h << 8 | (uint64_t)k;                               // sketch.c:114; kmer_span == k once l >= k
			h32 = (uint32_t)h;
		}
		X[i] = x, zs[i] = (uint8_t)z;
		if (K32) H32[i] = h32;
	}
	__syncthreads();

	// 3. the window in front of this thread's position: its newest minimal entry, and whether that hash is there more than once.
	//    The window AFTER the step (what the reference rescans when the minimum leaves, sketch.c:126-129) is the next thread's.
	const int i = HALO + tid, t = t0 + tid;
	key_t xm = NOKEY;
	int jm = i - w;
	bool ties = false;
#pragma unroll
	for (int j = i - w; j < i; ++j) {
		const key_t xj = K[j];
		if (xj <= xm) ties = xj == xm, xm = xj, jm = j;
	}
	jm_s[tid] = (int16_t)jm, ties_s[tid] = ties;
	__syncthreads();

	// 4. what mm_sketch pushes at this position
	Emit e;
	e.mask_s = e.mask_b = 0, e.first_m = e.last_e = -1;
	int cnt = 0;
	if (t < L) {
		const int l = run_len(i);
		const key_t xt = K[i];
		int after = jm;                                                     // the minimum after this step (shared index)
		if (l == w + k - 1 && xm != NOKEY && ties)
			for (int j = i - w + 1; j < i; ++j) if (K[j] == xm && j != jm) e.mask_s |= 1ull << (j - (i - w + 1));
		if (xt <= xm) {
			if (l >= w + k && xm != NOKEY) e.first_m = jm;
			after = i;
		} else if (jm == i - w) {
			if (l >= w + k - 1 && xm != NOKEY) e.first_m = jm;
			int jn;
			bool tn;
			if (tid + 1 < SKETCH_TILE) jn = jm_s[tid + 1], tn = ties_s[tid + 1];
			else {                                                          // the last position of the tile has no neighbour to ask
				key_t xq = NOKEY;
				jn = i - w + 1, tn = false;
				for (int j = i - w + 1; j <= i; ++j) {
					const key_t xj = K[j];
					if (xj <= xq) tn = xj == xq, xq = xj, jn = j;
				}
			}
			const key_t xn = K[jn];
			if (l >= w + k - 1 && xn != NOKEY && tn)
				for (int j = i - w + 1; j <= i; ++j) if (K[j] == xn && j != jn) e.mask_b |= 1ull << (j - (i - w + 1));
			after = jn;
		}
		if (t == L - 1 && K[after] != NOKEY) e.last_e = after;
		cnt = __popcll(e.mask_s) + (e.first_m >= 0) + __popcll(e.mask_b) + (e.last_e >= 0);
	}
	// block-wide exclusive scan of the counts
	int incl = cnt;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const int o = __shfl_up_sync(FULL, incl, d);
		if (lane >= d) incl += o;
	}
	if (lane == 31) warp_sum[warp] = incl;
	__syncthreads();
	int before, total;
	{	// every warp scans the 16 warp totals for itself (one shared-memory read and four shuffles, no second barrier)
		int v = warp_sum[lane & (SKETCH_TILE / 32 - 1)], sc = v;
#pragma unroll
		for (int d = 1; d < SKETCH_TILE / 32; d <<= 1) {
			const int o = __shfl_up_sync(FULL, sc, d);
			if ((lane & (SKETCH_TILE / 32 - 1)) >= d) sc += o;
		}
		before = __shfl_sync(FULL, sc - v, warp);
		total = __shfl_sync(FULL, sc, SKETCH_TILE / 32 - 1);
	}
	// decoupled look-back: how many minimizers the tiles before this one pushed.  Warp 0 looks at 32 predecessors at a time.
	if (warp == 0) {
		volatile unsigned long long *state = s.tile_state;
		const unsigned long long F1 = 1ull << 62, F2 = 2ull << 62, VAL = F1 - 1;
		long long excl = 0;
		if (tile > 0) {
			if (lane == 0) state[tile] = F1 | (unsigned long long)total;
			__syncwarp();
			for (int top = tile - 1; top >= 0; top -= 32) {
				const int p = top - lane;
				unsigned long long v = F2;                              // lanes in front of tile 0 count as a finished prefix of 0
				if (p >= 0) while (((v = state[p]) >> 62) == 0) { }
				const unsigned done = __ballot_sync(FULL, v >> 62 == 2);
				const int stop = done ? __ffs((int)done) - 1 : 32;          // nearest predecessor with an inclusive prefix
				long long part = lane <= stop ? (long long)(v & VAL) : 0;
#pragma unroll
				for (int d = 16; d; d >>= 1) part += __shfl_xor_sync(FULL, part, d);
				excl += part;
				if (done) break;
			}
		}
		if (lane == 0) {
			state[tile] = F2 | (unsigned long long)(excl + total);
			s.tile_excl[tile] = excl;
			if (tile == s.n_tiles - 1) s.tile_excl[s.n_tiles] = excl + total;
			excl_s = excl;
		}
	}
	__syncthreads();
	if (cnt == 0) return;
	const long long at = excl_s + before + (incl - cnt);
	if (at + cnt > s.mv_cap) return;                        // (the host sees the total and comes back with a larger buffer)
	ulonglong2 *dst = s.mv + at;
	auto push = [&](int j) { *dst++ = make_ulonglong2(X[j], (uint64_t)(uint32_t)(pb + j) << 1 | zs[j]); };     // sketch.c:115 (rid 0)
	const int j0 = i - w + 1;
	for (unsigned long long m = e.mask_s; m; m &= m - 1) push(j0 + __ffsll((long long)m) - 1);
	if (e.first_m >= 0) push(e.first_m);
	for (unsigned long long m = e.mask_b; m; m &= m - 1) push(j0 + __ffsll((long long)m) - 1);
	if (e.last_e >= 0) push(e.last_e);
}

// ---------------------------------------------------------------------------------------------------------------
// Sketch, eight positions per thread (w >= 8).  The same decisions as sketch_kernel, with the per-position overhead shared: a
// thread converts its 8 bases with one shared-memory word and a table, rolls the k-mer and its reverse complement from one
// position to the next (two shifts each, sketch.c:108-109 as written), keeps its 8 hashes in registers, and gets the nine windows
// around them from w + 16 combinations instead of 9 w comparisons: suffix minima over the w hashes in front of its first position
// (read from shared memory, newest kept on ties) combined with prefix minima over its own.  A CTA of 256 threads covers 1,920
// positions after a halo of 128.
// ---------------------------------------------------------------------------------------------------------------
constexpr int S8_P = 8;
constexpr int S8_HALO_THREADS = 16;
constexpr int S8_THREADS = SKETCH8_TILE / S8_P + S8_HALO_THREADS;      // 256
constexpr int S8_SPAN = S8_THREADS * S8_P;                             // 2,048 positions in shared memory
static_assert(S8_THREADS % 32 == 0, "the block scans take each warp's total from lane 31");
constexpr int S8_HALO = S8_HALO_THREADS * S8_P;                        // 128 >= w + k

#define KI(i_) ((i_) + ((i_) >> 3))                            // index into the padded hash array: a warp's stride-8 accesses hit 32 banks

template <class key_t> struct WinMin { key_t x; int j; bool ties; };

template <bool K32, int CTAS>
__global__ void __launch_bounds__(S8_THREADS, CTAS) sketch8_kernel(const SeedArgs s)
{
	typedef typename std::conditional<K32, uint32_t, uint64_t>::type key_t;
	constexpr key_t NOKEY = (key_t)~(key_t)0;
	__shared__ __align__(16) uint8_t raw[S8_SPAN + 16];     // the bases as they come, from the 16-byte boundary at or below the span's first
	__shared__ uint8_t nt4_tab[256];
	__shared__ __align__(8) uint16_t pk16[S8_THREADS + 4];  // 2-bit bases, 8 per entry, earlier base in the lower bits (4 entries of padding in front)
	__shared__ key_t K[S8_SPAN + S8_SPAN / 8];              // the hashes (all ones where invalid); one slot of padding per 8: KI()
	__shared__ uint8_t zs8[S8_THREADS];                     // strand bits of a thread's 8 positions
	__shared__ int warp_val[(S8_THREADS + 31) / 32];
	__shared__ int tile_s;
	__shared__ long long excl_s;

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	constexpr int N_WARPS = (S8_THREADS + 31) / 32;
	if (tid == 0) tile_s = atomicAdd(s.tile_ticket, 1);
	if (tid < 256) nt4_tab[tid] = (uint8_t)nt4((unsigned)tid);
	if (tid < 4) pk16[tid] = 0;
	__syncthreads();
	const int tile = tile_s;
	const int r = s.tile_read[tile];
	const int64_t so = s.seq_off[r];
	const int L = (int)(s.seq_off[r + 1] - so);
	const int t0 = (tile - s.tile_off[r]) * SKETCH8_TILE;
	const int pb = t0 - S8_HALO;                            // position of shared-memory index 0
	const int k = s.k, w = s.w;
	const uint64_t mask = (1ull << 2 * k) - 1;

	// 1. the span's bytes, 16 at a time from aligned addresses (bytes of the neighbouring reads come along and are masked below)
	const int64_t g0 = so + pb, g0a = g0 & ~15ll;           // offset of the span in the sub-batch's sequences; may be negative
	if (tid <= S8_SPAN / 16) {
		const int64_t c = g0a + 16 * tid;
		uint4 v = make_uint4(0x4e4e4e4eu, 0x4e4e4e4eu, 0x4e4e4e4eu, 0x4e4e4e4eu);
		if (c >= 0 && c < s.seq_len) v = __ldg((const uint4*)(s.seq + c));     // (the buffer is padded to 16 bytes past seq_len)
		*(uint4*)&raw[16 * tid] = v;
	}
	__syncthreads();
	// 2. this thread's 8 bases: codes, packed bases, ambiguity bits
	const int i0 = tid * S8_P, p0 = pb + i0;                // shared index / position of this thread's first base
	uint64_t eight;
	{
		const int o = (int)(g0 - g0a) + i0, sh = (o & 7) * 8;
		const uint64_t lo = *(const uint64_t*)&raw[o & ~7], hi = *(const uint64_t*)&raw[(o & ~7) + 8];
		eight = sh ? lo >> sh | hi << (64 - sh) : lo;
		if (p0 < 0 || p0 + S8_P > L) {                      // outside the read: 'N'
#pragma unroll
			for (int j = 0; j < S8_P; ++j)
				if (p0 + j < 0 || p0 + j >= L) eight = (eight & ~(0xffull << 8 * j)) | 0x4eull << 8 * j;
		}
	}
	unsigned two16 = 0, bad8 = 0;
#pragma unroll
	for (int j = 0; j < S8_P; ++j) {
		const unsigned c = nt4_tab[(eight >> 8 * j) & 0xff];
		two16 |= (c & 3) << 2 * j;
		bad8 |= (c >> 2) << j;
	}
	pk16[4 + tid] = (uint16_t)two16;
	// valid bases in front of this thread's first one: distance to the last ambiguous base before it (block-wide max scan), capped
	int last_bad = bad8 ? i0 + (31 - __clz(bad8)) : -1;     // shared index of this thread's last ambiguous base
	{
		int v = last_bad;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const int o = __shfl_up_sync(FULL, v, d);
			if (lane >= d && o > v) v = o;
		}
		if (lane == 31) warp_val[warp] = v;
		__syncthreads();                                    // (also: pk16 is complete)
		int before = -1;
		for (int q = 0; q < warp; ++q) before = warp_val[q] > before ? warp_val[q] : before;
		int excl = __shfl_up_sync(FULL, v, 1);
		if (lane == 0) excl = -1;
		last_bad = excl > before ? excl : before;           // last ambiguous base strictly before i0 (-1: none in the span)
	}
	int run = last_bad < 0 ? 128 : i0 - 1 - last_bad;       // 128: the run reaches past what is loaded (every threshold is <= w + k <= 92)
	run = run > 128 ? 128 : run;
	// 3. the 8 hashes, rolling the two k-mers from the k - 1 bases in front
	uint64_t fw, rv;
	{
		const uint64_t *pk64 = (const uint64_t*)pk16;       // entries tid .. tid + 3 of the padded array = the 32 bases in front of i0
		const int sh = (tid & 3) * 16;
		uint64_t prev = pk64[tid >> 2] >> sh;
		if (sh) prev |= pk64[(tid >> 2) + 1] << (64 - sh);
		const uint64_t f = k > 1 ? prev >> (64 - 2 * (k - 1)) : 0;     // the k - 1 bases before i0, oldest lowest
		rv = (f ^ (mask >> 2)) << 2;                        // sketch.c:109 after those bases: complemented, one slot left for the next
		uint64_t x = __brevll(f);
		x = ((x >> 1) & 0x5555555555555555ull) | ((x & 0x5555555555555555ull) << 1);
		fw = k > 1 ? x >> (64 - 2 * (k - 1)) : 0;           // sketch.c:108 after those bases: oldest on top
	}
	key_t xk[S8_P];
	int runs[S8_P];
	unsigned z8 = 0;
#pragma unroll
	for (int j = 0; j < S8_P; ++j) {
		const unsigned c = (two16 >> 2 * j) & 3;
		fw = (fw << 2 | c) & mask;                          // sketch.c:108
		rv = rv >> 2 | (uint64_t)(3 ^ c) << 2 * (k - 1);    // sketch.c:109
		run = (bad8 >> j & 1) ? 0 : (run < 128 ? run + 1 : 128);
		runs[j] = run;
		key_t x = NOKEY;
		if (run >= k) {
			const unsigned z = fw < rv ? 0 : 1;                // sketch.c:111
			const uint64_t km = z ? rv : fw;
			x = K32 ? (key_t)hash32((uint32_t)km, (uint32_t)mask) : (key_t)(hash64(km, mask) << 8 | (uint64_t)k);
			z8 |= z << j;
		}
		xk[j] = x;
		K[KI(i0 + j)] = x;
	}
	zs8[tid] = (uint8_t)z8;
	__syncthreads();

	// 4. windows.  Position j of this thread (j = 0..8, 8 = the next thread's first) looks at o_j .. o_{w-1}, v_0 .. v_{j-1}: the w hashes in
	//    front of i0 (shared memory) and the thread's own.  Newest minimum and "is that hash there twice" for each.
	WinMin<key_t> win[S8_P + 1];
	if (tid >= S8_HALO_THREADS) {
		WinMin<key_t> suf;                                  // newest minimum of o_i .. o_{w-1}, built from the newest end
		suf.x = NOKEY, suf.j = i0 - 1, suf.ties = false;    // (nothing yet; an invalid newest entry is its own minimum)
		auto older = [&](int i) {                           // an older entry joins: it takes over only if strictly smaller
			const key_t xo = K[KI(i0 - w + i)];
			if (xo < suf.x) suf.x = xo, suf.j = i0 - w + i, suf.ties = false;
			else if (xo == suf.x) suf.ties = true;
		};
		for (int i = w - 1; i > S8_P; --i) older(i);
#pragma unroll
		for (int i = S8_P; i >= 0; --i) {
			if (i < w) older(i);
			win[i] = suf;
		}
		WinMin<key_t> pre;                                  // newest minimum of v_0 .. v_{j-1}
		pre.x = xk[0], pre.j = i0, pre.ties = false;
#pragma unroll
		for (int j = 1; j <= S8_P; ++j) {
			WinMin<key_t> m = pre;                            // window j = suffix j (older) ++ prefix j (newer)
			if (j < w) {
				const WinMin<key_t> o = win[j];
				if (o.x < pre.x) m = o;
				else if (o.x == pre.x) m.ties = true;
			}
			win[j] = m;
			if (j < S8_P) {
				if (xk[j] <= pre.x) pre.ties = xk[j] == pre.x, pre.x = xk[j], pre.j = i0 + j;
			}
		}
	}

	// 5. what mm_sketch pushes at each of the 8 positions (see sketch_kernel): counted first, written after the offsets are known
	unsigned f_first = 0, f_last = 0, f_ties = 0;           // per position: A/B pushes the old minimum; E; the rare equal-hash lists must be walked
	int cnt = 0;
	if (tid >= S8_HALO_THREADS) {
#pragma unroll
		for (int j = 0; j < S8_P; ++j) {
			const int i = i0 + j, t = p0 + j;
			if (t >= L) break;
			const int l = runs[j];
			const key_t xm = win[j].x, xt = xk[j];
			const int jm = win[j].j;
			int after = jm;
			if (l == w + k - 1 && xm != NOKEY && win[j].ties) {
				f_ties |= 1u << j;
				for (int q = i - w + 1; q < i; ++q) cnt += K[KI(q)] == xm && q != jm;
			}
			if (xt <= xm) {
				if (l >= w + k && xm != NOKEY) f_first |= 1u << j, ++cnt;
				after = i;
			} else if (jm == i - w) {
				if (l >= w + k - 1 && xm != NOKEY) f_first |= 1u << j, ++cnt;
				const key_t xn = win[j + 1].x;
				const int jn = win[j + 1].j;
				if (l >= w + k - 1 && xn != NOKEY && win[j + 1].ties) {
					f_ties |= 1u << (8 + j);
					for (int q = i - w + 1; q <= i; ++q) cnt += K[KI(q)] == xn && q != jn;
				}
				after = jn;
			}
			if (t == L - 1 && K[KI(after)] != NOKEY) f_last |= 1u << j, ++cnt;
		}
	}
	// block-wide exclusive scan of the counts
	int incl = cnt;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const int o = __shfl_up_sync(FULL, incl, d);
		if (lane >= d) incl += o;
	}
	__syncthreads();                                        // (warp_val is reused)
	if (lane == 31) warp_val[warp] = incl;
	__syncthreads();
	int before = 0, total = 0;
	for (int q = 0; q < N_WARPS; ++q) {
		const int v = warp_val[q];
		if (q < warp) before += v;
		total += v;
	}
	// decoupled look-back: how many minimizers the tiles before this one pushed.  Warp 0 looks at 32 predecessors at a time.
	if (warp == 0) {
		volatile unsigned long long *state = s.tile_state;
		const unsigned long long F1 = 1ull << 62, F2 = 2ull << 62, VAL = F1 - 1;
		long long excl = 0;
		if (tile > 0) {
			if (lane == 0) state[tile] = F1 | (unsigned long long)total;
			__syncwarp();
			for (int top = tile - 1; top >= 0; top -= 32) {
				const int p = top - lane;
				unsigned long long v = F2;
				if (p >= 0) while (((v = state[p]) >> 62) == 0) { }
				const unsigned done = __ballot_sync(FULL, v >> 62 == 2);
				const int stop = done ? __ffs((int)done) - 1 : 32;
				long long part = lane <= stop ? (long long)(v & VAL) : 0;
#pragma unroll
				for (int d = 16; d; d >>= 1) part += __shfl_xor_sync(FULL, part, d);
				excl += part;
				if (done) break;
			}
		}
		if (lane == 0) {
			state[tile] = F2 | (unsigned long long)(excl + total);
			s.tile_excl[tile] = excl;
			if (tile == s.n_tiles - 1) s.tile_excl[s.n_tiles] = excl + total;
			excl_s = excl;
		}
	}
	__syncthreads();
	if (cnt == 0) return;
	const long long at = excl_s + before + (incl - cnt);
	if (at + cnt > s.mv_cap) return;                        // (the host sees the total and comes back with a larger buffer)
	ulonglong2 *dst = s.mv + at;
	auto push = [&](int q) {                                // sketch.c:115 (rid 0): x = hash << 8 | span, y = position << 1 | strand
		const uint64_t x = K32 ? (uint64_t)K[KI(q)] << 8 | (uint64_t)k : (uint64_t)K[KI(q)];
		*dst++ = make_ulonglong2(x, (uint64_t)(uint32_t)(pb + q) << 1 | (zs8[q >> 3] >> (q & 7) & 1));
	};
#pragma unroll
	for (int j = 0; j < S8_P; ++j) {
		const int i = i0 + j, t = p0 + j;
		if (t >= L) break;
		const key_t xm = win[j].x;
		const int jm = win[j].j, jn = win[j + 1].j;
		if (f_ties >> j & 1)
			for (int q = i - w + 1; q < i; ++q) if (K[KI(q)] == xm && q != jm) push(q);
		if (f_first >> j & 1) push(jm);
		if (f_ties >> (8 + j) & 1) {
			const key_t xn = win[j + 1].x;
			for (int q = i - w + 1; q <= i; ++q) if (K[KI(q)] == xn && q != jn) push(q);
		}
		if (f_last >> j & 1) {
			const key_t xt = xk[j];
			push(xt <= xm ? i : (jm == i - w ? jn : jm));
		}
	}
}

// ---------------------------------------------------------------------------------------------------------------
// Small utilities: exclusive prefix sum over the reads of a sub-batch (one CTA), per-read minimizer offsets
// ---------------------------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(1024) scan_kernel(const T *in, int64_t *out, int64_t n)
{
	__shared__ int64_t warp_sum[32];
	__shared__ int64_t carry_s;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) carry_s = 0;
	__syncthreads();
	for (int64_t base = 0; base < n; base += 1024) {
		const int64_t i = base + tid;
		const int64_t v = i < n ? (int64_t)in[i] : 0;
		int64_t incl = v;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const int64_t o = __shfl_up_sync(FULL, incl, d);
			if (lane >= d) incl += o;
		}
		if (lane == 31) warp_sum[warp] = incl;
		__syncthreads();
		int64_t before = carry_s;
		for (int q = 0; q < warp; ++q) before += warp_sum[q];
		if (i < n) out[i] = before + incl - v;
		__syncthreads();
		if (tid == 1023) carry_s = before + incl;
		__syncthreads();
	}
	if (tid == 0) out[n] = carry_s;
}

// Zeroing by a kernel rather than cudaMemsetAsync: small memsets can be routed to a copy engine, where they wait behind other
// sub-batches' output copies.
__global__ void zero_words_kernel(unsigned long long *p, int64_t n, unsigned long long *q, int64_t nq)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) p[i] = 0;
	if (i < nq) q[i] = 0;
}

// A few device scalars stored straight into mapped host memory.  The host needs them to size the next stage; as 8-byte D2H copies
// they queued behind other sub-batches' output copies on the copy engine and every pipeline context stalled there for milliseconds.
__global__ void export_scalars_kernel(volatile int64_t *host, const int64_t *a, const int64_t *b, const int *c)
{
	if (a) host[0] = *a;
	if (b) host[1] = *b;
	if (c) host[2] = *c;
	__threadfence_system();
}

__global__ void read_offsets_kernel(const SeedArgs s)
{
	const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r <= s.n_reads) s.mv_off[r] = s.n_tiles > 0 ? s.tile_excl[r < s.n_reads ? s.tile_off[r] : s.n_tiles] : 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Index: build and probe (mm_idx_get, index.c:81-98)
// ---------------------------------------------------------------------------------------------------------------
__global__ void index_insert_kernel(int64_t n_keys, const uint64_t *keys, const uint64_t *vals, uint64_t *tab_keys, uint64_t *tab_vals, int log2cap)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_keys) return;
	const uint64_t key = keys[i], cap_mask = (1ull << log2cap) - 1;
	uint64_t h = table_slot(key >> 1, log2cap);
	for (;;) {
		const unsigned long long old = atomicCAS((unsigned long long*)&tab_keys[h], (unsigned long long)EMPTY, (unsigned long long)key);
		if (old == EMPTY) { tab_vals[h] = vals[i]; return; }
		h = (h + 1) & cap_mask;                                 // (keys are distinct: an occupied slot is somebody else's)
	}
}

__device__ __forceinline__ int index_get(const DeviceIndex &ix, uint64_t minimizer, uint64_t &val)
{
	const uint64_t cap_mask = (1ull << ix.log2cap) - 1;
	uint64_t h = table_slot(minimizer, ix.log2cap);
	for (;;) {
		const uint64_t key = __ldg(&ix.tab_keys[h]);
		if (key == EMPTY) { val = 0; return 0; }
		if (key >> 1 == minimizer) {
			val = __ldg(&ix.tab_vals[h]);
			return (key & 1) ? 1 : (int)(uint32_t)val;              // index.c:90-96
		}
		h = (h + 1) & cap_mask;
	}
}

__global__ void __launch_bounds__(256) index_lookup_kernel(const DeviceIndex ix, int64_t n, const ulonglong2 *mv, const uint64_t *raw, int32_t *occ, uint64_t *hv)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const uint64_t minimizer = mv ? __ldg(&mv[i].x) >> 8 : raw[i];      // map.c:103
	uint64_t v;
	occ[i] = index_get(ix, minimizer, v);
	hv[i] = v;
}

// ---------------------------------------------------------------------------------------------------------------
// collect_matches (map.c:90-123), one warp per read
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) matches_kernel(const SeedArgs s)
{
	const int lane = threadIdx.x & 31;
	const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
	for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < s.n_reads; r += n_warps) {
		const int64_t m0 = s.mv_off[r];
		const int n = (int)(s.mv_off[r + 1] - m0);
		const ulonglong2 *mv = s.mv + m0;
		int rep_len = 0, prev_en = 0, n_mp = 0;
		long long n_a = 0;
		ulonglong2 q_next = make_ulonglong2(0, 0);          // the next 32 minimizers are on their way while these are worked on
		int t_next = 0;
		if (lane < n) q_next = __ldg(mv + lane), t_next = __ldg(s.occ + m0 + lane);
		for (int base = 0; base < n; base += 32) {
			const int i = base + lane;
			const bool in = i < n;
			const uint64_t x = q_next.x, y = q_next.y;
			const int t = t_next;
			if (i + 32 < n) q_next = __ldg(mv + i + 32), t_next = __ldg(s.occ + m0 + i + 32);
			else q_next = make_ulonglong2(0, 0), t_next = 0;
			__syncwarp();
			const int q_span = (int)(x & 0xff), q_pos = (int)(uint32_t)y;
			const bool high = in && t >= s.max_occ, match = in && t < s.max_occ;
			// repeat length (map.c:104-110, :120): the union of the query intervals [en - q_span, en) of the skipped minimizers, taken in
			// order of en; each adds en - max(st, en of the previous one)
			const int en = (q_pos >> 1) + 1, st = en - q_span;
			const unsigned hm = __ballot_sync(FULL, high);
			const unsigned below = hm & lanemask_lt(lane);
			const int src = below ? 31 - __clz(below) : 0;
			int pe = __shfl_sync(FULL, en, src);
			if (!below) pe = prev_en;
			int add = high ? en - (st > pe ? st : pe) : 0;
#pragma unroll
			for (int d = 16; d; d >>= 1) add += __shfl_xor_sync(FULL, add, d);
			rep_len += add;
			if (hm) prev_en = __shfl_sync(FULL, en, 31 - __clz(hm));
			// matches: mini_pos (map.c:117) and the first anchor of each (exclusive prefix of the occurrence counts, map.c:116)
			const unsigned mm = __ballot_sync(FULL, match);
			if (match) s.mini_pos[m0 + n_mp + __popc(mm & lanemask_lt(lane))] = (uint32_t)(q_pos >> 1);
			int incl = match ? t : 0;
			const int mine = incl;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const int o = __shfl_up_sync(FULL, incl, d);
				if (lane >= d) incl += o;
			}
			if (in) s.arel[m0 + i] = match ? (int32_t)(n_a + incl - mine) : -1;
			__syncwarp();
			n_a += __shfl_sync(FULL, incl, 31);
			n_mp += __popc(mm);
		}
		if (lane == 0) s.rep_len[r] = rep_len, s.n_mini_pos[r] = n_mp, s.n_a[r] = n_a;
		__syncwarp();
	}
}

// ---------------------------------------------------------------------------------------------------------------
// collect_seed_hits (map.c:215-247): anchors of one read, written by its warp in the reference's order (match by match)
// ---------------------------------------------------------------------------------------------------------------
__device__ void expand_read(const SeedArgs &s, const DeviceIndex &ix, int64_t r, int lane, ulonglong2 *dst)
{
	const int64_t m0 = s.mv_off[r];
	const int n = (int)(s.mv_off[r + 1] - m0);
	const int qlen = (int)(s.seq_off[r + 1] - s.seq_off[r]);
	ulonglong2 *a = dst + s.a_off[r];
	for (int base = 0; base < n; base += 32) {
		const int i = base + lane;
		int t = 0, rel = -1;
		uint64_t x = 0, y = 0, v = 0;
		bool tandem = false;
		if (i < n) {
			rel = s.arel[m0 + i];
			if (rel >= 0) {
				const ulonglong2 q = __ldg(s.mv + m0 + i);
				x = q.x, y = q.y, t = s.occ[m0 + i], v = s.hv[m0 + i];
				// map.c:113-115: the same minimizer right before or after this one in the read's sketch (whatever its occurrence count)
				tandem = (i > 0 && __ldg(&s.mv[m0 + i - 1].x) >> 8 == x >> 8) || (i < n - 1 && __ldg(&s.mv[m0 + i + 1].x) >> 8 == x >> 8);
			}
		}
		__syncwarp();
		const uint32_t q_pos = (uint32_t)y;
		const uint64_t q_span = x & 0xff;
		for (int h = 0; h < t; ++h) {
			const uint64_t rr = t == 1 ? v : __ldg(&ix.pos[(v >> 32) + h]);                 // index.c:90-96
			const uint32_t rpos = (uint32_t)rr >> 1;
			uint64_t ax, ay;
			if ((rr & 1) == (q_pos & 1)) {                                                   // forward strand (map.c:232-234)
				ax = (rr & 0xffffffff00000000ull) | rpos;
				ay = q_span << 32 | (q_pos >> 1);
			} else {                                                                         // reverse strand (map.c:235-238)
				ax = 1ull << 63 | (rr & 0xffffffff00000000ull) | rpos;
				ay = q_span << 32 | (uint32_t)(qlen - (int)((q_pos >> 1) + 1 - (uint32_t)q_span) - 1);
			}
			ay |= (y >> 32) << 48;                                                           // seg_id (map.c:239): 0 on this path
			if (tandem) ay |= SEED_TANDEM;
			a[rel + h] = make_ulonglong2(ax, ay);
		}
		__syncwarp();
	}
}

// The whole sub-batch at once, one thread per minimizer: where a minimizer's anchors go is already known (arel, a_off), so nothing
// orders the threads; the read a minimizer belongs to is found by bisection over mv_off (a table that stays in L1/L2).
__global__ void __launch_bounds__(256) expand_kernel(const SeedArgs s, const DeviceIndex ix, int64_t n_mv)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_mv) return;
	const int rel = s.arel[i];
	if (rel < 0) return;
	int lo = 0, hi = (int)s.n_reads;                        // mv_off[lo] <= i < mv_off[hi]
	while (hi - lo > 1) {
		const int mid = (lo + hi) >> 1;
		if (__ldg(s.mv_off + mid) <= i) lo = mid; else hi = mid;
	}
	const int64_t r = lo, m0 = __ldg(s.mv_off + r), m1 = __ldg(s.mv_off + r + 1);
	const int qlen = (int)(s.seq_off[r + 1] - s.seq_off[r]);
	ulonglong2 *a = s.a_tmp + s.a_off[r] + rel;
	const ulonglong2 q = __ldg(s.mv + i);
	const uint64_t x = q.x, y = q.y, v = s.hv[i];
	const int t = s.occ[i];
	const bool tandem = (i > m0 && __ldg(&s.mv[i - 1].x) >> 8 == x >> 8) || (i < m1 - 1 && __ldg(&s.mv[i + 1].x) >> 8 == x >> 8);     // map.c:113-115
	const uint32_t q_pos = (uint32_t)y;
	const uint64_t q_span = x & 0xff;
	const uint64_t ay_f = q_span << 32 | (q_pos >> 1) | (y >> 32) << 48 | (tandem ? SEED_TANDEM : 0);                                // map.c:232-234, :239
	const uint64_t ay_r = q_span << 32 | (uint32_t)(qlen - (int)((q_pos >> 1) + 1 - (uint32_t)q_span) - 1) | (y >> 32) << 48 | (tandem ? SEED_TANDEM : 0);     // map.c:235-238
	const uint64_t *pos = ix.pos + (v >> 32);
	for (int h = 0; h < t; ++h) {
		const uint64_t rr = t == 1 ? v : __ldg(pos + h);   // index.c:90-96
		const uint32_t rpos = (uint32_t)rr >> 1;
		const bool fwd = (rr & 1) == (q_pos & 1);
		a[h] = make_ulonglong2((fwd ? 0 : 1ull << 63) | (rr & 0xffffffff00000000ull) | rpos, fwd ? ay_f : ay_r);
	}
}

// ---------------------------------------------------------------------------------------------------------------
// radix_sort_128x (map.c:245).  Stable LSD radix sort by one warp over the bytes of x that differ inside the read; the result is
// the reference's unless the read holds equal keys, and those reads are listed for the exact replay below.
// ---------------------------------------------------------------------------------------------------------------
__device__ void warp_sort_by_x(ulonglong2 *keys, ulonglong2 *tmp, int n, int *hist, int lane)
{
	uint64_t diff = 0;
	const uint64_t k0 = keys[0].x;
	for (int k = lane; k < n; k += 32) diff |= keys[k].x ^ k0;
	__syncwarp();
#pragma unroll
	for (int d = 16; d; d >>= 1) diff |= __shfl_xor_sync(FULL, diff, d);
	ulonglong2 *src = keys, *dst = tmp;
	for (int shift = 0; shift < 64; shift += 8) {
		if (((diff >> shift) & 0xff) == 0) continue;
		for (int b = lane; b < 256; b += 32) hist[b] = 0;
		__syncwarp();
		for (int base = 0; base < n; base += 32) {
			const int k = base + lane;
			const int dig = k < n ? (int)(src[k].x >> shift & 0xff) : 256;
			const unsigned peers = __match_any_sync(FULL, dig);
			if (k < n && (peers & lanemask_lt(lane)) == 0) hist[dig] += __popc(peers);
			__syncwarp();
		}
		{
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
		for (int base = 0; base < n; base += 32) {
			const int k = base + lane;
			const bool act = k < n;
			ulonglong2 rec = make_ulonglong2(0, 0);
			if (act) rec = src[k];
			const int dig = act ? (int)(rec.x >> shift & 0xff) : 256;
			const unsigned peers = __match_any_sync(FULL, dig);
			const int rank = __popc(peers & lanemask_lt(lane));
			int pos = 0;
			if (act) pos = hist[dig] + rank;
			__syncwarp();
			if (act) {
				dst[pos] = rec;
				if (rank == 0) hist[dig] += __popc(peers);
			}
			__syncwarp();
		}
		ulonglong2 *t = src; src = dst; dst = t;
	}
	if (src != keys) for (int k = lane; k < n; k += 32) keys[k] = src[k];
	__syncwarp();
}

// The expansion leaves a read's anchors in a_tmp; the sorted read is written to a.  Reads of up to SORT_SMEM anchors — all but the
// longest — are sorted inside shared memory: the keys are loaded once, a 16-bit index per anchor is what the radix passes move, and
// the 16-byte records are gathered once at the end.  Longer reads take the global-memory sort above.
constexpr int SORT_SMEM = 1024;
constexpr int SORT_WARPS = 2;

__global__ void __launch_bounds__(SORT_WARPS * 32) sort_kernel(const SeedArgs s)
{
	__shared__ uint64_t xs_all[SORT_WARPS][SORT_SMEM];
	__shared__ uint16_t idx_all[SORT_WARPS][2][SORT_SMEM];
	__shared__ int hist_all[SORT_WARPS][256];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint64_t *xs = xs_all[warp];
	int *hist = hist_all[warp];
	for (;;) {
		int r = 0;
		if (lane == 0) r = atomicAdd(&s.tie_count[1], 1);
		r = __shfl_sync(FULL, r, 0);
		if (r >= s.n_reads) break;
		const int64_t o = s.a_off[r], n64 = s.a_off[r + 1] - o;
		if (n64 <= 0) continue;
		const int n = (int)n64;
		const ulonglong2 *src = s.a_tmp + o;
		ulonglong2 *a = s.a + o;
		if (n <= 32) {                                      // one key per lane: rank by (x, position)
			const ulonglong2 rec = lane < n ? src[lane] : make_ulonglong2(EMPTY, 0);
			int rank = 0;
			bool tie = false;
			for (int q = 0; q < n; ++q) {
				const uint64_t xq = __shfl_sync(FULL, rec.x, q);
				rank += xq < rec.x || (xq == rec.x && q < lane);
				tie |= q != lane && xq == rec.x;
			}
			if (lane < n) a[rank] = rec;
			const bool any_tie = __any_sync(FULL, lane < n && tie);
			if (any_tie && lane == 0) s.tie_list[atomicAdd(&s.tie_count[0], 1)] = r;
			__syncwarp();
			continue;
		}
		bool tie = false;
		if (n <= SORT_SMEM) {
			uint16_t *ia = idx_all[warp][0], *ib = idx_all[warp][1];
			uint64_t diff = 0;
			const uint64_t k0 = src[0].x;
			for (int k = lane; k < n; k += 32) {
				const uint64_t x = src[k].x;
				xs[k] = x, ia[k] = (uint16_t)k, diff |= x ^ k0;
			}
#pragma unroll
			for (int d = 16; d; d >>= 1) diff |= __shfl_xor_sync(FULL, diff, d);
			__syncwarp();
			for (int shift = 0; shift < 64; shift += 8) {
				if (((diff >> shift) & 0xff) == 0) continue;
				for (int b = lane; b < 256; b += 32) hist[b] = 0;
				__syncwarp();
				for (int base = 0; base < n; base += 32) {
					const int k = base + lane;
					const int dig = k < n ? (int)(xs[ia[k]] >> shift & 0xff) : 256;
					const unsigned peers = __match_any_sync(FULL, dig);
					if (k < n && (peers & lanemask_lt(lane)) == 0) hist[dig] += __popc(peers);
					__syncwarp();
				}
				{
					int loc[8], sum = 0;
#pragma unroll
					for (int q = 0; q < 8; ++q) loc[q] = hist[lane * 8 + q], sum += loc[q];
					int incl = sum;
#pragma unroll
					for (int d = 1; d < 32; d <<= 1) {
						const int v = __shfl_up_sync(FULL, incl, d);
						if (lane >= d) incl += v;
					}
					int run = incl - sum;
#pragma unroll
					for (int q = 0; q < 8; ++q) hist[lane * 8 + q] = run, run += loc[q];
				}
				__syncwarp();
				for (int base = 0; base < n; base += 32) {
					const int k = base + lane;
					const bool act = k < n;
					const uint16_t id = act ? ia[k] : 0;
					const int dig = act ? (int)(xs[id] >> shift & 0xff) : 256;
					const unsigned peers = __match_any_sync(FULL, dig);
					const int rank = __popc(peers & lanemask_lt(lane));
					int pos = 0;
					if (act) pos = hist[dig] + rank;
					__syncwarp();
					if (act) {
						ib[pos] = id;
						if (rank == 0) hist[dig] += __popc(peers);
					}
					__syncwarp();
				}
				uint16_t *t = ia; ia = ib; ib = t;
			}
			for (int k = lane; k < n; k += 32) {
				const int id = ia[k];
				a[k] = src[id];
				if (k + 1 < n) tie |= xs[id] == xs[ia[k + 1]];
			}
		} else {
			for (int k = lane; k < n; k += 32) a[k] = src[k];
			__syncwarp();
			warp_sort_by_x(a, s.a_tmp + o, n, hist, lane);
			for (int k = lane; k + 1 < n; k += 32) tie |= a[k].x == a[k + 1].x;
		}
		__syncwarp();
		if (__any_sync(FULL, tie) && lane == 0) s.tie_list[atomicAdd(&s.tie_count[0], 1)] = r;
		__syncwarp();
	}
}

// Reads with equal keys: expand again, this time straight into a[], and replay radix_sort_128x itself — insertion sort up to 64
// elements, else the in-place MSD byte radix permutation (ksort.h:116-151).  Rare for reads against a unique reference (a minimizer
// has to repeat inside the query); the rule for tandem repeats.  The permutation of one level is a chain of dependent swaps and
// stays on one lane, but everything around it is done by the warp: the digit counts (match-any histogram), the bucket offsets,
// the insertion sorts of the small buckets (every lane takes the buckets of its 8 digits: the ranges are disjoint), and levels on
// which every key has the same digit are skipped outright (nothing would move).  Reads of up to REPLAY_SMEM anchors are replayed
// in shared memory, where a dependent access costs tens of cycles instead of hundreds.
constexpr int REPLAY_SMEM = 8192;                // 128 KB of shared memory: one replaying warp per SM is plenty, such reads are few
constexpr int REPLAY_WORK = 160;                // pending ranges of a read held in shared memory (a range has > 64 elements)

__device__ void warp_replay_sort(W16 *w, int n, int *sm /* 768 ints */, int3 *work, int work_cap, int lane)
{
	int *head = sm, *tail = sm + 256, *cnt = sm + 512;
	if (n <= 64) {
		if (lane == 0) insertion_by_x(w, n);
		__syncwarp();
		return;
	}
	int n_work = 1;
	if (lane == 0) work[0] = make_int3(0, n, 56);
	__syncwarp();
	while (n_work > 0) {
		const int3 job = work[--n_work];
		__syncwarp();
		W16 *a = w + job.x;
		const int m = job.y, shift = job.z;
		for (int b = lane; b < 256; b += 32) cnt[b] = 0;
		__syncwarp();
		for (int base = 0; base < m; base += 32) {
			const int k = base + lane;
			const int dig = k < m ? (int)(a[k].x >> shift & 0xff) : 256;
			const unsigned peers = __match_any_sync(FULL, dig);
			if (k < m && (peers & lanemask_lt(lane)) == 0) cnt[dig] += __popc(peers);
			__syncwarp();
		}
		int loc[8], sum = 0;
		bool one = false;
#pragma unroll
		for (int q = 0; q < 8; ++q) loc[q] = cnt[lane * 8 + q], sum += loc[q], one |= loc[q] == m;
		int incl = sum;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const int v = __shfl_up_sync(FULL, incl, d);
			if (lane >= d) incl += v;
		}
		int run = incl - sum;
#pragma unroll
		for (int q = 0; q < 8; ++q) head[lane * 8 + q] = run, run += loc[q], tail[lane * 8 + q] = run;
		const bool one_bucket = __any_sync(FULL, one);
		__syncwarp();
		if (!one_bucket && lane == 0) {                   // ksort.h:131-143, verbatim in effect
			for (int k = 0; k < 256;) {
				if (head[k] == tail[k]) { ++k; continue; }
				int l = (int)(a[head[k]].x >> shift & 0xff);
				if (l == k) { ++head[k]; continue; }
				W16 carry = a[head[k]];
				do {
					const W16 out = a[head[l]];
					a[head[l]++] = carry;
					carry = out;
					l = (int)(carry.x >> shift & 0xff);
				} while (l != k);
				a[head[k]++] = carry;
			}
		}
		__syncwarp();
		if (shift) {                                      // ksort.h:145-150: recurse into buckets > 64, insertion-sort the others
			const int next = shift > 8 ? shift - 8 : 0;
			int big = 0;
#pragma unroll
			for (int q = 0; q < 8; ++q) big += loc[q] > 64;
			int bincl = big;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const int v = __shfl_up_sync(FULL, bincl, d);
				if (lane >= d) bincl += v;
			}
			int slot = n_work + bincl - big;
			const int n_big = __shfl_sync(FULL, bincl, 31);
#pragma unroll
			for (int q = 0; q < 8; ++q) {
				const int c = loc[q], beg = tail[lane * 8 + q] - c;
				if (c > 64) {
					if (slot < work_cap) work[slot] = make_int3(job.x + beg, c, next);
					else insertion_by_x(a + beg, c);          // unreachable: work_cap >= n / 65 + 1 pending ranges always fit
					++slot;
				} else if (c > 1) insertion_by_x(a + beg, c);
			}
			n_work += n_big < work_cap - n_work ? n_big : work_cap - n_work;
		}
		__syncwarp();
	}
}

__global__ void __launch_bounds__(32) tie_replay_kernel(const SeedArgs s, const DeviceIndex ix)
{
	extern __shared__ __align__(16) unsigned char replay_smem[];
	W16 *buf = (W16*)replay_smem;
	int *sm = (int*)(buf + REPLAY_SMEM);
	int3 *work_s = (int3*)(sm + 768);
	const int lane = threadIdx.x & 31;
	int n_tie = 0;
	if (lane == 0) n_tie = s.tie_count[0];
	n_tie = __shfl_sync(FULL, n_tie, 0);
	for (;;) {
		int slot = 0;
		if (lane == 0) slot = atomicAdd(&s.tie_count[2], 1);
		slot = __shfl_sync(FULL, slot, 0);
		if (slot >= n_tie) break;
		const int r = s.tie_list[slot];
		expand_read(s, ix, r, lane, s.a);
		__syncwarp();
		const int64_t o = s.a_off[r];
		const int n = (int)(s.a_off[r + 1] - o);
		W16 *w = (W16*)(s.a + o);
		if (n <= REPLAY_SMEM) {
			for (int k = lane; k < n; k += 32) buf[k] = w[k];
			__syncwarp();
			warp_replay_sort(buf, n, sm, work_s, REPLAY_WORK, lane);
			for (int k = lane; k < n; k += 32) w[k] = buf[k];
		} else warp_replay_sort(w, n, sm, (int3*)(s.a_tmp + o), (int)((int64_t)n * 16 / (int64_t)sizeof(int3)), lane);
		__syncwarp();
	}
}

}  // anonymous namespace

// ---------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------
int launch_index_build(const DeviceIndex &ix, const uint64_t *d_keys, const uint64_t *d_vals, cudaStream_t stream)
{
	cudaMemsetAsync(ix.tab_keys, 0xff, sizeof(uint64_t) << ix.log2cap, stream);
	if (ix.n_keys <= 0) return 0;
	index_insert_kernel<<<(unsigned)((ix.n_keys + 255) / 256), 256, 0, stream>>>(ix.n_keys, d_keys, d_vals, ix.tab_keys, ix.tab_vals, ix.log2cap);
	return 1;
}

int launch_index_lookup(const DeviceIndex &ix, int64_t n, const ulonglong2 *mv, const uint64_t *raw, int32_t *occ, uint64_t *hv, cudaStream_t stream)
{
	if (n <= 0) return 0;
	index_lookup_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(ix, n, mv, raw, occ, hv);
	return 1;
}

int sketch_tile_positions(int w)
{
	static const bool off = getenv("MM2B_SKETCH8") && atoi(getenv("MM2B_SKETCH8")) == 0;
	return w >= S8_P && !off ? SKETCH8_TILE : SKETCH_TILE;
}

int launch_sketch(const SeedArgs &s, cudaStream_t stream)
{
	if (s.n_tiles <= 0) return 0;
	zero_words_kernel<<<(unsigned)((s.n_tiles + 255) / 256), 256, 0, stream>>>(s.tile_state, s.n_tiles, (unsigned long long*)s.tile_ticket, 1);    // (the ticket owns 8 bytes)
	const bool k32 = 2 * s.k < 32;
	if (sketch_tile_positions(s.w) == SKETCH8_TILE) {       // eight positions per thread
		static const bool occ3 = getenv("MM2B_SKETCH8_CTAS") && atoi(getenv("MM2B_SKETCH8_CTAS")) == 3;     // register budget for 3 or 4 CTAs per SM
		if (k32 && occ3) sketch8_kernel<true, 3><<<s.n_tiles, S8_THREADS, 0, stream>>>(s);
		else if (k32) sketch8_kernel<true, 4><<<s.n_tiles, S8_THREADS, 0, stream>>>(s);
		else if (occ3) sketch8_kernel<false, 3><<<s.n_tiles, S8_THREADS, 0, stream>>>(s);
		else sketch8_kernel<false, 4><<<s.n_tiles, S8_THREADS, 0, stream>>>(s);
		return 1;
	}
#define MM2B_SKETCH(W_) do { if (k32) sketch_kernel<true, W_><<<s.n_tiles, SKETCH_TILE, 0, stream>>>(s); else sketch_kernel<false, W_><<<s.n_tiles, SKETCH_TILE, 0, stream>>>(s); } while (0)
	switch (s.w) {                          // the presets' window sizes get an unrolled scan (options.c:82-150: 5, 10, 11, 19)
	case 5: MM2B_SKETCH(5); break;
	case 10: MM2B_SKETCH(10); break;
	case 11: MM2B_SKETCH(11); break;
	case 19: MM2B_SKETCH(19); break;
	default: MM2B_SKETCH(0); break;
	}
#undef MM2B_SKETCH
	return 1;
}

int launch_scan_i64(const int64_t *in, int64_t *out, int64_t n, cudaStream_t stream)
{
	scan_kernel<int64_t><<<1, 1024, 0, stream>>>(in, out, n);
	return 1;
}

int launch_export_scalars(int64_t *host_mapped, const int64_t *a, const int64_t *b, const int *c, cudaStream_t stream)
{
	export_scalars_kernel<<<1, 1, 0, stream>>>(host_mapped, a, b, c);
	return 1;
}

int launch_read_offsets(const SeedArgs &s, cudaStream_t stream)
{
	read_offsets_kernel<<<(unsigned)((s.n_reads + 1 + 255) / 256), 256, 0, stream>>>(s);
	return 1;
}

static int warp_grid(int64_t n_reads, int n_sms, int warps_per_cta)
{
	int64_t blocks = (n_reads + warps_per_cta - 1) / warps_per_cta;
	const int64_t cap = (int64_t)n_sms * 16;
	return (int)(blocks > cap ? cap : blocks < 1 ? 1 : blocks);
}

int launch_matches(const SeedArgs &s, int n_sms, cudaStream_t stream)
{
	if (s.n_reads <= 0) return 0;
	matches_kernel<<<warp_grid(s.n_reads, n_sms, 8), 256, 0, stream>>>(s);
	return 1;
}

int launch_expand(const SeedArgs &s, const DeviceIndex &ix, int64_t n_mv, cudaStream_t stream)
{
	if (s.n_reads <= 0 || n_mv <= 0) return 0;
	expand_kernel<<<(unsigned)((n_mv + 255) / 256), 256, 0, stream>>>(s, ix, n_mv);
	return 1;
}

int launch_sort(const SeedArgs &s, const DeviceIndex &ix, int n_sms, cudaStream_t stream)
{
	if (s.n_reads <= 0) return 0;
	zero_words_kernel<<<1, 32, 0, stream>>>((unsigned long long*)s.tie_count, 2, nullptr, 0);
	sort_kernel<<<warp_grid(s.n_reads, n_sms, SORT_WARPS), SORT_WARPS * 32, 0, stream>>>(s);
	constexpr int replay_bytes = REPLAY_SMEM * 16 + 768 * 4 + REPLAY_WORK * 12;
	static bool attr_set[64];
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 64 && !attr_set[dev]) {
		cudaFuncSetAttribute(tie_replay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, replay_bytes);
		attr_set[dev] = true;
	}
	tie_replay_kernel<<<n_sms, 32, replay_bytes, stream>>>(s, ix);
	return 2;
}

}  // namespace mm2b
