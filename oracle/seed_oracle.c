/* TEST INFRASTRUCTURE — CPU oracle for the seeding front end: a plain-C restatement of
 *   mm_sketch            /root/reference/sketch.c:77-143   (non-HPC)
 *   mm_idx_get           /root/reference/index.c:81-98     (over the flat index of include/mm2seed_b200.h)
 *   collect_matches      /root/reference/map.c:90-123
 *   collect_seed_hits    /root/reference/map.c:215-247     (with mm2o_sort_128x, chain_oracle.c, for radix_sort_128x at map.c:245)
 * Pinned against the reference itself: oracle/_ref/mm2-seed-ref records what the reference's own functions produce for the same
 * reads (tests/test_seed_oracle.py).  Only tests/, smoke() and bench.py's CPU legs may use this file; the product never does. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "chain_oracle.h"

static int base_code(unsigned char ch)                     /* seq_nt4_table, sketch.c:9-26 */
{
	switch (ch) {
	case 0: case 'A': case 'a': return 0;
	case 1: case 'C': case 'c': return 1;
	case 2: case 'G': case 'g': return 2;
	case 3: case 'T': case 't': case 'U': case 'u': return 3;
	default: return 4;
	}
}

static uint64_t mix64(uint64_t key, uint64_t mask)         /* hash64, sketch.c:28-38 */
{
	key = (~key + (key << 21)) & mask;
	key ^= key >> 24;
	key = (key + (key << 3) + (key << 8)) & mask;
	key ^= key >> 14;
	key = (key + (key << 2) + (key << 4)) & mask;
	key ^= key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}

#define NONE UINT64_MAX

/* Minimizers of seq[0..len) appended to out (capacity cap); returns their number, or -1 if cap is too small.
 * A ring of the last w k-mer records, the current minimum and the rules for when a record is written out (sketch.c:89-142). */
int64_t mm2o_sketch(const char *seq, int len, int w, int k, mm2o_anchor_t *out, int64_t cap)
{
	const uint64_t mask = (1ULL << 2 * k) - 1, top = 2 * (k - 1);
	uint64_t fw = 0, rv = 0;
	mm2o_anchor_t ring[256], best = {NONE, NONE};
	int64_t n = 0;
	int run = 0, slot = 0, best_slot = 0;
	memset(ring, 0xff, sizeof(ring));
#define EMIT(rec) do { if (n >= cap) return -1; out[n++] = (rec); } while (0)
	for (int i = 0; i < len; ++i) {
		const int c = base_code((unsigned char)seq[i]);
		mm2o_anchor_t cur = {NONE, NONE};
		if (c < 4) {
			const int span = run + 1 < k ? run + 1 : k;                       /* sketch.c:107 */
			fw = (fw << 2 | (uint64_t)c) & mask;                              /* sketch.c:108 */
			rv = rv >> 2 | (uint64_t)(3 ^ c) << top;                          /* sketch.c:109 */
			if (fw == rv) continue;                                           /* sketch.c:110: strand unknown, the position takes no slot */
			const int strand = fw < rv ? 0 : 1;
			++run;
			if (run >= k) {
				cur.x = mix64(strand ? rv : fw, mask) << 8 | (uint64_t)span;
				cur.y = (uint64_t)(uint32_t)i << 1 | (uint64_t)strand;         /* rid 0 */
			}
		} else run = 0;                                                       /* sketch.c:117 */
		ring[slot] = cur;
		if (run == w + k - 1 && best.x != NONE) {                             /* first full window: equal hashes not written yet (sketch.c:119-124) */
			for (int j = slot + 1; j < w; ++j) if (ring[j].x == best.x && ring[j].y != best.y) EMIT(ring[j]);
			for (int j = 0; j < slot; ++j) if (ring[j].x == best.x && ring[j].y != best.y) EMIT(ring[j]);
		}
		if (cur.x <= best.x) {                                                /* sketch.c:125-127 */
			if (run >= w + k && best.x != NONE) EMIT(best);
			best = cur, best_slot = slot;
		} else if (slot == best_slot) {                                       /* the minimum leaves the window (sketch.c:128-141) */
			if (run >= w + k - 1 && best.x != NONE) EMIT(best);
			best.x = NONE;
			for (int j = slot + 1; j < w; ++j) if (best.x >= ring[j].x) best = ring[j], best_slot = j;
			for (int j = 0; j <= slot; ++j) if (best.x >= ring[j].x) best = ring[j], best_slot = j;
			if (run >= w + k - 1 && best.x != NONE) {
				for (int j = slot + 1; j < w; ++j) if (ring[j].x == best.x && ring[j].y != best.y) EMIT(ring[j]);
				for (int j = 0; j <= slot; ++j) if (ring[j].x == best.x && ring[j].y != best.y) EMIT(ring[j]);
			}
		}
		if (++slot == w) slot = 0;
	}
	if (best.x != NONE) EMIT(best);
#undef EMIT
	return n;
}

/* ---- index (flat arrays) ------------------------------------------------------------------------------------------ */
typedef struct {
	int k, w;
	uint64_t cap_mask;
	uint64_t *tk, *tv;
	const uint64_t *pos;
} mm2o_index_t;

mm2o_index_t *mm2o_index_new(int k, int w, int64_t n_keys, const uint64_t *keys, const uint64_t *vals, const uint64_t *pos)
{
	mm2o_index_t *ix = (mm2o_index_t*)calloc(1, sizeof(*ix));
	uint64_t cap = 16;
	while (cap < (uint64_t)n_keys * 2) cap <<= 1;
	ix->k = k, ix->w = w, ix->cap_mask = cap - 1, ix->pos = pos;
	ix->tk = (uint64_t*)malloc(cap * 8), ix->tv = (uint64_t*)malloc(cap * 8);
	memset(ix->tk, 0xff, cap * 8);
	for (int64_t i = 0; i < n_keys; ++i) {
		uint64_t h = ((keys[i] >> 1) * 0xD6E8FEB86659FD93ULL >> 17) & ix->cap_mask;
		while (ix->tk[h] != NONE) h = (h + 1) & ix->cap_mask;
		ix->tk[h] = keys[i], ix->tv[h] = vals[i];
	}
	return ix;
}

void mm2o_index_free(mm2o_index_t *ix)
{
	if (ix) free(ix->tk), free(ix->tv), free(ix);
}

/* mm_idx_get (index.c:81-98): number of occurrences; *first points at them */
static int index_get(const mm2o_index_t *ix, uint64_t minimizer, const uint64_t **first)
{
	uint64_t h = (minimizer * 0xD6E8FEB86659FD93ULL >> 17) & ix->cap_mask;
	for (; ix->tk[h] != NONE; h = (h + 1) & ix->cap_mask)
		if (ix->tk[h] >> 1 == minimizer) {
			if (ix->tk[h] & 1) { *first = &ix->tv[h]; return 1; }
			*first = ix->pos + (ix->tv[h] >> 32);
			return (int)(uint32_t)ix->tv[h];
		}
	*first = 0;
	return 0;
}

int mm2o_index_get(const mm2o_index_t *ix, uint64_t minimizer, uint64_t *val)
{
	const uint64_t *p;
	const int n = index_get(ix, minimizer, &p);
	*val = n ? p[0] : 0;
	return n;
}

/* collect_matches + collect_seed_hits for one read (one segment, no skip flags): anchors sorted like map.c:245 into a (capacity cap_a),
 * mini_pos (capacity n_mv) as 64-bit q_span << 32 | pos.  Returns the number of anchors or -1. */
int64_t mm2o_seed(const mm2o_index_t *ix, int max_occ, int qlen, int64_t n_mv, const mm2o_anchor_t *mv, mm2o_anchor_t *a, int64_t cap_a,
                  int32_t *rep_len, int32_t *n_mini_pos, uint64_t *mini_pos)
{
	int64_t n_a = 0;
	int rep_st = 0, rep_en = 0, rep = 0, n_mp = 0;
	for (int64_t i = 0; i < n_mv; ++i) {
		const uint64_t minimizer = mv[i].x >> 8;
		const uint32_t q_pos = (uint32_t)mv[i].y, q_span = (uint32_t)(mv[i].x & 0xff);
		const uint64_t *hits;
		const int t = index_get(ix, minimizer, &hits);
		if (t >= max_occ) {                                                   /* map.c:104-110 */
			const int en = (int)(q_pos >> 1) + 1, st = en - (int)q_span;
			if (st > rep_en) rep += rep_en - rep_st, rep_st = st, rep_en = en;
			else rep_en = en;
			continue;
		}
		mini_pos[n_mp++] = (uint64_t)q_span << 32 | q_pos >> 1;               /* map.c:117 */
		const int tandem = (i > 0 && mv[i - 1].x >> 8 == minimizer) || (i + 1 < n_mv && mv[i + 1].x >> 8 == minimizer);   /* map.c:113-115 */
		for (int h = 0; h < t; ++h) {                                         /* map.c:226-243 */
			const uint64_t r = hits[h];
			mm2o_anchor_t p;
			if (n_a >= cap_a) return -1;
			if ((r & 1) == (q_pos & 1)) {
				p.x = (r & 0xffffffff00000000ULL) | (uint32_t)r >> 1;
				p.y = (uint64_t)q_span << 32 | q_pos >> 1;
			} else {
				p.x = 1ULL << 63 | (r & 0xffffffff00000000ULL) | (uint32_t)r >> 1;
				p.y = (uint64_t)q_span << 32 | (uint32_t)(qlen - (int)((q_pos >> 1) + 1 - q_span) - 1);
			}
			p.y |= (mv[i].y >> 32) << 48;
			if (tandem) p.y |= 1ULL << 42;                                    /* MM_SEED_TANDEM, mmpriv.h:19 */
			a[n_a++] = p;
		}
	}
	rep += rep_en - rep_st;                                                   /* map.c:120 */
	*rep_len = rep, *n_mini_pos = n_mp;
	mm2o_sort_128x(a, n_a);                                                   /* map.c:245 */
	return n_a;
}
