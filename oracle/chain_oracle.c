/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's *software* chaining, the parity target for the CUDA path:
 * minimap2 v2.18 mm_chain_dp (/root/reference/chain.c:29-423) with ENABLE_MAX_SKIP_ON_SW and
 * the HW/SW predictor removed.  Parity status: PINNED — tests/test_oracle.py checks
 * this file against the reference's own compiled chain.c (oracle/_ref/libmm2ref.so, built in place
 * from /root/reference by oracle/Makefile) and against tests/golden/ fixtures captured from the
 * reference CLI (f/p/v per anchor, u[] and b[] per read).  The reference tree itself stores no
 * golden vectors for this path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this.  The product (minimap2-fpga_b200/) never does.
 */
#include <stdlib.h>
#include <string.h>
#include "chain_oracle.h"

/* floor(log2(v)) for v > 0; chain.c:22-27 does it with a 256-entry table */
static inline int floor_log2_u32(uint32_t v)
{
	int r = 0;
	while (v >>= 1) ++r;
	return r;
}

/* chain.c:48-49 — the 8-bit q_span field only; double arithmetic rounded once to float */
float mm2o_avg_qspan_scaled(int64_t n, const mm2o_anchor_t *a)
{
	uint64_t sum = 0;
	int64_t i;
	float r;
	for (i = 0; i < n; ++i) sum += a[i].y >> 32 & 0xff;
	r = .01 * (float)sum / n;
	return r;
}

#define SEG_SHIFT 48                   /* MM_SEED_SEG_SHIFT, mmpriv.h:22 */
#define SEG_OF(y) ((int32_t)(((y) >> SEG_SHIFT) & 0xff))

/* chain.c:184-238: fill f (best score ending at i), p (predecessor), v (peak score on the path), t (visit stamps) */
void mm2o_dp_fill(const mm2o_params_t *par, int64_t n, const mm2o_anchor_t *a, int32_t *f, int32_t *p, int32_t *v, int32_t *t, mm2o_stats_t *stat)
{
	const int32_t max_dist_x = par->max_dist_x, max_dist_y = par->max_dist_y, bw = par->bw;
	const float avg = mm2o_avg_qspan_scaled(n, a);
	int64_t i, j, st = 0, cells = 0, wcells = 0;
	memset(t, 0, (size_t)n * 4);
	for (i = 0; i < n; ++i) {
		const uint64_t ri = a[i].x;
		const int32_t qi = (int32_t)a[i].y, q_span = (int32_t)(a[i].y >> 32 & 0xff);
		const int32_t sidi = SEG_OF(a[i].y);
		int64_t best_j = -1;
		int32_t best = q_span, n_skip = 0;
		while (st < i && ri > a[st].x + max_dist_x) ++st;          /* chain.c:192 */
		if (i - st > par->max_iter) st = i - par->max_iter;        /* chain.c:193 */
		wcells += i - st;
		for (j = i - 1; j >= st; --j) {
			const int64_t dr = (int64_t)(ri - a[j].x);
			const int32_t dq = qi - (int32_t)a[j].y;
			const int same = sidi == SEG_OF(a[j].y);
			int32_t dd, sc, lg, gap;
			++cells;
			if ((same && dr == 0) || dq <= 0) continue;                               /* :202 */
			if ((same && dq > max_dist_y) || dq > max_dist_x) continue;               /* :203 */
			dd = (int32_t)(dr > dq ? dr - dq : dq - dr);                              /* :204 */
			if (same && dd > bw) continue;                                            /* :205 */
			if (par->n_segs > 1 && !par->is_cdna && same && dr > max_dist_y) continue; /* :206 */
			sc = (int32_t)(dq < dr ? dq : dr);                                        /* :207-208 */
			if (sc > q_span) sc = q_span;
			lg = dd ? floor_log2_u32((uint32_t)dd) : 0;                               /* :209 */
			if (par->is_cdna || !same) {                                              /* :211-217 */
				const int c_lin = (int)(dd * avg), c_log = lg;
				gap = 0;
				if (!same && dr == 0) ++sc;
				else if (dr > dq || !same) gap = c_lin < c_log ? c_lin : c_log;
				else gap = c_lin + (c_log >> 1);
			} else gap = (int)(dd * avg) + (lg >> 1);                                 /* :218 */
			sc -= (int)((double)gap * par->gap_scale + .499);                         /* :219 */
			sc += f[j];                                                               /* :220 */
			if (sc > best) {                                                          /* :226-228 */
				best = sc, best_j = j;
				if (n_skip > 0) --n_skip;
			} else if (t[j] == i) {                                                   /* :229-232 */
				if (++n_skip > par->max_skip) break;
			}
			if (p[j] >= 0) t[p[j]] = (int32_t)i;                                      /* :233 */
		}
		f[i] = best, p[i] = (int32_t)best_j;
		v[i] = best_j >= 0 && v[best_j] > best ? v[best_j] : best;                    /* :237 */
	}
	if (stat) stat->cells += cells, stat->window_cells += wcells, stat->n_anchors += n;
}

/* ---- sorting helpers -------------------------------------------------------------------------- */

static int cmp_u64_desc(const void *pa, const void *pb)
{
	uint64_t x = *(const uint64_t*)pa, y = *(const uint64_t*)pb;
	return x < y ? 1 : x > y ? -1 : 0;
}

static void insertion_by_x(mm2o_anchor_t *a, int64_t n)  /* stable; ksort.h:106-115 */
{
	int64_t i, j;
	for (i = 1; i < n; ++i) {
		mm2o_anchor_t key = a[i];
		for (j = i; j > 0 && key.x < a[j - 1].x; --j) a[j] = a[j - 1];
		a[j] = key;
	}
}

/* In-place MSD byte radix sort by .x, ksort.h:116-146.  It is NOT stable and the chain order the
 * reference emits for equal keys is whatever this exact permutation scheme produces, so it is
 * restated operation for operation (bucket cursors as offsets instead of pointers). */
static void flag_sort_by_x(mm2o_anchor_t *a, int64_t n, int shift)
{
	int64_t head[256], tail[256], cnt[256], i;
	int k;
	memset(cnt, 0, sizeof(cnt));
	for (i = 0; i < n; ++i) ++cnt[a[i].x >> shift & 0xff];
	for (k = 0, i = 0; k < 256; ++k) head[k] = i, i += cnt[k], tail[k] = i;
	for (k = 0; k < 256;) {
		if (head[k] == tail[k]) { ++k; continue; }
		int l = (int)(a[head[k]].x >> shift & 0xff);
		if (l == k) { ++head[k]; continue; }
		mm2o_anchor_t carry = a[head[k]];
		do {                                   /* follow the displacement cycle until something lands in bucket k */
			mm2o_anchor_t out = a[head[l]];
			a[head[l]++] = carry;
			carry = out;
			l = (int)(carry.x >> shift & 0xff);
		} while (l != k);
		a[head[k]++] = carry;
	}
	if (shift) {
		const int next = shift > 8 ? shift - 8 : 0;
		for (k = 0; k < 256; ++k) {
			const int64_t beg = tail[k] - cnt[k];
			if (cnt[k] > 64) flag_sort_by_x(a + beg, cnt[k], next);
			else if (cnt[k] > 1) insertion_by_x(a + beg, cnt[k]);
		}
	}
}

void mm2o_sort_128x(mm2o_anchor_t *a, int64_t n)   /* radix_sort_128x, ksort.h:147-151, misc.c:155-156 */
{
	if (n <= 64) insertion_by_x(a, n);
	else flag_sort_by_x(a, n, 56);
}

/* ---- full path: DP fill + chain ends/peaks + sort + backtrack + compaction + order by ref pos ---- */

int mm2o_chain(const mm2o_params_t *par, int64_t n, const mm2o_anchor_t *a, int32_t *f_out, int32_t *p_out, int32_t *v_out,
               int32_t *n_u_out, uint64_t *u_out, int64_t *n_v_out, mm2o_anchor_t *b_out, mm2o_stats_t *stat)
{
	int32_t *f, *p, *v, *t, *path, n_u, n_kept, k;
	int64_t i, j, n_v;
	uint64_t *u;
	mm2o_anchor_t *w;

	*n_u_out = 0, *n_v_out = 0;
	if (n == 0 || a == 0) return MM2O_EMPTY;                                           /* chain.c:38-41 */
	f = (int32_t*)malloc((size_t)n * 4), p = (int32_t*)malloc((size_t)n * 4);
	v = (int32_t*)malloc((size_t)n * 4), t = (int32_t*)malloc((size_t)n * 4);
	mm2o_dp_fill(par, n, a, f, p, v, t, stat);
	if (f_out) memcpy(f_out, f, (size_t)n * 4);
	if (p_out) memcpy(p_out, p, (size_t)n * 4);
	if (v_out) memcpy(v_out, v, (size_t)n * 4);

	/* chain.c:348-367 — anchors nobody points to, whose path peak reaches min_sc, walked back to that peak */
	memset(t, 0, (size_t)n * 4);
	for (i = 0; i < n; ++i) if (p[i] >= 0) t[p[i]] = 1;
	for (i = 0, n_u = 0; i < n; ++i) if (t[i] == 0 && v[i] >= par->min_sc) ++n_u;
	if (n_u == 0) { free(f), free(p), free(v), free(t); return MM2O_NO_CHAIN; }       /* chain.c:355-358 */
	u = (uint64_t*)malloc((size_t)n_u * 8);
	for (i = 0, n_u = 0; i < n; ++i) {
		if (t[i] == 0 && v[i] >= par->min_sc) {
			j = i;
			while (j >= 0 && f[j] < v[j]) j = p[j];
			if (j < 0) j = i;
			u[n_u++] = (uint64_t)f[j] << 32 | (uint64_t)j;
		}
	}
	qsort(u, (size_t)n_u, 8, cmp_u64_desc);               /* chain.c:368-372: ascending + reverse == descending; equal values (two ends sharing a peak) are interchangeable */

	/* chain.c:374-391 — greedy backtrack in score order; marks persist even when a candidate is dropped */
	path = (int32_t*)malloc((size_t)n * 4);
	memset(t, 0, (size_t)n * 4);
	for (i = 0, n_v = 0, n_kept = 0; i < n_u; ++i) {
		const int64_t n_v0 = n_v;
		const int32_t k0 = n_kept, sc_peak = (int32_t)(u[i] >> 32);
		j = (int32_t)u[i];
		do {
			path[n_v++] = (int32_t)j;
			t[j] = 1;
			j = p[j];
		} while (j >= 0 && t[j] == 0);
		if (j < 0) {
			if (n_v - n_v0 >= par->min_cnt) u[n_kept++] = u[i] >> 32 << 32 | (uint64_t)(n_v - n_v0);
		} else if (sc_peak - f[j] >= par->min_sc) {
			if (n_v - n_v0 >= par->min_cnt) u[n_kept++] = ((u[i] >> 32) - (uint64_t)f[j]) << 32 | (uint64_t)(n_v - n_v0);
		}
		if (k0 == n_kept) n_v = n_v0;
	}
	n_u = n_kept;

	/* chain.c:396-422 — anchors of each kept chain in forward order, chains ordered by first-anchor x */
	w = (mm2o_anchor_t*)malloc((size_t)(n_u > 0 ? n_u : 1) * sizeof(mm2o_anchor_t));
	for (i = 0, k = 0; i < n_u; ++i) {
		const int32_t cnt = (int32_t)u[i];
		w[i].x = a[path[k + cnt - 1]].x;                 /* first anchor of chain i in forward order */
		w[i].y = (uint64_t)k << 32 | (uint64_t)i;
		k += cnt;
	}
	mm2o_sort_128x(w, n_u);
	for (i = 0, n_v = 0; i < n_u; ++i) {
		const int32_t src = (int32_t)w[i].y, cnt = (int32_t)u[src];
		const int64_t k0 = (int64_t)(w[i].y >> 32);
		u_out[i] = u[src];
		for (j = 0; j < cnt; ++j) b_out[n_v + j] = a[path[k0 + (cnt - 1 - j)]];
		n_v += cnt;
	}
	*n_u_out = n_u, *n_v_out = n_v;
	if (stat) stat->n_chains += n_u, stat->n_chained += n_v;
	free(f), free(p), free(v), free(t), free(u), free(path), free(w);
	return MM2O_OK;
}
