/* TEST INFRASTRUCTURE — capture shim around the REFERENCE mm_chain_dp.
 *
 * oracle/Makefile compiles /root/reference/chain.c with -Dmm_chain_dp=mm_chain_dp_ref, so the
 * reference CLI built into oracle/_ref/minimap2-sw resolves its two call sites
 * (map.c:316, map.c:338) to the wrapper below.  With MM2_DUMP=<path> in the environment every
 * call is appended to <path> in the format of oracle/dump_format.h; otherwise it is a pure
 * pass-through.  MM2_DUMP_NO_FPV=1 leaves the f/p/v arrays out (bench workloads: 12 B/anchor less).  The f/p/v arrays come from the hook the Makefile splices in after the DP fill
 * (chain.c:346).  Nothing here is on the product path.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <pthread.h>
#include "dump_format.h"

typedef struct { uint64_t x, y; } mm128_t;

#ifdef __cplusplus
extern "C" {
#endif
mm128_t *mm_chain_dp_ref(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
		float gap_scale, int is_cdna, int n_segs, int64_t n, mm128_t *a, int *n_u_, uint64_t **_u, void *km, int tid);
mm128_t *mm_chain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
		float gap_scale, int is_cdna, int n_segs, int64_t n, mm128_t *a, int *n_u_, uint64_t **_u, void *km, int tid);
void mm2_ref_hook_fpv(int64_t n, const int32_t *f, const int32_t *p, const int32_t *v);
#ifdef __cplusplus
}
#endif

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static FILE *g_fp;
static int g_state; /* 0 = unknown, 1 = dumping, -1 = off */
static int g_no_fpv;

static __thread int32_t *t_fpv;   /* 3*n int32 captured by the hook for the call in flight */
static __thread int64_t t_fpv_n;
static __thread int t_armed;

void mm2_ref_hook_fpv(int64_t n, const int32_t *f, const int32_t *p, const int32_t *v)
{
	if (!t_armed) return;
	t_fpv = (int32_t*)malloc((size_t)n * 12 + 1);
	memcpy(t_fpv, f, (size_t)n * 4);
	memcpy(t_fpv + n, p, (size_t)n * 4);
	memcpy(t_fpv + 2 * n, v, (size_t)n * 4);
	t_fpv_n = n;
}

static int dumping(void)
{
	if (g_state == 0) {
		pthread_mutex_lock(&g_lock);
		if (g_state == 0) {
			const char *fn = getenv("MM2_DUMP");
			const char *nf = getenv("MM2_DUMP_NO_FPV");
			g_no_fpv = nf && *nf && *nf != '0';
			if (fn && *fn && (g_fp = fopen(fn, "wb")) != 0) g_state = 1;
			else g_state = -1;
		}
		pthread_mutex_unlock(&g_lock);
	}
	return g_state == 1;
}

mm128_t *mm_chain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
		float gap_scale, int is_cdna, int n_segs, int64_t n, mm128_t *a, int *n_u_, uint64_t **_u, void *km, int tid)
{
	mm128_t *a_copy = 0, *b;
	mm2_dump_hdr_t h;
	int64_t i, n_v = 0;
	if (!dumping() || n <= 0 || a == 0)
		return mm_chain_dp_ref(max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc, gap_scale, is_cdna, n_segs, n, a, n_u_, _u, km, tid);
	a_copy = (mm128_t*)malloc((size_t)n * sizeof(mm128_t));
	memcpy(a_copy, a, (size_t)n * sizeof(mm128_t)); /* the reference frees `a` (chain.c:39,356,421) */
	t_armed = !g_no_fpv, t_fpv = 0, t_fpv_n = 0;
	b = mm_chain_dp_ref(max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc, gap_scale, is_cdna, n_segs, n, a, n_u_, _u, km, tid);
	t_armed = 0;
	memset(&h, 0, sizeof(h));
	h.magic = MM2_DUMP_MAGIC;
	h.max_dist_x = max_dist_x, h.max_dist_y = max_dist_y, h.bw = bw, h.max_skip = max_skip, h.max_iter = max_iter;
	h.min_cnt = min_cnt, h.min_sc = min_sc, h.is_cdna = is_cdna, h.n_segs = n_segs, h.gap_scale = gap_scale;
	h.n = n;
	h.n_u = *n_u_;
	if (t_fpv && t_fpv_n == n) h.flags |= MM2_DUMP_HAS_FPV;
	if (b == 0) h.flags |= MM2_DUMP_B_NULL;
	if (*_u == 0) h.flags |= MM2_DUMP_U_NULL;
	for (i = 0; i < h.n_u; ++i) n_v += (int32_t)(*_u)[i];
	h.n_v = (int32_t)n_v;
	pthread_mutex_lock(&g_lock);
	fwrite(&h, sizeof(h), 1, g_fp);
	fwrite(a_copy, sizeof(mm128_t), (size_t)n, g_fp);
	if (h.flags & MM2_DUMP_HAS_FPV) fwrite(t_fpv, 4, (size_t)n * 3, g_fp);
	if (h.n_u > 0) fwrite(*_u, 8, (size_t)h.n_u, g_fp);
	if (n_v > 0) fwrite(b, sizeof(mm128_t), (size_t)n_v, g_fp);
	fflush(g_fp);
	pthread_mutex_unlock(&g_lock);
	free(a_copy);
	free(t_fpv), t_fpv = 0;
	return b;
}
