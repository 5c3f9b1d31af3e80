/* TEST INFRASTRUCTURE — CPU oracle for the chaining hot path (see chain_oracle.c).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it. */
#ifndef MM2O_CHAIN_ORACLE_H
#define MM2O_CHAIN_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t x, y; } mm2o_anchor_t;               /* == mm128_t, minimap.h:53 */

typedef struct {                                               /* positional args of mm_chain_dp, chain.c:29 */
	int32_t max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc, is_cdna, n_segs;
	float gap_scale;
} mm2o_params_t;

typedef struct {
	int64_t cells;         /* iterations of the inner j loop the reference executes (chain.c:197), incl. `continue`d ones */
	int64_t window_cells;  /* sum_i (i - st_i) after the max_iter clamp (chain.c:192-193) */
	int64_t n_anchors, n_chains, n_chained;
} mm2o_stats_t;

enum { MM2O_EMPTY = 0, MM2O_NO_CHAIN = 1, MM2O_OK = 2 };

float mm2o_avg_qspan_scaled(int64_t n, const mm2o_anchor_t *a);
void mm2o_dp_fill(const mm2o_params_t *par, int64_t n, const mm2o_anchor_t *a, int32_t *f, int32_t *p, int32_t *v, int32_t *t, mm2o_stats_t *st);
int mm2o_chain(const mm2o_params_t *par, int64_t n, const mm2o_anchor_t *a, int32_t *f, int32_t *p, int32_t *v,
               int32_t *n_u, uint64_t *u, int64_t *n_v, mm2o_anchor_t *b, mm2o_stats_t *st);
void mm2o_sort_128x(mm2o_anchor_t *a, int64_t n);

/* batch replay over CSR-packed reads with n_threads pthreads; see replay.c */
typedef mm2o_anchor_t *(*mm2o_ref_fn_t)(int, int, int, int, int, int, int, float, int, int, int64_t, mm2o_anchor_t*, int*, uint64_t**, void*, int);
double mm2o_replay(const mm2o_params_t *par, int64_t n_reads, const int64_t *off, const mm2o_anchor_t *a, int n_threads,
                   mm2o_ref_fn_t ref_fn, int32_t *n_u, int64_t *n_v, uint64_t *u, mm2o_anchor_t *b, mm2o_stats_t *st);

#ifdef __cplusplus
}
#endif
#endif
