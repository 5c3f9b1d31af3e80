// TEST INFRASTRUCTURE — a CPU stand-in for the batch entry points of the chaining backend, so that the product's phase-split
// caller (minimap2-fpga_b200/host/map_batch.cpp) can be run under the reference CLI on a machine without a GPU:
// oracle/_ref/minimap2-batch-sw = the reference CLI + map_batch.cpp + THIS file, chaining every staged read with the reference's
// own chain.c (mm_chain_dp_ref).  Its PAF must be byte-identical to the reference's, which checks the staging, the CSR layout,
// the index gather, the kalloc discipline of the two halves of mm_map_frag and the second chaining pass on real mapping runs.
// Nothing here is on the product path; the product links libmm2chain_b200.so instead.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mm2chain_b200.h"

extern "C" mm2b_anchor_t *mm_chain_dp_ref(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc, float gap_scale,
                                          int is_cdna, int n_segs, int64_t n, mm2b_anchor_t *a, int *n_u_, uint64_t **_u, void *km, int tid);

extern "C" {

const char *mm2b_last_error(void) { return "software stand-in"; }
void *mm2b_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void mm2b_host_free(void *p) { free(p); }

int mm2b_chain_batch_ex(const mm2b_params_t *par, int64_t n_reads, const int64_t *off, const mm2b_anchor_t *a,
                        int32_t *n_u, int32_t *n_v, int32_t *status, int64_t *u_off, int64_t *b_off,
                        uint64_t *u, int64_t u_cap, mm2b_anchor_t *b, int32_t *bi, int64_t b_cap, unsigned flags, mm2b_stats_t *stats)
{
	(void)u_cap, (void)b_cap, (void)flags, (void)stats;
	if (b || !bi) { fprintf(stderr, "batch_sw_shim: only the index output is implemented\n"); exit(1); }
	for (int64_t r = 0; r < n_reads; ++r) {
		const int64_t o = off[r], n = off[r + 1] - o;
		mm2b_anchor_t *ac = 0;
		if (n > 0) { ac = (mm2b_anchor_t*)malloc((size_t)n * 16); memcpy(ac, a + o, (size_t)n * 16); }     // consumed by the reference (chain.c:421)
		int nu = 0;
		uint64_t *uu = 0;
		mm2b_anchor_t *bb = mm_chain_dp_ref(par->max_dist_x, par->max_dist_y, par->bw, par->max_skip, par->max_iter, par->min_cnt, par->min_sc, par->gap_scale,
		                                    par->is_cdna, par->n_segs, n, ac, &nu, &uu, 0, 0);
		u_off[r] = o, b_off[r] = o;
		n_u[r] = nu, n_v[r] = 0;
		status[r] = n == 0 ? MM2B_READ_EMPTY : (uu == 0 ? MM2B_READ_NO_CHAIN : MM2B_READ_OK);
		int64_t pos = 0;
		for (int c = 0; c < nu; ++c) {
			const int len = (int32_t)uu[c];
			u[o + c] = uu[c];
			int64_t k = 0;                          // a chain's anchors are in increasing index order
			for (int q = 0; q < len; ++q, ++k) {
				while (k < n && (a[o + k].x != bb[pos + q].x || a[o + k].y != bb[pos + q].y)) ++k;
				if (k >= n) { fprintf(stderr, "batch_sw_shim: chained anchor not found in its read\n"); exit(1); }
				bi[o + pos + q] = (int32_t)k;
			}
			pos += len;
		}
		n_v[r] = (int32_t)pos;
		free(uu), free(bb);
	}
	u_off[n_reads] = b_off[n_reads] = n_reads > 0 ? off[n_reads] : 0;
	return MM2B_OK;
}

}
