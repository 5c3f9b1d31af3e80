// TEST INFRASTRUCTURE — a CPU stand-in for the batch entry points of the chaining backend, so that the product's phase-split
// caller (minimap2-fpga_b200/host/map_batch.cpp) can be run under the reference CLI on a machine without a GPU:
// oracle/_ref/minimap2-batch-sw = the reference CLI + map_batch.cpp + THIS file, chaining every staged read with the reference's
// own chain.c (mm_chain_dp_ref).  Its PAF must be byte-identical to the reference's, which checks the staging, the CSR layout,
// the index gather, the kalloc discipline of the two halves of mm_map_frag and the second chaining pass on real mapping runs.
// The seeding front end (mm2b_map_batch) is stood in for as well: minimizers and anchors from the seeding oracle (seed_oracle.c),
// chains from the reference's chain.c — so the front-end path of map_batch.cpp (sequence staging, result segments, mini_pos
// expansion) is checked under the real CLI too.
// Nothing here is on the product path; the product links libmm2chain_b200.so instead.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mm2seed_b200.h"
#include "chain_oracle.h"

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


// ---- seeding front end, CPU stand-in ---------------------------------------------------------------------------------------
}   // extern "C"

extern "C" {
typedef struct mm2o_index_s mm2o_index_t;
mm2o_index_t *mm2o_index_new(int k, int w, int64_t n_keys, const uint64_t *keys, const uint64_t *vals, const uint64_t *pos);
void mm2o_index_free(mm2o_index_t *ix);
int64_t mm2o_sketch(const char *seq, int len, int w, int k, mm2o_anchor_t *out, int64_t cap);
int64_t mm2o_seed(const mm2o_index_t *ix, int max_occ, int qlen, int64_t n_mv, const mm2o_anchor_t *mv, mm2o_anchor_t *a, int64_t cap_a,
                  int32_t *rep_len, int32_t *n_mini_pos, uint64_t *mini_pos);
}

struct mm2b_index {
	int k, w;
	uint64_t *keys, *vals, *pos;
	mm2o_index_t *ix;
};

#include <vector>
struct SwResult {
	mm2b_map_result_t res;
	std::vector<int32_t> status, n_u, n_v, rep_len, n_mini_pos, n_mini, seg;
	std::vector<int64_t> n_a, u_off, b_off, mp_off;
	std::vector<uint64_t> u;
	std::vector<mm2b_anchor_t> b;
	std::vector<uint32_t> mp;
	uint64_t *seg_u;
	mm2b_anchor_t *seg_b;
	uint32_t *seg_mp;
};

extern "C" {

int mm2b_map_supported(int k, int w, int is_hpc, int n_segs, int64_t map_flag, int sdust_thres)
{
	const int64_t unsupported = 0x001 | 0x002 | 0x100000 | 0x200000 | 0x400000;
	return k >= 1 && k <= 28 && (k & 1) && w >= 1 && w <= 64 && !is_hpc && n_segs == 1 && !(map_flag & unsupported) && sdust_thres <= 0;
}

mm2b_index_t *mm2b_index_create(const mm2b_index_desc_t *d)
{
	mm2b_index_t *x = new mm2b_index_t();
	x->k = d->k, x->w = d->w;
	x->keys = (uint64_t*)malloc((size_t)(d->n_keys + 1) * 8), x->vals = (uint64_t*)malloc((size_t)(d->n_keys + 1) * 8), x->pos = (uint64_t*)malloc((size_t)(d->n_pos + 1) * 8);
	memcpy(x->keys, d->keys, (size_t)d->n_keys * 8), memcpy(x->vals, d->vals, (size_t)d->n_keys * 8), memcpy(x->pos, d->pos, (size_t)d->n_pos * 8);
	x->ix = mm2o_index_new(d->k, d->w, d->n_keys, x->keys, x->vals, x->pos);
	return x;
}

void mm2b_index_destroy(mm2b_index_t *x)
{
	if (!x) return;
	mm2o_index_free(x->ix);
	free(x->keys), free(x->vals), free(x->pos);
	delete x;
}

int mm2b_map_batch(mm2b_index_t *idx, const mm2b_seed_params_t *seed, const mm2b_params_t *par,
                   int64_t n_reads, const int64_t *seq_off, const char *seq, mm2b_map_result_t **out)
{
	SwResult *R = new SwResult();
	const size_t n = (size_t)n_reads;
	R->status.assign(n, 0), R->n_u.assign(n, 0), R->n_v.assign(n, 0), R->rep_len.assign(n, 0), R->n_mini_pos.assign(n, 0), R->n_mini.assign(n, 0), R->seg.assign(n, 0);
	R->n_a.assign(n, 0), R->u_off.assign(n, 0), R->b_off.assign(n, 0), R->mp_off.assign(n, 0);
	for (int64_t r = 0; r < n_reads; ++r) {
		const int len = (int)(seq_off[r + 1] - seq_off[r]);
		std::vector<mm2o_anchor_t> mv((size_t)len + 8);
		const int64_t n_mv = len > 0 ? mm2o_sketch(seq + seq_off[r], len, idx->w, idx->k, mv.data(), (int64_t)mv.size()) : 0;
		std::vector<uint64_t> mp((size_t)n_mv + 1);
		int64_t cap = 4096 + 64 * n_mv, n_a;
		mm2o_anchor_t *a = 0;
		int32_t rep = 0, nmp = 0;
		for (;;) {
			a = (mm2o_anchor_t*)malloc((size_t)cap * 16);
			n_a = mm2o_seed(idx->ix, seed->max_occ, len, n_mv, mv.data(), a, cap, &rep, &nmp, mp.data());
			if (n_a >= 0) break;
			free(a), cap *= 8;
		}
		R->n_mini[r] = (int32_t)n_mv, R->n_a[r] = n_a, R->rep_len[r] = rep, R->n_mini_pos[r] = nmp;
		R->mp_off[r] = (int64_t)R->mp.size();
		for (int32_t k = 0; k < nmp; ++k) R->mp.push_back((uint32_t)mp[(size_t)k]);
		int nu = 0;
		uint64_t *uu = 0;
		if (n_a == 0) free(a), a = 0;
		mm2b_anchor_t *bb = mm_chain_dp_ref(par->max_dist_x, par->max_dist_y, par->bw, par->max_skip, par->max_iter, par->min_cnt, par->min_sc, par->gap_scale,
		                                    par->is_cdna, par->n_segs, n_a, (mm2b_anchor_t*)a, &nu, &uu, 0, 0);      // consumes a
		R->status[r] = n_a == 0 ? MM2B_READ_EMPTY : (uu == 0 ? MM2B_READ_NO_CHAIN : MM2B_READ_OK);
		R->u_off[r] = (int64_t)R->u.size(), R->b_off[r] = (int64_t)R->b.size();
		int64_t nv = 0;
		for (int c = 0; c < nu; ++c) R->u.push_back(uu[c]), nv += (int32_t)uu[c];
		for (int64_t k = 0; k < nv; ++k) R->b.push_back(bb[k]);
		R->n_u[r] = nu, R->n_v[r] = (int32_t)nv;
		free(uu), free(bb);
	}
	mm2b_map_result_t &res = R->res;
	memset(&res, 0, sizeof(res));
	res.n_reads = n_reads;
	res.status = R->status.data(), res.n_u = R->n_u.data(), res.n_v = R->n_v.data(), res.rep_len = R->rep_len.data(), res.n_mini_pos = R->n_mini_pos.data();
	res.n_mini = R->n_mini.data(), res.seg = R->seg.data(), res.n_a = R->n_a.data(), res.u_off = R->u_off.data(), res.b_off = R->b_off.data(), res.mp_off = R->mp_off.data();
	R->seg_u = R->u.data(), R->seg_b = R->b.data(), R->seg_mp = R->mp.data();
	res.n_segs = 1, res.seg_u = &R->seg_u, res.seg_b = &R->seg_b, res.seg_mini_pos = &R->seg_mp;
	res.priv = R;
	*out = &R->res;
	return MM2B_OK;
}

void mm2b_map_result_release(mm2b_map_result_t *res)
{
	if (res) delete (SwResult*)res->priv;
}

}
