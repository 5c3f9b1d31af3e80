// The reference's index as flat arrays for the GPU: SURVEY.md 8(f) next-1, host side.
//
// mm_idx_t keeps its minimizer table hidden (minimap.h:71 `struct mm_idx_bucket_s *B; // index (hidden)`): 2^b buckets, each a khash from
// `minimizer >> b << 1 | single` to either one position or (first << 32 | n) into the bucket's position array (index.c:27-32, filled
// at index.c:205-238, read by mm_idx_get, index.c:81-98).  The bucket and hash types are private to index.c, so — like host/map_batch.cpp
// does with map.c — this file IS the translation unit of the reference's index.c: it includes index.c unchanged, from where it lies
// (-I /root/reference, nothing copied), and appends one function that walks the buckets.  A maintainer of the reference would add that
// function at the end of index.c.
#include "index.c"

#define MM2B_HOST_DECLARES_MM_CHAIN_DP          /* mmpriv.h:65 already declares it (with mm128_t) */
#include "mm2seed_b200.h"

// Flatten `mi` into malloc'd arrays (free with mm2b_index_flat_free).  Returns 0, or -1 when out of memory.
extern "C" int mm2b_index_flatten(const mm_idx_t *mi, mm2b_index_desc_t *out)
{
	const uint32_t n_buckets = 1u << mi->b;
	int64_t n_keys = 0, n_pos = 0;
	for (uint32_t i = 0; i < n_buckets; ++i) {
		const idxhash_t *h = (const idxhash_t*)mi->B[i].h;
		if (h) n_keys += kh_size(h);
		n_pos += mi->B[i].n;
	}
	uint64_t *keys = (uint64_t*)malloc((size_t)(n_keys ? n_keys : 1) * 8), *vals = (uint64_t*)malloc((size_t)(n_keys ? n_keys : 1) * 8);
	uint64_t *pos = (uint64_t*)malloc((size_t)(n_pos ? n_pos : 1) * 8);
	if (!keys || !vals || !pos) { free(keys), free(vals), free(pos); return -1; }
	int64_t nk = 0, np = 0;
	for (uint32_t i = 0; i < n_buckets; ++i) {
		const mm_idx_bucket_t *b = &mi->B[i];
		const idxhash_t *h = (const idxhash_t*)b->h;
		if (h)
			for (khint_t k = 0; k != kh_end(h); ++k) {
				if (!kh_exist(h, k)) continue;
				const uint64_t key = kh_key(h, k), val = kh_val(h, k);
				const uint64_t minimizer = key >> 1 << mi->b | i;                       // index.c:84-88 read backwards
				keys[nk] = minimizer << 1 | (key & 1);
				vals[nk] = (key & 1) ? val : ((val >> 32) + (uint64_t)np) << 32 | (uint32_t)val;   // index.c:90-96
				++nk;
			}
		if (b->n > 0) memcpy(pos + np, b->p, (size_t)b->n * 8);
		np += b->n;
	}
	out->k = mi->k, out->w = mi->w, out->is_hpc = !!(mi->flag & MM_I_HPC), out->n_seq = (int32_t)mi->n_seq;
	out->n_keys = nk, out->n_pos = np, out->keys = keys, out->vals = vals, out->pos = pos;
	return np < (1ll << 32) ? 0 : -1;          // `first` has 32 bits
}

extern "C" void mm2b_index_flat_free(mm2b_index_desc_t *d)
{
	free((void*)d->keys), free((void*)d->vals), free((void*)d->pos);
	d->keys = d->vals = d->pos = 0;
}
