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

namespace {
struct FlattenJob {
	const mm_idx_t *mi;
	const int64_t *key_off, *pos_off;       // per bucket: where its keys / positions start in the flat arrays
	uint64_t *keys, *vals, *pos;
};

void flatten_bucket(void *data, long i, int tid)
{
	(void)tid;
	const FlattenJob &j = *(const FlattenJob*)data;
	const mm_idx_t *mi = j.mi;
	const mm_idx_bucket_t *b = &mi->B[i];
	const idxhash_t *h = (const idxhash_t*)b->h;
	int64_t nk = j.key_off[i];
	const uint64_t np = (uint64_t)j.pos_off[i];
	if (h)
		for (khint_t k = 0; k != kh_end(h); ++k) {
			if (!kh_exist(h, k)) continue;
			const uint64_t key = kh_key(h, k), val = kh_val(h, k);
			const uint64_t minimizer = key >> 1 << mi->b | (uint64_t)i;                     // index.c:84-88 read backwards
			j.keys[nk] = minimizer << 1 | (key & 1);
			j.vals[nk] = (key & 1) ? val : ((val >> 32) + np) << 32 | (uint32_t)val;         // index.c:90-96
			++nk;
		}
	if (b->n > 0) memcpy(j.pos + np, b->p, (size_t)b->n * 8);
}
}  // namespace

// Flatten `mi` into malloc'd arrays (free with mm2b_index_flat_free), buckets in parallel on n_threads threads (kt_for, kthread.c:54).
// Returns 0, or -1 when out of memory or when the position lists do not fit the 32-bit `first` field.
extern "C" int mm2b_index_flatten_mt(const mm_idx_t *mi, mm2b_index_desc_t *out, int n_threads)
{
	const uint32_t n_buckets = 1u << mi->b;
	int64_t *key_off = (int64_t*)malloc(((size_t)n_buckets + 1) * 8), *pos_off = (int64_t*)malloc(((size_t)n_buckets + 1) * 8);
	if (!key_off || !pos_off) { free(key_off), free(pos_off); return -1; }
	int64_t n_keys = 0, n_pos = 0;
	for (uint32_t i = 0; i < n_buckets; ++i) {
		const idxhash_t *h = (const idxhash_t*)mi->B[i].h;
		key_off[i] = n_keys, pos_off[i] = n_pos;
		if (h) n_keys += kh_size(h);
		n_pos += mi->B[i].n;
	}
	uint64_t *keys = (uint64_t*)malloc((size_t)(n_keys ? n_keys : 1) * 8), *vals = (uint64_t*)malloc((size_t)(n_keys ? n_keys : 1) * 8);
	uint64_t *pos = (uint64_t*)malloc((size_t)(n_pos ? n_pos : 1) * 8);
	if (!keys || !vals || !pos || n_pos >= (1ll << 32)) { free(keys), free(vals), free(pos), free(key_off), free(pos_off); return -1; }
	FlattenJob job = {mi, key_off, pos_off, keys, vals, pos};
	kt_for(n_threads < 1 ? 1 : n_threads, flatten_bucket, &job, (long)n_buckets);
	free(key_off), free(pos_off);
	out->k = mi->k, out->w = mi->w, out->is_hpc = !!(mi->flag & MM_I_HPC), out->n_seq = (int32_t)mi->n_seq;
	out->n_keys = n_keys, out->n_pos = n_pos, out->keys = keys, out->vals = vals, out->pos = pos;
	return 0;
}

extern "C" int mm2b_index_flatten(const mm_idx_t *mi, mm2b_index_desc_t *out) { return mm2b_index_flatten_mt(mi, out, 1); }

extern "C" void mm2b_index_flat_free(mm2b_index_desc_t *d)
{
	free((void*)d->keys), free((void*)d->vals), free((void*)d->pos);
	d->keys = d->vals = d->pos = 0;
}
