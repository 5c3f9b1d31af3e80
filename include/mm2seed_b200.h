/* mm2seed_b200 — C ABI of the B200-native seeding front end: SURVEY.md 8(f) next-4 and next-1 in front of the chaining path.
 * Read sequences go in; minimizers, seed hits and the anchor sort run on the GPU, the anchors are born in HBM and go straight
 * into the chaining kernels of mm2chain_b200.h; final chains come back.  Plain C: pointers and sizes only.
 *
 * What each piece replaces in the reference (/root/reference):
 *
 *   mm2b_index_flatten / mm2b_index_create   mm_idx_t's hidden buckets (index.c:27-32, khash tables filled at index.c:205-238) and
 *                                            mm_idx_get (index.c:81-98): one open-addressing table in HBM per bound device
 *   sketch kernels                           mm_sketch (sketch.c:77-143) as called by collect_minimizers (map.c:64-78), non-HPC
 *   seed kernels                             collect_matches (map.c:90-123) and collect_seed_hits (map.c:215-247) including
 *                                            radix_sort_128x's order of equal keys (map.c:245, ksort.h:116-151)
 *   mm2b_map_batch                           the first half of mm_map_frag (map.c:287-316) for a whole mini-batch: sequences in,
 *                                            what mm_chain_dp returns (u[], b[]) plus rep_len and mini_pos out
 *
 * Scope of the device path (anything else is the caller's business — the phase-split caller in host/map_batch.cpp keeps the
 * reference's own per-read code for it): one segment per read (no paired reads), no homopolymer compression, odd k <= 28
 * (no strand-symmetric k-mers), w <= 64, no SDUST masking, none of MM_F_NO_DIAG / MM_F_NO_DUAL / MM_F_FOR_ONLY / MM_F_REV_ONLY /
 * MM_F_HEAP_SORT.  mm2b_map_supported() answers for a given configuration.
 */
#ifndef MM2SEED_B200_H
#define MM2SEED_B200_H

#include "mm2chain_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* The index as flat arrays (host memory).  keys[i] = minimizer << 1 | single, where `minimizer` is the full hash value
 * mm_sketch stores in mm128_t::x >> 8 (the reference splits it into a bucket id and a khash key, index.c:84-88);
 * vals[i] = the position word itself (rid << 32 | pos << 1 | strand) when `single`, else first << 32 | n: the n positions
 * pos[first .. first + n), sorted as the reference keeps them (index.c:230). */
typedef struct {
	int32_t k, w, is_hpc, n_seq;
	int64_t n_keys, n_pos;
	const uint64_t *keys, *vals, *pos;
} mm2b_index_desc_t;

typedef struct mm2b_index mm2b_index_t;

/* Copy the flat index to every device bound by mm2b_init and build the device hash tables.  NULL on failure (mm2b_last_error). */
mm2b_index_t *mm2b_index_create(const mm2b_index_desc_t *desc);
void mm2b_index_destroy(mm2b_index_t *idx);
/* mm_idx_get (index.c:81): occurrences of `minimizer`; *val receives the table's value word (see mm2b_index_desc_t).  Test access;
 * synchronous, one lookup kernel on the first bound device. */
int mm2b_index_lookup(mm2b_index_t *idx, int64_t n, const uint64_t *minimizers, int32_t *n_occ, uint64_t *val);

/* The seeding arguments of mm_map_frag that the device path honours. */
typedef struct {
	int32_t max_occ;        /* mm_mapopt_t::mid_occ as passed to collect_seed_hits (map.c:296): minimizers with >= max_occ hits are skipped */
	int32_t reserved;
} mm2b_seed_params_t;

/* 1 when (k, w, is_hpc, n_segs, map flags, sdust threshold) is a configuration the device path reproduces, else 0. */
int mm2b_map_supported(int k, int w, int is_hpc, int n_segs, int64_t map_flag, int sdust_thres);

/* Results of one mm2b_map_batch call.  How many anchors a read seeds is only known on the device, so the variable-length outputs
 * come back in SEGMENTS — one per sub-batch of reads, each a set of pinned host buffers owned by the library (pooled across
 * calls) — and every read says which segment holds its results.  Valid until mm2b_map_result_release.  Per read r, with s = seg[r]:
 *   status[r], n_u[r], n_v[r]                  as mm_chain_dp returns them for the read's anchors (mm2chain_b200.h)
 *   seg_u[s][u_off[r] .. + n_u[r])             == the reference's final u[] for the read   (chain.c:419)
 *   seg_b[s][b_off[r] .. + n_v[r])             == the reference's final b[] for the read   (chain.c:420)
 *   n_a[r]                                     anchors collect_seed_hits produced (map.c:246)
 *   rep_len[r]                                 collect_matches' *rep_len (map.c:104-120)
 *   n_mini_pos[r], seg_mini_pos[s][mp_off[r] ..)  collect_matches' mini_pos (map.c:117) as 32-bit query positions; the reference's
 *                                              64-bit entries are q_span << 32 | pos, and q_span == k on this path
 *   n_mini[r]                                  minimizers mm_sketch produced for the read */
typedef struct {
	int64_t n_reads;
	int32_t *status, *n_u, *n_v, *rep_len, *n_mini_pos, *n_mini, *seg;
	int64_t *n_a, *u_off, *b_off, *mp_off;
	int32_t n_segs;
	uint64_t **seg_u;
	mm2b_anchor_t **seg_b;
	uint32_t **seg_mini_pos;
	/* totals and timings of the call */
	int64_t tot_mini, tot_anchors, tot_chains, tot_chained, n_tie_reads;
	int64_t h2d_bytes, d2h_bytes;
	double sketch_ms, seed_ms, sort_ms, chain_ms;      /* CUDA-event time summed over sub-batches */
	int64_t cells_ref;                                  /* reference-semantics DP cells when counting is on (mm2b_set_counting), else 0 */
	void *priv;
} mm2b_map_result_t;

/* Sketch, seed, sort and chain `n_reads` reads.  Read r is seq[seq_off[r] .. seq_off[r+1]) as the ASCII bases bseq.c hands to
 * mm_map_frag (any byte that is not ACGTacgt counts as ambiguous, sketch.c:9-26).  Reads are sharded over the bound devices in
 * sub-batches; results come back in input order.  Blocking; thread-safe.  Returns MM2B_OK or an error code; *out is NULL on error.
 * Environment overrides (tuning / diagnosis): MM2B_MAP_CTX=n pipeline contexts per device (default 6), MM2B_MAP_SUB_BYTES=n bytes of
 * sequence per sub-batch (default 64 MiB), MM2B_MAP_RAMP=0 equal sub-batches from the start, MM2B_MAP_TRACE=1 per-sub-batch timeline on
 * stderr, MM2B_MAP_ONE_STREAM=1 copies and kernels of a context on one stream (slower: kept for comparison), MM2B_SKETCH8=0 the
 * one-position-per-thread sketch kernel for every window size, MM2B_SKETCH8_CTAS=3 its 80-register build. */
int mm2b_map_batch(mm2b_index_t *idx, const mm2b_seed_params_t *seed, const mm2b_params_t *chain,
                   int64_t n_reads, const int64_t *seq_off, const char *seq, mm2b_map_result_t **out);
void mm2b_map_result_release(mm2b_map_result_t *res);

/* Test / oracle access to the intermediate products of one small batch (synchronous, first bound device):
 *   mini      minimizers as mm_sketch emits them (x = hash << 8 | span, y = pos << 1 | strand), mini_off[n_reads + 1]
 *   anchors   the sorted anchors collect_seed_hits returns, a_off[n_reads + 1]
 * Buffers are malloc'd by the library; free them with mm2b_free. */
int mm2b_seed_debug(mm2b_index_t *idx, const mm2b_seed_params_t *seed, int64_t n_reads, const int64_t *seq_off, const char *seq,
                    int64_t **mini_off, mm2b_anchor_t **mini, int64_t **a_off, mm2b_anchor_t **anchors,
                    int32_t **rep_len, int32_t **n_mini_pos, uint32_t **mini_pos, int64_t *n_tie_reads);
void mm2b_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
