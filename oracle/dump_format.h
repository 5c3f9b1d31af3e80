/* TEST INFRASTRUCTURE — on-disk record format for captured mm_chain_dp calls.
 *
 * One record per call of the reference's mm_chain_dp (chain.c:29), little-endian, records
 * concatenated.  Used by oracle/dump_shim.c (writer, wraps the *reference* function),
 * tests/ (reader: tests/dumpio.py) and tests/golden/ fixtures.
 *
 *   mm2_dump_hdr_t           64 bytes
 *   mm128  a[n]              input anchors exactly as passed in (chain.c:29 `a`)
 *   int32  f[n], p[n], v[n]  only if MM2_DUMP_HAS_FPV: state after the DP fill (chain.c:238)
 *   uint64 u[n_u]            final chain list (score<<32 | n_anchors), order of chain.c:419
 *   mm128  b[n_v]            final chained anchors, order of chain.c:420
 */
#ifndef MM2_DUMP_FORMAT_H
#define MM2_DUMP_FORMAT_H
#include <stdint.h>

#define MM2_DUMP_MAGIC 0x4443324du /* "M2CD" */
#define MM2_DUMP_HAS_FPV 1u
#define MM2_DUMP_B_NULL  2u  /* mm_chain_dp returned NULL */
#define MM2_DUMP_U_NULL  4u  /* *_u was NULL on return */

typedef struct {
	uint32_t magic, flags;
	int32_t max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc, is_cdna, n_segs;
	float gap_scale;
	int64_t n;
	int32_t n_u, n_v;
} mm2_dump_hdr_t;

#endif
