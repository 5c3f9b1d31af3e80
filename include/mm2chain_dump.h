/* mm2chain_dump — wire format of captured chaining calls ("anchor dumps"), the input of the batched replay caller
 * (minimap2-fpga_b200/host/replay_main.cpp, binary mm2b-replay).
 *
 * A dump is what the seeding half of mm_map_frag (map.c:272-316) hands to mm_chain_dp, recorded once per call: the chaining
 * arguments (chain.c:29), the sorted anchors (map.c:245) and, optionally, what the reference returned for them.  Replaying a
 * dump feeds the chaining backend whole mini-batches from pinned host memory without re-running the seeding (SURVEY.md §8d
 * config 5, §8f next-3).  Little-endian, records concatenated, optionally gzip-compressed as a whole:
 *
 *   mm2b_dump_hdr_t          64 bytes
 *   mm2b_anchor_t a[n]       input anchors exactly as passed in
 *   int32 f[n], p[n], v[n]   only if MM2B_DUMP_HAS_FPV: DP state after the fill (chain.c:238); skipped by the replay
 *   uint64 u[n_u]            the reference's final chain list (score<<32 | n_anchors), order of chain.c:419
 *   mm2b_anchor_t b[n_v]     the reference's final chained anchors, order of chain.c:420
 *
 * n_u = n_v = 0 with MM2B_DUMP_U_NULL set is the "no chain" return (chain.c:355-358).
 */
#ifndef MM2CHAIN_DUMP_H
#define MM2CHAIN_DUMP_H

#include <stdint.h>

#define MM2B_DUMP_MAGIC   0x4443324du /* "M2CD" */
#define MM2B_DUMP_HAS_FPV 1u
#define MM2B_DUMP_B_NULL  2u          /* mm_chain_dp returned NULL */
#define MM2B_DUMP_U_NULL  4u          /* *_u was NULL on return */

typedef struct {
	uint32_t magic, flags;
	int32_t max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc, is_cdna, n_segs;
	float gap_scale;
	int64_t n;                        /* anchors in this call */
	int32_t n_u, n_v;                 /* recorded result sizes */
} mm2b_dump_hdr_t;

#endif
