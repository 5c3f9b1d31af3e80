// Internal declarations shared by the CUDA kernels (chain_kernels.cu) and the C-ABI shim (chain_api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "mm2chain_b200.h"

namespace mm2b {

// Per-read scratch layout in HBM.  Read r with n anchors starting at global anchor offset o owns the
// SCRATCH_BYTES_PER_ANCHOR*n bytes at scratch + SCRATCH_BYTES_PER_ANCHOR*o, carved as
//   [F 4n][P 4n][X 8n][V 4n][T 4n][U 8n]
//   F,P,V  DP state (chain.c:236-237);  T visit stamps / marks;  U chain-end keys, then kept chains
//   X      second buffer of the radix sort of U;  W (16 B per chain, final order sort) aliases F+P+X once the backtrack is done
//   V -> PATH (backtrack index list);  T -> worklist of the final sort when a read has more than 64 chains
constexpr int SCRATCH_BYTES_PER_ANCHOR = 32;

struct BatchArgs {
	mm2b_params_t par;
	int64_t n_reads;
	const int64_t *off;
	const mm2b_anchor_t *a;
	uint8_t *scratch;
	int32_t *n_u, *n_v, *status;
	// packed output, filled by each read's own warp (no separate scan / gather pass): the warp reserves its share from the two
	// cursors and records where it wrote in u_off[r] / b_off[r]
	unsigned long long *out_cursor; // [0] entries of u handed out, [1] entries of b / bi handed out
	int64_t *u_off, *b_off;
	uint64_t *u;
	mm2b_anchor_t *b;               // chained anchors (16 B each), or nullptr when ...
	int32_t *bi;                    // ... their indices inside the read (4 B each) are wanted instead
	const int32_t *order;           // processing order (longest reads first) or nullptr
	int *work_counter;              // persistent-warp work queue
	unsigned long long *counters;   // [0] chunks issued, [1] reads on the general path, [2] reference-semantics cells, [3] window cells, [4] reads taken by the heavy-read kernel, [5] reads cut into pieces
	int32_t *dbg_fpv;               // optional 3 x n_anchors int32 (f, p, v) copy for tests, or nullptr
	int64_t n_anchors;
	int count_cells;                // tally reference-semantics cells / issued chunks (statistics only; costs kernel time)
	// heavy-read kernel (one CTA per read with long windows); all null / 0 when it is not used for this batch
	uint8_t *heavy_flag;            // per read: taken by the heavy-read kernel
	int32_t *heavy_list;            // the reads it takes
	int *heavy_count, *heavy_counter;   // adjacent ints: length of heavy_list, its work-queue cursor
	long long heavy_min_cells;      // a read is heavy when its estimated window cells reach this (and its mean window is long)
	int heavy_cap;                  // at most this many reads per batch (one wave of CTAs): the kernel buys latency, not throughput
	// long reads cut at their x-gap cut points (chain.c:192) into pieces that are filled by warps of their own; all null / 0 when off.
	// A read that is cut gets heavy_flag = 2, so the warp-per-read kernel leaves it alone.
	struct SegRead *seg_reads;      // the reads that are cut (at most seg_cap)
	int32_t *seg_items;             // work items: index into seg_reads << 8 | piece
	int *seg_ctl;                   // [0] reads cut, [1] work items, [2] work-queue cursor
	int32_t *seg_done;              // per cut read: pieces filled so far
	int seg_cap, seg_min_read, seg_min_piece;
};

constexpr int SEG_MAX_PIECES = 16;
struct SegRead {
	int32_t read, n_pieces;
	float avg;                      // avg_qspan_scaled of the whole read (chain.c:49)
	int32_t flags;                  // 1: low words suffice for the window search, 2: general scoring path
	int32_t start[SEG_MAX_PIECES + 1];
};

// launchers (all asynchronous on `stream`); each returns the number of kernels it launched
int launch_order(int64_t n_reads, const int64_t *off, int32_t *order, int *bucket_scratch, cudaStream_t stream);
int launch_chain(const BatchArgs &args, int n_sms, cudaStream_t stream);
// anchors packed to 8 B by the host (low words + runs of equal high words, see host/chain_backend.cpp) -> mm128_t in HBM
int launch_unpack(int64_t n_anchors, const uint2 *lo, const uint2 *xruns, int n_xruns, const uint2 *yruns, int n_yruns, mm2b_anchor_t *a,
                  int n_sms, cudaStream_t stream);
double measure_int32_peak(int device);
int heavy_ring_slots();          // ring capacity of the heavy-read kernel: it needs max_iter + 64 <= this
int heavy_min_window();          // ... and only pays off for windows longer than this
unsigned debug_flags();          // range-check violations seen so far (0x80000000 | codes) in the -DMM2B_DEBUG_CHECKS build, else 0

}  // namespace mm2b
