// Internal declarations shared by the seeding kernels (seed_kernels.cu) and the host pipeline (host/map_backend.cpp).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "mm2seed_b200.h"

namespace mm2b {

constexpr int SKETCH_TILE = 512;        // positions per CTA of the one-position-per-thread sketch kernel (w < 8)
constexpr int SKETCH8_TILE = 1920;      // ... of the eight-positions-per-thread kernel (240 threads + 16 for the halo = 8 full warps)
constexpr int SKETCH_MAX_W = 64;        // window sizes the emission masks cover
constexpr int SKETCH_MAX_K = 28;        // mm_sketch's own limit (sketch.c:84)

// The index of one device: open-addressing table (linear probing) over the flat index of mm2b_index_desc_t.
struct DeviceIndex {
	int device;
	int k, w;
	int log2cap;                // table capacity = 1 << log2cap, load factor <= 0.5
	uint64_t *tab_keys;         // minimizer << 1 | single, or ~0 for an empty slot
	uint64_t *tab_vals;         // see mm2b_index_desc_t::vals
	uint64_t *pos;              // concatenated position lists
	int64_t n_keys, n_pos;
};

// One sub-batch of reads in HBM, stage by stage.
struct SeedArgs {
	int64_t n_reads;
	const uint8_t *seq;         // concatenated ASCII bases (16-byte aligned, readable for 16 bytes past seq_len)
	int64_t seq_len;
	const int64_t *seq_off;     // [n_reads + 1]
	const int32_t *tile_off;    // [n_reads + 1] first sketch tile of every read
	const int32_t *tile_read;   // [n_tiles] the read every tile belongs to
	int32_t n_tiles;
	int k, w, max_occ;
	// sketch
	unsigned long long *tile_state; // [n_tiles] look-back words: flag << 62 | count
	int *tile_ticket;           // tiles are taken in ticket order
	int64_t *tile_excl;         // [n_tiles + 1] minimizers pushed by the tiles before each tile; entry n_tiles = the total
	ulonglong2 *mv;             // minimizers: x = hash << 8 | span, y = pos << 1 | strand
	int64_t mv_cap;             // capacity of mv: beyond it the kernel only counts
	int64_t *mv_off;            // [n_reads + 1]
	// lookup + matches
	int32_t *occ;               // per minimizer: occurrences in the index (mm_idx_get's *n)
	uint64_t *hv;               // per minimizer: the table's value word
	int32_t *arel;              // per minimizer: first anchor of this minimizer inside its read (matches only)
	uint32_t *mini_pos;         // per read at mv_off[r]: query positions of the matches (map.c:117)
	int32_t *rep_len, *n_mini_pos;
	int64_t *n_a;               // per read
	const int64_t *a_off;       // [n_reads + 1] exclusive prefix of n_a
	ulonglong2 *a, *a_tmp;      // anchors and the second buffer of their sort
	int32_t *tie_list;          // reads whose sorted anchors hold equal keys
	int *tie_count;             // [0] length of tie_list, [1] work-queue cursor of the sort, [2] cursor of the replay
};

int launch_index_build(const DeviceIndex &ix, const uint64_t *d_keys, const uint64_t *d_vals, cudaStream_t stream);
int launch_index_lookup(const DeviceIndex &ix, int64_t n, const ulonglong2 *mv, const uint64_t *raw_minimizers, int32_t *occ, uint64_t *hv, cudaStream_t stream);
int sketch_tile_positions(int w);        // positions per tile of the kernel launch_sketch will use for this window size
int launch_sketch(const SeedArgs &s, cudaStream_t stream);
int launch_scan_i64(const int64_t *in, int64_t *out, int64_t n, cudaStream_t stream);      // out[0..n]: exclusive prefix, out[n] = total
int launch_export_scalars(int64_t *host_mapped, const int64_t *a, const int64_t *b, const int *c, cudaStream_t stream);   // host_mapped[0..2] = *a, *b, *c (null: left alone)
int launch_read_offsets(const SeedArgs &s, cudaStream_t stream);                            // mv_off[r] = tile_excl[tile_off[r]]
int launch_matches(const SeedArgs &s, int n_sms, cudaStream_t stream);                      // collect_matches per read
int launch_expand(const SeedArgs &s, const DeviceIndex &ix, int64_t n_mv, cudaStream_t stream);      // collect_seed_hits into a_tmp, one thread per minimizer
int launch_sort(const SeedArgs &s, const DeviceIndex &ix, int n_sms, cudaStream_t stream);  // stable sort + exact replay of the reads with equal keys

}  // namespace mm2b
