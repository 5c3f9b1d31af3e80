// C-ABI shim over the CUDA kernels: workspaces and the device-buffer batch call (include/mm2chain_b200.h).
// The host-buffer batch call, the device worker threads and the mm_chain_dp drop-in live in host/chain_backend.cpp.
#include "chain_kernels.cuh"
#include "shim_internal.h"
#include <atomic>
#include <mutex>
#include <unordered_map>
#include <utility>
#include <vector>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace mm2b {

static thread_local char g_err[512];
static std::atomic<int64_t> g_launches{0};
static std::mutex g_pin_mu;
static std::vector<std::pair<size_t, void*>> g_pin_free;        // pooled pinned blocks: (bytes, pointer)
static std::unordered_map<void*, size_t> g_pin_size;            // every live block handed out by mm2b_host_alloc
static size_t g_pin_pooled = 0, g_pin_pool_max = (size_t)16 << 30;
static std::atomic<int64_t> g_pin_misses{0};                    // requests the pool could not serve (each one a cudaHostAlloc)

void set_error(const char *fmt, const char *a, const char *b)
{
	snprintf(g_err, sizeof(g_err), fmt, a ? a : "", b ? b : "");
}
bool cuda_ok(cudaError_t e, const char *what)
{
	if (e == cudaSuccess) return true;
	set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
	return false;
}
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long pin_pool_misses() { return (long)g_pin_misses.load(std::memory_order_relaxed); }

}  // namespace mm2b

using namespace mm2b;

static constexpr int SEG_CAP = 16384;       // reads cut into pieces per batch at most; further long reads stay whole

struct mm2b_workspace {
	int device, n_sms;
	int64_t max_anchors, max_reads;
	uint8_t *scratch;           // SCRATCH_BYTES_PER_ANCHOR * max_anchors
	int32_t *order;             // max_reads
	int32_t *heavy_list;        // max_reads: reads routed to the heavy-read kernel
	uint8_t *heavy_flag;        // max_reads
	int heavy_on;               // MM2B_HEAVY (default 1)
	long long heavy_min_cells;  // MM2B_HEAVY_MIN_CELLS: estimated window cells from which a read counts as heavy
	int64_t longest_hint;       // mm2b_ws_set_longest_read: longest read of the next batch, or -1
	int seg_on, seg_min_read, seg_min_piece;    // MM2B_SEG (default 1), MM2B_SEG_MIN_READ, MM2B_SEG_MIN_PIECE: long-read segmenting
	SegRead *seg_reads;         // SEG_CAP reads cut per batch at most
	int32_t *seg_items, *seg_done;
	int *small;                 // [0] work counter, [2] heavy-read count, [3] heavy-read cursor, [64..320) length buckets
	unsigned long long *counters;   // [0..5) statistics, [6..8) output cursors
	int32_t *dbg_fpv;           // 3 * max_anchors when MM2B_KEEP_FPV=1
	size_t bytes;
	cudaEvent_t ev_k1[2];       // around the chaining kernel of the last batch
	int count_cells;
	int64_t last_reads, last_anchors;
	const int32_t *last_n_u, *last_n_v;
	const int64_t *last_u_off, *last_b_off;
};

extern "C" {

const char *mm2b_last_error(void) { return g_err; }
int mm2b_abi_version(void) { return MM2B_ABI_VERSION; }
int64_t mm2b_launch_count(void) { return g_launches.load(); }

int mm2b_cuda_device_count(void)
{
	int n = 0;
	if (!cuda_ok(cudaGetDeviceCount(&n), "cudaGetDeviceCount")) return -1;
	return n;
}

// Pinned host memory comes from a small pool: cudaHostAlloc costs about half a millisecond per megabyte on the hosts measured, which
// is more than the whole GPU side of a mini-batch, so blocks are kept when they are freed and handed out again (best fit, at most
// twice the size asked for), and mm2b_host_reserve can fill the pool ahead of time on a background thread.
void *mm2b_host_alloc(size_t bytes)
{
	if (bytes == 0) bytes = 1;
	{
		std::lock_guard<std::mutex> lk(g_pin_mu);
		size_t best = (size_t)-1;
		for (size_t i = 0; i < g_pin_free.size(); ++i)
			if (g_pin_free[i].first >= bytes && g_pin_free[i].first <= 2 * bytes + (1u << 20) && (best == (size_t)-1 || g_pin_free[i].first < g_pin_free[best].first)) best = i;
		if (best != (size_t)-1) {
			void *p = g_pin_free[best].second;
			g_pin_pooled -= g_pin_free[best].first;
			g_pin_free.erase(g_pin_free.begin() + (long)best);
			return p;
		}
	}
	void *p = 0;
	g_pin_misses.fetch_add(1, std::memory_order_relaxed);
	if (!cuda_ok(cudaHostAlloc(&p, bytes, cudaHostAllocPortable), "cudaHostAlloc")) return 0;
	std::lock_guard<std::mutex> lk(g_pin_mu);
	g_pin_size[p] = bytes;
	return p;
}
void mm2b_host_free(void *p)
{
	if (!p) return;
	{
		std::lock_guard<std::mutex> lk(g_pin_mu);
		auto it = g_pin_size.find(p);
		if (it != g_pin_size.end() && g_pin_pooled + it->second <= g_pin_pool_max) {
			g_pin_free.push_back(std::make_pair(it->second, p));
			g_pin_pooled += it->second;
			return;
		}
		if (it != g_pin_size.end()) g_pin_size.erase(it);
	}
	cudaFreeHost(p);
}
void mm2b_host_reserve(size_t bytes, int n_blocks)
{
	std::vector<void*> got;
	for (int i = 0; i < n_blocks; ++i) {
		void *p = 0;
		if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) break;
		{ std::lock_guard<std::mutex> lk(g_pin_mu); g_pin_size[p] = bytes ? bytes : 1; }
		got.push_back(p);
	}
	for (void *p : got) mm2b_host_free(p);
}
void mm2b_host_pool_trim(void)
{
	std::vector<void*> blocks;
	{
		std::lock_guard<std::mutex> lk(g_pin_mu);
		for (auto &b : g_pin_free) blocks.push_back(b.second), g_pin_size.erase(b.second);
		g_pin_free.clear(), g_pin_pooled = 0;
	}
	for (void *p : blocks) cudaFreeHost(p);
}

mm2b_workspace_t *mm2b_ws_create(int device, int64_t max_anchors, int64_t max_reads)
{
	if (max_anchors < 1) max_anchors = 1;
	if (max_reads < 1) max_reads = 1;
	if (!cuda_ok(cudaSetDevice(device), "cudaSetDevice")) return 0;
	mm2b_workspace_t *ws = (mm2b_workspace_t*)calloc(1, sizeof(*ws));
	ws->device = device, ws->max_anchors = max_anchors, ws->max_reads = max_reads, ws->longest_hint = -1;
	{ const char *e = getenv("MM2B_COUNT_CELLS"); ws->count_cells = e && atoi(e) > 0; }
	{ const char *e = getenv("MM2B_HEAVY"); ws->heavy_on = e ? atoi(e) != 0 : 1; }
	{ const char *e = getenv("MM2B_HEAVY_MIN_CELLS"); ws->heavy_min_cells = e ? atoll(e) : 16ll << 20; }
	{ const char *e = getenv("MM2B_SEG"); ws->seg_on = e ? atoi(e) != 0 : 1; }
	{ const char *e = getenv("MM2B_SEG_MIN_READ"); ws->seg_min_read = e && atoi(e) > 0 ? atoi(e) : 8192; }
	{ const char *e = getenv("MM2B_SEG_MIN_PIECE"); ws->seg_min_piece = e && atoi(e) > 0 ? atoi(e) : 2048; }
	cudaDeviceGetAttribute(&ws->n_sms, cudaDevAttrMultiProcessorCount, device);
	const size_t sz_scratch = (size_t)max_anchors * SCRATCH_BYTES_PER_ANCHOR;
	const size_t sz_order = (size_t)max_reads * sizeof(int32_t);
	const char *keep = getenv("MM2B_KEEP_FPV");
	bool ok = cuda_ok(cudaMalloc(&ws->scratch, sz_scratch), "cudaMalloc(scratch)")
	       && cuda_ok(cudaMalloc(&ws->order, sz_order), "cudaMalloc(order)")
	       && cuda_ok(cudaMalloc(&ws->heavy_list, sz_order), "cudaMalloc(heavy_list)")
	       && cuda_ok(cudaMalloc(&ws->heavy_flag, (size_t)max_reads), "cudaMalloc(heavy_flag)")
	       && cuda_ok(cudaMalloc(&ws->small, 512 * sizeof(int)), "cudaMalloc(small)")
	       && cuda_ok(cudaMalloc(&ws->counters, 8 * sizeof(unsigned long long)), "cudaMalloc(counters)")
	       && cuda_ok(cudaMalloc(&ws->seg_reads, SEG_CAP * sizeof(SegRead)), "cudaMalloc(seg_reads)")
	       && cuda_ok(cudaMalloc(&ws->seg_items, SEG_CAP * SEG_MAX_PIECES * sizeof(int32_t)), "cudaMalloc(seg_items)")
	       && cuda_ok(cudaMalloc(&ws->seg_done, SEG_CAP * sizeof(int32_t)), "cudaMalloc(seg_done)");
	ws->bytes = sz_scratch + 2 * sz_order + (size_t)max_reads + 512 * sizeof(int) + 64;
	if (ok && keep && atoi(keep) > 0) {
		ok = cuda_ok(cudaMalloc(&ws->dbg_fpv, (size_t)max_anchors * 12), "cudaMalloc(dbg_fpv)");
		ws->bytes += (size_t)max_anchors * 12;
	}
	if (ok) ok = cuda_ok(cudaMemset(ws->counters, 0, 8 * sizeof(unsigned long long)), "cudaMemset");
	if (ok) ok = cuda_ok(cudaEventCreate(&ws->ev_k1[0]), "cudaEventCreate") && cuda_ok(cudaEventCreate(&ws->ev_k1[1]), "cudaEventCreate");
	if (!ok) { mm2b_ws_destroy(ws); return 0; }
	return ws;
}

void mm2b_ws_destroy(mm2b_workspace_t *ws)
{
	if (!ws) return;
	cudaSetDevice(ws->device);
	if (ws->ev_k1[0]) cudaEventDestroy(ws->ev_k1[0]);
	if (ws->ev_k1[1]) cudaEventDestroy(ws->ev_k1[1]);
	cudaFree(ws->scratch), cudaFree(ws->order), cudaFree(ws->heavy_list), cudaFree(ws->heavy_flag), cudaFree(ws->small), cudaFree(ws->counters), cudaFree(ws->dbg_fpv);
	cudaFree(ws->seg_reads), cudaFree(ws->seg_items), cudaFree(ws->seg_done);
	free(ws);
}

size_t mm2b_ws_bytes(const mm2b_workspace_t *ws) { return ws ? ws->bytes : 0; }
void mm2b_ws_set_counting(mm2b_workspace_t *ws, int on) { if (ws) ws->count_cells = on != 0; }
void mm2b_ws_set_longest_read(mm2b_workspace_t *ws, int64_t n_anchors) { if (ws) ws->longest_hint = n_anchors; }
const unsigned long long *mm2b_ws_counters_dev(const mm2b_workspace_t *ws) { return ws ? ws->counters : 0; }

static int chain_device(mm2b_workspace_t *ws, const mm2b_params_t *par, int64_t n_reads, int64_t n_anchors,
                        const int64_t *d_off, const mm2b_anchor_t *d_a,
                        int32_t *d_n_u, int32_t *d_n_v, int32_t *d_status, int64_t *d_u_off, int64_t *d_b_off,
                        uint64_t *d_u, mm2b_anchor_t *d_b, int32_t *d_bi, void *stream_)
{
	if (!ws || !par || n_reads < 0 || n_anchors < 0) { set_error("%s%s", "mm2b_chain_batch_device: bad argument", ""); return MM2B_ERR_ARG; }
	if (n_reads > ws->max_reads || n_anchors > ws->max_anchors || n_reads >= (1ll << 31)) {
		set_error("%s%s", "mm2b_chain_batch_device: batch exceeds workspace capacity", "");
		return MM2B_ERR_CAPACITY;
	}
	if (n_anchors > 0 && !d_b && !d_bi) { set_error("%s%s", "mm2b_chain_batch_device: no output buffer for the chained anchors", ""); return MM2B_ERR_ARG; }
	cudaStream_t stream = (cudaStream_t)stream_;
	int prev = -1;
	cudaGetDevice(&prev);
	if (prev != ws->device && !cuda_ok(cudaSetDevice(ws->device), "cudaSetDevice")) return MM2B_ERR_CUDA;
	int launches = 0;
	launches += launch_order(n_reads, d_off, ws->order, ws->small + 64, stream);
	BatchArgs ba;
	memset(&ba, 0, sizeof(ba));
	ba.par = *par, ba.n_reads = n_reads, ba.off = d_off, ba.a = d_a, ba.scratch = ws->scratch;
	ba.n_u = d_n_u, ba.n_v = d_n_v, ba.status = d_status, ba.order = ws->order, ba.work_counter = ws->small;
	ba.out_cursor = ws->counters + 6, ba.u_off = d_u_off, ba.b_off = d_b_off, ba.u = d_u, ba.b = d_bi ? nullptr : d_b, ba.bi = d_bi;
	ba.counters = ws->counters, ba.dbg_fpv = ws->dbg_fpv, ba.n_anchors = ws->max_anchors, ba.count_cells = ws->count_cells;
	// Reads with long windows (tandem repeats) go to the heavy-read kernel when the batch's arguments allow its scoring path
	// (same-segment genomic cost) and its ring holds a whole window; the cell tally is a warp-per-read feature.
	// a window has at most max_iter cells: a batch whose longest read has fewer than heavy_min_cells / max_iter anchors has no heavy read
	const bool may_have_heavy = ws->longest_hint < 0 || (ws->longest_hint >= 64 && ws->longest_hint * (int64_t)par->max_iter >= ws->heavy_min_cells);
	// Long reads with usable x-gap cut points are cut into pieces (long-read segmenting, chain_kernels.cu); like the heavy-read kernel
	// this is off while cells are tallied or f/p/v are kept for tests.
	const bool may_have_long = ws->longest_hint < 0 || ws->longest_hint >= ws->seg_min_read;
	ws->longest_hint = -1;
	if (ws->heavy_on && may_have_heavy && !ws->count_cells && !par->is_cdna && par->gap_scale == 1.0f && par->n_segs <= 1 && par->bw >= 0 && par->bw < (1 << 24)
	    && par->max_dist_x > 0 && par->max_dist_y > 0 && par->max_iter > heavy_min_window() && (int64_t)par->max_iter + 64 <= heavy_ring_slots()) {
		ba.heavy_flag = ws->heavy_flag, ba.heavy_list = ws->heavy_list, ba.heavy_count = ws->small + 2, ba.heavy_counter = ws->small + 3;
		ba.heavy_min_cells = ws->heavy_min_cells, ba.heavy_cap = ws->n_sms;
	}
	if (ws->seg_on && may_have_long && !ws->count_cells && !ws->dbg_fpv && par->max_dist_x > 0) {
		ba.heavy_flag = ws->heavy_flag;
		ba.seg_reads = ws->seg_reads, ba.seg_items = ws->seg_items, ba.seg_done = ws->seg_done, ba.seg_ctl = ws->small + 8;
		ba.seg_cap = SEG_CAP, ba.seg_min_read = ws->seg_min_read, ba.seg_min_piece = ws->seg_min_piece;
	}
	cudaEventRecord(ws->ev_k1[0], stream);
	launches += launch_chain(ba, ws->n_sms, stream);
	cudaEventRecord(ws->ev_k1[1], stream);
	// totals: the cursors' final values close the offset arrays (entry n_reads), as the prefix sums of the old layout did
	cudaMemcpyAsync(d_u_off + n_reads, ba.out_cursor, 8, cudaMemcpyDeviceToDevice, stream);
	cudaMemcpyAsync(d_b_off + n_reads, ba.out_cursor + 1, 8, cudaMemcpyDeviceToDevice, stream);
	count_launches(launches);
	ws->last_reads = n_reads, ws->last_anchors = n_anchors;
	ws->last_n_u = d_n_u, ws->last_n_v = d_n_v, ws->last_u_off = d_u_off, ws->last_b_off = d_b_off;
	const bool ok = cuda_ok(cudaGetLastError(), "kernel launch");
	if (prev >= 0 && prev != ws->device) cudaSetDevice(prev);
	return ok ? MM2B_OK : MM2B_ERR_CUDA;
}

int mm2b_chain_batch_device(mm2b_workspace_t *ws, const mm2b_params_t *par, int64_t n_reads, int64_t n_anchors,
                            const int64_t *d_off, const mm2b_anchor_t *d_a,
                            int32_t *d_n_u, int32_t *d_n_v, int32_t *d_status, int64_t *d_u_off, int64_t *d_b_off,
                            uint64_t *d_u, mm2b_anchor_t *d_b, void *stream)
{
	return chain_device(ws, par, n_reads, n_anchors, d_off, d_a, d_n_u, d_n_v, d_status, d_u_off, d_b_off, d_u, d_b, nullptr, stream);
}

int mm2b_chain_batch_device_idx(mm2b_workspace_t *ws, const mm2b_params_t *par, int64_t n_reads, int64_t n_anchors,
                                const int64_t *d_off, const mm2b_anchor_t *d_a,
                                int32_t *d_n_u, int32_t *d_n_v, int32_t *d_status, int64_t *d_u_off, int64_t *d_b_off,
                                uint64_t *d_u, int32_t *d_bi, void *stream)
{
	return chain_device(ws, par, n_reads, n_anchors, d_off, d_a, d_n_u, d_n_v, d_status, d_u_off, d_b_off, d_u, nullptr, d_bi, stream);
}

int mm2b_unpack_anchors_device(int device, int64_t n_anchors, const void *d_lo, const void *d_xruns, int32_t n_xruns, const void *d_yruns, int32_t n_yruns,
                               mm2b_anchor_t *d_a, void *stream)
{
	if (n_anchors < 0 || n_anchors >= (1ll << 31) || (n_anchors > 0 && (!d_lo || !d_xruns || !d_yruns || !d_a || n_xruns < 1 || n_yruns < 1))) {
		set_error("%s%s", "mm2b_unpack_anchors_device: bad argument", "");
		return MM2B_ERR_ARG;
	}
	int prev = -1, n_sms = 148;
	cudaGetDevice(&prev);
	if (prev != device && !cuda_ok(cudaSetDevice(device), "cudaSetDevice")) return MM2B_ERR_CUDA;
	cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, device);
	count_launches(launch_unpack(n_anchors, (const uint2*)d_lo, (const uint2*)d_xruns, n_xruns, (const uint2*)d_yruns, n_yruns, d_a, n_sms, (cudaStream_t)stream));
	const bool ok = cuda_ok(cudaGetLastError(), "kernel launch");
	if (prev >= 0 && prev != device) cudaSetDevice(prev);
	return ok ? MM2B_OK : MM2B_ERR_CUDA;
}

int mm2b_ws_stats(mm2b_workspace_t *ws, void *stream_, mm2b_stats_t *st)
{
	if (!ws || !st) return MM2B_ERR_ARG;
	cudaStream_t stream = (cudaStream_t)stream_;
	int prev = -1;
	cudaGetDevice(&prev);
	if (prev != ws->device) cudaSetDevice(ws->device);
	unsigned long long c[6] = {0, 0, 0, 0, 0, 0};
	int64_t tot[2] = {0, 0};
	bool ok = cuda_ok(cudaStreamSynchronize(stream), "cudaStreamSynchronize")
	       && cuda_ok(cudaMemcpy(c, ws->counters, sizeof(c), cudaMemcpyDeviceToHost), "cudaMemcpy(counters)");
	if (ok && ws->last_u_off && ws->last_reads >= 0) {
		ok = cuda_ok(cudaMemcpy(&tot[0], ws->last_u_off + ws->last_reads, 8, cudaMemcpyDeviceToHost), "cudaMemcpy(u_off)")
		  && cuda_ok(cudaMemcpy(&tot[1], ws->last_b_off + ws->last_reads, 8, cudaMemcpyDeviceToHost), "cudaMemcpy(b_off)");
	}
	memset(st, 0, sizeof(*st));
	st->n_reads = ws->last_reads, st->n_anchors = ws->last_anchors;
	st->n_chains = tot[0], st->n_chained = tot[1];
	st->cells_issued = (int64_t)c[0] * 32, st->n_general_reads = (int64_t)c[1], st->cells_ref = (int64_t)c[2], st->window_cells = (int64_t)c[3];
	st->n_heavy_reads = (int64_t)c[4], st->n_cut_reads = (int64_t)c[5];
	if (prev >= 0 && prev != ws->device) cudaSetDevice(prev);
	return ok ? MM2B_OK : MM2B_ERR_CUDA;
}

int mm2b_ws_copy_fpv(mm2b_workspace_t *ws, void *stream_, int64_t n_anchors, int32_t *h_f, int32_t *h_p, int32_t *h_v)
{
	if (!ws || !ws->dbg_fpv) { set_error("%s%s", "mm2b_ws_copy_fpv: workspace was not created with MM2B_KEEP_FPV=1", ""); return MM2B_ERR_ARG; }
	if (n_anchors > ws->max_anchors) return MM2B_ERR_CAPACITY;
	int prev = -1;
	cudaGetDevice(&prev);
	if (prev != ws->device) cudaSetDevice(ws->device);
	const size_t sz = (size_t)n_anchors * 4;
	bool ok = cuda_ok(cudaStreamSynchronize((cudaStream_t)stream_), "cudaStreamSynchronize")
	       && cuda_ok(cudaMemcpy(h_f, ws->dbg_fpv, sz, cudaMemcpyDeviceToHost), "cudaMemcpy(f)")
	       && cuda_ok(cudaMemcpy(h_p, ws->dbg_fpv + ws->max_anchors, sz, cudaMemcpyDeviceToHost), "cudaMemcpy(p)")
	       && cuda_ok(cudaMemcpy(h_v, ws->dbg_fpv + 2 * ws->max_anchors, sz, cudaMemcpyDeviceToHost), "cudaMemcpy(v)");
	if (prev >= 0 && prev != ws->device) cudaSetDevice(prev);
	return ok ? MM2B_OK : MM2B_ERR_CUDA;
}

double mm2b_ws_chain_kernel_ms(mm2b_workspace_t *ws)
{
	if (!ws || ws->last_reads <= 0) return 0.0;
	float ms = 0;
	if (!cuda_ok(cudaEventSynchronize(ws->ev_k1[1]), "cudaEventSynchronize") ||
	    !cuda_ok(cudaEventElapsedTime(&ms, ws->ev_k1[0], ws->ev_k1[1]), "cudaEventElapsedTime")) return -1.0;
	return (double)ms;
}

unsigned mm2b_debug_flags(void) { return debug_flags(); }

double mm2b_measure_int32_peak(int device)
{
	count_launches(6);
	return measure_int32_peak(device);
}

}  // extern "C"
