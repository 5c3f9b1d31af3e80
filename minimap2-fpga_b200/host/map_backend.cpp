// Host side of the seeding front end (include/mm2seed_b200.h): the index on every bound device and the batch call that takes read
// sequences through sketch -> seed -> sort -> chain on the GPU.  It replaces, for a whole mini-batch at once, the first half of the
// reference's mm_map_frag (/root/reference/map.c:287-316): collect_minimizers, collect_seed_hits and the mm_chain_dp call.
//
// Reads are cut into sub-batches by sequence bytes; per device six contexts (stream + device buffers that only ever grow) take
// sub-batches from a shared counter, so H2D, kernels and D2H of neighbouring sub-batches overlap.  How much the later stages need
// (minimizers, anchors, chained anchors) is only known on the device, so a sub-batch has three short host round trips: the
// minimizer total after the count pass of the sketch, the anchor total after the matches, the output totals after chaining.
// No collective, no NCCL: nothing is exchanged between devices.
#include <cuda_runtime_api.h>
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <vector>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mm2seed_b200.h"
#include "../csrc/seed_kernels.cuh"
#include "../csrc/shim_internal.h"

extern "C" int mm2b_device_id(int i);       // chain_backend.cpp: CUDA id of the i-th bound device
extern "C" int mm2b_counting(void);         // chain_backend.cpp: the statistics switch of mm2b_set_counting

using namespace mm2b;

struct mm2b_index {
	int k = 0, w = 0;
	std::vector<DeviceIndex> dev;
};

namespace {

template <class T> struct DevBuf {          // device buffer that only grows
	T *p = nullptr;
	int64_t cap = 0;
	bool ensure(int64_t n, const char *what)
	{
		if (n <= cap) return true;
		cudaFree(p), p = nullptr, cap = 0;
		const int64_t want = n + n / 4 + 64;
		if (!cuda_ok(cudaMalloc((void**)&p, (size_t)want * sizeof(T)), what)) return false;
		cap = want;
		return true;
	}
	void release() { cudaFree(p), p = nullptr, cap = 0; }
};

struct PinBuf {                             // pinned host buffer that only grows
	void *p = nullptr;
	size_t cap = 0;
	bool ensure(size_t bytes)
	{
		if (bytes <= cap) return true;
		mm2b_host_free(p), p = nullptr, cap = 0;
		const size_t want = bytes + bytes / 4 + 4096;
		p = mm2b_host_alloc(want);                  // from the library's pinned pool when it has a block of about this size
		if (!p) return false;
		cap = want;
		return true;
	}
	void release() { mm2b_host_free(p), p = nullptr, cap = 0; }
};

// One pipeline context: a stream and everything a sub-batch needs on its device
struct Ctx {
	int device = -1, n_sms = 0, id = 0;
	// Kernels, input copies and output copies each have their own stream.  With everything on one stream the inputs of a sub-batch were
	// ordered after the same context's previous output copy, and the marker that releases them sat in the device-to-host copy queue
	// behind every other context's pending outputs: no input moved until all outputs had (profiles/r3g_trace_one_stream.txt).
	cudaStream_t stream = nullptr, in_stream = nullptr, out_stream = nullptr;
	cudaEvent_t ev_in = nullptr;
	cudaEvent_t ev[6] = {};
	DevBuf<uint8_t> seq;
	DevBuf<int64_t> seq_off, tile_excl, mv_off, n_a, a_off, u_off, b_off;
	DevBuf<unsigned long long> tile_state;
	DevBuf<int32_t> tile_off, tile_read, occ, arel, rep_len, n_mini_pos, tie_list, n_u, n_v, status;
	DevBuf<uint32_t> mini_pos;
	DevBuf<uint64_t> hv, u;
	DevBuf<ulonglong2> mv, a, a_tmp, b;
	DevBuf<int> small;
	mm2b_workspace_t *ws = nullptr;
	int64_t ws_anchors = 0, ws_reads = 0;
	PinBuf h_small, h_read, h_tiles;    // offsets in; per-read results; tile -> read map
	int64_t *h_scal = nullptr, *d_scal = nullptr;      // 4 scalars the kernels store straight into host memory (mapped), and the device's address of them
	bool create(int dev)
	{
		device = dev;
		if (!cuda_ok(cudaSetDevice(dev), "cudaSetDevice")) return false;
		cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
		if (!cuda_ok(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
		if (!cuda_ok(cudaStreamCreateWithFlags(&in_stream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
		if (!cuda_ok(cudaStreamCreateWithFlags(&out_stream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
		if (!cuda_ok(cudaEventCreateWithFlags(&ev_in, cudaEventDisableTiming), "cudaEventCreate")) return false;
		for (auto &e : ev) if (!cuda_ok(cudaEventCreate(&e), "cudaEventCreate")) return false;
		if (!cuda_ok(cudaHostAlloc((void**)&h_scal, 64, cudaHostAllocMapped), "cudaHostAlloc") || !cuda_ok(cudaHostGetDevicePointer((void**)&d_scal, h_scal, 0), "cudaHostGetDevicePointer")) return false;
		memset(h_scal, 0, 64);
		return small.ensure(16, "cudaMalloc");
	}
	void destroy()
	{
		if (device < 0) return;
		cudaSetDevice(device);
		if (stream) cudaStreamSynchronize(stream);
		if (in_stream) cudaStreamSynchronize(in_stream), cudaStreamDestroy(in_stream), in_stream = nullptr;
		if (out_stream) cudaStreamSynchronize(out_stream), cudaStreamDestroy(out_stream), out_stream = nullptr;
		if (ev_in) cudaEventDestroy(ev_in), ev_in = nullptr;
		seq.release(), seq_off.release(), tile_excl.release(), tile_state.release(), mv_off.release(), n_a.release(), a_off.release(), u_off.release(), b_off.release();
		tile_off.release(), tile_read.release(), occ.release(), arel.release(), rep_len.release(), n_mini_pos.release(), tie_list.release();
		n_u.release(), n_v.release(), status.release(), mini_pos.release(), hv.release(), u.release(), mv.release(), a.release(), a_tmp.release(), b.release(), small.release();
		mm2b_ws_destroy(ws), ws = nullptr;
		h_small.release(), h_read.release(), h_tiles.release();
		if (h_scal) cudaFreeHost(h_scal), h_scal = d_scal = nullptr;
		for (auto &e : ev) if (e) cudaEventDestroy(e);
		if (stream) cudaStreamDestroy(stream);
		device = -1;
	}
};

struct Segment { PinBuf u, b, mp; };

std::mutex g_mu;
std::vector<Ctx*> g_free_ctx;               // idle contexts of all devices
std::vector<Segment*> g_free_seg;

Ctx *ctx_acquire(int device)
{
	{
		std::lock_guard<std::mutex> lk(g_mu);
		for (size_t i = 0; i < g_free_ctx.size(); ++i)
			if (g_free_ctx[i]->device == device) {
				Ctx *c = g_free_ctx[i];
				g_free_ctx.erase(g_free_ctx.begin() + (long)i);
				return c;
			}
	}
	static std::atomic<int> n_made{0};
	Ctx *c = new Ctx();
	c->id = n_made.fetch_add(1);
	if (!c->create(device)) { c->destroy(); delete c; return nullptr; }
	return c;
}
void ctx_release(Ctx *c) { std::lock_guard<std::mutex> lk(g_mu); g_free_ctx.push_back(c); }

Segment *seg_acquire()
{
	std::lock_guard<std::mutex> lk(g_mu);
	if (g_free_seg.empty()) return new Segment();
	Segment *s = g_free_seg.back();
	g_free_seg.pop_back();
	return s;
}

struct SubBatch { int64_t r0, r1; };

struct ResultPriv {
	std::vector<Segment*> segs;
	std::vector<uint64_t*> seg_u;
	std::vector<mm2b_anchor_t*> seg_b;
	std::vector<uint32_t*> seg_mp;
	void *per_read = nullptr;
};

struct DebugOut {                           // mm2b_seed_debug: copies of the intermediate products of one sub-batch
	std::vector<int64_t> mv_off, a_off;
	std::vector<mm2b_anchor_t> mv, a;
	std::vector<int32_t> rep_len, n_mini_pos;
	std::vector<uint32_t> mini_pos;
	int64_t n_tie = 0;
};

struct Call {
	const mm2b_index *idx;
	mm2b_seed_params_t seed;
	const mm2b_params_t *chain;             // null: stop after the sort (debug)
	const int64_t *seq_off;
	const char *seq;
	mm2b_map_result_t *res;
	ResultPriv *priv;
	std::vector<SubBatch> subs;
	std::atomic<int> next{0}, failed{0};
	char err[512] = {0};
	std::mutex mu;
	double sketch_ms = 0, seed_ms = 0, sort_ms = 0, chain_ms = 0;
	int64_t tot_mini = 0, tot_anchors = 0, tot_chains = 0, tot_chained = 0, n_tie = 0, h2d = 0, d2h = 0, cells = 0;
	DebugOut *dbg = nullptr;
	// MM2B_MAP_TRACE=1: where every sub-batch spent its time (host clock at the syncs, device clock at the stage events)
	bool trace = false;
	std::chrono::steady_clock::time_point t0;
	cudaEvent_t base_ev = nullptr;
	std::vector<std::array<float, 15>> rows;
	long misses0 = 0;
};

#define CK(expr, what) do { if (!cuda_ok((expr), what)) return false; } while (0)

// One sub-batch, start to finish, on context c.  Everything is enqueued on c.stream; the three syncs are where host decisions
// (buffer sizes) need device results.
bool run_sub(Call &call, Ctx &c, int si)
{
	const SubBatch sb = call.subs[si];
	const int64_t R = sb.r1 - sb.r0, s0 = call.seq_off[sb.r0], S = call.seq_off[sb.r1] - s0;
	const DeviceIndex *ix = nullptr;
	for (const DeviceIndex &d : call.idx->dev) if (d.device == c.device) ix = &d;
	if (!ix) { set_error("%s%s", "mm2b_map_batch: the index is not resident on this device", ""); return false; }
	mm2b_map_result_t *res = call.res;
	cudaStream_t st = c.stream;
	CK(cudaSetDevice(c.device), "cudaSetDevice");
	static const bool use_ev = !(getenv("MM2B_MAP_EVENTS") && atoi(getenv("MM2B_MAP_EVENTS")) == 0);
	float host_t[7] = {0, 0, 0, 0, 0, 0, 0};
	auto stamp = [&](int i) { if (call.trace) host_t[i] = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - call.t0).count(); };
	stamp(0);

	// ---- offsets and tiles (host), sequences to the device
	const int64_t tile_pos = sketch_tile_positions(ix->w);
	const int64_t max_tiles = S / tile_pos + R + 1;
	if (max_tiles >= (1ll << 31) || S >= (1ll << 40)) { set_error("%s%s", "mm2b_map_batch: sub-batch too large", ""); return false; }
	if (!c.h_small.ensure((size_t)(R + 1) * 12 + 64) || !c.h_tiles.ensure((size_t)max_tiles * 4)) return false;
	int64_t *h_seq_off = (int64_t*)c.h_small.p;
	int32_t *h_tile_off = (int32_t*)(h_seq_off + R + 1), *h_tile_read = (int32_t*)c.h_tiles.p;
	int64_t n_tiles64 = 0;
	for (int64_t r = 0; r <= R; ++r) {
		h_seq_off[r] = call.seq_off[sb.r0 + r] - s0;
		h_tile_off[r] = (int32_t)n_tiles64;
		if (r < R) {
			const int64_t nt = (call.seq_off[sb.r0 + r + 1] - call.seq_off[sb.r0 + r] + tile_pos - 1) / tile_pos;
			for (int64_t q = 0; q < nt; ++q) h_tile_read[n_tiles64 + q] = (int32_t)r;
			n_tiles64 += nt;
		}
	}
	const int32_t n_tiles = (int32_t)n_tiles64;
	if (!c.seq.ensure(S + 32, "cudaMalloc(seq)") || !c.seq_off.ensure(R + 1, "cudaMalloc") || !c.tile_off.ensure(R + 1, "cudaMalloc") ||
	    !c.tile_read.ensure(n_tiles + 1, "cudaMalloc") || !c.tile_state.ensure(n_tiles + 1, "cudaMalloc") || !c.tile_excl.ensure(n_tiles + 2, "cudaMalloc") || !c.mv_off.ensure(R + 2, "cudaMalloc") ||
	    !c.rep_len.ensure(R + 1, "cudaMalloc") || !c.n_mini_pos.ensure(R + 1, "cudaMalloc") || !c.n_a.ensure(R + 1, "cudaMalloc") || !c.a_off.ensure(R + 2, "cudaMalloc") ||
	    !c.tie_list.ensure(R + 1, "cudaMalloc") || !c.n_u.ensure(R + 1, "cudaMalloc") || !c.n_v.ensure(R + 1, "cudaMalloc") || !c.status.ensure(R + 1, "cudaMalloc") ||
	    !c.u_off.ensure(R + 2, "cudaMalloc") || !c.b_off.ensure(R + 2, "cudaMalloc")) return false;
	static const bool one_stream = getenv("MM2B_MAP_ONE_STREAM") && atoi(getenv("MM2B_MAP_ONE_STREAM")) > 0;     // the old layout, for comparison
	cudaStream_t st_in = one_stream ? st : c.in_stream, st_out = one_stream ? st : c.out_stream;
	if (use_ev) CK(cudaEventRecord(c.ev[0], st_in), "cudaEventRecord");
	if (S > 0) CK(cudaMemcpyAsync(c.seq.p, call.seq + s0, (size_t)S, cudaMemcpyHostToDevice, st_in), "H2D sequences");
	CK(cudaMemcpyAsync(c.seq_off.p, h_seq_off, (size_t)(R + 1) * 8, cudaMemcpyHostToDevice, st_in), "H2D seq_off");
	CK(cudaMemcpyAsync(c.tile_off.p, h_tile_off, (size_t)(R + 1) * 4, cudaMemcpyHostToDevice, st_in), "H2D tile_off");
	if (n_tiles) CK(cudaMemcpyAsync(c.tile_read.p, h_tile_read, (size_t)n_tiles * 4, cudaMemcpyHostToDevice, st_in), "H2D tile_read");
	if (!one_stream) {
		CK(cudaEventRecord(c.ev_in, st_in), "cudaEventRecord");
		CK(cudaStreamWaitEvent(st, c.ev_in, 0), "cudaStreamWaitEvent");
	}
	stamp(1);

	SeedArgs a;
	memset(&a, 0, sizeof(a));
	a.n_reads = R, a.seq = c.seq.p, a.seq_len = S, a.seq_off = c.seq_off.p, a.tile_off = c.tile_off.p, a.tile_read = c.tile_read.p, a.n_tiles = n_tiles;
	a.k = ix->k, a.w = ix->w, a.max_occ = call.seed.max_occ;
	a.tile_state = c.tile_state.p, a.tile_ticket = c.small.p + 8, a.tile_excl = c.tile_excl.p, a.mv_off = c.mv_off.p;
	a.rep_len = c.rep_len.p, a.n_mini_pos = c.n_mini_pos.p, a.n_a = c.n_a.p, a.a_off = c.a_off.p, a.tie_list = c.tie_list.p, a.tie_count = c.small.p;

	// ---- sketch: one pass into a buffer sized for the usual density of minimizers (2 / (w + 1) per base, and half as much again);
	//      (sync: how many there are) and once more into a larger buffer in the rare case that was not enough
	int launches = 0;
	volatile int64_t *h_tot = c.h_scal;         // totals arrive by a store from the device, not by a copy (see export_scalars_kernel)
	int64_t n_mv = 0;
	if (!c.mv.ensure(std::max<int64_t>(c.mv.cap, 3 * S / (ix->w + 1) + 1024), "cudaMalloc(mv)")) return false;
	static const bool tiny_first = getenv("MM2B_TEST_SMALL_MV") && atoi(getenv("MM2B_TEST_SMALL_MV")) > 0;     // test hook: force the second attempt
	for (int attempt = 0; attempt < 2; ++attempt) {
		a.mv = c.mv.p, a.mv_cap = tiny_first && attempt == 0 ? std::min<int64_t>(c.mv.cap, 1000) : c.mv.cap;
		h_tot[0] = 0;
		if (n_tiles > 0) {
			launches += launch_sketch(a, st);
			launches += launch_export_scalars(c.d_scal, c.tile_excl.p + n_tiles, nullptr, nullptr, st);
		}
		CK(cudaStreamSynchronize(st), "sketch");
		stamp(2);
		n_mv = h_tot[0];
		if (n_mv <= a.mv_cap) break;
		if (attempt == 1 || !c.mv.ensure(n_mv + 1, "cudaMalloc(mv)")) { if (attempt == 1) set_error("%s%s", "mm2b_map_batch: minimizer buffer", ""); return false; }
	}
	if (!c.occ.ensure(n_mv + 1, "cudaMalloc") || !c.hv.ensure(n_mv + 1, "cudaMalloc") || !c.arel.ensure(n_mv + 1, "cudaMalloc") || !c.mini_pos.ensure(n_mv + 1, "cudaMalloc")) return false;
	a.occ = c.occ.p, a.hv = c.hv.p, a.arel = c.arel.p, a.mini_pos = c.mini_pos.p;
	launches += launch_read_offsets(a, st);
	if (use_ev) CK(cudaEventRecord(c.ev[1], st), "cudaEventRecord");

	// ---- index probes, matches, (sync: how many anchors)
	launches += launch_index_lookup(*ix, n_mv, c.mv.p, nullptr, c.occ.p, c.hv.p, st);
	launches += launch_matches(a, c.n_sms, st);
	launches += launch_scan_i64(c.n_a.p, c.a_off.p, R, st);
	launches += launch_export_scalars(c.d_scal, c.a_off.p + R, nullptr, nullptr, st);
	CK(cudaStreamSynchronize(st), "matches");
	stamp(3);
	const int64_t n_anchors = h_tot[0];
	if (n_anchors >= (1ll << 31)) { set_error("%s%s", "mm2b_map_batch: more than 2^31 anchors in one sub-batch", ""); return false; }
	if (!c.a.ensure(n_anchors + 1, "cudaMalloc(a)") || !c.a_tmp.ensure(n_anchors + 1, "cudaMalloc(a_tmp)")) return false;
	a.a = c.a.p, a.a_tmp = c.a_tmp.p;
	launches += launch_expand(a, *ix, n_mv, st);
	if (use_ev) CK(cudaEventRecord(c.ev[2], st), "cudaEventRecord");
	launches += launch_sort(a, *ix, c.n_sms, st);
	if (use_ev) CK(cudaEventRecord(c.ev[3], st), "cudaEventRecord");

	// ---- per-read results staging: [n_u][n_v][status][rep_len][n_mini_pos] int32, then [n_a][mv_off][u_off][b_off] int64 (+1 entries)
	if (!c.h_read.ensure((size_t)(R + 2) * (5 * 4 + 4 * 8) + 64)) return false;
	int32_t *h_i32 = (int32_t*)c.h_read.p;
	int64_t *h_i64 = (int64_t*)(((uintptr_t)(h_i32 + 5 * (R + 1)) + 7) & ~(uintptr_t)7);
	int32_t *h_n_u = h_i32, *h_n_v = h_i32 + (R + 1), *h_status = h_i32 + 2 * (R + 1), *h_rep = h_i32 + 3 * (R + 1), *h_nmp = h_i32 + 4 * (R + 1);
	int64_t *h_n_a = h_i64, *h_mv_off = h_i64 + (R + 1), *h_u_off = h_i64 + 2 * (R + 1), *h_b_off = h_i64 + 3 * (R + 1);

	if (call.chain) {
		if (n_anchors > c.ws_anchors || R > c.ws_reads) {
			CK(cudaStreamSynchronize(st), "before workspace growth");
			mm2b_ws_destroy(c.ws);
			c.ws_anchors = n_anchors + n_anchors / 4 + 1024, c.ws_reads = R + R / 4 + 64;
			c.ws = mm2b_ws_create(c.device, c.ws_anchors, c.ws_reads);
			if (!c.ws) { c.ws_anchors = c.ws_reads = 0; return false; }
		}
		if (!c.u.ensure(n_anchors + 1, "cudaMalloc(u)") || !c.b.ensure(n_anchors + 1, "cudaMalloc(b)")) return false;
		mm2b_ws_set_counting(c.ws, mm2b_counting());
		if (mm2b_chain_batch_device(c.ws, call.chain, R, n_anchors, c.a_off.p, (const mm2b_anchor_t*)c.a.p, c.n_u.p, c.n_v.p, c.status.p, c.u_off.p, c.b_off.p,
		                            c.u.p, (mm2b_anchor_t*)c.b.p, st) != MM2B_OK) return false;
		if (use_ev) CK(cudaEventRecord(c.ev[4], st), "cudaEventRecord");
		launches += launch_export_scalars(c.d_scal, c.u_off.p + R, c.b_off.p + R, c.small.p, st);
	} else {
		if (use_ev) CK(cudaEventRecord(c.ev[4], st), "cudaEventRecord");
		launches += launch_export_scalars(c.d_scal, nullptr, nullptr, c.small.p, st);
	}
	CK(cudaStreamSynchronize(st), "chain");         // (sync: how many chains and chained anchors come back)
	const int64_t tot_u = call.chain ? h_tot[0] : 0, tot_b = call.chain ? h_tot[1] : 0;
	const int n_tie_reads = (int)h_tot[2];
	stamp(4);
	CK(cudaGetLastError(), "kernels");
	count_launches(launches);

	int64_t cells = 0;
	if (call.chain && mm2b_counting()) {
		mm2b_stats_t cs;
		if (mm2b_ws_stats(c.ws, st, &cs) == MM2B_OK) cells = cs.cells_ref;
	}

	if (call.dbg) {                             // debug: bring the intermediate products back
		DebugOut &d = *call.dbg;
		CK(cudaMemcpy(h_rep, c.rep_len.p, (size_t)R * 4, cudaMemcpyDeviceToHost), "D2H rep_len");
		CK(cudaMemcpy(h_nmp, c.n_mini_pos.p, (size_t)R * 4, cudaMemcpyDeviceToHost), "D2H n_mini_pos");
		CK(cudaMemcpy(h_mv_off, c.mv_off.p, (size_t)(R + 1) * 8, cudaMemcpyDeviceToHost), "D2H mv_off");
		d.mv_off.assign(h_mv_off, h_mv_off + R + 1);
		d.a_off.resize((size_t)R + 1);
		d.mv.resize((size_t)n_mv), d.a.resize((size_t)n_anchors), d.mini_pos.resize((size_t)n_mv);
		d.rep_len.assign(h_rep, h_rep + R), d.n_mini_pos.assign(h_nmp, h_nmp + R), d.n_tie = n_tie_reads;
		CK(cudaMemcpy(d.a_off.data(), c.a_off.p, (size_t)(R + 1) * 8, cudaMemcpyDeviceToHost), "D2H a_off");
		if (n_mv) CK(cudaMemcpy(d.mv.data(), c.mv.p, (size_t)n_mv * 16, cudaMemcpyDeviceToHost), "D2H mv");
		if (n_mv) CK(cudaMemcpy(d.mini_pos.data(), c.mini_pos.p, (size_t)n_mv * 4, cudaMemcpyDeviceToHost), "D2H mini_pos");
		if (n_anchors) CK(cudaMemcpy(d.a.data(), c.a.p, (size_t)n_anchors * 16, cudaMemcpyDeviceToHost), "D2H a");
		return true;
	}

	// ---- outputs into this sub-batch's segment (the kernels are done: the host has just waited for their totals)
	Segment *seg = call.priv->segs[(size_t)si];
	if (!seg->u.ensure((size_t)(tot_u + 1) * 8) || !seg->b.ensure((size_t)(tot_b + 1) * 16) || !seg->mp.ensure((size_t)(n_mv + 1) * 4)) return false;
	if (tot_u) CK(cudaMemcpyAsync(seg->u.p, c.u.p, (size_t)tot_u * 8, cudaMemcpyDeviceToHost, st_out), "D2H u");
	if (tot_b) CK(cudaMemcpyAsync(seg->b.p, c.b.p, (size_t)tot_b * 16, cudaMemcpyDeviceToHost, st_out), "D2H b");
	if (n_mv) CK(cudaMemcpyAsync(seg->mp.p, c.mini_pos.p, (size_t)n_mv * 4, cudaMemcpyDeviceToHost, st_out), "D2H mini_pos");
	if (call.chain) {
		CK(cudaMemcpyAsync(h_n_u, c.n_u.p, (size_t)R * 4, cudaMemcpyDeviceToHost, st_out), "D2H n_u");
		CK(cudaMemcpyAsync(h_n_v, c.n_v.p, (size_t)R * 4, cudaMemcpyDeviceToHost, st_out), "D2H n_v");
		CK(cudaMemcpyAsync(h_status, c.status.p, (size_t)R * 4, cudaMemcpyDeviceToHost, st_out), "D2H status");
		CK(cudaMemcpyAsync(h_u_off, c.u_off.p, (size_t)(R + 1) * 8, cudaMemcpyDeviceToHost, st_out), "D2H u_off");
		CK(cudaMemcpyAsync(h_b_off, c.b_off.p, (size_t)(R + 1) * 8, cudaMemcpyDeviceToHost, st_out), "D2H b_off");
	}
	CK(cudaMemcpyAsync(h_rep, c.rep_len.p, (size_t)R * 4, cudaMemcpyDeviceToHost, st_out), "D2H rep_len");
	CK(cudaMemcpyAsync(h_nmp, c.n_mini_pos.p, (size_t)R * 4, cudaMemcpyDeviceToHost, st_out), "D2H n_mini_pos");
	CK(cudaMemcpyAsync(h_n_a, c.n_a.p, (size_t)R * 8, cudaMemcpyDeviceToHost, st_out), "D2H n_a");
	CK(cudaMemcpyAsync(h_mv_off, c.mv_off.p, (size_t)(R + 1) * 8, cudaMemcpyDeviceToHost, st_out), "D2H mv_off");
	if (use_ev) CK(cudaEventRecord(c.ev[5], st_out), "cudaEventRecord");
	stamp(5);
	CK(cudaStreamSynchronize(st_out), "outputs");
	stamp(6);
	call.priv->seg_u[(size_t)si] = (uint64_t*)seg->u.p, call.priv->seg_b[(size_t)si] = (mm2b_anchor_t*)seg->b.p, call.priv->seg_mp[(size_t)si] = (uint32_t*)seg->mp.p;
	for (int64_t r = 0; r < R; ++r) {
		const int64_t g = sb.r0 + r;
		res->status[g] = h_status[r], res->n_u[g] = h_n_u[r], res->n_v[g] = h_n_v[r], res->rep_len[g] = h_rep[r], res->n_mini_pos[g] = h_nmp[r];
		res->n_mini[g] = (int32_t)(h_mv_off[r + 1] - h_mv_off[r]), res->seg[g] = si;
		res->n_a[g] = h_n_a[r], res->u_off[g] = h_u_off[r], res->b_off[g] = h_b_off[r], res->mp_off[g] = h_mv_off[r];
	}
	float t01 = 0, t12 = 0, t23 = 0, t34 = 0;
	if (use_ev) {
		cudaEventElapsedTime(&t01, c.ev[0], c.ev[1]), cudaEventElapsedTime(&t12, c.ev[1], c.ev[2]);
		cudaEventElapsedTime(&t23, c.ev[2], c.ev[3]), cudaEventElapsedTime(&t34, c.ev[3], c.ev[4]);
	}
	std::lock_guard<std::mutex> lk(call.mu);
	if (call.trace && call.base_ev) {
		std::array<float, 15> row;
		row[0] = (float)si, row[1] = (float)c.id;
		for (int i = 0; i < 7; ++i) row[2 + i] = host_t[i];
		for (int i = 0; i < 6; ++i) { row[9 + i] = 0; if (use_ev) cudaEventElapsedTime(&row[9 + i], call.base_ev, c.ev[i]); }
		call.rows.push_back(row);
	}
	call.sketch_ms += t01, call.seed_ms += t12, call.sort_ms += t23, call.chain_ms += t34;
	call.tot_mini += n_mv, call.tot_anchors += n_anchors, call.tot_chains += tot_u, call.tot_chained += tot_b, call.n_tie += n_tie_reads, call.cells += cells;
	call.h2d += S + (R + 1) * 12, call.d2h += tot_u * 8 + tot_b * 16 + n_mv * 4 + R * 28 + (R + 1) * 24;
	return true;
}

void worker(Call *call, int device)
{
	Ctx *c = ctx_acquire(device);
	if (!c) {
		std::lock_guard<std::mutex> lk(call->mu);
		if (!call->failed.exchange(1)) snprintf(call->err, sizeof(call->err), "%s", mm2b_last_error());
		return;
	}
	for (;;) {
		if (call->failed.load()) break;
		const int si = call->next.fetch_add(1);
		if (si >= (int)call->subs.size()) break;
		if (!run_sub(*call, *c, si)) {
			cudaStreamSynchronize(c->in_stream), cudaStreamSynchronize(c->stream), cudaStreamSynchronize(c->out_stream);
			std::lock_guard<std::mutex> lk(call->mu);
			if (!call->failed.exchange(1)) snprintf(call->err, sizeof(call->err), "%s", mm2b_last_error());
			break;
		}
	}
	ctx_release(c);
}

int64_t env_ll(const char *name, int64_t dflt)
{
	const char *s = getenv(name);
	if (!s) return dflt;
	const long long v = atoll(s);
	return v > 0 ? v : dflt;
}

void cut_subs(Call &call, int64_t n_reads)
{
	const int64_t sub_bytes = env_ll("MM2B_MAP_SUB_BYTES", 64ll << 20), sub_reads = 1 << 18;     // (kernels over a few thousand reads do not fill the GPU)
	// The first sub-batches of a large call are smaller: the contexts all start at once, and until the first input copy has landed
	// the kernels have nothing to do (with equal sizes the first 7 ms of a call were input copies only).
	static const bool ramp = !(getenv("MM2B_MAP_RAMP") && atoi(getenv("MM2B_MAP_RAMP")) == 0);
	const bool large = ramp && call.seq_off[n_reads] - call.seq_off[0] > 6 * sub_bytes;
	for (int64_t r0 = 0; r0 < n_reads;) {
		const size_t k = call.subs.size();
		const int64_t lim = call.seq_off[r0] + (large && k < 2 ? sub_bytes / 4 : large && k < 4 ? sub_bytes / 2 : sub_bytes);
		int64_t r1 = std::upper_bound(call.seq_off + r0 + 1, call.seq_off + n_reads + 1, lim) - call.seq_off - 1;
		if (r1 <= r0) r1 = r0 + 1;
		if (r1 - r0 > sub_reads) r1 = r0 + sub_reads;
		call.subs.push_back(SubBatch{r0, r1});
		r0 = r1;
	}
}

void run_call(Call &call)
{
	const int n_dev = (int)call.idx->dev.size();
	const int per_dev = (int)env_ll("MM2B_MAP_CTX", 6);      // measured: 2 / 3 / 4 / 6 contexts -> 1.09 / 1.19 / 1.27 / 1.37 M reads/s (profiles/r2r_front_sweep.txt)
	const int n_workers = (int)std::min<int64_t>((int64_t)n_dev * per_dev, (int64_t)call.subs.size());
	static const bool trace = getenv("MM2B_MAP_TRACE") && atoi(getenv("MM2B_MAP_TRACE")) > 0;
	cudaStream_t base_st = nullptr;
	if (trace && n_dev == 1 && !call.dbg && cudaSetDevice(call.idx->dev[0].device) == cudaSuccess && cudaStreamCreateWithFlags(&base_st, cudaStreamNonBlocking) == cudaSuccess) {
		call.trace = true, call.t0 = std::chrono::steady_clock::now(), call.misses0 = pin_pool_misses();
		cudaEventCreate(&call.base_ev);
		cudaEventRecord(call.base_ev, base_st);
		cudaStreamSynchronize(base_st);
	}
	std::vector<std::thread> th;
	for (int i = 1; i < n_workers; ++i) th.emplace_back(worker, &call, call.idx->dev[(size_t)(i % n_dev)].device);
	if (n_workers > 0) worker(&call, call.idx->dev[0].device);
	for (auto &t : th) t.join();
	if (call.trace) {
		std::sort(call.rows.begin(), call.rows.end());
		fprintf(stderr, "[mm2b trace] sub ctx | host: start inputs-enqueued sketch-sync matches-sync chain-sync outputs-enqueued outputs-sync | device: h2d-start sketch-end seed-end sort-end chain-end d2h-end (ms since the call began); pinned-pool misses during the call: %ld\n", pin_pool_misses() - call.misses0);
		for (const auto &r : call.rows)
			fprintf(stderr, "[mm2b trace] %3d %2d | %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f | %6.2f %6.2f %6.2f %6.2f %6.2f %6.2f\n", (int)r[0], (int)r[1], r[2], r[3], r[4], r[5], r[6], r[7], r[8], r[9], r[10], r[11], r[12], r[13], r[14]);
		cudaEventDestroy(call.base_ev);
		cudaStreamDestroy(base_st);
	}
}

}  // namespace

extern "C" {

int mm2b_map_supported(int k, int w, int is_hpc, int n_segs, int64_t map_flag, int sdust_thres)
{
	const int64_t unsupported = 0x001 | 0x002 | 0x100000 | 0x200000 | 0x400000;     // MM_F_NO_DIAG, NO_DUAL, FOR_ONLY, REV_ONLY, HEAP_SORT (minimap.h:8-30)
	return k >= 1 && k <= SKETCH_MAX_K && (k & 1) && w >= 1 && w <= SKETCH_MAX_W && !is_hpc && n_segs == 1 && !(map_flag & unsupported) && sdust_thres <= 0;
}

mm2b_index_t *mm2b_index_create(const mm2b_index_desc_t *d)
{
	if (!d || d->n_keys < 0 || d->n_pos < 0 || (d->n_keys > 0 && (!d->keys || !d->vals)) || (d->n_pos > 0 && !d->pos)) { set_error("%s%s", "mm2b_index_create: bad argument", ""); return nullptr; }
	if (mm2b_num_devices() <= 0 && mm2b_init(0, nullptr) != MM2B_OK) return nullptr;
	mm2b_index_t *idx = new mm2b_index_t();
	idx->k = d->k, idx->w = d->w;
	int log2cap = 4;
	while ((1ll << log2cap) < 2 * d->n_keys) ++log2cap;
	bool ok = true;
	for (int i = 0; i < mm2b_num_devices() && ok; ++i) {
		DeviceIndex ix;
		memset(&ix, 0, sizeof(ix));
		ix.device = mm2b_device_id(i), ix.k = d->k, ix.w = d->w, ix.log2cap = log2cap, ix.n_keys = d->n_keys, ix.n_pos = d->n_pos;
		uint64_t *d_keys = nullptr, *d_vals = nullptr;
		ok = cuda_ok(cudaSetDevice(ix.device), "cudaSetDevice")
		  && cuda_ok(cudaMalloc((void**)&ix.tab_keys, sizeof(uint64_t) << log2cap), "cudaMalloc(index keys)")
		  && cuda_ok(cudaMalloc((void**)&ix.tab_vals, sizeof(uint64_t) << log2cap), "cudaMalloc(index values)")
		  && cuda_ok(cudaMalloc((void**)&ix.pos, (size_t)std::max<int64_t>(d->n_pos, 1) * 8), "cudaMalloc(index positions)")
		  && cuda_ok(cudaMalloc((void**)&d_keys, (size_t)std::max<int64_t>(d->n_keys, 1) * 8), "cudaMalloc") && cuda_ok(cudaMalloc((void**)&d_vals, (size_t)std::max<int64_t>(d->n_keys, 1) * 8), "cudaMalloc");
		idx->dev.push_back(ix);                 // (pushed before the copies so that a failure below still frees it)
		if (ok && d->n_keys) ok = cuda_ok(cudaMemcpy(d_keys, d->keys, (size_t)d->n_keys * 8, cudaMemcpyHostToDevice), "H2D index keys")
		                       && cuda_ok(cudaMemcpy(d_vals, d->vals, (size_t)d->n_keys * 8, cudaMemcpyHostToDevice), "H2D index values");
		if (ok && d->n_pos) ok = cuda_ok(cudaMemcpy(ix.pos, d->pos, (size_t)d->n_pos * 8, cudaMemcpyHostToDevice), "H2D index positions");
		if (ok) {
			count_launches(launch_index_build(ix, d_keys, d_vals, nullptr));
			ok = cuda_ok(cudaDeviceSynchronize(), "index build");
		}
		cudaFree(d_keys), cudaFree(d_vals);
	}
	if (!ok) { mm2b_index_destroy(idx); return nullptr; }
	return idx;
}

void mm2b_index_destroy(mm2b_index_t *idx)
{
	if (!idx) return;
	for (DeviceIndex &ix : idx->dev) {
		cudaSetDevice(ix.device);
		cudaFree(ix.tab_keys), cudaFree(ix.tab_vals), cudaFree(ix.pos);
	}
	delete idx;
}

int mm2b_index_lookup(mm2b_index_t *idx, int64_t n, const uint64_t *minimizers, int32_t *n_occ, uint64_t *val)
{
	if (!idx || idx->dev.empty() || n < 0 || (n > 0 && (!minimizers || !n_occ || !val))) { set_error("%s%s", "mm2b_index_lookup: bad argument", ""); return MM2B_ERR_ARG; }
	if (n == 0) return MM2B_OK;
	const DeviceIndex &ix = idx->dev[0];
	uint64_t *d_m = nullptr, *d_v = nullptr;
	int32_t *d_n = nullptr;
	bool ok = cuda_ok(cudaSetDevice(ix.device), "cudaSetDevice") && cuda_ok(cudaMalloc((void**)&d_m, (size_t)n * 8), "cudaMalloc") && cuda_ok(cudaMalloc((void**)&d_v, (size_t)n * 8), "cudaMalloc")
	       && cuda_ok(cudaMalloc((void**)&d_n, (size_t)n * 4), "cudaMalloc") && cuda_ok(cudaMemcpy(d_m, minimizers, (size_t)n * 8, cudaMemcpyHostToDevice), "H2D");
	if (ok) {
		count_launches(launch_index_lookup(ix, n, nullptr, d_m, d_n, d_v, nullptr));
		ok = cuda_ok(cudaMemcpy(n_occ, d_n, (size_t)n * 4, cudaMemcpyDeviceToHost), "D2H") && cuda_ok(cudaMemcpy(val, d_v, (size_t)n * 8, cudaMemcpyDeviceToHost), "D2H");
	}
	cudaFree(d_m), cudaFree(d_v), cudaFree(d_n);
	return ok ? MM2B_OK : MM2B_ERR_CUDA;
}

int mm2b_map_batch(mm2b_index_t *idx, const mm2b_seed_params_t *seed, const mm2b_params_t *chain,
                   int64_t n_reads, const int64_t *seq_off, const char *seq, mm2b_map_result_t **out)
{
	if (out) *out = nullptr;
	if (!idx || idx->dev.empty() || !seed || !chain || n_reads < 0 || !seq_off || !out || (n_reads > 0 && seq_off[n_reads] > 0 && !seq)) {
		set_error("%s%s", "mm2b_map_batch: bad argument", "");
		return MM2B_ERR_ARG;
	}
	if (!mm2b_map_supported(idx->k, idx->w, 0, 1, 0, 0)) { set_error("%s%s", "mm2b_map_batch: this (k, w) is outside the device path (odd k <= 28, w <= 64)", ""); return MM2B_ERR_ARG; }
	Call call;
	call.idx = idx, call.seed = *seed, call.chain = chain, call.seq_off = seq_off, call.seq = seq;
	cut_subs(call, n_reads);
	mm2b_map_result_t *res = (mm2b_map_result_t*)calloc(1, sizeof(*res));
	ResultPriv *priv = new ResultPriv();
	const size_t nr = (size_t)std::max<int64_t>(n_reads, 1);
	char *blk = (char*)malloc(nr * (7 * 4 + 4 * 8) + 64);
	priv->per_read = blk;
	res->n_reads = n_reads, res->priv = priv;
	res->n_a = (int64_t*)blk, res->u_off = res->n_a + nr, res->b_off = res->u_off + nr, res->mp_off = res->b_off + nr;
	res->status = (int32_t*)(res->mp_off + nr), res->n_u = res->status + nr, res->n_v = res->n_u + nr, res->rep_len = res->n_v + nr;
	res->n_mini_pos = res->rep_len + nr, res->n_mini = res->n_mini_pos + nr, res->seg = res->n_mini + nr;
	const size_t ns = call.subs.size();
	priv->segs.resize(ns), priv->seg_u.assign(ns, nullptr), priv->seg_b.assign(ns, nullptr), priv->seg_mp.assign(ns, nullptr);
	for (size_t i = 0; i < ns; ++i) priv->segs[i] = seg_acquire();
	call.res = res, call.priv = priv;
	run_call(call);
	res->n_segs = (int32_t)ns;
	res->seg_u = priv->seg_u.data(), res->seg_b = priv->seg_b.data(), res->seg_mini_pos = priv->seg_mp.data();
	res->tot_mini = call.tot_mini, res->tot_anchors = call.tot_anchors, res->tot_chains = call.tot_chains, res->tot_chained = call.tot_chained, res->n_tie_reads = call.n_tie;
	res->h2d_bytes = call.h2d, res->d2h_bytes = call.d2h, res->cells_ref = call.cells;
	res->sketch_ms = call.sketch_ms, res->seed_ms = call.seed_ms, res->sort_ms = call.sort_ms, res->chain_ms = call.chain_ms;
	if (call.failed.load()) {
		mm2b_map_result_release(res);
		set_error("%s%s", call.err, "");
		return MM2B_ERR_CUDA;
	}
	*out = res;
	return MM2B_OK;
}

void mm2b_map_result_release(mm2b_map_result_t *res)
{
	if (!res) return;
	ResultPriv *priv = (ResultPriv*)res->priv;
	{
		std::lock_guard<std::mutex> lk(g_mu);
		for (Segment *s : priv->segs) g_free_seg.push_back(s);
	}
	free(priv->per_read);
	delete priv;
	free(res);
}

int mm2b_seed_debug(mm2b_index_t *idx, const mm2b_seed_params_t *seed, int64_t n_reads, const int64_t *seq_off, const char *seq,
                    int64_t **mini_off, mm2b_anchor_t **mini, int64_t **a_off, mm2b_anchor_t **anchors,
                    int32_t **rep_len, int32_t **n_mini_pos, uint32_t **mini_pos, int64_t *n_tie_reads)
{
	if (!idx || idx->dev.empty() || !seed || n_reads <= 0 || !seq_off || !seq) { set_error("%s%s", "mm2b_seed_debug: bad argument", ""); return MM2B_ERR_ARG; }
	Call call;
	DebugOut dbg;
	call.idx = idx, call.seed = *seed, call.chain = nullptr, call.seq_off = seq_off, call.seq = seq, call.dbg = &dbg;
	call.subs.push_back(SubBatch{0, n_reads});
	call.res = nullptr, call.priv = nullptr;
	worker(&call, idx->dev[0].device);
	if (call.failed.load()) { set_error("%s%s", call.err, ""); return MM2B_ERR_CUDA; }
	auto dup = [](const void *p, size_t bytes) { void *q = malloc(bytes ? bytes : 1); if (bytes) memcpy(q, p, bytes); return q; };
	if (mini_off) *mini_off = (int64_t*)dup(dbg.mv_off.data(), dbg.mv_off.size() * 8);
	if (mini) *mini = (mm2b_anchor_t*)dup(dbg.mv.data(), dbg.mv.size() * 16);
	if (a_off) *a_off = (int64_t*)dup(dbg.a_off.data(), dbg.a_off.size() * 8);
	if (anchors) *anchors = (mm2b_anchor_t*)dup(dbg.a.data(), dbg.a.size() * 16);
	if (rep_len) *rep_len = (int32_t*)dup(dbg.rep_len.data(), dbg.rep_len.size() * 4);
	if (n_mini_pos) *n_mini_pos = (int32_t*)dup(dbg.n_mini_pos.data(), dbg.n_mini_pos.size() * 4);
	if (mini_pos) *mini_pos = (uint32_t*)dup(dbg.mini_pos.data(), dbg.mini_pos.size() * 4);
	if (n_tie_reads) *n_tie_reads = dbg.n_tie;
	return MM2B_OK;
}

void mm2b_free(void *p) { free(p); }

// called by mm2b_shutdown (chain_backend.cpp): the pooled contexts and segments hold device and pinned memory
void mm2b_map_backend_shutdown(void)
{
	std::lock_guard<std::mutex> lk(g_mu);
	for (Ctx *c : g_free_ctx) { c->destroy(); delete c; }
	g_free_ctx.clear();
	for (Segment *s : g_free_seg) { s->u.release(), s->b.release(), s->mp.release(); delete s; }
	g_free_seg.clear();
}

}  // extern "C"
