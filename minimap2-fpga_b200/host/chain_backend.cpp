// Host backend of the B200 chaining offload: the C++ side above the C-ABI device calls.
//
// It replaces the reference's OpenCL host code (/root/reference/chain_hardware.cpp):
//   hardware_init  (chain_hardware.cpp:278-400)  ->  mm2b_init       bind CUDA devices, start one worker thread per device
//   cleanup        (chain_hardware.cpp:403-441)  ->  mm2b_shutdown
//   run_chaining_on_hw (chain_hardware.cpp:27-205, one read, mutex-arbitrated single kernel, f/p out)
//                                                ->  mm2b_chain_batch many reads, sharded over devices, final chains out
//   mm_chain_dp    (chain.c:29)                  ->  mm_chain_dp     same signature; every read goes to the GPU
//
// Reads are independent, so a batch is cut into sub-batches (contiguous read ranges of <= MM2B_SUB_ANCHORS anchors) that the
// per-device worker threads pull from a shared counter; each worker keeps NSLOT sub-batches in flight on separate streams so
// host packing, H2D copies, kernels, D2H copies and host gathering of neighbouring sub-batches overlap.  Several calls may be
// in flight at once: a worker serves all of them.  No collective, no NCCL: nothing is exchanged between devices.
//
// PCIe is what bounds this path (16 B per anchor in, 16 B per chained anchor out in the reference's layout), so both
// directions are slimmed down:
//   in   helper threads pack each sub-batch, on its way into the pinned staging buffer, to 8 B per anchor ({x_lo, y_lo}) plus
//        run-length lists of the high words; a tiny kernel restores mm128_t in HBM
//   out  the kernel returns the INDEX of every chained anchor inside its read (4 B); the host, which still holds a[], gathers
//        b[] from them (or hands the indices to the caller)
#include <cuda_runtime_api.h>
#include <vector_functions.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "mm2chain_b200.h"
#include "../csrc/shim_internal.h"

// kalloc of the host application (kalloc.c); weak so that the library also loads stand-alone (tests, bench)
extern "C" void *kmalloc(void *km, size_t size) __attribute__((weak));
extern "C" void kfree(void *km, void *ptr) __attribute__((weak));

extern "C" void mm2b_map_backend_shutdown(void);     // host/map_backend.cpp: pooled contexts of the seeding front end

namespace {

using mm2b::cuda_ok;
using mm2b::set_error;

constexpr int NSLOT = 6;

template <class T> cudaError_t dmalloc(T **p, size_t bytes) { return cudaMalloc((void**)p, bytes ? bytes : 1); }
template <class T> cudaError_t hmalloc(T **p, size_t bytes, unsigned flags) { return cudaHostAlloc((void**)p, bytes ? bytes : 1, flags); }

double now_ms()
{
	return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void *host_kmalloc(void *km, size_t size)
{
	if (kmalloc) return kmalloc(km, size);
	if (km) { fprintf(stderr, "[mm2b] a kalloc arena was passed but the host application exports no kmalloc\n"); exit(1); }
	return size ? malloc(size) : 0;     // kalloc.c:133-134
}
void host_kfree(void *km, void *p)
{
	if (kfree) { kfree(km, p); return; }
	if (km) { fprintf(stderr, "[mm2b] a kalloc arena was passed but the host application exports no kfree\n"); exit(1); }
	free(p);
}

// ---- helper threads: packing on the way in, gathering on the way out ---------------------------------------------------
// A plain task queue.  The tasks are memory-bound passes over one chunk of a sub-batch; the device workers never run them
// themselves (they only drive the streams), so a slow host cannot stall the copy engines of sub-batches already issued.
struct Pool {
	std::vector<std::thread> th;
	std::mutex mu;
	std::condition_variable cv;
	std::deque<std::function<void()>> q;
	bool stop = false;
	void start(int n)
	{
		stop = false;
		for (int i = 0; i < n; ++i) th.emplace_back([this] { loop(); });
	}
	void loop()
	{
		for (;;) {
			std::function<void()> fn;
			{
				std::unique_lock<std::mutex> lk(mu);
				cv.wait(lk, [&] { return stop || !q.empty(); });
				if (q.empty()) return;
				fn = std::move(q.front());
				q.pop_front();
			}
			fn();
		}
	}
	void submit(std::function<void()> fn)
	{
		{ std::lock_guard<std::mutex> lk(mu); q.push_back(std::move(fn)); }
		cv.notify_one();
	}
	void shutdown()
	{
		{ std::lock_guard<std::mutex> lk(mu); stop = true; }
		cv.notify_all();
		for (auto &t : th) if (t.joinable()) t.join();
		th.clear();
		q.clear();
	}
};

struct Device;

// Device + pinned buffers for one sub-batch in flight
struct Slot {
	int device = -1;
	Device *owner = nullptr;
	// Kernels, input copies and output copies each have their own stream.  On one stream, a sub-batch's input copy is ordered after
	// the slot's previous output copy, and the marker that releases it waits in the device-to-host copy queue behind the other
	// slots' pending outputs; small count copies wait there too.  Measured on the seeding pipeline (profiles/r3g_trace_*.txt).
	cudaStream_t stream = nullptr, in_stream = nullptr, out_stream = nullptr;
	cudaEvent_t ev[6] = {};     // 0 start, 1 h2d done, 2 kernels done, 3 totals on the host, 4 out d2h start, 5 out d2h done
	int64_t *h_tot = nullptr, *d_tot = nullptr;         // mapped: chains / chained anchors of the sub-batch, stored by a kernel
	mm2b_workspace_t *ws = nullptr;
	int64_t cap_anchors = 0, cap_reads = 0, cap_runs = 0;
	int64_t *d_off = nullptr, *d_u_off = nullptr, *d_b_off = nullptr;
	mm2b_anchor_t *d_a = nullptr, *d_b = nullptr;       // d_b only exists once a caller asked for device-side gathering
	uint64_t *d_u = nullptr;
	int32_t *d_bi = nullptr;
	int32_t *d_n_u = nullptr, *d_n_v = nullptr, *d_status = nullptr;
	uint2 *d_lo = nullptr, *d_xruns = nullptr, *d_yruns = nullptr;
	int64_t *h_off = nullptr, *h_u_off = nullptr, *h_b_off = nullptr;   // pinned
	unsigned long long *h_cnt = nullptr;                                // pinned copy of the workspace counters
	uint2 *h_lo = nullptr, *h_xruns = nullptr, *h_yruns = nullptr;      // pinned: packed anchors of the sub-batch
	int32_t *h_bi = nullptr;                                            // pinned: indices coming back when the host gathers b[]
	// the sub-batch currently occupying the slot
	struct Job *job = nullptr;
	int sub = -1;
	int stage = 0;              // 0 free, 1 packing, 2 kernels + counts in flight, 3 outputs in flight, 4 gathering
	std::atomic<int> sig{0};    // set by the stream callbacks: 1 = counts are on the host, 2 = outputs are on the host
	std::atomic<int> host_left{0};      // helper tasks of the current stage still running
	std::atomic<int> pack_overflow{0};  // some chunk had more runs than fit: this sub-batch goes over raw
	std::vector<int> n_xr, n_yr;        // runs found per chunk
	int n_xruns = 0, n_yruns = 0;
	bool packed = false, ring_copied = false;   // ring_copied: the helper threads have already enqueued the copies of the packed words
	double t_host0 = 0;

	bool create(int dev, Device *own)
	{
		device = dev, owner = own;
		if (!cuda_ok(cudaSetDevice(dev), "cudaSetDevice")) return false;
		if (!cuda_ok(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
		static const bool one = getenv("MM2B_ONE_STREAM") && atoi(getenv("MM2B_ONE_STREAM")) > 0;      // the old layout, for comparison
		if (one) in_stream = out_stream = stream;
		else if (!cuda_ok(cudaStreamCreateWithFlags(&in_stream, cudaStreamNonBlocking), "cudaStreamCreate") ||
		         !cuda_ok(cudaStreamCreateWithFlags(&out_stream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
		if (!cuda_ok(cudaHostAlloc((void**)&h_tot, 64, cudaHostAllocMapped | cudaHostAllocPortable), "cudaHostAlloc") ||
		    !cuda_ok(cudaHostGetDevicePointer((void**)&d_tot, h_tot, 0), "cudaHostGetDevicePointer")) return false;
		for (auto &e : ev) if (!cuda_ok(cudaEventCreate(&e), "cudaEventCreate")) return false;
		return true;
	}
	void sync_all() { cudaStreamSynchronize(in_stream), cudaStreamSynchronize(stream), cudaStreamSynchronize(out_stream); }
	void release()
	{
		if (device < 0) return;
		cudaSetDevice(device);
		mm2b_ws_destroy(ws), ws = nullptr;
		cudaFree(d_off), cudaFree(d_u_off), cudaFree(d_b_off), cudaFree(d_a), cudaFree(d_b), cudaFree(d_u), cudaFree(d_bi);
		cudaFree(d_n_u), cudaFree(d_n_v), cudaFree(d_status), cudaFree(d_lo), cudaFree(d_xruns), cudaFree(d_yruns);
		cudaFreeHost(h_off), cudaFreeHost(h_u_off), cudaFreeHost(h_b_off), cudaFreeHost(h_cnt);
		cudaFreeHost(h_lo), cudaFreeHost(h_xruns), cudaFreeHost(h_yruns), cudaFreeHost(h_bi);
		h_cnt = nullptr;
		d_off = d_u_off = d_b_off = nullptr, d_a = d_b = nullptr, d_u = nullptr, d_bi = nullptr, d_n_u = d_n_v = d_status = nullptr;
		d_lo = d_xruns = d_yruns = nullptr, h_lo = h_xruns = h_yruns = nullptr, h_bi = nullptr;
		h_off = h_u_off = h_b_off = nullptr;
		cap_anchors = cap_reads = cap_runs = 0;
	}
	void destroy()
	{
		release();
		if (device < 0) return;
		for (auto &e : ev) if (e) cudaEventDestroy(e);
		if (in_stream && in_stream != stream) cudaStreamDestroy(in_stream);
		if (out_stream && out_stream != stream) cudaStreamDestroy(out_stream);
		if (stream) cudaStreamDestroy(stream);
		if (h_tot) cudaFreeHost(h_tot), h_tot = d_tot = nullptr;
		device = -1;
	}
	bool ensure(int64_t n_anchors, int64_t n_reads)
	{
		if (n_anchors <= cap_anchors && n_reads <= cap_reads) return true;
		const int64_t na = std::max<int64_t>(n_anchors + n_anchors / 4, std::max<int64_t>(cap_anchors, 1024));
		const int64_t nr = std::max<int64_t>(n_reads + n_reads / 4, std::max<int64_t>(cap_reads, 64));
		const int64_t nruns = na / 4 + 64;
		sync_all();
		release();
		ws = mm2b_ws_create(device, na, nr);
		if (!ws) return false;
		bool ok = cuda_ok(dmalloc(&d_off, (nr + 1) * 8), "cudaMalloc") && cuda_ok(dmalloc(&d_u_off, (nr + 1) * 8), "cudaMalloc")
		       && cuda_ok(dmalloc(&d_b_off, (nr + 1) * 8), "cudaMalloc") && cuda_ok(dmalloc(&d_a, na * 16), "cudaMalloc")
		       && cuda_ok(dmalloc(&d_bi, na * 4), "cudaMalloc") && cuda_ok(dmalloc(&d_u, na * 8), "cudaMalloc")
		       && cuda_ok(dmalloc(&d_lo, na * 8), "cudaMalloc") && cuda_ok(dmalloc(&d_xruns, nruns * 8), "cudaMalloc")
		       && cuda_ok(dmalloc(&d_yruns, nruns * 8), "cudaMalloc")
		       && cuda_ok(dmalloc(&d_n_u, nr * 4), "cudaMalloc") && cuda_ok(dmalloc(&d_n_v, nr * 4), "cudaMalloc")
		       && cuda_ok(dmalloc(&d_status, nr * 4), "cudaMalloc")
		       && cuda_ok(hmalloc(&h_off, (nr + 1) * 8, cudaHostAllocPortable), "cudaHostAlloc")
		       && cuda_ok(hmalloc(&h_u_off, (nr + 1) * 8, cudaHostAllocPortable), "cudaHostAlloc")
		       && cuda_ok(hmalloc(&h_b_off, (nr + 1) * 8, cudaHostAllocPortable), "cudaHostAlloc")
		       && cuda_ok(hmalloc(&h_cnt, 64, cudaHostAllocPortable), "cudaHostAlloc")
		       && cuda_ok(hmalloc(&h_lo, na * 8, cudaHostAllocPortable), "cudaHostAlloc")
		       && cuda_ok(hmalloc(&h_xruns, nruns * 8, cudaHostAllocPortable), "cudaHostAlloc")
		       && cuda_ok(hmalloc(&h_yruns, nruns * 8, cudaHostAllocPortable), "cudaHostAlloc")
		       && cuda_ok(hmalloc(&h_bi, na * 4, cudaHostAllocPortable), "cudaHostAlloc");
		if (!ok) return false;
		cap_anchors = na, cap_reads = nr, cap_runs = nruns;
		return true;
	}
	bool ensure_b()
	{
		if (d_b) return true;
		return cuda_ok(dmalloc(&d_b, cap_anchors * 16), "cudaMalloc(d_b)");
	}
};

struct SubBatch { int64_t r0, r1; };     // reads [r0, r1)

struct Job {
	const mm2b_params_t *par;
	int64_t n_reads;
	const int64_t *off;
	const mm2b_anchor_t *a;
	int32_t *n_u, *n_v, *status;
	int64_t *u_off, *b_off;
	uint64_t *u;
	mm2b_anchor_t *b;
	int32_t *bi;
	bool pack = true, device_gather = false;
	std::vector<SubBatch> subs;
	std::atomic<int> next{0};
	std::atomic<int> failed{0};
	char err[512] = {0};
	std::mutex mu;
	std::condition_variable cv;
	int workers_left = 0;       // guarded by mu; the caller leaves when it reaches 0
	// stats
	std::atomic<int64_t> n_chains{0}, n_chained{0}, cells_issued{0}, cells_ref{0}, window_cells{0}, n_general{0}, n_heavy{0}, n_cut{0};
	std::atomic<int64_t> h2d_bytes{0}, d2h_bytes{0}, n_packed{0}, n_raw{0};
	double h2d_ms = 0, kernel_ms = 0, d2h_ms = 0, pack_ms = 0, gather_ms = 0;   // guarded by mu
};

struct ActiveJob { Job *job; int in_flight; bool exhausted; };

struct Device {
	int id = -1;
	Slot slots[NSLOT];
	std::thread worker;
	std::mutex mu;
	std::condition_variable cv;
	std::deque<Job*> queue;
	uint64_t wake = 0;          // bumped (under mu) by stream callbacks and helper tasks: something changed, look again
	bool stop = false;
	void poke()
	{
		{ std::lock_guard<std::mutex> lk(mu); ++wake; }
		cv.notify_one();
	}
};

struct Backend {
	std::vector<Device*> devs;
	std::mutex mu;              // guards init/shutdown
	bool up = false;
	int64_t sub_anchors = 2 << 20;
	int64_t pack_chunk = 256 << 10;     // anchors per helper task (= per staging piece: 2 MB packed; measured best of 16k..256k, profiles/r2b_e2e_sweep.txt)
	bool want_stats = true;
	bool trace = false;
	bool default_pack = true, default_device_gather = true;
	// Packing is a pass over host memory (24 B of host traffic per anchor against 8 B saved on the link).  At most `pack_inflight`
	// sub-batches are being packed at any time; a sub-batch that finds the limit reached goes over raw (16 B/anchor) right away.
	// Measured on this pool's hosts (16 vCPUs, ~160 GB/s of memcpy traffic, 55 GB/s H2D; profiles/r2b_e2e_sweep.txt): packing
	// everything wins (15.4 ms per 100k reads against 19.4 ms raw; limits of 1-2 in flight were slower than either), so that is
	// the default.  0 = never pack.
	int pack_inflight = 99;
	std::atomic<int> packing_now{0};
	// Cache-resident staging: every helper thread packs its chunk into a small pinned piece of its own (two per thread, taking
	// turns) and enqueues that piece's H2D copy itself, so the copy engine reads the packed words out of the CPU's cache: they
	// are never written to nor read from host DRAM, which is what the packing pass is bound by.  0 = one big staging buffer
	// per sub-batch, written with non-temporal stores and copied in one piece.
	bool pack_ring = true;
	std::atomic<int> count_cells{0};
	Pool pool;
	cudaEvent_t trace_ev0[64] = {};
};
// heap-allocated and never destroyed: a host that exits without mm2b_shutdown() (main.c returns early on several error paths; a
// script that forgets) must not run the destructors of joinable std::threads at exit (std::terminate -> abort -> a core dump of a
// process with a CUDA context mapped)
Backend &g = *new Backend();

void job_fail(Job *job)
{
	std::lock_guard<std::mutex> lk(job->mu);
	if (!job->failed.exchange(1)) snprintf(job->err, sizeof(job->err), "%s", mm2b_last_error());
}

void CUDART_CB slot_signal_counts(void *p) { Slot *s = (Slot*)p; s->sig.store(1, std::memory_order_release); s->owner->poke(); }
void CUDART_CB slot_signal_outputs(void *p) { Slot *s = (Slot*)p; s->sig.store(2, std::memory_order_release); s->owner->poke(); }

// ---- packing ----------------------------------------------------------------------------------------------------------
// One chunk [i0, i1) of a sub-batch: low words to lo[], and a run of {first index, high word} whenever a high word differs
// from the previous anchor's (every chunk opens its own runs, so chunks are independent).  Returns false when the runs do not
// fit `cap` entries — then the anchors' high words are too varied for this format and the sub-batch goes over as it is.
struct PackState { uint32_t px, py; int nx, ny; };

// anchors [i, e) one at a time; false when the run lists are full
inline bool pack_scalar(const mm2b_anchor_t *a, int64_t i, int64_t e, uint2 *lo, uint2 *xr, uint2 *yr, int cap, PackState &st)
{
	for (; i < e; ++i) {
		const uint64_t x = a[i].x, y = a[i].y;
		lo[i] = make_uint2((uint32_t)x, (uint32_t)y);
		const uint32_t xh = (uint32_t)(x >> 32), yh = (uint32_t)(y >> 32);
		if (xh != st.px) {
			if (st.nx == cap) return false;
			xr[st.nx++] = make_uint2((uint32_t)i, xh), st.px = xh;
		}
		if (yh != st.py) {
			if (st.ny == cap) return false;
			yr[st.ny++] = make_uint2((uint32_t)i, yh), st.py = yh;
		}
	}
	return true;
}

#if defined(__x86_64__)
// Four anchors per step: the low words leave with one non-temporal 32-byte store (the packed buffer is written once and read
// by the copy engine, so it should not displace the source in the cache nor be read for ownership), the high words are compared
// with their predecessors' in one go and only a change drops to the scalar code.  The pass is bound by host memory bandwidth.
template <bool NT>
__attribute__((target("avx2"))) bool pack_avx2(const mm2b_anchor_t *a, int64_t i, int64_t e, uint2 *lo, uint2 *xr, uint2 *yr, int cap, PackState &st)
{
	while (i < e && ((uintptr_t)(lo + i) & 31)) {                  // up to the first 32-byte boundary of the output
		if (!pack_scalar(a, i, i + 1, lo, xr, yr, cap, st)) return false;
		++i;
	}
	const __m256i pick_lo = _mm256_setr_epi32(0, 2, 4, 6, 0, 2, 4, 6), pick_hi = _mm256_setr_epi32(1, 3, 5, 7, 1, 3, 5, 7);
	for (; i + 4 <= e; i += 4) {
		const __m256i v0 = _mm256_loadu_si256((const __m256i*)(a + i)), v1 = _mm256_loadu_si256((const __m256i*)(a + i + 2));
		const __m256i l0 = _mm256_permutevar8x32_epi32(v0, pick_lo), l1 = _mm256_permutevar8x32_epi32(v1, pick_lo);
		if (NT) _mm256_stream_si256((__m256i*)(lo + i), _mm256_permute2x128_si256(l0, l1, 0x20));
		else _mm256_store_si256((__m256i*)(lo + i), _mm256_permute2x128_si256(l0, l1, 0x20));
		const __m256i h0 = _mm256_permutevar8x32_epi32(v0, pick_hi), h1 = _mm256_permutevar8x32_epi32(v1, pick_hi);
		const __m256i hi = _mm256_permute2x128_si256(h0, h1, 0x20);       // xh0 yh0 xh1 yh1 xh2 yh2 xh3 yh3
		const __m256i first = _mm256_setr_epi32((int)st.px, (int)st.py, 0, 0, 0, 0, 0, 0);
		// the predecessors' high words: hi shifted up by one anchor, the state's words in front
		const __m256i prev = _mm256_blend_epi32(_mm256_permutevar8x32_epi32(hi, _mm256_setr_epi32(0, 1, 0, 1, 2, 3, 4, 5)), first, 0x03);
		if (_mm256_movemask_epi8(_mm256_cmpeq_epi32(hi, prev)) != -1) {   // some high word changes inside these four
			PackState t = st;
			// (the low words are already stored; the scalar pass stores them again, which is harmless)
			if (!pack_scalar(a, i, i + 4, lo, xr, yr, cap, t)) return false;
			st = t;
		}
	}
	if (NT) _mm_sfence();
	return pack_scalar(a, i, e, lo, xr, yr, cap, st);
}
#endif

// nt: the packed words leave with non-temporal stores (a big staging buffer that only the copy engine reads) or with ordinary
// ones (a small staging piece that should still be in the cache when the copy engine comes for it)
bool pack_chunk(const mm2b_anchor_t *a, int64_t i0, int64_t i1, uint2 *lo, uint2 *xr, int &nx, uint2 *yr, int &ny, int cap, bool nt = true)
{
	nx = ny = 0;
	if (i0 >= i1) return true;
	PackState st;
	st.px = ~(uint32_t)(a[i0].x >> 32), st.py = ~(uint32_t)(a[i0].y >> 32), st.nx = st.ny = 0;
	bool ok;
#if defined(__x86_64__)
	static const bool have_avx2 = __builtin_cpu_supports("avx2");
	if (have_avx2) ok = nt ? pack_avx2<true>(a, i0, i1, lo, xr, yr, cap, st) : pack_avx2<false>(a, i0, i1, lo, xr, yr, cap, st);
	else
#endif
	ok = pack_scalar(a, i0, i1, lo, xr, yr, cap, st);
	nx = st.nx, ny = st.ny;
	return ok;
}

int n_chunks_of(int64_t na) { return (int)std::max<int64_t>(1, (na + g.pack_chunk - 1) / g.pack_chunk); }

// the staging pieces of one helper thread (cache-resident staging)
struct Pieces {
	uint2 *buf[2] = {nullptr, nullptr};
	int64_t cap = 0;
	cudaEvent_t ev[2][64] = {};             // per piece and device: recorded behind the piece's last copy
	int busy_dev[2] = {-1, -1};
	int next = 0;
	~Pieces()
	{
		for (int k = 0; k < 2; ++k) {
			cudaFreeHost(buf[k]);
			for (auto &e : ev[k]) if (e) cudaEventDestroy(e);
		}
	}
};

void start_pack(Slot &s, Job *job, int si)
{
	const SubBatch sb = job->subs[si];
	const int64_t a0 = job->off[sb.r0], na = job->off[sb.r1] - a0;
	const int nc = n_chunks_of(na);
	const int cap = (int)(s.cap_runs / nc);         // every chunk owns an equal share of the run buffers
	s.n_xr.assign(nc, 0), s.n_yr.assign(nc, 0);
	s.pack_overflow.store(0);
	s.host_left.store(nc);
	s.t_host0 = now_ms();
	g.packing_now.fetch_add(1);
	const mm2b_anchor_t *src = job->a + a0;
	const bool ring = g.pack_ring && s.device < 64;
	s.ring_copied = ring;
	for (int c = 0; c < nc; ++c) {
		g.pool.submit([&s, src, na, nc, cap, c, ring] {
			const int64_t i0 = na * c / nc, i1 = na * (c + 1) / nc;
			if (!ring) {
				if (!s.pack_overflow.load(std::memory_order_relaxed) &&
				    !pack_chunk(src, i0, i1, s.h_lo, s.h_xruns + (int64_t)c * cap, s.n_xr[c], s.h_yruns + (int64_t)c * cap, s.n_yr[c], cap))
					s.pack_overflow.store(1);
			} else if (!s.pack_overflow.load(std::memory_order_relaxed)) {
				static thread_local Pieces pc;
				const int64_t n = i1 - i0;
				if (n > pc.cap) {
					for (int k = 0; k < 2; ++k) {
						if (pc.busy_dev[k] >= 0) cudaEventSynchronize(pc.ev[k][pc.busy_dev[k]]), pc.busy_dev[k] = -1;
						cudaFreeHost(pc.buf[k]), pc.buf[k] = nullptr;
					}
					pc.cap = std::max<int64_t>(n, g.pack_chunk + 64);
					if (!cuda_ok(hmalloc(&pc.buf[0], pc.cap * 8, cudaHostAllocPortable), "cudaHostAlloc") || !cuda_ok(hmalloc(&pc.buf[1], pc.cap * 8, cudaHostAllocPortable), "cudaHostAlloc"))
						s.pack_overflow.store(2);
				}
				const int k = pc.next ^= 1;
				if (pc.buf[k] && pc.busy_dev[k] >= 0) cudaEventSynchronize(pc.ev[k][pc.busy_dev[k]]), pc.busy_dev[k] = -1;     // its previous copy has left
				if (pc.buf[k] && pack_chunk(src, i0, i1, pc.buf[k] - i0, s.h_xruns + (int64_t)c * cap, s.n_xr[c], s.h_yruns + (int64_t)c * cap, s.n_yr[c], cap, false)) {
					cudaSetDevice(s.device);
					if (!pc.ev[k][s.device]) cudaEventCreateWithFlags(&pc.ev[k][s.device], cudaEventDisableTiming);
					if (cudaMemcpyAsync(s.d_lo + i0, pc.buf[k], (size_t)n * 8, cudaMemcpyHostToDevice, s.in_stream) != cudaSuccess ||
					    cudaEventRecord(pc.ev[k][s.device], s.in_stream) != cudaSuccess) s.pack_overflow.store(2);
					else pc.busy_dev[k] = s.device;
				} else s.pack_overflow.store(1);
			}
			if (s.host_left.fetch_sub(1, std::memory_order_acq_rel) == 1) s.owner->poke();
		});
	}
}

// close the gaps between the chunks' run lists (a few thousand entries); returns false if the sub-batch must go over raw
bool finish_pack(Slot &s, int64_t na)
{
	g.packing_now.fetch_sub(1);
	if (s.pack_overflow.load()) return false;
	const int nc = (int)s.n_xr.size();
	const int cap = (int)(s.cap_runs / nc);
	int nx = s.n_xr[0], ny = s.n_yr[0];
	for (int c = 1; c < nc; ++c) {
		memmove(s.h_xruns + nx, s.h_xruns + (int64_t)c * cap, (size_t)s.n_xr[c] * 8), nx += s.n_xr[c];
		memmove(s.h_yruns + ny, s.h_yruns + (int64_t)c * cap, (size_t)s.n_yr[c] * 8), ny += s.n_yr[c];
	}
	s.n_xruns = nx, s.n_yruns = ny;
	return na == 0 || (nx > 0 && ny > 0);
}

// enqueue H2D + kernels + D2H of the per-read counts for the slot's sub-batch
bool stage_issue(Slot &s, Job *job, int si, bool packed)
{
	const SubBatch sb = job->subs[si];
	const int64_t nr = sb.r1 - sb.r0, a0 = job->off[sb.r0], na = job->off[sb.r1] - a0;
	mm2b_ws_set_counting(s.ws, g.count_cells.load());
	int64_t longest = 0;
	for (int64_t r = 0; r <= nr; ++r) {
		s.h_off[r] = job->off[sb.r0 + r] - a0;
		if (r > 0 && s.h_off[r] - s.h_off[r - 1] > longest) longest = s.h_off[r] - s.h_off[r - 1];
	}
	mm2b_ws_set_longest_read(s.ws, longest);          // lets the device call skip the heavy-read kernel when no read can qualify
	cudaStream_t st = s.stream, sti = s.in_stream;
	s.packed = packed;
	bool ok = cuda_ok(cudaEventRecord(s.ev[0], sti), "cudaEventRecord")
	       && cuda_ok(cudaMemcpyAsync(s.d_off, s.h_off, (nr + 1) * 8, cudaMemcpyHostToDevice, sti), "H2D off");
	int64_t h2d = (nr + 1) * 8;
	if (ok && na > 0) {
		if (packed) {
			ok = (s.ring_copied || cuda_ok(cudaMemcpyAsync(s.d_lo, s.h_lo, (size_t)na * 8, cudaMemcpyHostToDevice, sti), "H2D packed anchors"))
			  && cuda_ok(cudaMemcpyAsync(s.d_xruns, s.h_xruns, (size_t)s.n_xruns * 8, cudaMemcpyHostToDevice, sti), "H2D x runs")
			  && cuda_ok(cudaMemcpyAsync(s.d_yruns, s.h_yruns, (size_t)s.n_yruns * 8, cudaMemcpyHostToDevice, sti), "H2D y runs");
			h2d += na * 8 + ((int64_t)s.n_xruns + s.n_yruns) * 8;
		} else {
			ok = cuda_ok(cudaMemcpyAsync(s.d_a, job->a + a0, (size_t)na * 16, cudaMemcpyHostToDevice, sti), "H2D anchors");
			h2d += na * 16;
		}
	}
	ok = ok && cuda_ok(cudaEventRecord(s.ev[1], sti), "cudaEventRecord")
	        && (sti == st || cuda_ok(cudaStreamWaitEvent(st, s.ev[1], 0), "cudaStreamWaitEvent"));
	if (ok && na > 0 && packed) ok = mm2b_unpack_anchors_device(s.device, na, s.d_lo, s.d_xruns, s.n_xruns, s.d_yruns, s.n_yruns, s.d_a, st) == MM2B_OK;
	if (!ok) return false;
	int rc;
	if (job->device_gather) {
		if (!s.ensure_b()) return false;
		rc = mm2b_chain_batch_device(s.ws, job->par, nr, na, s.d_off, s.d_a, s.d_n_u, s.d_n_v, s.d_status, s.d_u_off, s.d_b_off, s.d_u, s.d_b, st);
	} else {
		rc = mm2b_chain_batch_device_idx(s.ws, job->par, nr, na, s.d_off, s.d_a, s.d_n_u, s.d_n_v, s.d_status, s.d_u_off, s.d_b_off, s.d_u, s.d_bi, st);
	}
	if (rc != MM2B_OK) return false;
	s.sig.store(0);
	// the two totals the host sizes the output copies by are stored into mapped host memory by a kernel; the per-read counts
	// travel with the outputs
	ok = cuda_ok(cudaEventRecord(s.ev[2], st), "cudaEventRecord");
	mm2b::count_launches(mm2b::launch_export_scalars(s.d_tot, s.d_u_off + nr, s.d_b_off + nr, nullptr, st));
	ok = ok && cuda_ok(cudaEventRecord(s.ev[3], st), "cudaEventRecord")
	  && cuda_ok(cudaLaunchHostFunc(st, slot_signal_counts, &s), "cudaLaunchHostFunc");
	job->h2d_bytes += h2d, job->d2h_bytes += nr * 12 + (nr + 1) * 16 + 40;
	if (packed) job->n_packed += 1; else job->n_raw += 1;
	return ok;
}

// the totals are on the host: enqueue the D2H of the per-read counts and of exactly the packed u[] and b[] / bi[] bytes
bool stage_outputs(Slot &s, Job *job)
{
	const SubBatch sb = job->subs[s.sub];
	const int64_t nr = sb.r1 - sb.r0, a0 = job->off[sb.r0];
	const int64_t tot_u = ((volatile int64_t*)s.h_tot)[0], tot_b = ((volatile int64_t*)s.h_tot)[1];
	cudaStream_t st = s.out_stream;
	bool ok = cuda_ok(cudaEventRecord(s.ev[4], st), "cudaEventRecord")
	  && cuda_ok(cudaMemcpyAsync(job->n_u + sb.r0, s.d_n_u, nr * 4, cudaMemcpyDeviceToHost, st), "D2H n_u")
	  && cuda_ok(cudaMemcpyAsync(job->n_v + sb.r0, s.d_n_v, nr * 4, cudaMemcpyDeviceToHost, st), "D2H n_v")
	  && cuda_ok(cudaMemcpyAsync(job->status + sb.r0, s.d_status, nr * 4, cudaMemcpyDeviceToHost, st), "D2H status")
	  && cuda_ok(cudaMemcpyAsync(s.h_u_off, s.d_u_off, (nr + 1) * 8, cudaMemcpyDeviceToHost, st), "D2H u_off")
	  && cuda_ok(cudaMemcpyAsync(s.h_b_off, s.d_b_off, (nr + 1) * 8, cudaMemcpyDeviceToHost, st), "D2H b_off")
	  && cuda_ok(cudaMemcpyAsync(s.h_cnt, mm2b_ws_counters_dev(s.ws), 48, cudaMemcpyDeviceToHost, st), "D2H counters")
	  && (tot_u == 0 || cuda_ok(cudaMemcpyAsync(job->u + a0, s.d_u, (size_t)tot_u * 8, cudaMemcpyDeviceToHost, st), "D2H u"));
	int64_t d2h = tot_u * 8;
	if (ok && tot_b > 0) {
		if (job->device_gather) {
			ok = cuda_ok(cudaMemcpyAsync(job->b + a0, s.d_b, (size_t)tot_b * 16, cudaMemcpyDeviceToHost, st), "D2H b");
			d2h += tot_b * 16;
		} else {
			int32_t *dst = job->bi ? job->bi + a0 : s.h_bi;           // indices straight to the caller, or staged for the host gather
			ok = cuda_ok(cudaMemcpyAsync(dst, s.d_bi, (size_t)tot_b * 4, cudaMemcpyDeviceToHost, st), "D2H bi");
			d2h += tot_b * 4;
		}
	}
	ok = ok && cuda_ok(cudaEventRecord(s.ev[5], st), "cudaEventRecord")
	        && cuda_ok(cudaLaunchHostFunc(st, slot_signal_outputs, &s), "cudaLaunchHostFunc");
	job->n_chains += tot_u, job->n_chained += tot_b, job->d2h_bytes += d2h;
	return ok;
}

// b[] = a[bi[]] per read (chain.c:412-420's copies, done where a[] still lives), in chunks of reads on the helper threads
void start_gather(Slot &s, Job *job)
{
	const SubBatch sb = job->subs[s.sub];
	const int64_t nr = sb.r1 - sb.r0, a0 = job->off[sb.r0], tot_b = s.h_b_off[nr];
	const int nc = (int)std::max<int64_t>(1, std::min<int64_t>(nr, (tot_b + g.pack_chunk - 1) / g.pack_chunk));
	const int32_t *idx = job->bi ? job->bi + a0 : s.h_bi;
	s.host_left.store(nc);
	s.t_host0 = now_ms();
	for (int c = 0; c < nc; ++c) {
		g.pool.submit([&s, job, sb, a0, nr, nc, c, idx] {
			const int64_t r0 = nr * c / nc, r1 = nr * (c + 1) / nc;
			for (int64_t r = r0; r < r1; ++r) {
				const int32_t nv = job->n_v[sb.r0 + r];
				if (nv <= 0) continue;
				const mm2b_anchor_t *src = job->a + job->off[sb.r0 + r];
				const int32_t *ix = idx + s.h_b_off[r];
				mm2b_anchor_t *dst = job->b + a0 + s.h_b_off[r];
				for (int32_t k = 0; k < nv; ++k) dst[k] = src[ix[k]];
			}
			if (s.host_left.fetch_sub(1, std::memory_order_acq_rel) == 1) s.owner->poke();
		});
	}
}

void stage_finish(Slot &s, Job *job)
{
	if (g.trace) {              // MM2B_TRACE=1: timeline of this sub-batch relative to the first event of the job on this slot's device
		float t[6];
		for (int i = 0; i < 6; ++i) cudaEventElapsedTime(&t[i], g.trace_ev0[s.device], s.ev[i]);
		const SubBatch sb = job->subs[s.sub];
		fprintf(stderr, "[mm2b trace] dev %d sub %3d reads %6lld anchors %8lld %s | start %8.3f h2d_done %8.3f kern_done %8.3f cnt_done %8.3f out_start %8.3f out_done %8.3f ms\n",
		        s.device, s.sub, (long long)(sb.r1 - sb.r0), (long long)(job->off[sb.r1] - job->off[sb.r0]), s.packed ? "packed" : "raw   ", t[0], t[1], t[2], t[3], t[4], t[5]);
	}
	if (g.want_stats) {
		float h2d = 0, ker = 0, d2h0 = 0, d2h1 = 0;
		cudaEventElapsedTime(&h2d, s.ev[0], s.ev[1]), cudaEventElapsedTime(&ker, s.ev[1], s.ev[2]);
		cudaEventElapsedTime(&d2h0, s.ev[2], s.ev[3]), cudaEventElapsedTime(&d2h1, s.ev[4], s.ev[5]);
		job->cells_issued += (int64_t)s.h_cnt[0] * 32, job->n_general += (int64_t)s.h_cnt[1], job->cells_ref += (int64_t)s.h_cnt[2], job->window_cells += (int64_t)s.h_cnt[3];
		job->n_heavy += (int64_t)s.h_cnt[4], job->n_cut += (int64_t)s.h_cnt[5];
		std::lock_guard<std::mutex> lk(job->mu);
		job->h2d_ms += h2d, job->kernel_ms += ker, job->d2h_ms += d2h0 + d2h1;
	}
}

// One worker per device.  It never blocks on the GPU or on the helper threads: stream callbacks and helper tasks bump
// `wake`, the worker sleeps on the condition variable in between (no polling loop, no busy core).
void device_worker(Device *d)
{
	cudaSetDevice(d->id);
	std::vector<ActiveJob> active;
	int in_flight = 0;
	uint64_t seen = 0;
	auto release_slot = [&](Slot &s) {
		for (auto &aj : active) if (aj.job == s.job) --aj.in_flight;
		s.stage = 0, s.sub = -1, s.job = nullptr;
		--in_flight;
	};
	auto fail_slot = [&](Slot &s) {              // nothing of this sub-batch may still be in flight towards the caller's buffers
		job_fail(s.job);
		s.sync_all();
		while (s.host_left.load() > 0) std::this_thread::yield();
		release_slot(s);
	};
	for (;;) {
		{
			std::unique_lock<std::mutex> lk(d->mu);
			if (in_flight == 0 && active.empty())
				d->cv.wait(lk, [&] { return d->stop || !d->queue.empty(); });
			else
				d->cv.wait_for(lk, std::chrono::milliseconds(2), [&] { return d->stop || !d->queue.empty() || d->wake != seen; });
			seen = d->wake;
			while (!d->queue.empty()) {
				active.push_back(ActiveJob{d->queue.front(), 0, false});
				d->queue.pop_front();
			}
			if (d->stop && active.empty() && in_flight == 0) return;
		}
		bool progressed = true;
		while (progressed) {
			progressed = false;
			// advance whatever is ready
			for (int k = 0; k < NSLOT; ++k) {
				Slot &s = d->slots[k];
				if (s.stage == 0) continue;
				Job *job = s.job;
				if (s.stage == 1) {                                         // packing
					if (s.host_left.load(std::memory_order_acquire) > 0) continue;
					progressed = true;
					const SubBatch sb = job->subs[s.sub];
					{
						std::lock_guard<std::mutex> lk(job->mu);
						job->pack_ms += now_ms() - s.t_host0;
					}
					const bool packed = finish_pack(s, job->off[sb.r1] - job->off[sb.r0]);
					if (job->failed.load() || !stage_issue(s, job, s.sub, packed)) { fail_slot(s); continue; }
					s.stage = 2;
				} else if (s.stage == 2) {                                  // kernels + counts
					if (s.sig.load(std::memory_order_acquire) < 1) continue;
					progressed = true;
					if (job->failed.load() || !cuda_ok(cudaEventQuery(s.ev[3]), "totals") || !stage_outputs(s, job)) { fail_slot(s); continue; }
					s.stage = 3;
				} else if (s.stage == 3) {                                  // outputs
					if (s.sig.load(std::memory_order_acquire) < 2) continue;
					progressed = true;
					if (job->failed.load() || !cuda_ok(cudaEventQuery(s.ev[5]), "output copy")) { fail_slot(s); continue; }
					stage_finish(s, job);
					const SubBatch sb = job->subs[s.sub];
					{       // sub-batch outputs are packed from the sub-batch's own anchor offset: they always fit there since n_v <= n per read
						const int64_t a0 = job->off[sb.r0];
						for (int64_t r = 0; r < sb.r1 - sb.r0; ++r) job->u_off[sb.r0 + r] = a0 + s.h_u_off[r], job->b_off[sb.r0 + r] = a0 + s.h_b_off[r];
					}
					if (job->b && !job->device_gather && s.h_b_off[sb.r1 - sb.r0] > 0) {
						start_gather(s, job);
						s.stage = 4;
					} else release_slot(s);
				} else {                                                    // gathering
					if (s.host_left.load(std::memory_order_acquire) > 0) continue;
					progressed = true;
					{
						std::lock_guard<std::mutex> lk(job->mu);
						job->gather_ms += now_ms() - s.t_host0;
					}
					release_slot(s);
				}
			}
			// keep the pipeline fed: every free slot gets the next sub-batch of the oldest call that still has some
			for (int k = 0; k < NSLOT && in_flight < NSLOT; ++k) {
				Slot &s = d->slots[k];
				if (s.stage != 0) continue;
				ActiveJob *pick = nullptr;
				int si = -1;
				for (auto &aj : active) {
					if (aj.exhausted) continue;
					if (aj.job->failed.load()) { aj.exhausted = true; continue; }
					si = aj.job->next.fetch_add(1);
					if (si >= (int)aj.job->subs.size()) { aj.exhausted = true; continue; }
					pick = &aj;
					break;
				}
				if (!pick) break;
				Job *job = pick->job;
				if (g.trace && !g.trace_ev0[d->id]) {
					cudaEventCreate(&g.trace_ev0[d->id]);
					cudaEventRecord(g.trace_ev0[d->id], s.stream);
				}
				const SubBatch sb = job->subs[si];
				const int64_t na = job->off[sb.r1] - job->off[sb.r0];
				s.job = job, s.sub = si;
				++pick->in_flight, ++in_flight;
				progressed = true;
				if (!s.ensure(na, sb.r1 - sb.r0)) { s.stage = 2; fail_slot(s); continue; }
				if (job->pack && na > 0 && na < (1ll << 31) && g.packing_now.load() < g.pack_inflight) {
					s.stage = 1;
					start_pack(s, job, si);
				} else {
					s.stage = 2;
					if (!stage_issue(s, job, si, false)) fail_slot(s);
				}
			}
			// calls this device has nothing left to do for: tell the caller (under its lock: it may leave as soon as it sees 0)
			for (size_t i = 0; i < active.size();) {
				ActiveJob &aj = active[i];
				if (aj.exhausted && aj.in_flight == 0) {
					Job *job = aj.job;
					active.erase(active.begin() + (long)i);
					std::lock_guard<std::mutex> lk(job->mu);
					if (--job->workers_left == 0) job->cv.notify_all();
				} else ++i;
			}
		}
	}
}

int parse_device_list(const char *s, std::vector<int> &out)
{
	while (s && *s) {
		char *e;
		long v = strtol(s, &e, 10);
		if (e == s) break;
		out.push_back((int)v);
		s = *e == ',' ? e + 1 : e;
	}
	return (int)out.size();
}

[[noreturn]] void fatal(const char *what)
{
	fprintf(stderr, "[mm2b] fatal: %s: %s\n", what, mm2b_last_error());     // same behaviour as checkError (chain_hardware.cpp:208)
	exit(EXIT_FAILURE);
}

// ---- cross-thread batcher behind mm_chain_dp ------------------------------------------------------------------------
// mm_chain_dp is a synchronous per-read call made by n_threads kt_for workers (map.c:561).  One GPU launch per read would
// be dominated by launch/sync latency, so concurrent callers are aggregated: each caller copies its anchors into the open
// flight's pinned buffer and sleeps; a dispatcher thread per device closes the flight as soon as the GPU is free, runs it as
// ONE device batch and wakes the callers, which gather their own chains from the anchors they still hold (only the 4-byte
// indices come back over PCIe).  While a flight is on the GPU the next one fills,
// so the batch size adapts to the load (1 read with -t 1, hundreds with an oversubscribed -t).  This is the CUDA counterpart
// of the reference's hw_queue / mutex arbitration (chain_hardware.cpp:45-98), which admitted ONE read at a time.
struct Req {
	int64_t n = 0, a_off = 0;
	int32_t n_u = 0, n_v = 0, status = 0;
	int64_t u_off = 0, b_off = 0;
};

struct Flight {
	Slot slot;
	mm2b_anchor_t *h_a = nullptr;
	int32_t *h_bi = nullptr;
	uint64_t *h_u = nullptr;
	int32_t *h_cnt = nullptr;       // 3 x max_reqs: n_u, n_v, status
	int64_t cap = 0;                // anchors
	mm2b_params_t par;
	std::vector<Req*> reqs;
	int64_t used = 0;
	int copies_pending = 0, consumers_pending = 0;
	int state = 0;                  // 0 filling, 1 closed / on the GPU, 2 results ready
	uint64_t epoch = 0;
};

constexpr int FLIGHT_MAX_REQS = 4096;

struct Batcher {
	int dev = -1;
	std::mutex mu;
	std::condition_variable cv_callers, cv_disp;
	Flight fl[2];
	int open = 0;
	bool stop = false;
	std::thread th;
	std::atomic<int64_t> n_flights{0}, n_reqs{0};
};
std::vector<Batcher*> g_batchers;
std::atomic<int> g_batcher_rr{0};

bool flight_grow(Flight &f, int dev, int64_t need)
{
	const int64_t cap = std::max<int64_t>(need + need / 4, 1 << 21);
	cudaSetDevice(dev);
	cudaFreeHost(f.h_a), cudaFreeHost(f.h_bi), cudaFreeHost(f.h_u);
	f.h_a = nullptr, f.h_bi = nullptr, f.h_u = nullptr;
	if (!cuda_ok(hmalloc(&f.h_a, cap * 16, cudaHostAllocPortable), "cudaHostAlloc") || !cuda_ok(hmalloc(&f.h_bi, cap * 4, cudaHostAllocPortable), "cudaHostAlloc") ||
	    !cuda_ok(hmalloc(&f.h_u, cap * 8, cudaHostAllocPortable), "cudaHostAlloc")) return false;
	if (!f.h_cnt && !cuda_ok(hmalloc(&f.h_cnt, FLIGHT_MAX_REQS * 12, cudaHostAllocPortable), "cudaHostAlloc")) return false;
	f.cap = cap;
	return true;
}

bool same_par(const mm2b_params_t &a, const mm2b_params_t &b) { return memcmp(&a, &b, sizeof(a)) == 0; }

void flight_run(Batcher *bt, Flight &f)          // dispatcher thread, no lock held
{
	Slot &s = f.slot;
	cudaSetDevice(bt->dev);
	const int64_t nr = (int64_t)f.reqs.size(), na = f.used;
	if (!s.ensure(na, nr)) fatal("workspace allocation");
	mm2b_ws_set_counting(s.ws, 0);
	for (int64_t r = 0; r < nr; ++r) s.h_off[r] = f.reqs[r]->a_off;
	s.h_off[nr] = na;
	{
		int64_t longest = 0;
		for (int64_t r = 0; r < nr; ++r) longest = std::max<int64_t>(longest, s.h_off[r + 1] - s.h_off[r]);
		mm2b_ws_set_longest_read(s.ws, longest);
	}
	cudaStream_t st = s.stream;
	bool ok = cuda_ok(cudaMemcpyAsync(s.d_off, s.h_off, (nr + 1) * 8, cudaMemcpyHostToDevice, st), "H2D off")
	       && cuda_ok(cudaMemcpyAsync(s.d_a, f.h_a, (size_t)na * 16, cudaMemcpyHostToDevice, st), "H2D anchors");
	if (!ok || mm2b_chain_batch_device_idx(s.ws, &f.par, nr, na, s.d_off, s.d_a, s.d_n_u, s.d_n_v, s.d_status, s.d_u_off, s.d_b_off, s.d_u, s.d_bi, st) != MM2B_OK)
		fatal("enqueue");
	ok = cuda_ok(cudaMemcpyAsync(f.h_cnt, s.d_n_u, nr * 4, cudaMemcpyDeviceToHost, st), "D2H n_u")
	  && cuda_ok(cudaMemcpyAsync(f.h_cnt + FLIGHT_MAX_REQS, s.d_n_v, nr * 4, cudaMemcpyDeviceToHost, st), "D2H n_v")
	  && cuda_ok(cudaMemcpyAsync(f.h_cnt + 2 * FLIGHT_MAX_REQS, s.d_status, nr * 4, cudaMemcpyDeviceToHost, st), "D2H status")
	  && cuda_ok(cudaMemcpyAsync(s.h_u_off, s.d_u_off, (nr + 1) * 8, cudaMemcpyDeviceToHost, st), "D2H u_off")
	  && cuda_ok(cudaMemcpyAsync(s.h_b_off, s.d_b_off, (nr + 1) * 8, cudaMemcpyDeviceToHost, st), "D2H b_off")
	  && cuda_ok(cudaStreamSynchronize(st), "cudaStreamSynchronize");
	if (!ok) fatal("chain");
	const int64_t tot_u = s.h_u_off[nr], tot_b = s.h_b_off[nr];
	ok = (tot_u == 0 || cuda_ok(cudaMemcpyAsync(f.h_u, s.d_u, (size_t)tot_u * 8, cudaMemcpyDeviceToHost, st), "D2H u"))
	  && (tot_b == 0 || cuda_ok(cudaMemcpyAsync(f.h_bi, s.d_bi, (size_t)tot_b * 4, cudaMemcpyDeviceToHost, st), "D2H bi"))
	  && cuda_ok(cudaStreamSynchronize(st), "cudaStreamSynchronize");
	if (!ok) fatal("copy back");
	for (int64_t r = 0; r < nr; ++r) {
		Req *q = f.reqs[r];
		q->n_u = f.h_cnt[r], q->n_v = f.h_cnt[FLIGHT_MAX_REQS + r], q->status = f.h_cnt[2 * FLIGHT_MAX_REQS + r];
		q->u_off = s.h_u_off[r], q->b_off = s.h_b_off[r];
	}
	bt->n_flights += 1, bt->n_reqs += nr;
}

void batcher_loop(Batcher *bt)
{
	cudaSetDevice(bt->dev);
	std::unique_lock<std::mutex> lk(bt->mu);
	for (;;) {
		bt->cv_disp.wait(lk, [&] { return bt->stop || (bt->fl[bt->open].state == 0 && !bt->fl[bt->open].reqs.empty()); });
		if (bt->stop) return;
		Flight &f = bt->fl[bt->open];
		f.state = 1;                    // closed: later callers go to the other flight (and wait there if it is still being read out)
		bt->open ^= 1;
		bt->cv_callers.notify_all();
		bt->cv_disp.wait(lk, [&] { return f.copies_pending == 0; });
		lk.unlock();
		flight_run(bt, f);
		lk.lock();
		f.consumers_pending = (int)f.reqs.size();
		f.state = 2;
		bt->cv_callers.notify_all();
	}
}

// one synchronous read through the batcher; returns the flight holding the results (caller must release it)
Flight *batcher_submit(Batcher *bt, const mm2b_params_t &par, int64_t n, const mm2b_anchor_t *a, Req &req)
{
	std::unique_lock<std::mutex> lk(bt->mu);
	Flight *f = nullptr;
	for (;;) {
		f = &bt->fl[bt->open];
		if (f->state == 0) {
			if (f->reqs.empty() && n > f->cap) { if (!flight_grow(*f, bt->dev, n)) fatal("pinned staging"); }
			if ((f->reqs.empty() || same_par(f->par, par)) && f->used + n <= f->cap && (int)f->reqs.size() < FLIGHT_MAX_REQS) break;
		}
		bt->cv_disp.notify_one();
		bt->cv_callers.wait(lk);
	}
	req.n = n, req.a_off = f->used;
	f->used += n, f->par = par;
	f->reqs.push_back(&req);
	++f->copies_pending;
	const uint64_t epoch = f->epoch;
	lk.unlock();
	bt->cv_disp.notify_one();
	memcpy(f->h_a + req.a_off, a, (size_t)n * 16);
	lk.lock();
	if (--f->copies_pending == 0) bt->cv_disp.notify_one();
	bt->cv_callers.wait(lk, [&] { return f->state == 2 && f->epoch == epoch; });
	return f;
}

void batcher_release(Batcher *bt, Flight *f)
{
	std::lock_guard<std::mutex> lk(bt->mu);
	if (--f->consumers_pending == 0) {
		f->reqs.clear();
		f->used = 0, f->state = 0, ++f->epoch;
		bt->cv_callers.notify_all();
		bt->cv_disp.notify_one();
	}
}

// undo whatever mm2b_init built so far (g.mu held): a failed start-up leaves no device without a worker and no thread behind
// release = false: the process is about to exit (mm2b_shutdown_at_exit): stop and join every thread of the library so that none of them
// is inside CUDA when the runtime goes down, but leave device and pinned memory to the operating system — un-pinning gigabytes of
// staging takes 0.4-1.2 s on the hosts measured (profiles/r3d_cli_wall.txt), which is as long as mapping a mini-batch.
void teardown_locked(bool release = true)
{
	if (release) mm2b_map_backend_shutdown();
	for (Device *d : g.devs) {
		{ std::lock_guard<std::mutex> l2(d->mu); d->stop = true; }
		d->cv.notify_all();
	}
	for (Device *d : g.devs) {
		if (d->worker.joinable()) d->worker.join();
		if (release) {
			for (auto &s : d->slots) s.destroy();
			delete d;
		}
	}
	g.devs.clear();
	for (Batcher *bt : g_batchers) {
		{ std::lock_guard<std::mutex> l2(bt->mu); bt->stop = true; }
		bt->cv_disp.notify_all();
		if (bt->th.joinable()) bt->th.join();
		if (g.trace) fprintf(stderr, "[mm2b trace] batcher dev %d: %lld reads in %lld flights\n", bt->dev, (long long)bt->n_reqs.load(), (long long)bt->n_flights.load());
		if (release) {
			for (auto &f : bt->fl) {
				f.slot.destroy();
				cudaFreeHost(f.h_a), cudaFreeHost(f.h_bi), cudaFreeHost(f.h_u), cudaFreeHost(f.h_cnt);
			}
			delete bt;
		}
	}
	g_batchers.clear();
	g.pool.shutdown();
	if (release) {
		for (auto &e : g.trace_ev0) if (e) { cudaEventDestroy(e); e = nullptr; }
		mm2b_host_pool_trim();
	}
	g.up = false;
}

}  // namespace

extern "C" {

int mm2b_init(int n_devices, const int *devices)
{
	std::lock_guard<std::mutex> lk(g.mu);
	if (g.up) return MM2B_OK;
	// Each in-flight sub-batch owns a stream (6 pipeline slots + 2 batcher flights per device); with the default of 8 hardware
	// queues streams alias and pick up false dependencies (measured: copies stalled behind other sub-batches' kernels).  16 is
	// enough and keeps context creation fast (measured on this pool: 0.3 s at 8, 0.5 s at 16, 1.5-3.4 s at 32).  Only takes
	// effect if the CUDA context does not exist yet, which is the case for the minimap2 CLI and for binding.load().
	setenv("CUDA_DEVICE_MAX_CONNECTIONS", "16", 0);
	if (const char *s = getenv("MM2B_TRACE")) g.trace = atoi(s) > 0;
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	auto lap = [&](const char *what) {
		if (!g.trace) return;
		clock_gettime(CLOCK_MONOTONIC, &t1);
		fprintf(stderr, "[mm2b trace] init: %-28s %.3f s\n", what, (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec));
		t0 = t1;
	};
	int visible = 0;
	if (!cuda_ok(cudaGetDeviceCount(&visible), "cudaGetDeviceCount") || visible <= 0) {
		if (visible <= 0) set_error("%s%s", "mm2b_init: no CUDA device visible", "");
		return MM2B_ERR_CUDA;
	}
	lap("cudaGetDeviceCount");
	std::vector<int> ids;
	if (n_devices > 0 && devices) ids.assign(devices, devices + n_devices);
	else if (n_devices > 0) for (int i = 0; i < n_devices; ++i) ids.push_back(i);
	else if (!parse_device_list(getenv("MM2B_DEVICES"), ids)) for (int i = 0; i < visible; ++i) ids.push_back(i);
	for (int id : ids) if (id < 0 || id >= visible) { set_error("%s%s", "mm2b_init: device id out of range", ""); return MM2B_ERR_ARG; }
	bool use_batcher = true;
	if (const char *s = getenv("MM2B_BATCHER")) use_batcher = atoi(s) != 0;
	if (const char *s = getenv("MM2B_COUNT_CELLS")) g.count_cells.store(atoi(s) > 0);
	if (const char *s = getenv("MM2B_SUB_ANCHORS")) { const long long v = atoll(s); if (v > 0) g.sub_anchors = v; }
	if (const char *s = getenv("MM2B_PACK_CHUNK")) { const long long v = atoll(s); if (v > 0) g.pack_chunk = v; }
	if (const char *s = getenv("MM2B_PACK")) g.default_pack = atoi(s) != 0;
	if (const char *s = getenv("MM2B_PACK_INFLIGHT")) g.pack_inflight = atoi(s);
	if (const char *s = getenv("MM2B_PACK_RING")) g.pack_ring = atoi(s) != 0;
	if (const char *s = getenv("MM2B_GATHER")) g.default_device_gather = strcmp(s, "host") != 0;
	int n_helpers = (int)std::thread::hardware_concurrency() - 2;
	n_helpers = std::max(2, std::min(n_helpers, 16));
	if (const char *s = getenv("MM2B_HOST_THREADS")) { const int v = atoi(s); if (v > 0) n_helpers = std::min(v, 256); }
	bool ok = true;
	for (int id : ids) {
		Device *d = new Device();
		d->id = id;
		g.devs.push_back(d);
		for (auto &s : d->slots) if (!s.create(id, d)) { ok = false; break; }
		if (!ok) break;
	}
	lap("contexts, streams, events");
	if (ok) {
		for (Device *d : g.devs) d->worker = std::thread(device_worker, d);
		g.pool.start(n_helpers);
		if (use_batcher) {
			for (Device *d : g.devs) {
				Batcher *bt = new Batcher();
				bt->dev = d->id;
				g_batchers.push_back(bt);
				for (auto &f : bt->fl) if (!f.slot.create(d->id, d)) { ok = false; break; }
				if (!ok) break;
				bt->th = std::thread(batcher_loop, bt);
			}
		}
	}
	if (!ok) {                       // leave nothing half-built behind: a later call starts from scratch
		char keep[512];
		snprintf(keep, sizeof(keep), "%s", mm2b_last_error());
		teardown_locked();
		set_error("%s%s", keep, "");
		return MM2B_ERR_CUDA;
	}
	lap("worker + helper + batcher threads");
	g.up = true;
	return MM2B_OK;
}

// hardware_init() replacement for hosts that do other start-up work next (the minimap2 CLI loads its index right after,
// main.c:367-371): bring the devices up on a background thread; the first call that needs them waits for it.
// heap-allocated and never destroyed: a host that exits without cleanup() (main.c returns early on several error paths) must
// not run a joinable std::thread's destructor (std::terminate)
static std::thread &g_init_thread = *new std::thread();
static std::mutex g_init_mu;
static int g_init_rc = MM2B_OK;
static std::atomic<size_t> g_reserve_seq_bytes{0};

static void reserve_for_mapping_now(size_t seq_bytes)
{
	if (seq_bytes == 0) return;
	// what host/map_batch.cpp and host/map_backend.cpp will ask for: one block for the staged sequences, and per sub-batch of
	// sequence (MM2B_MAP_SUB_BYTES, 64 MB) the chained anchors (~0.6 B per base) and mini_pos (~0.75 B per base) coming back
	const size_t sub = (size_t)64 << 20;
	mm2b_host_reserve(seq_bytes + seq_bytes / 8 + 4096, 1);
	mm2b_host_reserve(sub, 2 * (int)((seq_bytes + sub - 1) / sub) + 2);
}

int mm2b_init_async(int n_devices, const int *devices)
{
	std::lock_guard<std::mutex> lk(g_init_mu);
	if (g.up || g_init_thread.joinable()) return MM2B_OK;
	int visible = 0;
	if (!cuda_ok(cudaGetDeviceCount(&visible), "cudaGetDeviceCount") || visible <= 0) {      // fail now, loudly, if there is no GPU at all
		if (visible <= 0) set_error("%s%s", "mm2b_init: no CUDA device visible", "");
		return MM2B_ERR_CUDA;
	}
	std::vector<int> ids;
	if (n_devices > 0 && devices) ids.assign(devices, devices + n_devices);
	g_init_thread = std::thread([n_devices, ids] {
		g_init_rc = mm2b_init(n_devices, ids.empty() ? nullptr : ids.data());
		if (g_init_rc == MM2B_OK) reserve_for_mapping_now(g_reserve_seq_bytes.exchange(0));
	});
	return MM2B_OK;
}

void mm2b_reserve_for_mapping(size_t seq_bytes)
{
	std::lock_guard<std::mutex> lk(g_init_mu);
	if (g_init_thread.joinable()) { g_reserve_seq_bytes.store(seq_bytes); return; }     // picked up when the devices are up
	if (g.up) reserve_for_mapping_now(seq_bytes);
}

static int ensure_up(void)
{
	{
		std::lock_guard<std::mutex> lk(g_init_mu);
		if (g_init_thread.joinable()) g_init_thread.join();
	}
	if (g.up) return MM2B_OK;
	if (g_init_rc != MM2B_OK) return g_init_rc;
	return mm2b_init(0, nullptr);
}

void mm2b_shutdown(void)
{
	{
		std::lock_guard<std::mutex> lk0(g_init_mu);
		if (g_init_thread.joinable()) g_init_thread.join();
	}
	std::lock_guard<std::mutex> lk(g.mu);
	if (!g.up) return;
	const double t0 = now_ms();
	teardown_locked();
	if (g.trace) fprintf(stderr, "[mm2b trace] shutdown: %.3f s\n", (now_ms() - t0) * 1e-3);
}

void mm2b_shutdown_at_exit(void)
{
	{
		std::lock_guard<std::mutex> lk0(g_init_mu);
		if (g_init_thread.joinable()) g_init_thread.join();
	}
	std::lock_guard<std::mutex> lk(g.mu);
	if (!g.up) return;
	const double t0 = now_ms();
	teardown_locked(false);
	if (g.trace) fprintf(stderr, "[mm2b trace] shutdown at exit: %.3f s\n", (now_ms() - t0) * 1e-3);
}

int mm2b_num_devices(void) { return g.up ? (int)g.devs.size() : 0; }
// (internal, for host/map_backend.cpp) CUDA id of the i-th bound device; the statistics switch
int mm2b_device_id(int i) { return g.up && i >= 0 && i < (int)g.devs.size() ? g.devs[(size_t)i]->id : -1; }
int mm2b_counting(void) { return g.count_cells.load(); }

double mm2b_measure_host_copy(int n_threads, size_t bytes_per_thread)
{
	if (n_threads < 1) n_threads = 1;
	if (bytes_per_thread < (1u << 20)) bytes_per_thread = 1u << 20;
	std::vector<char*> src((size_t)n_threads), dst((size_t)n_threads);
	for (int t = 0; t < n_threads; ++t) {
		src[t] = (char*)malloc(bytes_per_thread), dst[t] = (char*)malloc(bytes_per_thread);
		if (!src[t] || !dst[t]) return -1.0;
		memset(src[t], 1, bytes_per_thread), memset(dst[t], 2, bytes_per_thread);      // touch every page first
	}
	double best = 0;
	for (int rep = 0; rep < 3; ++rep) {
		std::vector<std::thread> th;
		const double t0 = now_ms();
		for (int t = 0; t < n_threads; ++t) th.emplace_back([&, t] { memcpy(dst[t], src[t], bytes_per_thread); });
		for (auto &x : th) x.join();
		const double gbs = 2.0 * (double)bytes_per_thread * n_threads / ((now_ms() - t0) * 1e-3) / 1e9;     // bytes read + bytes written
		if (gbs > best) best = gbs;
	}
	for (int t = 0; t < n_threads; ++t) free(src[t]), free(dst[t]);
	return best;
}

int mm2b_pack_anchors(const mm2b_anchor_t *a, int64_t n, void *lo, void *xruns, int32_t *n_xruns, void *yruns, int32_t *n_yruns, int32_t cap_runs)
{
	if (n < 0 || n >= (1ll << 31) || (n > 0 && (!a || !lo)) || !xruns || !yruns || !n_xruns || !n_yruns || cap_runs < 1) {
		set_error("%s%s", "mm2b_pack_anchors: bad argument", "");
		return MM2B_ERR_ARG;
	}
	int nx = 0, ny = 0;
	if (!pack_chunk(a, 0, n, (uint2*)lo, (uint2*)xruns, nx, (uint2*)yruns, ny, cap_runs)) {
		set_error("%s%s", "mm2b_pack_anchors: more runs of high words than cap_runs", "");
		return MM2B_ERR_CAPACITY;
	}
	*n_xruns = nx, *n_yruns = ny;
	return MM2B_OK;
}
void mm2b_set_counting(int on) { g.count_cells.store(on != 0); }

int mm2b_chain_batch_ex(const mm2b_params_t *par, int64_t n_reads, const int64_t *off, const mm2b_anchor_t *a,
                        int32_t *n_u, int32_t *n_v, int32_t *status, int64_t *u_off, int64_t *b_off,
                        uint64_t *u, int64_t u_cap, mm2b_anchor_t *b, int32_t *bi, int64_t b_cap, unsigned flags, mm2b_stats_t *stats)
{
	if (!g.up) {
		const int rc = ensure_up();
		if (rc != MM2B_OK) return rc;
	}
	if (!par || n_reads < 0 || !off || !n_u || !n_v || !status || !u_off || !b_off) { set_error("%s%s", "mm2b_chain_batch: NULL argument", ""); return MM2B_ERR_ARG; }
	const int64_t n_anchors = n_reads > 0 ? off[n_reads] : 0;
	if (n_anchors > 0 && (!a || !u || (!b && !bi))) { set_error("%s%s", "mm2b_chain_batch: NULL buffer", ""); return MM2B_ERR_ARG; }
	if (u_cap < n_anchors || b_cap < n_anchors) { set_error("%s%s", "mm2b_chain_batch: u_cap and b_cap must be >= off[n_reads]", ""); return MM2B_ERR_CAPACITY; }
	Job job;
	job.par = par, job.n_reads = n_reads, job.off = off, job.a = a;
	job.n_u = n_u, job.n_v = n_v, job.status = status, job.u_off = u_off, job.b_off = b_off, job.u = u, job.b = b, job.bi = bi;
	job.pack = !(flags & MM2B_F_RAW_INPUT);
	job.device_gather = b && !bi && ((flags & MM2B_F_DEVICE_GATHER) || !(flags & MM2B_F_HOST_GATHER));
	// Cut into sub-batches of <= sub_anchors anchors (a larger single read stands alone).  The first and last few are smaller:
	// the first copy and the last kernel + copy-back are the only stages nothing overlaps with, so they should be short.
	for (int64_t r0 = 0; r0 < n_reads;) {
		int64_t size = g.sub_anchors;
		const int64_t done = off[r0], left = n_anchors - off[r0];
		if (n_anchors > 6 * g.sub_anchors) {
			if (done < g.sub_anchors / 4 || left <= g.sub_anchors / 2) size = g.sub_anchors / 4;
			else if (done < g.sub_anchors || left <= 3 * g.sub_anchors / 2) size = g.sub_anchors / 2;
		}
		const int64_t lim = off[r0] + size;
		int64_t r1 = std::upper_bound(off + r0 + 1, off + n_reads + 1, lim) - off - 1;
		if (r1 <= r0) r1 = r0 + 1;
		if (r1 - r0 > (1 << 20)) r1 = r0 + (1 << 20);
		job.subs.push_back(SubBatch{r0, r1});
		r0 = r1;
	}
	u_off[n_reads] = n_anchors, b_off[n_reads] = n_anchors;
	if (!job.subs.empty()) {
		const int n_workers = (int)std::min<size_t>(g.devs.size(), job.subs.size());
		job.workers_left = n_workers;
		for (int i = 0; i < n_workers; ++i) {
			Device *d = g.devs[i];
			{ std::lock_guard<std::mutex> lk(d->mu); d->queue.push_back(&job); }
			d->cv.notify_one();
		}
		std::unique_lock<std::mutex> lk(job.mu);
		job.cv.wait(lk, [&] { return job.workers_left == 0; });
	}
	if (stats) {
		memset(stats, 0, sizeof(*stats));
		stats->n_reads = n_reads, stats->n_anchors = n_anchors, stats->n_chains = job.n_chains, stats->n_chained = job.n_chained;
		stats->cells_issued = job.cells_issued, stats->cells_ref = job.cells_ref, stats->window_cells = job.window_cells, stats->n_general_reads = job.n_general;
		stats->n_heavy_reads = job.n_heavy, stats->n_cut_reads = job.n_cut;
		stats->h2d_ms = job.h2d_ms, stats->kernel_ms = job.kernel_ms, stats->d2h_ms = job.d2h_ms;
		stats->h2d_bytes = job.h2d_bytes, stats->d2h_bytes = job.d2h_bytes, stats->n_packed_subs = job.n_packed, stats->n_raw_subs = job.n_raw;
		stats->pack_ms = job.pack_ms, stats->gather_ms = job.gather_ms;
	}
	if (job.failed.load()) { set_error("%s%s", job.err, ""); return MM2B_ERR_CUDA; }
	return MM2B_OK;
}

int mm2b_chain_batch(const mm2b_params_t *par, int64_t n_reads, const int64_t *off, const mm2b_anchor_t *a,
                     int32_t *n_u, int32_t *n_v, int32_t *status, int64_t *u_off, int64_t *b_off,
                     uint64_t *u, int64_t u_cap, mm2b_anchor_t *b, int64_t b_cap, mm2b_stats_t *stats)
{
	// b[] wanted as 16-byte anchors: the copy back (16 B per chained anchor) already loads the host's memory system, and packing the
	// input next to it made the call slower and erratic on the hosts measured (20-33 ms per 100k reads packed against 23-24 ms raw,
	// profiles/r2b_e2e_sweep.txt, profiles/r2e_bench.json), so this entry point sends mm128_t as it is unless MM2B_PACK=1 asks otherwise.
	// Callers that take indices (mm2b_chain_batch_ex with bi) get the packed input: there it wins (15-21 ms against 19 ms).
	const char *pe = getenv("MM2B_PACK");
	const bool pack_b = pe && atoi(pe) > 0;
	const unsigned flags = (pack_b && g.default_pack ? 0u : (unsigned)MM2B_F_RAW_INPUT) | (g.default_device_gather ? (unsigned)MM2B_F_DEVICE_GATHER : (unsigned)MM2B_F_HOST_GATHER);
	if (!b && n_reads > 0 && off && off[n_reads] > 0) { set_error("%s%s", "mm2b_chain_batch: NULL buffer", ""); return MM2B_ERR_ARG; }
	return mm2b_chain_batch_ex(par, n_reads, off, a, n_u, n_v, status, u_off, b_off, u, u_cap, b, nullptr, b_cap, flags, stats);
}

mm2b_anchor_t *mm_chain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
                           float gap_scale, int is_cdna, int n_segs, int64_t n, mm2b_anchor_t *a, int *n_u_, uint64_t **_u,
                           void *km, int tid)
{
	(void)tid;
	if (_u) *_u = 0, *n_u_ = 0;
	if (n == 0 || a == 0) {                                           // chain.c:38-41
		host_kfree(km, a);
		return 0;
	}
	if (!g.up && ensure_up() != MM2B_OK) fatal("mm2b_init");
	const mm2b_params_t par = {max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc, is_cdna, n_segs, gap_scale};
	mm2b_anchor_t *b = nullptr;
	if (!g_batchers.empty()) {
		static thread_local int my_batcher = -1;
		if (my_batcher < 0 || my_batcher >= (int)g_batchers.size()) my_batcher = g_batcher_rr.fetch_add(1) % (int)g_batchers.size();
		Batcher *bt = g_batchers[my_batcher];
		Req req;
		Flight *f = batcher_submit(bt, par, n, a, req);
		if (req.status == MM2B_READ_OK) {                             // otherwise chain.c:355-358: NULL, *_u = NULL, *n_u_ = 0
			// chain.c:359, :397: u has the size of the chain-END list in the reference; only its first n_u entries are defined, so
			// an allocation of max(n_u,1) entries is indistinguishable to the caller (map.c:344-379 reads u[0..n_u) and frees it)
			uint64_t *u = (uint64_t*)host_kmalloc(km, (size_t)std::max(req.n_u, 1) * 8);
			b = (mm2b_anchor_t*)host_kmalloc(km, (size_t)req.n_v * 16);   // kmalloc(km, 0) == NULL, as at chain.c:397
			if (req.n_u > 0) memcpy(u, f->h_u + req.u_off, (size_t)req.n_u * 8);
			const int32_t *ix = f->h_bi + req.b_off;
			for (int32_t k = 0; k < req.n_v; ++k) b[k] = a[ix[k]];      // chain.c:412-420: the caller still holds a[]
			*n_u_ = req.n_u, *_u = u;
		}
		batcher_release(bt, f);
		host_kfree(km, a);                                            // chain.c:356 / :421 — `a` is consumed on every path
		return b;
	}
	// MM2B_BATCHER=0: no aggregation across threads; every call is a one-read batch through the sub-batch pipeline
	const int64_t off[2] = {0, n};
	int32_t nu = 0, nv = 0, status = 0;
	int64_t uo[2], bo[2];
	std::vector<uint64_t> u_tmp((size_t)n);
	std::vector<int32_t> bi((size_t)n);
	if (mm2b_chain_batch_ex(&par, 1, off, a, &nu, &nv, &status, uo, bo, u_tmp.data(), n, nullptr, bi.data(), n, MM2B_F_RAW_INPUT, nullptr) != MM2B_OK) fatal("mm2b_chain_batch");
	if (status == MM2B_READ_OK) {
		uint64_t *u = (uint64_t*)host_kmalloc(km, (size_t)std::max(nu, 1) * 8);
		b = (mm2b_anchor_t*)host_kmalloc(km, (size_t)nv * 16);
		if (nu > 0) memcpy(u, u_tmp.data() + uo[0], (size_t)nu * 8);
		for (int32_t k = 0; k < nv; ++k) b[k] = a[bi[(size_t)bo[0] + (size_t)k]];
		*n_u_ = nu, *_u = u;
	}
	host_kfree(km, a);
	return b;
}

}  // extern "C"
