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
// H2D copies, kernels and D2H copies of neighbouring sub-batches overlap.  No collective, no NCCL: nothing is exchanged
// between devices.  Outputs land in the caller's arrays in input order.
#include <cuda_runtime_api.h>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "mm2chain_b200.h"
#include "../csrc/shim_internal.h"
#include "fiber_for.h"

// kalloc of the host application (kalloc.c); weak so that the library also loads stand-alone (tests, bench)
extern "C" void *kmalloc(void *km, size_t size) __attribute__((weak));
extern "C" void kfree(void *km, void *ptr) __attribute__((weak));

namespace {

using mm2b::cuda_ok;
using mm2b::set_error;

constexpr int NSLOT = 6;

template <class T> cudaError_t dmalloc(T **p, size_t bytes) { return cudaMalloc((void**)p, bytes ? bytes : 1); }
template <class T> cudaError_t hmalloc(T **p, size_t bytes, unsigned flags) { return cudaHostAlloc((void**)p, bytes ? bytes : 1, flags); }

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

// Device + pinned buffers for one sub-batch in flight
struct Slot {
	int device = -1;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev[6] = {};     // 0 start, 1 h2d done, 2 kernels done, 3 counts d2h done, 4 out d2h start, 5 out d2h done
	mm2b_workspace_t *ws = nullptr;
	int64_t cap_anchors = 0, cap_reads = 0;
	int64_t *d_off = nullptr, *d_u_off = nullptr, *d_b_off = nullptr;
	mm2b_anchor_t *d_a = nullptr, *d_b = nullptr;
	uint64_t *d_u = nullptr;
	int32_t *d_n_u = nullptr, *d_n_v = nullptr, *d_status = nullptr;
	int64_t *h_off = nullptr, *h_u_off = nullptr, *h_b_off = nullptr;   // pinned
	unsigned long long *h_cnt = nullptr;                                // pinned copy of the workspace counters
	// the sub-batch currently occupying the slot
	int sub = -1;
	int stage = 0;              // 0 free, 1 kernels + counts in flight, 2 outputs in flight

	bool create(int dev)
	{
		device = dev;
		if (!cuda_ok(cudaSetDevice(dev), "cudaSetDevice")) return false;
		if (!cuda_ok(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
		for (auto &e : ev) if (!cuda_ok(cudaEventCreate(&e), "cudaEventCreate")) return false;
		return true;
	}
	void release()
	{
		if (device < 0) return;
		cudaSetDevice(device);
		mm2b_ws_destroy(ws), ws = nullptr;
		cudaFree(d_off), cudaFree(d_u_off), cudaFree(d_b_off), cudaFree(d_a), cudaFree(d_b), cudaFree(d_u);
		cudaFree(d_n_u), cudaFree(d_n_v), cudaFree(d_status);
		cudaFreeHost(h_off), cudaFreeHost(h_u_off), cudaFreeHost(h_b_off), cudaFreeHost(h_cnt);
		h_cnt = nullptr;
		d_off = d_u_off = d_b_off = nullptr, d_a = d_b = nullptr, d_u = nullptr, d_n_u = d_n_v = d_status = nullptr;
		h_off = h_u_off = h_b_off = nullptr;
		cap_anchors = cap_reads = 0;
	}
	void destroy()
	{
		release();
		if (device < 0) return;
		for (auto &e : ev) if (e) cudaEventDestroy(e);
		if (stream) cudaStreamDestroy(stream);
		device = -1;
	}
	bool ensure(int64_t n_anchors, int64_t n_reads)
	{
		if (n_anchors <= cap_anchors && n_reads <= cap_reads) return true;
		const int64_t na = std::max<int64_t>(n_anchors + n_anchors / 4, std::max<int64_t>(cap_anchors, 1024));
		const int64_t nr = std::max<int64_t>(n_reads + n_reads / 4, std::max<int64_t>(cap_reads, 64));
		cudaStreamSynchronize(stream);
		release();
		ws = mm2b_ws_create(device, na, nr);
		if (!ws) return false;
		bool ok = cuda_ok(dmalloc(&d_off, (nr + 1) * 8), "cudaMalloc") && cuda_ok(dmalloc(&d_u_off, (nr + 1) * 8), "cudaMalloc")
		       && cuda_ok(dmalloc(&d_b_off, (nr + 1) * 8), "cudaMalloc") && cuda_ok(dmalloc(&d_a, na * 16), "cudaMalloc")
		       && cuda_ok(dmalloc(&d_b, na * 16), "cudaMalloc") && cuda_ok(dmalloc(&d_u, na * 8), "cudaMalloc")
		       && cuda_ok(dmalloc(&d_n_u, nr * 4), "cudaMalloc") && cuda_ok(dmalloc(&d_n_v, nr * 4), "cudaMalloc")
		       && cuda_ok(dmalloc(&d_status, nr * 4), "cudaMalloc")
		       && cuda_ok(hmalloc(&h_off, (nr + 1) * 8, cudaHostAllocPortable), "cudaHostAlloc")
		       && cuda_ok(hmalloc(&h_u_off, (nr + 1) * 8, cudaHostAllocPortable), "cudaHostAlloc")
		       && cuda_ok(hmalloc(&h_b_off, (nr + 1) * 8, cudaHostAllocPortable), "cudaHostAlloc")
		       && cuda_ok(hmalloc(&h_cnt, 40, cudaHostAllocPortable), "cudaHostAlloc");
		if (!ok) return false;
		cap_anchors = na, cap_reads = nr;
		return true;
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
	std::vector<SubBatch> subs;
	std::atomic<int> next{0};
	std::atomic<int> failed{0};
	char err[512] = {0};
	std::mutex mu;
	std::condition_variable cv;
	int workers_left = 0;
	// stats
	std::atomic<int64_t> n_chains{0}, n_chained{0}, cells_issued{0}, cells_ref{0}, window_cells{0}, n_general{0}, n_heavy{0};
	double h2d_ms = 0, kernel_ms = 0, d2h_ms = 0;   // guarded by mu
};

struct Device {
	int id = -1;
	Slot slots[NSLOT];
	std::thread worker;
	std::mutex mu;
	std::condition_variable cv;
	std::deque<Job*> queue;
	bool stop = false;
};

struct Backend {
	std::vector<Device*> devs;
	std::mutex mu;              // guards init/shutdown
	bool up = false;
	int64_t sub_anchors = 2 << 20;
	bool want_stats = true;
	bool trace = false;
	std::atomic<int> count_cells{0};
	bool use_batcher = true;       // MM2B_BATCHER=0: every mm_chain_dp caller drives its own stream instead
	cudaEvent_t trace_ev0[64] = {};
} g;

void job_fail(Job *job)
{
	std::lock_guard<std::mutex> lk(job->mu);
	if (!job->failed.exchange(1)) snprintf(job->err, sizeof(job->err), "%s", mm2b_last_error());
}

// enqueue H2D + kernels + D2H of the per-read counts for sub-batch `si` on `s`
bool stage_issue(Slot &s, Job *job, int si)
{
	const SubBatch sb = job->subs[si];
	const int64_t nr = sb.r1 - sb.r0, a0 = job->off[sb.r0], na = job->off[sb.r1] - a0;
	if (!s.ensure(na, nr)) return false;
	mm2b_ws_set_counting(s.ws, g.count_cells.load());
	int64_t longest = 0;
	for (int64_t r = 0; r <= nr; ++r) {
		s.h_off[r] = job->off[sb.r0 + r] - a0;
		if (r > 0 && s.h_off[r] - s.h_off[r - 1] > longest) longest = s.h_off[r] - s.h_off[r - 1];
	}
	mm2b_ws_set_longest_read(s.ws, longest);          // lets the device call skip the heavy-read kernel when no read can qualify
	cudaStream_t st = s.stream;
	bool ok = cuda_ok(cudaEventRecord(s.ev[0], st), "cudaEventRecord")
	       && cuda_ok(cudaMemcpyAsync(s.d_off, s.h_off, (nr + 1) * 8, cudaMemcpyHostToDevice, st), "H2D off")
	       && (na == 0 || cuda_ok(cudaMemcpyAsync(s.d_a, job->a + a0, (size_t)na * 16, cudaMemcpyHostToDevice, st), "H2D anchors"))
	       && cuda_ok(cudaEventRecord(s.ev[1], st), "cudaEventRecord");
	if (!ok) return false;
	if (mm2b_chain_batch_device(s.ws, job->par, nr, na, s.d_off, s.d_a, s.d_n_u, s.d_n_v, s.d_status, s.d_u_off, s.d_b_off, s.d_u, s.d_b, st) != MM2B_OK)
		return false;
	ok = cuda_ok(cudaEventRecord(s.ev[2], st), "cudaEventRecord")
	  && cuda_ok(cudaMemcpyAsync(job->n_u + sb.r0, s.d_n_u, nr * 4, cudaMemcpyDeviceToHost, st), "D2H n_u")
	  && cuda_ok(cudaMemcpyAsync(job->n_v + sb.r0, s.d_n_v, nr * 4, cudaMemcpyDeviceToHost, st), "D2H n_v")
	  && cuda_ok(cudaMemcpyAsync(job->status + sb.r0, s.d_status, nr * 4, cudaMemcpyDeviceToHost, st), "D2H status")
	  && cuda_ok(cudaMemcpyAsync(s.h_u_off, s.d_u_off, (nr + 1) * 8, cudaMemcpyDeviceToHost, st), "D2H u_off")
	  && cuda_ok(cudaMemcpyAsync(s.h_b_off, s.d_b_off, (nr + 1) * 8, cudaMemcpyDeviceToHost, st), "D2H b_off")
	  && cuda_ok(cudaMemcpyAsync(s.h_cnt, mm2b_ws_counters_dev(s.ws), 40, cudaMemcpyDeviceToHost, st), "D2H counters")
	  && cuda_ok(cudaEventRecord(s.ev[3], st), "cudaEventRecord");
	s.sub = si, s.stage = 1;
	return ok;
}

// counts are on the host: publish offsets, enqueue the D2H of exactly the packed u[] / b[] bytes
bool stage_outputs(Slot &s, Job *job)
{
	const SubBatch sb = job->subs[s.sub];
	const int64_t nr = sb.r1 - sb.r0, a0 = job->off[sb.r0];
	if (!cuda_ok(cudaEventSynchronize(s.ev[3]), "cudaEventSynchronize")) return false;
	const int64_t tot_u = s.h_u_off[nr], tot_b = s.h_b_off[nr];
	// sub-batch outputs are packed from the sub-batch's own anchor offset: they always fit there since n_v <= n per read
	for (int64_t r = 0; r < nr; ++r) job->u_off[sb.r0 + r] = a0 + s.h_u_off[r], job->b_off[sb.r0 + r] = a0 + s.h_b_off[r];
	cudaStream_t st = s.stream;
	bool ok = cuda_ok(cudaEventRecord(s.ev[4], st), "cudaEventRecord")
	       && (tot_u == 0 || cuda_ok(cudaMemcpyAsync(job->u + a0, s.d_u, (size_t)tot_u * 8, cudaMemcpyDeviceToHost, st), "D2H u"))
	       && (tot_b == 0 || cuda_ok(cudaMemcpyAsync(job->b + a0, s.d_b, (size_t)tot_b * 16, cudaMemcpyDeviceToHost, st), "D2H b"))
	       && cuda_ok(cudaEventRecord(s.ev[5], st), "cudaEventRecord");
	job->n_chains += tot_u, job->n_chained += tot_b;
	s.stage = 2;
	return ok;
}

bool stage_finish(Slot &s, Job *job)
{
	if (!cuda_ok(cudaEventSynchronize(s.ev[5]), "cudaEventSynchronize")) return false;
	if (g.trace) {              // MM2B_TRACE=1: timeline of this sub-batch relative to the first event of the job on this slot's device
		float t[6];
		for (int i = 0; i < 6; ++i) cudaEventElapsedTime(&t[i], g.trace_ev0[s.device], s.ev[i]);
		const SubBatch sb = job->subs[s.sub];
		fprintf(stderr, "[mm2b trace] dev %d sub %3d reads %6lld anchors %8lld | start %8.3f h2d_done %8.3f kern_done %8.3f cnt_done %8.3f out_start %8.3f out_done %8.3f ms\n",
		        s.device, s.sub, (long long)(sb.r1 - sb.r0), (long long)(job->off[sb.r1] - job->off[sb.r0]), t[0], t[1], t[2], t[3], t[4], t[5]);
	}
	if (g.want_stats) {
		float h2d = 0, ker = 0, d2h0 = 0, d2h1 = 0;
		cudaEventElapsedTime(&h2d, s.ev[0], s.ev[1]), cudaEventElapsedTime(&ker, s.ev[1], s.ev[2]);
		cudaEventElapsedTime(&d2h0, s.ev[2], s.ev[3]), cudaEventElapsedTime(&d2h1, s.ev[4], s.ev[5]);
		job->cells_issued += (int64_t)s.h_cnt[0] * 32, job->n_general += (int64_t)s.h_cnt[1], job->cells_ref += (int64_t)s.h_cnt[2], job->window_cells += (int64_t)s.h_cnt[3];
		job->n_heavy += (int64_t)s.h_cnt[4];
		std::lock_guard<std::mutex> lk(job->mu);
		job->h2d_ms += h2d, job->kernel_ms += ker, job->d2h_ms += d2h0 + d2h1;
	}
	s.stage = 0, s.sub = -1;
	return true;
}

void run_job_on_device(Device *d, Job *job)
{
	cudaSetDevice(d->id);
	if (g.trace) {
		if (!g.trace_ev0[d->id]) cudaEventCreate(&g.trace_ev0[d->id]);
		cudaEventRecord(g.trace_ev0[d->id], d->slots[0].stream);
	}
	const int n_subs = (int)job->subs.size();
	bool more = true;
	int in_flight = 0;
	for (;;) {
		bool progressed = false;
		// advance whatever is ready, oldest sub-batch first, without blocking: a finished count copy turns into the output
		// copies, a finished output copy frees the slot
		for (int k = 0; k < NSLOT; ++k) {
			Slot &s = d->slots[k];
			if (s.stage == 0) continue;
			if (job->failed.load()) { cudaStreamSynchronize(s.stream); s.stage = 0, --in_flight, progressed = true; continue; }
			const cudaError_t q = cudaEventQuery(s.stage == 1 ? s.ev[3] : s.ev[5]);
			if (q == cudaErrorNotReady) continue;
			if (q != cudaSuccess) { mm2b::cuda_ok(q, "cudaEventQuery"); job_fail(job); s.stage = 0, --in_flight; continue; }
			progressed = true;
			if (s.stage == 1) {
				if (!stage_outputs(s, job)) { job_fail(job); s.stage = 0, --in_flight; }
			} else {
				if (!stage_finish(s, job)) job_fail(job);
				s.stage = 0, --in_flight;
			}
		}
		// keep the copy engines fed: every free slot gets the next sub-batch right away
		while (more && in_flight < NSLOT && !job->failed.load()) {
			const int si = job->next.fetch_add(1);
			if (si >= n_subs) { more = false; break; }
			int k = 0;
			while (d->slots[k].stage != 0) ++k;
			if (!stage_issue(d->slots[k], job, si)) { job_fail(job); d->slots[k].stage = 0; break; }
			++in_flight, progressed = true;
		}
		if (in_flight == 0 && (!more || job->failed.load())) break;
		if (!progressed) {          // nothing ready yet: events of different slots complete in no fixed order, so poll all of them
			struct timespec ts = {0, 20000};                                   // 20 us
			nanosleep(&ts, nullptr);
		}
	}
}

void device_worker(Device *d)
{
	cudaSetDevice(d->id);
	for (;;) {
		Job *job = nullptr;
		{
			std::unique_lock<std::mutex> lk(d->mu);
			d->cv.wait(lk, [&] { return d->stop || !d->queue.empty(); });
			if (d->queue.empty()) return;
			job = d->queue.front();
			d->queue.pop_front();
		}
		run_job_on_device(d, job);
		{
			std::lock_guard<std::mutex> lk(job->mu);
			--job->workers_left;
		}
		job->cv.notify_all();
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

// ---- per-thread single-read path (mm_chain_dp) -------------------------------------------------------------------
struct ThreadCtx {
	Slot slot;
	int32_t *h_cnt = nullptr;           // pinned: n_u, n_v, status
	mm2b_anchor_t *h_a = nullptr, *h_b = nullptr;   // pinned staging, cap_anchors
	uint64_t *h_u = nullptr;
	int64_t h_cap = 0;
	bool live = false;
};
std::mutex g_tctx_mu;
std::vector<ThreadCtx*> g_tctx;
std::atomic<int> g_tctx_rr{0};

ThreadCtx *thread_ctx()
{
	static thread_local ThreadCtx *t = nullptr;
	if (t && t->live) return t;
	t = new ThreadCtx();
	const int dev = g.devs[g_tctx_rr.fetch_add(1) % g.devs.size()]->id;
	if (!t->slot.create(dev)) { fprintf(stderr, "[mm2b] %s\n", mm2b_last_error()); exit(1); }
	hmalloc(&t->h_cnt, 64, cudaHostAllocPortable);
	t->live = true;
	std::lock_guard<std::mutex> lk(g_tctx_mu);
	g_tctx.push_back(t);
	return t;
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
// ONE device batch and wakes the callers, which copy their own results out.  While a flight is on the GPU the next one fills,
// so the batch size adapts to the load (1 read with -t 1, hundreds with an oversubscribed -t).  This is the CUDA counterpart
// of the reference's hw_queue / mutex arbitration (chain_hardware.cpp:45-98), which admitted ONE read at a time.
struct Req {
	int64_t n = 0, a_off = 0;
	int32_t n_u = 0, n_v = 0, status = 0;
	int64_t u_off = 0, b_off = 0;
};

struct Flight {
	Slot slot;
	mm2b_anchor_t *h_a = nullptr, *h_b = nullptr;
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
	cudaFreeHost(f.h_a), cudaFreeHost(f.h_b), cudaFreeHost(f.h_u);
	f.h_a = f.h_b = nullptr, f.h_u = nullptr;
	if (!cuda_ok(hmalloc(&f.h_a, cap * 16, cudaHostAllocPortable), "cudaHostAlloc") || !cuda_ok(hmalloc(&f.h_b, cap * 16, cudaHostAllocPortable), "cudaHostAlloc") ||
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
	if (!ok || mm2b_chain_batch_device(s.ws, &f.par, nr, na, s.d_off, s.d_a, s.d_n_u, s.d_n_v, s.d_status, s.d_u_off, s.d_b_off, s.d_u, s.d_b, st) != MM2B_OK)
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
	  && (tot_b == 0 || cuda_ok(cudaMemcpyAsync(f.h_b, s.d_b, (size_t)tot_b * 16, cudaMemcpyDeviceToHost, st), "D2H b"))
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
	if (const char *s = getenv("MM2B_TRACE")) g.trace = atoi(s) > 0;
	if (const char *s = getenv("MM2B_BATCHER")) g.use_batcher = atoi(s) != 0;
	if (const char *s = getenv("MM2B_COUNT_CELLS")) g.count_cells.store(atoi(s) > 0);
	if (const char *s = getenv("MM2B_SUB_ANCHORS")) { const long long v = atoll(s); if (v > 0) g.sub_anchors = v; }
	for (int id : ids) {
		Device *d = new Device();
		d->id = id;
		for (auto &s : d->slots) if (!s.create(id)) return MM2B_ERR_CUDA;
		g.devs.push_back(d);
	}
	lap("contexts, streams, events");
	for (Device *d : g.devs) d->worker = std::thread(device_worker, d);
	if (g.use_batcher) {
		for (Device *d : g.devs) {
			Batcher *bt = new Batcher();
			bt->dev = d->id;
			for (auto &f : bt->fl) if (!f.slot.create(d->id)) return MM2B_ERR_CUDA;
			bt->th = std::thread(batcher_loop, bt);
			g_batchers.push_back(bt);
		}
	}
	lap("worker + batcher threads");
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
	g_init_thread = std::thread([n_devices, ids] { g_init_rc = mm2b_init(n_devices, ids.empty() ? nullptr : ids.data()); });
	return MM2B_OK;
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
	for (Device *d : g.devs) {
		{ std::lock_guard<std::mutex> l2(d->mu); d->stop = true; }
		d->cv.notify_all();
	}
	for (Device *d : g.devs) {
		if (d->worker.joinable()) d->worker.join();
		for (auto &s : d->slots) s.destroy();
		delete d;
	}
	g.devs.clear();
	for (Batcher *bt : g_batchers) {
		{ std::lock_guard<std::mutex> l2(bt->mu); bt->stop = true; }
		bt->cv_disp.notify_all();
		if (bt->th.joinable()) bt->th.join();
		if (g.trace) fprintf(stderr, "[mm2b trace] batcher dev %d: %lld reads in %lld flights\n", bt->dev, (long long)bt->n_reqs.load(), (long long)bt->n_flights.load());
		for (auto &f : bt->fl) {
			f.slot.destroy();
			cudaFreeHost(f.h_a), cudaFreeHost(f.h_b), cudaFreeHost(f.h_u), cudaFreeHost(f.h_cnt);
		}
		delete bt;
	}
	g_batchers.clear();
	{
		std::lock_guard<std::mutex> l3(g_tctx_mu);
		for (ThreadCtx *t : g_tctx) {
			t->slot.destroy();
			cudaFreeHost(t->h_cnt), cudaFreeHost(t->h_a), cudaFreeHost(t->h_b), cudaFreeHost(t->h_u);
			t->live = false;    // the owning thread re-creates it on next use
		}
		g_tctx.clear();
	}
	g.up = false;
}

int mm2b_num_devices(void) { return g.up ? (int)g.devs.size() : 0; }
void mm2b_set_counting(int on) { g.count_cells.store(on != 0); }

int mm2b_chain_batch(const mm2b_params_t *par, int64_t n_reads, const int64_t *off, const mm2b_anchor_t *a,
                     int32_t *n_u, int32_t *n_v, int32_t *status, int64_t *u_off, int64_t *b_off,
                     uint64_t *u, int64_t u_cap, mm2b_anchor_t *b, int64_t b_cap, mm2b_stats_t *stats)
{
	if (!g.up) {
		const int rc = ensure_up();
		if (rc != MM2B_OK) return rc;
	}
	if (!par || n_reads < 0 || !off || !n_u || !n_v || !status || !u_off || !b_off) { set_error("%s%s", "mm2b_chain_batch: NULL argument", ""); return MM2B_ERR_ARG; }
	const int64_t n_anchors = n_reads > 0 ? off[n_reads] : 0;
	if (n_anchors > 0 && (!a || !u || !b)) { set_error("%s%s", "mm2b_chain_batch: NULL buffer", ""); return MM2B_ERR_ARG; }
	if (u_cap < n_anchors || b_cap < n_anchors) { set_error("%s%s", "mm2b_chain_batch: u_cap and b_cap must be >= off[n_reads]", ""); return MM2B_ERR_CAPACITY; }
	Job job;
	job.par = par, job.n_reads = n_reads, job.off = off, job.a = a;
	job.n_u = n_u, job.n_v = n_v, job.status = status, job.u_off = u_off, job.b_off = b_off, job.u = u, job.b = b;
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
		stats->n_heavy_reads = job.n_heavy;
		stats->h2d_ms = job.h2d_ms, stats->kernel_ms = job.kernel_ms, stats->d2h_ms = job.d2h_ms;
	}
	if (job.failed.load()) { set_error("%s%s", job.err, ""); return MM2B_ERR_CUDA; }
	return MM2B_OK;
}

}  // extern "C"

namespace {

// Batches of the fiber-based kt_for() (fiber_for.h): all reads parked on this OS thread, grouped by chaining arguments, one
// mm2b_chain_batch call per group from pinned staging that belongs to the thread.
struct FiberStage {
	mm2b_anchor_t *a = nullptr, *b = nullptr;
	uint64_t *u = nullptr;
	int64_t *off = nullptr, *u_off = nullptr, *b_off = nullptr;
	int32_t *n_u = nullptr, *n_v = nullptr, *status = nullptr;
	int64_t cap_a = 0, cap_r = 0;
	void reserve(int64_t na, int64_t nr)
	{
		if (na > cap_a) {
			mm2b_host_free(a), mm2b_host_free(b), mm2b_host_free(u);
			cap_a = std::max<int64_t>(na + na / 2, 1 << 20);
			a = (mm2b_anchor_t*)mm2b_host_alloc((size_t)cap_a * 16), b = (mm2b_anchor_t*)mm2b_host_alloc((size_t)cap_a * 16);
			u = (uint64_t*)mm2b_host_alloc((size_t)cap_a * 8);
		}
		if (nr > cap_r) {
			mm2b_host_free(off), mm2b_host_free(u_off), mm2b_host_free(b_off), mm2b_host_free(n_u), mm2b_host_free(n_v), mm2b_host_free(status);
			cap_r = std::max<int64_t>(2 * nr, 1024);
			off = (int64_t*)mm2b_host_alloc((size_t)cap_r * 16), u_off = (int64_t*)mm2b_host_alloc((size_t)cap_r * 16);
			b_off = (int64_t*)mm2b_host_alloc((size_t)cap_r * 16);          // (2 entries per read: every group needs one more than it has reads)
			n_u = (int32_t*)mm2b_host_alloc((size_t)cap_r * 4), n_v = (int32_t*)mm2b_host_alloc((size_t)cap_r * 4);
			status = (int32_t*)mm2b_host_alloc((size_t)cap_r * 4);
		}
		if (!a || !b || !u || !off || !u_off || !b_off || !n_u || !n_v || !status) fatal("pinned staging for kt_for batches");
	}
};

// kt_for()'s OS threads live for one call (one mini-batch of reads); the pinned staging outlives them in a pool
std::mutex g_stage_mu;
std::vector<FiberStage*> g_stage_pool;

struct StageLease {                          // held by a thread for as long as it lives: its results stay valid until its next flush
	FiberStage *st = nullptr;
	FiberStage *get()
	{
		if (!st) {
			std::lock_guard<std::mutex> lk(g_stage_mu);
			if (!g_stage_pool.empty()) st = g_stage_pool.back(), g_stage_pool.pop_back();
		}
		if (!st) st = new FiberStage();
		return st;
	}
	~StageLease()
	{
		if (!st) return;
		std::lock_guard<std::mutex> lk(g_stage_mu);
		g_stage_pool.push_back(st);
	}
};

void fiber_flush(mm2b::FiberReq **reqs, int n)
{
	static thread_local StageLease lease;
	FiberStage &st = *lease.get();
	int64_t na = 0;
	for (int r = 0; r < n; ++r) na += reqs[r]->n;
	st.reserve(na, n);
	std::vector<char> done((size_t)n, 0);
	int64_t a_base = 0, r_base = 0, o_base = 0;
	for (int first = 0; first < n; ++first) {
		if (done[first]) continue;
		const mm2b_params_t par = reqs[first]->par;
		int64_t cnt = 0, ga = 0;
		int64_t *off = st.off + o_base;
		off[0] = 0;
		std::vector<int> members;
		for (int r = first; r < n; ++r) {
			if (done[r] || !same_par(reqs[r]->par, par)) continue;
			done[r] = 1;
			members.push_back(r);
			memcpy(st.a + a_base + ga, reqs[r]->a, (size_t)reqs[r]->n * 16);
			ga += reqs[r]->n;
			off[++cnt] = ga;
		}
		if (mm2b_chain_batch(&par, cnt, off, st.a + a_base, st.n_u + r_base, st.n_v + r_base, st.status + r_base, st.u_off + o_base, st.b_off + o_base,
		                     st.u + a_base, std::max<int64_t>(ga, 1), st.b + a_base, std::max<int64_t>(ga, 1), nullptr) != MM2B_OK) fatal("mm2b_chain_batch");
		for (int64_t k = 0; k < cnt; ++k) {
			mm2b::FiberReq *q = reqs[members[(size_t)k]];
			q->n_u = st.n_u[r_base + k], q->n_v = st.n_v[r_base + k], q->status = st.status[r_base + k];
			q->u = st.u + a_base + st.u_off[o_base + k], q->b = st.b + a_base + st.b_off[o_base + k];
		}
		a_base += ga, r_base += cnt, o_base += cnt + 1;
	}
}

// MM2B_FIBER_ASYNC=1: the blocking batch call runs on a helper thread of the OS thread, which keeps seeding its other fibers
// meanwhile (fiber_for.h: submit / wait).  Off by default until it has been timed on the GPU box.
struct FiberFlusher {
	std::thread th;
	std::mutex mu;
	std::condition_variable cv;
	std::vector<mm2b::FiberReq*> reqs;
	bool busy = false, quit = false;
	void loop()
	{
		std::unique_lock<std::mutex> lk(mu);
		for (;;) {
			cv.wait(lk, [&] { return busy || quit; });
			if (quit) return;
			lk.unlock();
			fiber_flush(reqs.data(), (int)reqs.size());
			lk.lock();
			busy = false;
			cv.notify_all();
		}
	}
	~FiberFlusher()
	{
		if (!th.joinable()) return;
		{ std::lock_guard<std::mutex> lk(mu); quit = true; }
		cv.notify_all();
		th.join();
	}
};

void *fiber_submit(mm2b::FiberReq **reqs, int n)
{
	static thread_local FiberFlusher fl;     // dies with the OS thread (end of the kt_for() call): joins its helper
	if (!fl.th.joinable()) fl.th = std::thread([p = &fl] { p->loop(); });
	{
		std::lock_guard<std::mutex> lk(fl.mu);
		fl.reqs.assign(reqs, reqs + n);
		fl.busy = true;
	}
	fl.cv.notify_all();
	return &fl;
}

void fiber_wait(void *ticket)
{
	FiberFlusher *fl = (FiberFlusher*)ticket;
	std::unique_lock<std::mutex> lk(fl->mu);
	fl->cv.wait(lk, [&] { return !fl->busy; });
}

struct FiberFlushInstaller {
	FiberFlushInstaller()
	{
		mm2b::fiber_set_flush(fiber_flush);
		const char *e = getenv("MM2B_FIBER_ASYNC");
		if (e && atoi(e) > 0) mm2b::fiber_set_async(fiber_submit, fiber_wait);
	}
} g_fiber_flush_installer;

}  // namespace

extern "C" {

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
	if (mm2b::fiber_active()) {                                       // under the fiber-based kt_for(): park, get chained with the others
		mm2b::FiberReq req;
		req.par = mm2b_params_t{max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc, is_cdna, n_segs, gap_scale};
		req.n = n, req.a = a, req.n_u = req.n_v = 0, req.status = MM2B_READ_NO_CHAIN, req.u = nullptr, req.b = nullptr;
		mm2b::fiber_chain(&req);
		host_kfree(km, a);                                            // chain.c:356 / :421 — `a` is consumed on every path
		mm2b_anchor_t *b = nullptr;
		if (req.status == MM2B_READ_OK) {
			uint64_t *u = (uint64_t*)host_kmalloc(km, (size_t)std::max(req.n_u, 1) * 8);
			b = (mm2b_anchor_t*)host_kmalloc(km, (size_t)req.n_v * 16);
			if (req.n_u > 0) memcpy(u, req.u, (size_t)req.n_u * 8);
			if (req.n_v > 0) memcpy(b, req.b, (size_t)req.n_v * 16);
			*n_u_ = req.n_u, *_u = u;
		}
		return b;
	}
	if (g.use_batcher) {
		static thread_local int my_batcher = -1;
		if (my_batcher < 0 || my_batcher >= (int)g_batchers.size()) my_batcher = g_batcher_rr.fetch_add(1) % (int)g_batchers.size();
		Batcher *bt = g_batchers[my_batcher];
		const mm2b_params_t par = {max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc, is_cdna, n_segs, gap_scale};
		Req req;
		Flight *f = batcher_submit(bt, par, n, a, req);
		host_kfree(km, a);                                            // chain.c:356 / :421 — `a` is consumed on every path
		mm2b_anchor_t *b = nullptr;
		if (req.status == MM2B_READ_OK) {                             // otherwise chain.c:355-358: NULL, *_u = NULL, *n_u_ = 0
			uint64_t *u = (uint64_t*)host_kmalloc(km, (size_t)std::max(req.n_u, 1) * 8);
			b = (mm2b_anchor_t*)host_kmalloc(km, (size_t)req.n_v * 16);   // kmalloc(km, 0) == NULL, as at chain.c:397
			if (req.n_u > 0) memcpy(u, f->h_u + req.u_off, (size_t)req.n_u * 8);
			if (req.n_v > 0) memcpy(b, f->h_b + req.b_off, (size_t)req.n_v * 16);
			*n_u_ = req.n_u, *_u = u;
		}
		batcher_release(bt, f);
		return b;
	}
	ThreadCtx *t = thread_ctx();
	Slot &s = t->slot;
	cudaSetDevice(s.device);
	if (!s.ensure(n, 1)) fatal("workspace allocation");
	if (n > t->h_cap) {
		cudaFreeHost(t->h_a), cudaFreeHost(t->h_b), cudaFreeHost(t->h_u);
		t->h_cap = s.cap_anchors;
		if (!cuda_ok(hmalloc(&t->h_a, t->h_cap * 16, cudaHostAllocPortable), "cudaHostAlloc") ||
		    !cuda_ok(hmalloc(&t->h_b, t->h_cap * 16, cudaHostAllocPortable), "cudaHostAlloc") ||
		    !cuda_ok(hmalloc(&t->h_u, t->h_cap * 8, cudaHostAllocPortable), "cudaHostAlloc")) fatal("pinned staging");
	}
	const mm2b_params_t par = {max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc, is_cdna, n_segs, gap_scale};
	memcpy(t->h_a, a, (size_t)n * 16);
	s.h_off[0] = 0, s.h_off[1] = n;
	mm2b_ws_set_longest_read(s.ws, n);
	cudaStream_t st = s.stream;
	bool ok = cuda_ok(cudaMemcpyAsync(s.d_off, s.h_off, 16, cudaMemcpyHostToDevice, st), "H2D off")
	       && cuda_ok(cudaMemcpyAsync(s.d_a, t->h_a, (size_t)n * 16, cudaMemcpyHostToDevice, st), "H2D anchors");
	if (!ok || mm2b_chain_batch_device(s.ws, &par, 1, n, s.d_off, s.d_a, s.d_n_u, s.d_n_v, s.d_status, s.d_u_off, s.d_b_off, s.d_u, s.d_b, st) != MM2B_OK)
		fatal("enqueue");
	ok = cuda_ok(cudaMemcpyAsync(t->h_cnt, s.d_n_u, 4, cudaMemcpyDeviceToHost, st), "D2H")
	  && cuda_ok(cudaMemcpyAsync(t->h_cnt + 1, s.d_n_v, 4, cudaMemcpyDeviceToHost, st), "D2H")
	  && cuda_ok(cudaMemcpyAsync(t->h_cnt + 2, s.d_status, 4, cudaMemcpyDeviceToHost, st), "D2H")
	  && cuda_ok(cudaStreamSynchronize(st), "cudaStreamSynchronize");
	if (!ok) fatal("chain");
	const int n_u = t->h_cnt[0], n_v = t->h_cnt[1], status = t->h_cnt[2];
	if (n_u > 0) {
		ok = cuda_ok(cudaMemcpyAsync(t->h_u, s.d_u, (size_t)n_u * 8, cudaMemcpyDeviceToHost, st), "D2H u")
		  && cuda_ok(cudaMemcpyAsync(t->h_b, s.d_b, (size_t)n_v * 16, cudaMemcpyDeviceToHost, st), "D2H b")
		  && cuda_ok(cudaStreamSynchronize(st), "cudaStreamSynchronize");
		if (!ok) fatal("copy back");
	}
	host_kfree(km, a);                                                // chain.c:356 / :421 — `a` is consumed on every path
	if (status != MM2B_READ_OK) return 0;                             // chain.c:355-358
	// chain.c:359, :397: u has the size of the chain-END list in the reference; only its first n_u entries are defined, so
	// an allocation of max(n_u,1) entries is indistinguishable to the caller (map.c:344-379 reads u[0..n_u) and frees it)
	uint64_t *u = (uint64_t*)host_kmalloc(km, (size_t)std::max(n_u, 1) * 8);
	mm2b_anchor_t *b = (mm2b_anchor_t*)host_kmalloc(km, (size_t)n_v * 16);      // kmalloc(km, 0) == NULL, as at chain.c:397
	if (n_u > 0) memcpy(u, t->h_u, (size_t)n_u * 8);
	if (n_v > 0) memcpy(b, t->h_b, (size_t)n_v * 16);
	*n_u_ = n_u, *_u = u;
	return b;
}

}  // extern "C"
