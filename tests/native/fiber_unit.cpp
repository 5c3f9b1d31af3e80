// Unit test of the fiber-based kt_for() (minimap2-fpga_b200/host/fiber_for.cpp) without minimap2: a synthetic callback that
// "chains" 0-3 times per item with varying request sizes, against a flush that squares numbers.  Checks that every item runs
// exactly once, that a tid is never used by two items at once, that results reach the right fiber, and that nothing is left
// parked; for the synchronous and the asynchronous (submit / wait) protocol.  Built and run by tests/test_fiber_kt_for.py.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <thread>
#include <vector>

#include "fiber_for.h"

namespace {

struct Job {
	std::vector<std::atomic<int>> runs, tid_busy;
	std::vector<long> result;
	std::atomic<long> chained{0}, bad{0};
	int n_threads;
	Job(long n, int t) : runs(n), tid_busy(t), result(n, 0), n_threads(t) {}
};

// the "backend": u[0] = a[0].x squared, one output word per request, kept alive until the same thread flushes again
void square(mm2b::FiberReq **reqs, int n, std::vector<uint64_t> &out)
{
	out.resize((size_t)n);
	for (int r = 0; r < n; ++r) {
		out[(size_t)r] = reqs[r]->a[0].x * reqs[r]->a[0].x;
		reqs[r]->n_u = 1, reqs[r]->n_v = 0, reqs[r]->status = MM2B_READ_OK, reqs[r]->u = &out[(size_t)r], reqs[r]->b = nullptr;
	}
}

void flush_sync(mm2b::FiberReq **reqs, int n)
{
	static thread_local std::vector<uint64_t> out;
	square(reqs, n, out);
}

struct Ticket { std::future<void> done; std::vector<mm2b::FiberReq*> reqs; std::vector<uint64_t> out; };

void *submit_async(mm2b::FiberReq **reqs, int n)
{
	static thread_local std::vector<Ticket*> old;
	if (old.size() >= 2) { delete old.front(); old.erase(old.begin()); }
	Ticket *t = new Ticket;
	t->reqs.assign(reqs, reqs + n);
	t->done = std::async(std::launch::async, [t] {
		std::this_thread::sleep_for(std::chrono::microseconds(200));      // the backend really is away for a while
		square(t->reqs.data(), (int)t->reqs.size(), t->out);
	});
	old.push_back(t);
	return t;
}
void wait_async(void *t) { ((Ticket*)t)->done.get(); }

void work(void *data, long i, int tid)
{
	Job *j = (Job*)data;
	if (tid < 0 || tid >= j->n_threads || j->tid_busy[(size_t)tid].fetch_add(1) != 0) j->bad++;      // one item per tid at a time
	j->runs[(size_t)i]++;
	long acc = 0;
	const int n_calls = (int)((i * 2654435761u >> 7) % 4);               // 0..3 chaining calls for this item
	for (int c = 0; c < n_calls; ++c) {
		if (!mm2b::fiber_active()) { j->bad++; break; }
		mm2b_anchor_t a = {(uint64_t)(i + 1000 * c), 0};
		mm2b::FiberReq req;
		memset(&req, 0, sizeof(req));
		req.n = 1, req.a = &a;
		volatile char pad[2048];                                          // some stack in use across the switch
		pad[0] = (char)i, pad[2047] = (char)c;
		mm2b::fiber_chain(&req);
		if (req.status != MM2B_READ_OK || req.n_u != 1 || req.u[0] != a.x * a.x || pad[0] != (char)i || pad[2047] != (char)c) j->bad++;
		acc += (long)req.u[0];
		j->chained++;
	}
	j->result[(size_t)i] = acc;
	j->tid_busy[(size_t)tid].fetch_sub(1);
}

int run_case(long n, int n_threads, bool async)
{
	if (async) mm2b::fiber_set_async(submit_async, wait_async);
	else mm2b::fiber_set_async(nullptr, nullptr);
	Job j(n, n_threads);
	kt_for(n_threads, work, &j, n);
	long bad = j.bad.load(), expect_calls = 0;
	for (long i = 0; i < n; ++i) {
		if (j.runs[(size_t)i].load() != 1) ++bad;
		const int n_calls = (int)((i * 2654435761u >> 7) % 4);
		long acc = 0;
		for (int c = 0; c < n_calls; ++c) acc += (i + 1000 * c) * (i + 1000 * c);
		if (n_threads > 1 && n > 1 && j.result[(size_t)i] != acc) ++bad;
		expect_calls += n_calls;
	}
	if (n_threads > 1 && n > 1 && j.chained.load() != expect_calls) ++bad;
	printf("n=%ld threads=%d %s: %s\n", n, n_threads, async ? "async" : "sync", bad ? "FAILED" : "ok");
	return bad ? 1 : 0;
}

}  // namespace

int main()
{
	mm2b::fiber_set_flush(flush_sync);
	int rc = 0;
	for (int async = 0; async < 2; ++async)
		for (long n : {0L, 1L, 2L, 7L, 100L, 5000L})
			for (int t : {2, 3, 16, 257})
				rc |= run_case(n, t, async != 0);
	return rc;
}
