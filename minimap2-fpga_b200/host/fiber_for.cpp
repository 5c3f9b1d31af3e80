// kt_for() on fibers: see fiber_for.h.
#include "fiber_for.h"

#include <sys/mman.h>
#include <ucontext.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

namespace mm2b {
namespace {

struct Sched;

struct Fiber {
	enum State { IDLE, RUNNING, PARKED };
	ucontext_t ctx;
	void *stack = nullptr;
	Sched *sched = nullptr;
	int tid = 0;
	long job = -1;                       // index handed to the callback
	State state = IDLE;
	FiberReq *req = nullptr;             // the parked request
};

struct Shared {                          // one kt_for() call
	void (*func)(void*, long, int);
	void *data;
	long n;
	std::atomic<long> next{0};
};

struct Sched {                           // one OS thread and the fibers it multiplexes
	ucontext_t ctx;
	Shared *sh = nullptr;
	std::vector<Fiber*> fibers;
	Fiber *cur = nullptr;
	long n_flushes = 0, n_parked = 0;    // for MM2B_TRACE
};

thread_local Sched *tl_sched = nullptr;
std::atomic<FiberFlushFn> g_flush{nullptr};
std::atomic<FiberSubmitFn> g_submit{nullptr};
std::atomic<FiberWaitFn> g_wait{nullptr};

size_t stack_bytes()
{
	static const size_t v = [] {
		const char *e = getenv("MM2B_FIBER_STACK_KB");
		const long kb = e ? atol(e) : 1024;
		return (size_t)std::max(kb, 64L) << 10;
	}();
	return v;
}

// fiber stacks are recycled between kt_for() calls (one call per mini-batch of reads)
std::mutex g_pool_mu;
std::vector<void*> g_stack_pool;

void *stack_get()
{
	{
		std::lock_guard<std::mutex> lk(g_pool_mu);
		if (!g_stack_pool.empty()) {
			void *p = g_stack_pool.back();
			g_stack_pool.pop_back();
			return p;
		}
	}
	void *p = mmap(nullptr, stack_bytes(), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
	if (p == MAP_FAILED) {
		fprintf(stderr, "[mm2chain_b200] kt_for: cannot map a %zu-byte fiber stack\n", stack_bytes());
		exit(1);
	}
	mprotect(p, 4096, PROT_NONE);        // guard page: running off the stack faults instead of corrupting a neighbour
	return p;
}

void stack_put(void *p)
{
	std::lock_guard<std::mutex> lk(g_pool_mu);
	g_stack_pool.push_back(p);
}

// makecontext() passes ints: the fiber pointer travels in two halves
void fiber_main(unsigned lo, unsigned hi)
{
	Fiber *f = (Fiber*)(((uintptr_t)hi << 32) | (uintptr_t)lo);
	for (;;) {
		f->sched->sh->func(f->sched->sh->data, f->job, f->tid);
		f->state = Fiber::IDLE;
		swapcontext(&f->ctx, &f->sched->ctx);
	}
}

inline void run(Sched &s, Fiber *f, std::vector<Fiber*> &parked)
{
	f->state = Fiber::RUNNING;
	s.cur = f;
	swapcontext(&s.ctx, &f->ctx);        // until the callback returns or parks in fiber_chain()
	s.cur = nullptr;
	if (f->state == Fiber::PARKED) parked.push_back(f);
}

void sched_loop(Sched &s)
{
	tl_sched = &s;
	const FiberSubmitFn submit = g_submit.load();
	const FiberWaitFn wait = g_wait.load();
	const bool async = submit != nullptr && wait != nullptr && s.fibers.size() >= 2;
	// with an asynchronous backend a batch goes out as soon as half of the fibers are parked, and the other half keeps the
	// core busy while it is away; otherwise the batch goes out when nothing else can run
	const size_t submit_at = async ? (s.fibers.size() + 1) / 2 : s.fibers.size();
	std::vector<Fiber*> parked, away;        // parked: waiting to be sent; away: sent, not back yet
	std::vector<FiberReq*> reqs;
	void *ticket = nullptr;
	bool more = true;
	auto send = [&] {
		reqs.clear();
		for (Fiber *f : parked) reqs.push_back(f->req);
		++s.n_flushes, s.n_parked += (long)reqs.size();
		if (async) {
			ticket = submit(reqs.data(), (int)reqs.size());
			away.swap(parked);
			parked.clear();
		} else {
			g_flush.load()(reqs.data(), (int)reqs.size());
			std::vector<Fiber*> batch;
			batch.swap(parked);
			for (Fiber *f : batch) run(s, f, parked);          // a fiber may park again (map.c:338 chains a second time)
		}
	};
	auto receive = [&] {
		wait(ticket);
		ticket = nullptr;
		std::vector<Fiber*> batch;
		batch.swap(away);
		for (Fiber *f : batch) run(s, f, parked);
	};
	for (;;) {
		if (more) {                                            // start work on every idle fiber
			for (Fiber *f : s.fibers) {
				if (f->state != Fiber::IDLE) continue;
				const long i = s.sh->next.fetch_add(1);
				if (i >= s.sh->n) { more = false; break; }
				f->job = i;
				run(s, f, parked);
				if (async && away.empty() && parked.size() >= submit_at) send();
			}
		}
		if (!away.empty()) {                                   // nothing else to start right now: take the batch back
			receive();
			if (away.empty() && !parked.empty() && (parked.size() >= submit_at || !more)) send();
			continue;
		}
		if (parked.empty()) {
			if (!more) break;                                  // nothing parked, nothing away, nothing left to start
			continue;
		}
		send();                                                // every runnable fiber is parked in mm_chain_dp (or the work ran out)
	}
	tl_sched = nullptr;
}

void fiber_init(Fiber &f)                  // (a function of its own: getcontext() is a returns-twice call)
{
	getcontext(&f.ctx);
	f.ctx.uc_stack.ss_sp = f.stack, f.ctx.uc_stack.ss_size = stack_bytes(), f.ctx.uc_link = nullptr;
	const uintptr_t p = (uintptr_t)&f;
	makecontext(&f.ctx, (void (*)())fiber_main, 2, (unsigned)(p & 0xffffffffu), (unsigned)(p >> 32));
}

int os_threads(int n_threads)
{
	const char *e = getenv("MM2B_FIBER_OS_THREADS");
	int hw = e ? atoi(e) : (int)std::thread::hardware_concurrency();
	if (hw < 1) hw = 1;
	return std::min(hw, n_threads);
}

}  // namespace

bool fiber_active() { return g_flush.load() != nullptr && tl_sched != nullptr && tl_sched->cur != nullptr; }

void fiber_set_flush(FiberFlushFn fn) { g_flush.store(fn); }
void fiber_set_async(FiberSubmitFn submit, FiberWaitFn wait) { g_submit.store(submit), g_wait.store(wait); }

void fiber_chain(FiberReq *req)
{
	Fiber *f = tl_sched->cur;
	f->req = req;
	f->state = Fiber::PARKED;
	swapcontext(&f->ctx, &f->sched->ctx);                      // back in sched_loop(); returns here after the flush
}

}  // namespace mm2b

// Same contract as kthread.c:54: func(data, i, tid) for every i in [0, n), tid in [0, n_threads), no two concurrent calls with
// the same tid.  The tids are fibers; min(n_threads, cores) OS threads run them (MM2B_FIBER_OS_THREADS overrides the count).
extern "C" void kt_for(int n_threads, void (*func)(void*, long, int), void *data, long n)
{
	using namespace mm2b;
	if (n_threads <= 1 || n <= 1) {
		for (long j = 0; j < n; ++j) func(data, j, 0);
		return;
	}
	Shared sh;
	sh.func = func, sh.data = data, sh.n = n;
	const int W = os_threads(n_threads);
	std::vector<Sched> scheds(W);
	std::vector<Fiber> fibers(n_threads);
	for (int t = 0; t < n_threads; ++t) {
		Fiber &f = fibers[t];
		Sched &s = scheds[t % W];
		f.sched = &s, f.tid = t, f.stack = stack_get();
		fiber_init(f);
		s.fibers.push_back(&f);
	}
	for (Sched &s : scheds) s.sh = &sh;
	std::vector<std::thread> th;
	for (int w = 1; w < W; ++w) th.emplace_back(sched_loop, std::ref(scheds[w]));
	sched_loop(scheds[0]);
	for (std::thread &t : th) t.join();
	for (Fiber &f : fibers) stack_put(f.stack);
	if (const char *e = getenv("MM2B_TRACE")) {
		if (atoi(e) > 0) {
			long fl = 0, pk = 0;
			for (const Sched &s : scheds) fl += s.n_flushes, pk += s.n_parked;
			fprintf(stderr, "[mm2b trace] kt_for: %ld items on %d fibers / %d OS threads, %ld chaining calls in %ld batches\n", n, n_threads, W, pk, fl);
		}
	}
}
