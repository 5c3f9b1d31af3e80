// Fibers under kt_for(): how the per-read boundary gets batches without touching the reference's mapping code.
//
// The reference maps reads with kt_for(n_threads, worker_for, ...) (map.c:561, kthread.c:54): n_threads OS threads, each
// running mm_map_frag for one read at a time, each blocking in mm_chain_dp (map.c:316).  For an accelerator that is the worst
// shape: one synchronous round trip per read, and "more reads in flight" means more OS threads than cores.
//
// This file provides a kt_for() with the same signature in which the n_threads workers are *fibers* (ucontext) multiplexed
// on as many OS threads as the machine has cores.  Every fiber has its own tid, hence its own mm_tbuf_t / kalloc arena
// (map.c:434, :565-567), exactly as a reference worker thread would.  When a fiber reaches mm_chain_dp, the drop-in parks the
// request and switches to the next fiber; when an OS thread has nothing left to start, it chains all parked requests in
// one batch call and resumes their fibers.  `-t 512` then means 512 reads in flight per batch, not 512 threads.
//
// Integration: compile the reference's kthread.c with -Dkt_for=kt_for_per_thread (a build flag, no source change) and link
// this library; see INTEGRATION.md.  Without that flag the reference's own kt_for is used and nothing here runs.
#pragma once
#include <stdint.h>
#include "mm2chain_b200.h"

namespace mm2b {

struct FiberReq {                        // one parked mm_chain_dp call
	mm2b_params_t par;
	int64_t n;
	const mm2b_anchor_t *a;              // the caller's anchors; stay valid until the fiber is resumed
	// filled by the flush; u / b point into buffers that stay valid until the same OS thread flushes again
	int32_t n_u, n_v, status;
	const uint64_t *u;
	const mm2b_anchor_t *b;
};

// Chains all requests (blocking).  Called on the OS thread that owns the fibers.
typedef void (*FiberFlushFn)(FiberReq **reqs, int n);

// Optional asynchronous pair: submit() starts chaining the requests and returns a ticket at once, wait() blocks until they
// are done.  With it an OS thread keeps running its other fibers while a batch is on the accelerator (two groups of fibers
// take turns).  Both are called on the OS thread that owns the fibers.
typedef void *(*FiberSubmitFn)(FiberReq **reqs, int n);
typedef void (*FiberWaitFn)(void *ticket);
void fiber_set_async(FiberSubmitFn submit, FiberWaitFn wait);

bool fiber_active();                     // is the caller running on one of kt_for()'s fibers?
void fiber_chain(FiberReq *req);         // park the request, run other fibers, return once it has been chained
void fiber_set_flush(FiberFlushFn fn);   // who chains the batches (the backend installs its own at start-up)

}  // namespace mm2b

extern "C" void kt_for(int n_threads, void (*func)(void*, long, int), void *data, long n);
