/* TEST INFRASTRUCTURE — CPU stand-in for the backend under the fiber-based kt_for() (minimap2-fpga_b200/host/fiber_for.h).
 *
 * oracle/Makefile links the reference CLI with kthread.c's kt_for renamed away, the product's fiber_for.cpp in its place, and
 * this file as mm_chain_dp: on a fiber the call parks exactly like the product's drop-in does, and the batch is "chained"
 * by calling the REFERENCE's own mm_chain_dp (chain.c, compiled as mm_chain_dp_ref) once per parked read.  The resulting
 * binary (oracle/_ref/minimap2-fiber-sw) needs no GPU; its PAF must be byte-identical to the reference's, which checks the
 * scheduler, the tid / kalloc-arena discipline and the park / resume protocol on real mapping runs.  Nothing here ships.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "fiber_for.h"

typedef mm2b_anchor_t mm128_t;

extern "C" {
mm128_t *mm_chain_dp_ref(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
                         float gap_scale, int is_cdna, int n_segs, int64_t n, mm128_t *a, int *n_u_, uint64_t **_u, void *km, int tid);
void *kmalloc(void *km, size_t size);
void kfree(void *km, void *p);
void mm2_ref_hook_fpv(int64_t n, const int32_t *f, const int32_t *p, const int32_t *v) { (void)n, (void)f, (void)p, (void)v; }
}

static void sw_flush(mm2b::FiberReq **reqs, int n)
{
	static thread_local std::vector<void*> keep;      // results of the previous flush of this OS thread
	for (void *p : keep) free(p);
	keep.clear();
	for (int r = 0; r < n; ++r) {
		mm2b::FiberReq *q = reqs[r];
		mm128_t *copy = (mm128_t*)malloc((size_t)q->n * sizeof(mm128_t));      // the reference consumes its input (chain.c:421)
		memcpy(copy, q->a, (size_t)q->n * sizeof(mm128_t));
		int n_u = 0;
		uint64_t *u = 0;
		mm128_t *b = mm_chain_dp_ref(q->par.max_dist_x, q->par.max_dist_y, q->par.bw, q->par.max_skip, q->par.max_iter, q->par.min_cnt, q->par.min_sc,
		                             q->par.gap_scale, q->par.is_cdna, q->par.n_segs, q->n, copy, &n_u, &u, 0 /* km = NULL: malloc */, 0);
		int64_t n_v = 0;
		for (int i = 0; i < n_u; ++i) n_v += (int32_t)u[i];
		q->status = u ? MM2B_READ_OK : MM2B_READ_NO_CHAIN;
		q->n_u = n_u, q->n_v = (int32_t)n_v, q->u = u, q->b = b;
		if (u) keep.push_back(u);
		if (b) keep.push_back(b);
	}
}

/* Asynchronous variant (MM2_FIBER_SHIM_ASYNC=1): the same work on a helper thread per ticket, so that the scheduler's two-group
 * pipelining runs against a backend that really is away while other fibers run. */
#include <future>
struct Ticket { std::future<void> done; std::vector<mm2b::FiberReq*> reqs; std::vector<void*> keep; };
static void *sw_submit(mm2b::FiberReq **reqs, int n)
{
	static thread_local std::vector<Ticket*> old;            // results must outlive the fibers' copies: free two tickets late
	if (old.size() >= 2) {
		for (void *p : old.front()->keep) free(p);
		delete old.front();
		old.erase(old.begin());
	}
	Ticket *t = new Ticket;
	t->reqs.assign(reqs, reqs + n);
	t->done = std::async(std::launch::async, [t] {
		for (mm2b::FiberReq *q : t->reqs) {
			mm128_t *copy = (mm128_t*)malloc((size_t)q->n * sizeof(mm128_t));
			memcpy(copy, q->a, (size_t)q->n * sizeof(mm128_t));
			int n_u = 0;
			uint64_t *u = 0;
			mm128_t *b = mm_chain_dp_ref(q->par.max_dist_x, q->par.max_dist_y, q->par.bw, q->par.max_skip, q->par.max_iter, q->par.min_cnt, q->par.min_sc,
			                             q->par.gap_scale, q->par.is_cdna, q->par.n_segs, q->n, copy, &n_u, &u, 0, 0);
			int64_t n_v = 0;
			for (int i = 0; i < n_u; ++i) n_v += (int32_t)u[i];
			q->status = u ? MM2B_READ_OK : MM2B_READ_NO_CHAIN;
			q->n_u = n_u, q->n_v = (int32_t)n_v, q->u = u, q->b = b;
			if (u) t->keep.push_back(u);
			if (b) t->keep.push_back(b);
		}
	});
	old.push_back(t);
	return t;
}
static void sw_wait(void *ticket) { ((Ticket*)ticket)->done.get(); }

static struct Installer {
	Installer()
	{
		mm2b::fiber_set_flush(sw_flush);
		const char *e = getenv("MM2_FIBER_SHIM_ASYNC");
		if (e && atoi(e) > 0) mm2b::fiber_set_async(sw_submit, sw_wait);
	}
} g_installer;

extern "C" mm128_t *mm_chain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc,
                                float gap_scale, int is_cdna, int n_segs, int64_t n, mm128_t *a, int *n_u_, uint64_t **_u, void *km, int tid)
{
	if (!mm2b::fiber_active() || n == 0 || a == 0)
		return mm_chain_dp_ref(max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc, gap_scale, is_cdna, n_segs, n, a, n_u_, _u, km, tid);
	*_u = 0, *n_u_ = 0;
	mm2b::FiberReq req;
	req.par = mm2b_params_t{max_dist_x, max_dist_y, bw, max_skip, max_iter, min_cnt, min_sc, is_cdna, n_segs, gap_scale};
	req.n = n, req.a = a, req.n_u = req.n_v = 0, req.status = MM2B_READ_NO_CHAIN, req.u = 0, req.b = 0;
	mm2b::fiber_chain(&req);
	kfree(km, a);
	mm128_t *b = 0;
	if (req.status == MM2B_READ_OK) {
		uint64_t *u = (uint64_t*)kmalloc(km, (size_t)(req.n_u > 0 ? req.n_u : 1) * 8);
		b = (mm128_t*)kmalloc(km, (size_t)req.n_v * 16);
		if (req.n_u > 0) memcpy(u, req.u, (size_t)req.n_u * 8);
		if (req.n_v > 0) memcpy(b, req.b, (size_t)req.n_v * 16);
		*n_u_ = req.n_u, *_u = u;
	}
	return b;
}
