/* TEST INFRASTRUCTURE — multi-threaded CPU replay of a batch of reads through the chaining oracle
 * (or through the reference's own compiled mm_chain_dp when `ref_fn` is given).
 *
 * This is the CPU arm that bench.py times next to the GPU (cpu_baseline / --impl reference):
 * reads are independent, so like the reference's kt_for(worker_for) loop (map.c:561, kthread.c:54-72)
 * each thread pulls the next read from a shared counter and chains it.  Wall time is returned.
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>
#include "chain_oracle.h"

typedef struct {
	const mm2o_params_t *par;
	int64_t n_reads;
	const int64_t *off;
	const mm2o_anchor_t *a;
	mm2o_ref_fn_t ref_fn;
	int32_t *n_u;
	int64_t *n_v;
	uint64_t *u;
	mm2o_anchor_t *b;
	mm2o_anchor_t **a_copy;     /* per read: a private malloc'd copy of its anchors, made BEFORE the clock starts (the reference consumes `a`) */
	volatile int64_t next;
} job_t;

typedef struct { job_t *job; int tid; mm2o_stats_t st; } worker_t;

static void *worker(void *arg)
{
	worker_t *w = (worker_t*)arg;
	job_t *jb = w->job;
	for (;;) {
		const int64_t r = __sync_fetch_and_add(&jb->next, 1);
		if (r >= jb->n_reads) break;
		const int64_t o = jb->off[r], n = jb->off[r + 1] - o;
		int32_t n_u = 0;
		int64_t n_v = 0;
		if (jb->ref_fn) {
			/* the reference consumes `a` (chain.c:421: kfree) and returns kmalloc'd b/u; km == NULL => malloc/free (kalloc.c:133-149).
			 * The copy it consumes was made before the timed region (mm2o_replay), as the seeding stage would have left it. */
			const mm2o_params_t *q = jb->par;
			mm2o_anchor_t *ac = jb->a_copy[r], *b;
			uint64_t *u = 0;
			int i;
			b = jb->ref_fn(q->max_dist_x, q->max_dist_y, q->bw, q->max_skip, q->max_iter, q->min_cnt, q->min_sc, q->gap_scale,
			               q->is_cdna, q->n_segs, n, ac, &n_u, &u, 0, w->tid);
			for (i = 0; i < n_u; ++i) n_v += (int32_t)u[i];
			if (jb->u && n_u > 0) memcpy(jb->u + o, u, (size_t)n_u * 8);
			if (jb->b && n_v > 0) memcpy(jb->b + o, b, (size_t)n_v * 16);
			free(u), free(b);
			w->st.n_anchors += n, w->st.n_chains += n_u, w->st.n_chained += n_v;
		} else if (jb->u && jb->b) {
			mm2o_chain(jb->par, n, jb->a + o, 0, 0, 0, &n_u, jb->u + o, &n_v, jb->b + o, &w->st);
		} else if (n > 0) {
			uint64_t *u = (uint64_t*)malloc((size_t)n * 8);
			mm2o_anchor_t *b = (mm2o_anchor_t*)malloc((size_t)n * 16);
			mm2o_chain(jb->par, n, jb->a + o, 0, 0, 0, &n_u, u, &n_v, b, &w->st);
			free(u), free(b);
		}
		if (jb->n_u) jb->n_u[r] = n_u;
		if (jb->n_v) jb->n_v[r] = n_v;
	}
	return 0;
}

/* Outputs are written at each read's own anchor offset (u[off[r]..], b[off[r]..]; capacity n per read). */
double mm2o_replay(const mm2o_params_t *par, int64_t n_reads, const int64_t *off, const mm2o_anchor_t *a, int n_threads,
                   mm2o_ref_fn_t ref_fn, int32_t *n_u, int64_t *n_v, uint64_t *u, mm2o_anchor_t *b, mm2o_stats_t *st)
{
	job_t jb;
	worker_t *w;
	pthread_t *th;
	struct timespec t0, t1;
	int i;
	if (n_threads < 1) n_threads = 1;
	memset(&jb, 0, sizeof(jb));
	jb.par = par, jb.n_reads = n_reads, jb.off = off, jb.a = a, jb.ref_fn = ref_fn;
	jb.n_u = n_u, jb.n_v = n_v, jb.u = u, jb.b = b;
	w = (worker_t*)calloc((size_t)n_threads, sizeof(worker_t));
	th = (pthread_t*)calloc((size_t)n_threads, sizeof(pthread_t));
	if (ref_fn) {           /* untimed: the per-read input copies the reference will free */
		int64_t r;
		jb.a_copy = (mm2o_anchor_t**)calloc((size_t)(n_reads > 0 ? n_reads : 1), sizeof(mm2o_anchor_t*));
		for (r = 0; r < n_reads; ++r) {
			const int64_t n = off[r + 1] - off[r];
			if (n > 0) { jb.a_copy[r] = (mm2o_anchor_t*)malloc((size_t)n * 16); memcpy(jb.a_copy[r], a + off[r], (size_t)n * 16); }
		}
	}
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (i = 0; i < n_threads; ++i) { w[i].job = &jb, w[i].tid = i; pthread_create(&th[i], 0, worker, &w[i]); }
	for (i = 0; i < n_threads; ++i) pthread_join(th[i], 0);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	if (st) {
		memset(st, 0, sizeof(*st));
		for (i = 0; i < n_threads; ++i) {
			st->cells += w[i].st.cells, st->window_cells += w[i].st.window_cells;
			st->n_anchors += w[i].st.n_anchors, st->n_chains += w[i].st.n_chains, st->n_chained += w[i].st.n_chained;
		}
	}
	free(w), free(th), free(jb.a_copy);
	return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
