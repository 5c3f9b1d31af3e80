// Phase-split mapping of a mini-batch: SURVEY.md 8(f) next-3, the batched caller of the chaining backend inside minimap2 itself.
//
// The reference maps a mini-batch with kt_for(n_threads, worker_for, step, n_frag) (map.c:561): every worker runs mm_map_frag
// (map.c:272-392) for one read at a time and blocks in mm_chain_dp (map.c:316) — for an accelerator the worst shape, one
// synchronous round trip per read.  Here the same mini-batch goes through three phases instead:
//
//   A  seed      kt_for over the reads: the first half of mm_map_frag — sketch, seed hits, sort (map.c:287-314) — and the
//                anchors are copied into pinned staging blocks (one atomic cursor per block, no lock on the common path)
//   B  chain     ONE mm2b_chain_batch_ex call per staging block (all blocks concurrently): every read of the mini-batch is on
//                the GPU at once; 4-byte indices of the chained anchors come back
//   C  finish    kt_for over the reads: the second half of mm_map_frag — mm_gen_regs ... mm_set_mapq (map.c:318-376) — on the
//                chains gathered from the staged anchors, then worker_for's own epilogue (map.c:454-466)
//
// How it is built: this file IS the translation unit of the reference's map.c — it includes map.c unchanged, from where it
// lies (-I /root/reference; nothing is copied into this repository), with the one kt_for call of worker_pipeline renamed to the
// function at the end of this file.  Every static helper of map.c (collect_minimizers, collect_seed_hits, chain_post,
// align_regs ...) and step_t are therefore in scope, and no line of the reference changes.  A maintainer of the reference
// would make the same edit by hand: replace `kt_for(p->n_threads, worker_for, in, n_frag)` at map.c:561 by
// `mm2b_map_frags(p->n_threads, worker_for, in, n_frag)` and append this file's code to map.c.
//
// With the seeding front end (include/mm2seed_b200.h) phase A shrinks to copying the read sequences into
// one staging buffer and phase B becomes ONE mm2b_map_batch call: sketch, seed hits, sort and chaining all run on the GPU, and
// phase C receives, per read, exactly what the first half of mm_map_frag would have left it: the chains (u[], b[]), rep_len and
// mini_pos.  That path is taken when the configuration is one the device reproduces (mm2b_map_supported: one segment per read,
// no HPC, odd k, no SDUST, none of the seed-skipping flags) and the chaining arguments do not depend on the read; otherwise the
// anchors are seeded on the host as described above.  MM2B_FRONT=0 switches it off.
//
// The kalloc discipline of the reference is kept: nothing allocated from a thread's arena (mm_tbuf_t::km) outlives the phase
// that allocated it, so the leak check of map.c:386 holds at the end of BOTH halves (it is restated in each).
// What is not batched falls back to the reference's own per-read worker (and from there to the per-read mm_chain_dp drop-in):
// independent-segment mode, the debugging dumps of map.c:298-303/350-354, and the rare second chaining pass of the short-read
// preset (map.c:318-340), which phase C runs synchronously for the reads that need it.
#define kt_for mm2b_map_frags                 /* map.c:561 */
#include "map.c"
#undef kt_for

#include <algorithm>
#include <atomic>
#include <mutex>
#include <thread>
#include <vector>
#include <time.h>
#define MM2B_HOST_DECLARES_MM_CHAIN_DP          /* mmpriv.h:65 already declares it (with mm128_t) */
#include "mm2seed_b200.h"

extern "C" int mm2b_index_flatten_mt(const mm_idx_t *mi, mm2b_index_desc_t *out, int n_threads);      // host/idx_flatten.cpp (index.c + the bucket walk)
extern "C" void mm2b_index_flat_free(mm2b_index_desc_t *d);

extern "C" void kt_for(int n_threads, void (*func)(void*, long, int), void *data, long n);      // kthread.c:54

namespace {

constexpr int64_t BLOCK_ANCHORS = 4 << 20;          // 64 MB of pinned anchors per staging block

struct Block {                                      // anchors of many reads, back to back, plus room for their results
	mm2b_params_t par;
	int64_t cap = 0;
	mm2b_anchor_t *a = nullptr;                     // pinned
	uint64_t *u = nullptr;
	int32_t *bi = nullptr;
	std::atomic<int64_t> cursor{0};
	std::atomic<bool> closed{false};
	// phase B
	std::vector<int64_t> off, u_off, b_off;
	std::vector<int32_t> n_u, n_v, status;
	std::vector<long> frag;                         // fragment index of every read in the block, in staging order
	void reserve(int64_t n)
	{
		if (n <= cap) return;
		mm2b_host_free(a), mm2b_host_free(u), mm2b_host_free(bi);
		cap = n;
		a = (mm2b_anchor_t*)mm2b_host_alloc((size_t)cap * 16), u = (uint64_t*)mm2b_host_alloc((size_t)cap * 8), bi = (int32_t*)mm2b_host_alloc((size_t)cap * 4);
		if (!a || !u || !bi) { fprintf(stderr, "[mm2b] pinned staging for a mini-batch: %s\n", mm2b_last_error()); exit(1); }
	}
};

struct Frag {                                       // what the first half of mm_map_frag leaves for the second
	int block = -1;                                 // -1: nothing was seeded (empty / over-long query: map.c:284-285)
	int64_t pos = 0, n_a = 0;
	int32_t slot = 0;                               // index of the read inside its block (phase B)
	int rep_len = 0, n_mini_pos = 0, qlen_sum = 0, gap_ref = 0, gap_qry = 0;
	uint32_t hash = 0;
	uint64_t *mini_pos = nullptr;                   // malloc'd copy (the arena copy dies with phase A)
	// seeding front end: the read's row in the mm2b_map_batch result (-1: not on that path)
	int64_t front = -1;
};

constexpr int MAX_BLOCKS = 4096;

struct Stage {
	// seeding front end: the device copy of the index this process is mapping against, the sequence staging buffer, the result
	const mm_idx_t *front_mi = nullptr;
	const void *front_mi_S = nullptr;
	int front_mi_part = -1;
	mm2b_index_t *front_idx = nullptr;
	char *seq = nullptr;
	int64_t seq_cap = 0;
	std::vector<int64_t> seq_off;
	mm2b_map_result_t *front_res = nullptr;
	std::mutex mu;
	Block *blocks[MAX_BLOCKS];                      // blocks of the current mini-batch: appended under `mu`, read without it
	std::atomic<int> n_blocks{0};
	std::vector<Block*> pool;                       // spare blocks kept across mini-batches (their pinned memory is reused)
	std::vector<Frag> frags;
	step_t *step = nullptr;
};

bool same_par(const mm2b_params_t &a, const mm2b_params_t &b) { return memcmp(&a, &b, sizeof(a)) == 0; }

// room for n anchors in a block with these chaining arguments
mm2b_anchor_t *stage_alloc(Stage &st, const mm2b_params_t &par, int64_t n, int &block, int64_t &pos)
{
	static thread_local int last = -1;              // the block this thread used last (per mini-batch: checked against the list)
	for (;;) {
		Block *b = nullptr;
		int bi = last;
		if (bi >= 0 && bi < st.n_blocks.load(std::memory_order_acquire) && !st.blocks[bi]->closed.load() && same_par(st.blocks[bi]->par, par)) b = st.blocks[bi];
		if (!b) {
			std::lock_guard<std::mutex> lk(st.mu);
			const int nb = st.n_blocks.load();
			for (bi = nb - 1; bi >= 0; --bi)
				if (!st.blocks[bi]->closed.load() && same_par(st.blocks[bi]->par, par)) { b = st.blocks[bi]; break; }
			if (!b) {
				if (nb == MAX_BLOCKS) { fprintf(stderr, "[mm2b] more than %d staging blocks in one mini-batch\n", MAX_BLOCKS); exit(1); }
				if (!st.pool.empty()) b = st.pool.back(), st.pool.pop_back();
				else b = new Block();
				b->reserve(std::max<int64_t>(BLOCK_ANCHORS, n));
				b->par = par, b->cursor.store(0), b->closed.store(false);
				st.blocks[nb] = b;
				st.n_blocks.store(nb + 1, std::memory_order_release);
				bi = nb;
			}
		}
		const int64_t p = b->cursor.fetch_add(n);
		if (p + n <= b->cap) { last = block = bi, pos = p; return b->a + p; }
		b->closed.store(true);                      // full: later reads open a new block (the cursor stays past the end)
		last = -1;
	}
}

void leak_check(mm_tbuf_t *b, const char *qname, int qlen_sum)      // map.c:380-391
{
	km_stat_t kmst;
	if (!b->km) return;
	km_stat(b->km, &kmst);
	if (mm_dbg_flag & MM_DBG_PRINT_QNAME)
		fprintf(stderr, "QM\t%s\t%d\tcap=%ld,nCore=%ld,largest=%ld\n", qname, qlen_sum, kmst.capacity, kmst.n_cores, kmst.largest);
	assert(kmst.n_blocks == kmst.n_cores);          // otherwise, there is a memory leak
	if (kmst.largest > 1U << 28) {
		km_destroy(b->km);
		b->km = km_init();
	}
}

// worker_for's prologue (map.c:430-441): segment lengths and sequences of fragment i, mates flipped as the pairing orientation asks
int frag_segments(step_t *s, long i, int *qlens, const char **qseqs, bool flip)
{
	const int off = s->seg_off[i], pe_ori = s->p->opt->pe_ori, n = s->n_seg[i];
	assert(n <= MM_MAX_SEG);
	for (int j = 0; j < n; ++j) {
		if (flip && n == 2 && ((j == 0 && (pe_ori >> 1 & 1)) || (j == 1 && (pe_ori & 1))))
			mm_revcomp_bseq(&s->seq[off + j]);
		qlens[j] = s->seq[off + j].l_seq;
		qseqs[j] = s->seq[off + j].seq;
	}
	return n;
}

void chain_gaps(const mm_mapopt_t *opt, int qlen_sum, int &gap_ref, int &gap_qry)      // map.c:305-314
{
	if (opt->flag & MM_F_SR) gap_qry = qlen_sum > opt->max_gap ? qlen_sum : opt->max_gap;
	else gap_qry = opt->max_gap;
	if (opt->max_gap_ref > 0) gap_ref = opt->max_gap_ref;
	else if (opt->max_frag_len > 0) {
		gap_ref = opt->max_frag_len - qlen_sum;
		if (gap_ref < opt->max_gap) gap_ref = opt->max_gap;
	} else gap_ref = opt->max_gap;
}

mm2b_params_t chain_params(const mm_mapopt_t *opt, int gap_ref, int gap_qry, int n_segs)    // the arguments of the call at map.c:316
{
	mm2b_params_t p;
	p.max_dist_x = gap_ref, p.max_dist_y = gap_qry, p.bw = opt->bw, p.max_skip = opt->max_chain_skip, p.max_iter = opt->max_chain_iter;
	p.min_cnt = opt->min_cnt, p.min_sc = opt->min_chain_score, p.is_cdna = !!(opt->flag & MM_F_SPLICE), p.n_segs = n_segs, p.gap_scale = opt->chain_gap_scale;
	return p;
}

// ---- phase A: map.c:278-314, then stage the anchors --------------------------------------------------------------------
void seed_worker(void *data, long i, int tid)
{
	Stage &st = *(Stage*)data;
	step_t *s = st.step;
	const mm_mapopt_t *opt = s->p->opt;
	const mm_idx_t *mi = s->p->mi;
	mm_tbuf_t *b = s->buf[tid];
	Frag &f = st.frags[i];
	int qlens[MM_MAX_SEG];
	const char *qseqs[MM_MAX_SEG];
	const int off = s->seg_off[i], n_segs = frag_segments(s, i, qlens, qseqs, true);
	const char *qname = s->seq[off].name;
	int qlen_sum = 0;
	for (int j = 0; j < n_segs; ++j) qlen_sum += qlens[j], s->n_reg[off + j] = 0, s->reg[off + j] = 0;
	f.qlen_sum = qlen_sum;
	if (qlen_sum == 0 || n_segs <= 0 || n_segs > MM_MAX_SEG) return;
	if (opt->max_qlen > 0 && qlen_sum > opt->max_qlen) return;

	f.hash = qname ? __ac_X31_hash_string(qname) : 0;
	f.hash ^= __ac_Wang_hash(qlen_sum) + __ac_Wang_hash(opt->seed);
	f.hash = __ac_Wang_hash(f.hash);

	mm128_v mv = {0, 0, 0};
	int64_t n_a = 0;
	uint64_t *mini_pos = 0;
	collect_minimizers(b->km, opt, mi, n_segs, qlens, qseqs, &mv);
	mm128_t *a = (opt->flag & MM_F_HEAP_SORT)
	           ? collect_seed_hits_heap(b->km, opt, opt->mid_occ, mi, qname, &mv, qlen_sum, &n_a, &f.rep_len, &f.n_mini_pos, &mini_pos)
	           : collect_seed_hits(b->km, opt, opt->mid_occ, mi, qname, &mv, qlen_sum, &n_a, &f.rep_len, &f.n_mini_pos, &mini_pos);
	chain_gaps(opt, qlen_sum, f.gap_ref, f.gap_qry);
	f.n_a = n_a;
	mm2b_anchor_t *dst = stage_alloc(st, chain_params(opt, f.gap_ref, f.gap_qry, n_segs), n_a, f.block, f.pos);
	if (n_a > 0) memcpy(dst, a, (size_t)n_a * sizeof(mm128_t));
	if (f.n_mini_pos > 0) {
		f.mini_pos = (uint64_t*)malloc((size_t)f.n_mini_pos * 8);
		memcpy(f.mini_pos, mini_pos, (size_t)f.n_mini_pos * 8);
	}
	kfree(b->km, mv.a);
	kfree(b->km, a);
	kfree(b->km, mini_pos);
	leak_check(b, qname, qlen_sum);
}

// ---- phase C: map.c:318-392 on the chains of fragment i, then worker_for's epilogue (map.c:442-466) ---------------------
void finish_worker(void *data, long i, int tid)
{
	Stage &st = *(Stage*)data;
	step_t *s = st.step;
	const mm_mapopt_t *opt = s->p->opt;
	const mm_idx_t *mi = s->p->mi;
	mm_tbuf_t *b = s->buf[tid];
	Frag &f = st.frags[i];
	int qlens[MM_MAX_SEG];
	const char *qseqs[MM_MAX_SEG];
	const int off = s->seg_off[i], n_segs = frag_segments(s, i, qlens, qseqs, false), pe_ori = opt->pe_ori;
	const char *qname = s->seq[off].name;
	const int is_sr = !!(opt->flag & MM_F_SR), is_splice = !!(opt->flag & MM_F_SPLICE), qlen_sum = f.qlen_sum;
	if (f.block >= 0 || f.front >= 0) {
		int n_regs0 = 0, rep_len = f.rep_len, n_mini_pos = f.n_mini_pos, j;
		uint64_t *u = 0, *mini_pos = f.mini_pos;
		mm128_t *a = 0;
		bool mini_pos_in_km = false;
		if (f.front >= 0) {
			// seeding front end: the chains, rep_len and mini_pos of this read came back from the GPU (mm2b_map_batch)
			const mm2b_map_result_t *res = st.front_res;
			const int64_t r = f.front;
			const int sg = res->seg[r];
			rep_len = res->rep_len[r], n_mini_pos = res->n_mini_pos[r];
			if (n_mini_pos > 0) {                                             // map.c:117: q_span << 32 | position; q_span == k on this path
				const uint32_t *mp = res->seg_mini_pos[sg] + res->mp_off[r];
				const uint64_t span = (uint64_t)mi->k << 32;
				mini_pos = (uint64_t*)malloc((size_t)n_mini_pos * 8);
				for (int32_t k = 0; k < n_mini_pos; ++k) mini_pos[k] = span | mp[k];
			}
			if (res->status[r] == MM2B_READ_OK) {
				const int32_t nu = res->n_u[r], nv = res->n_v[r];
				u = (uint64_t*)kmalloc(b->km, (size_t)(nu > 0 ? nu : 1) * 8);
				a = (mm128_t*)kmalloc(b->km, (size_t)nv * sizeof(mm128_t));
				if (nu > 0) memcpy(u, res->seg_u[sg] + res->u_off[r], (size_t)nu * 8);
				if (nv > 0) memcpy(a, res->seg_b[sg] + res->b_off[r], (size_t)nv * sizeof(mm128_t));
				n_regs0 = nu;
			}
		} else {
		// what mm_chain_dp would have returned (chain.c:396-422): u[] copied, b[] gathered from the staged anchors by index
		Block &blk = *st.blocks[f.block];
		if (blk.status[f.slot] == MM2B_READ_OK) {
			const int32_t nu = blk.n_u[f.slot], nv = blk.n_v[f.slot];
			const mm128_t *src = (const mm128_t*)blk.a + f.pos;
			const int32_t *ix = blk.bi + blk.b_off[f.slot];
			u = (uint64_t*)kmalloc(b->km, (size_t)(nu > 0 ? nu : 1) * 8);
			a = (mm128_t*)kmalloc(b->km, (size_t)nv * sizeof(mm128_t));
			if (nu > 0) memcpy(u, blk.u + blk.u_off[f.slot], (size_t)nu * 8);
			for (int32_t k = 0; k < nv; ++k) a[k] = src[ix[k]];
			n_regs0 = nu;
		}
		}
		if (opt->max_occ > opt->mid_occ && rep_len > 0) {                 // map.c:318-340, verbatim in effect: rare, done per read
			int rechain = 0;
			if (n_regs0 > 0) {
				int n_chained_segs = 1, max = 0, max_i = -1, max_off = -1, o = 0;
				for (j = 0; j < n_regs0; ++j) {
					if (max < (int)(u[j] >> 32)) max = u[j] >> 32, max_i = j, max_off = o;
					o += (uint32_t)u[j];
				}
				for (j = 1; j < (int32_t)u[max_i]; ++j)
					if ((a[max_off + j].y & MM_SEED_SEG_MASK) != (a[max_off + j - 1].y & MM_SEED_SEG_MASK)) ++n_chained_segs;
				if (n_chained_segs < n_segs) rechain = 1;
			} else rechain = 1;
			if (rechain) {
				mm128_v mv = {0, 0, 0};
				int64_t n_a;
				kfree(b->km, a);
				kfree(b->km, u);
				free(mini_pos), mini_pos = 0, f.mini_pos = 0;
				collect_minimizers(b->km, opt, mi, n_segs, qlens, qseqs, &mv);      // (the sketch of phase A died with its arena)
				if (opt->flag & MM_F_HEAP_SORT) a = collect_seed_hits_heap(b->km, opt, opt->max_occ, mi, qname, &mv, qlen_sum, &n_a, &rep_len, &n_mini_pos, &mini_pos);
				else a = collect_seed_hits(b->km, opt, opt->max_occ, mi, qname, &mv, qlen_sum, &n_a, &rep_len, &n_mini_pos, &mini_pos);
				mini_pos_in_km = true;
				kfree(b->km, mv.a);
				a = mm_chain_dp(f.gap_ref, f.gap_qry, opt->bw, opt->max_chain_skip, opt->max_chain_iter, opt->min_cnt, opt->min_chain_score, opt->chain_gap_scale,
				                is_splice, n_segs, n_a, a, &n_regs0, &u, b->km, tid);
			}
		}
		b->frag_gap = f.gap_ref;
		b->rep_len = rep_len;

		mm_reg1_t *regs0 = mm_gen_regs(b->km, f.hash, qlen_sum, n_regs0, u, a);
		if (mi->n_alt) {
			mm_mark_alt(mi, n_regs0, regs0);
			mm_hit_sort(b->km, &n_regs0, regs0, opt->alt_drop);
		}
		chain_post(opt, f.gap_ref, mi, b->km, qlen_sum, n_segs, qlens, &n_regs0, regs0, a);
		if (!is_sr) mm_est_err(mi, qlen_sum, n_regs0, regs0, a, n_mini_pos, mini_pos);

		int *n_regs = &s->n_reg[off];
		mm_reg1_t **regs = &s->reg[off];
		if (n_segs == 1) {                              // uni-segment
			regs0 = align_regs(opt, mi, b->km, qlens[0], qseqs[0], &n_regs0, regs0, a);
			mm_set_mapq(b->km, n_regs0, regs0, opt->min_chain_score, opt->a, rep_len, is_sr);
			n_regs[0] = n_regs0, regs[0] = regs0;
		} else {                                        // multi-segment
			mm_seg_t *seg = mm_seg_gen(b->km, f.hash, n_segs, qlens, n_regs0, regs0, n_regs, regs, a);
			free(regs0);
			for (j = 0; j < n_segs; ++j) {
				mm_set_parent(b->km, opt->mask_level, opt->mask_len, n_regs[j], regs[j], opt->a * 2 + opt->b, opt->flag & MM_F_HARD_MLEVEL, opt->alt_drop);
				regs[j] = align_regs(opt, mi, b->km, qlens[j], qseqs[j], &n_regs[j], regs[j], seg[j].a);
				mm_set_mapq(b->km, n_regs[j], regs[j], opt->min_chain_score, opt->a, rep_len, is_sr);
			}
			mm_seg_free(b->km, n_segs, seg);
			if (n_segs == 2 && opt->pe_ori >= 0 && (opt->flag & MM_F_CIGAR))
				mm_pair(b->km, f.gap_ref, opt->pe_bonus, opt->a * 2 + opt->b, opt->a, qlens, n_regs, regs);
		}
		kfree(b->km, a);
		kfree(b->km, u);
		if (mini_pos_in_km) kfree(b->km, mini_pos);
		else free(mini_pos);
		f.mini_pos = 0;
		leak_check(b, qname, qlen_sum);
	}
	for (int j = 0; j < n_segs; ++j) s->rep_len[off + j] = b->rep_len, s->frag_gap[off + j] = b->frag_gap;      // map.c:450-453
	for (int j = 0; j < n_segs; ++j)                  // flip the query strand and coordinate back to the original read strand (map.c:454-466)
		if (n_segs == 2 && ((j == 0 && (pe_ori >> 1 & 1)) || (j == 1 && (pe_ori & 1)))) {
			mm_revcomp_bseq(&s->seq[off + j]);
			for (int k = 0; k < s->n_reg[off + j]; ++k) {
				mm_reg1_t *r = &s->reg[off + j][k];
				const int t = r->qs;
				r->qs = qlens[j] - r->qe;
				r->qe = qlens[j] - t;
				r->rev = !r->rev;
			}
		}
}

// ---- seeding front end: phase A is a copy of the sequences, phase B one mm2b_map_batch call ------------------------------------
void stage_seq_worker(void *data, long i, int tid)
{
	(void)tid;
	Stage &st = *(Stage*)data;
	step_t *s = st.step;
	const mm_mapopt_t *opt = s->p->opt;
	Frag &f = st.frags[i];
	const int off = s->seg_off[i];
	const mm_bseq1_t *q = &s->seq[off];
	s->n_reg[off] = 0, s->reg[off] = 0;
	f.qlen_sum = q->l_seq;
	if (st.seq_off[i + 1] == st.seq_off[i]) return;                       // empty or over-long query (map.c:284-285): nothing to map
	f.front = i;
	f.hash = q->name ? __ac_X31_hash_string(q->name) : 0;                 // map.c:287-289
	f.hash ^= __ac_Wang_hash(f.qlen_sum) + __ac_Wang_hash(opt->seed);
	f.hash = __ac_Wang_hash(f.hash);
	chain_gaps(opt, f.qlen_sum, f.gap_ref, f.gap_qry);
	memcpy(st.seq + st.seq_off[i], q->seq, (size_t)q->l_seq);
}

double wall_s()
{
	struct timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	return t.tv_sec + 1e-9 * t.tv_nsec;
}

// the device copy of the index being mapped against (one per process at a time: the CLI maps against one index part at a time)
mm2b_index_t *front_index(Stage &st, const mm_idx_t *mi, int n_threads)
{
	if (st.front_idx && st.front_mi == mi && st.front_mi_S == (const void*)mi->S && st.front_mi_part == mi->index) return st.front_idx;
	mm2b_index_destroy(st.front_idx), st.front_idx = nullptr;
	mm2b_index_desc_t d;
	const double t0 = wall_s();
	if (mm2b_index_flatten_mt(mi, &d, n_threads) != 0) return nullptr;
	const double t1 = wall_s();
	st.front_idx = mm2b_index_create(&d);
	if (getenv("MM2B_TRACE") && atoi(getenv("MM2B_TRACE")) > 0)
		fprintf(stderr, "[mm2b trace] front end: index flattened in %.3f s (%lld minimizers, %lld positions), on the device in %.3f s\n", t1 - t0, (long long)d.n_keys, (long long)d.n_pos, wall_s() - t1);
	mm2b_index_flat_free(&d);
	if (!st.front_idx) { fprintf(stderr, "[mm2b] fatal: index upload: %s\n", mm2b_last_error()); exit(EXIT_FAILURE); }
	st.front_mi = mi, st.front_mi_S = (const void*)mi->S, st.front_mi_part = mi->index;
	return st.front_idx;
}

bool front_applies(const step_t *s, long n)
{
	static const bool off = getenv("MM2B_FRONT") && atoi(getenv("MM2B_FRONT")) == 0;
	const mm_mapopt_t *opt = s->p->opt;
	const mm_idx_t *mi = s->p->mi;
	if (off || !mm2b_map_supported(mi->k, mi->w, mi->flag & MM_I_HPC, 1, opt->flag, opt->sdust_thres)) return false;
	if ((opt->flag & MM_F_SR) || (opt->max_gap_ref <= 0 && opt->max_frag_len > 0)) return false;      // chaining gaps depend on the read (map.c:305-314)
	for (long i = 0; i < n; ++i) if (s->n_seg[i] != 1) return false;
	return true;
}

bool map_frags_front(Stage &st, int n_threads, long n)
{
	step_t *s = st.step;
	const mm_mapopt_t *opt = s->p->opt;
	static const bool trace = getenv("MM2B_TRACE") && atoi(getenv("MM2B_TRACE")) > 0;
	double t0 = wall_s(), t1;
	auto lap = [&](const char *what) {
		if (!trace) return;
		t1 = wall_s();
		fprintf(stderr, "[mm2b trace] front end: %-22s %.3f s\n", what, t1 - t0);
		t0 = t1;
	};
	mm2b_index_t *idx = front_index(st, s->p->mi, n_threads);
	lap("index on the device");
	if (!idx) return false;                                                 // (an index the flat format cannot hold: seed on the host)
	st.seq_off.assign((size_t)n + 1, 0);
	for (long i = 0; i < n; ++i) {
		const int len = s->seq[s->seg_off[i]].l_seq;
		const bool skip = len <= 0 || (opt->max_qlen > 0 && len > opt->max_qlen);
		st.seq_off[(size_t)i + 1] = st.seq_off[(size_t)i] + (skip ? 0 : len);
	}
	if (st.seq_off[(size_t)n] > st.seq_cap) {
		mm2b_host_free(st.seq);
		st.seq_cap = st.seq_off[(size_t)n] + st.seq_off[(size_t)n] / 8 + 4096;
		st.seq = (char*)mm2b_host_alloc((size_t)st.seq_cap);
		if (!st.seq) { fprintf(stderr, "[mm2b] pinned staging for the sequences: %s\n", mm2b_last_error()); exit(1); }
	}
	lap("staging buffer");
	kt_for(n_threads, stage_seq_worker, &st, n);                             // A
	lap("A: sequences staged");
	int gap_ref = 0, gap_qry = 0;
	chain_gaps(opt, 0, gap_ref, gap_qry);
	const mm2b_params_t par = chain_params(opt, gap_ref, gap_qry, 1);
	const mm2b_seed_params_t sp = {opt->mid_occ, 0};
	if (mm2b_map_batch(idx, &sp, &par, n, st.seq_off.data(), st.seq, &st.front_res) != MM2B_OK) {      // B
		fprintf(stderr, "[mm2b] fatal: mapping a mini-batch: %s\n", mm2b_last_error());                  // same behaviour as checkError (chain_hardware.cpp:208)
		exit(EXIT_FAILURE);
	}
	lap("B: mm2b_map_batch");
	if (trace) {
		const mm2b_map_result_t *r = st.front_res;
		fprintf(stderr, "[mm2b trace] front end: %ld reads, %lld bases -> %lld minimizers, %lld anchors (%lld reads with equal keys), %lld chains; sketch %.2f seed %.2f sort %.2f chain %.2f ms, %d segments\n",
		        n, (long long)st.seq_off[(size_t)n], (long long)r->tot_mini, (long long)r->tot_anchors, (long long)r->n_tie_reads, (long long)r->tot_chains,
		        r->sketch_ms, r->seed_ms, r->sort_ms, r->chain_ms, r->n_segs);
	}
	kt_for(n_threads, finish_worker, &st, n);                                // C
	lap("C: reads finished");
	mm2b_map_result_release(st.front_res), st.front_res = nullptr;
	return true;
}

// ---- phase B: every staging block in one batch call, all blocks at once -------------------------------------------------------
void chain_block(Stage &st, int bi)
{
	Block &b = *st.blocks[bi];
	const int64_t n = (int64_t)b.frag.size();
	b.n_u.assign((size_t)n, 0), b.n_v.assign((size_t)n, 0), b.status.assign((size_t)n, 0);
	b.u_off.assign((size_t)n + 1, 0), b.b_off.assign((size_t)n + 1, 0);
	if (mm2b_chain_batch_ex(&b.par, n, b.off.data(), b.a, b.n_u.data(), b.n_v.data(), b.status.data(), b.u_off.data(), b.b_off.data(),
	                        b.u, b.cap, nullptr, b.bi, b.cap, 0, nullptr) != MM2B_OK) {
		fprintf(stderr, "[mm2b] fatal: chaining a mini-batch: %s\n", mm2b_last_error());      // same behaviour as checkError (chain_hardware.cpp:208)
		exit(EXIT_FAILURE);
	}
}

Stage &the_stage()
{
	static Stage *st = new Stage();             // (kept for the life of the process: its pinned blocks are reused by every mini-batch)
	return *st;
}
std::mutex g_stage_mu;                          // step 1 of the pipeline runs one mini-batch at a time (kt_pipeline, kthread.c:128-136); kept explicit

}  // namespace

// The call worker_pipeline makes for step 1 (map.c:561), phase-split.  `func` is the reference's own worker_for: used as it is
// for the modes that are not batched.
extern "C" void mm2b_map_frags(int n_threads, void (*func)(void*, long, int), void *data, long n)
{
	step_t *s = (step_t*)data;
	const mm_mapopt_t *opt = s->p->opt;
	static const bool off = getenv("MM2B_PHASE_SPLIT") && atoi(getenv("MM2B_PHASE_SPLIT")) == 0;
	if (off || n <= 0 || (opt->flag & MM_F_INDEPEND_SEG) || (mm_dbg_flag & (MM_DBG_PRINT_SEED | MM_DBG_PRINT_QNAME))) {
		kt_for(n_threads, func, data, n);
		return;
	}
	std::lock_guard<std::mutex> guard(g_stage_mu);
	Stage &st = the_stage();
	st.step = s;
	st.frags.assign((size_t)n, Frag());
	if (front_applies(s, n) && map_frags_front(st, n_threads, n)) return;
	kt_for(n_threads, seed_worker, &st, n);                                  // A
	const int nb = st.n_blocks.load();
	for (int k = 0; k < nb; ++k) st.blocks[k]->off.clear(), st.blocks[k]->frag.clear();
	{	// staging order of every block -> CSR offsets (reads sit back to back: the cursor only moves forward)
		std::vector<std::vector<long>> by_block((size_t)nb);
		for (long i = 0; i < n; ++i) if (st.frags[i].block >= 0) by_block[(size_t)st.frags[i].block].push_back(i);
		for (int k = 0; k < nb; ++k) {
			std::vector<long> &v = by_block[k];
			std::sort(v.begin(), v.end(), [&](long x, long y) {
				const Frag &fx = st.frags[x], &fy = st.frags[y];
				return fx.pos != fy.pos ? fx.pos < fy.pos : fx.n_a < fy.n_a;      // (reads without anchors share a position with their successor: empty first)
			});
			Block &b = *st.blocks[k];
			int64_t end = 0;
			for (size_t r = 0; r < v.size(); ++r) {
				Frag &f = st.frags[v[r]];
				f.slot = (int32_t)r;
				b.off.push_back(f.pos);
				end = f.pos + f.n_a;
			}
			b.off.push_back(end);
			b.frag.swap(v);
		}
	}
	{	// B
		std::vector<std::thread> th;
		for (int k = 1; k < nb; ++k) th.emplace_back(chain_block, std::ref(st), k);
		if (nb > 0) chain_block(st, 0);
		for (auto &t : th) t.join();
	}
	kt_for(n_threads, finish_worker, &st, n);                                // C
	{
		std::lock_guard<std::mutex> lk(st.mu);
		for (int k = 0; k < nb; ++k) st.pool.push_back(st.blocks[k]);
		st.n_blocks.store(0);
	}
}
