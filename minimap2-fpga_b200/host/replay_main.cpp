// mm2b-replay — batched caller for anchor dumps (include/mm2chain_dump.h).
//
// The reference chains one read per mm_chain_dp call from n_threads workers (map.c:316, kthread.c:65).  This caller is the
// other shape: it reads the chaining inputs of many reads, recorded once, packs every group of reads that shares its chaining
// arguments into one CSR batch in pinned host memory and hands it to mm2b_chain_batch — the way a phase-split mm_map_frag
// would feed a whole -K mini-batch (§8d config 5).  With --check the results are compared, bit for bit,
// with what the reference returned when the dump was recorded.
//
//   mm2b-replay [-g 0,1,..] [-r repeats] [-B reads_per_call] [--check] [--quiet] dump[.gz] ...
//
// Exit status: 0 ok, 1 results differ from the recorded ones, 2 usage / unreadable dump, 3 backend error (e.g. no GPU: there is
// no CPU fallback).
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mm2chain_b200.h"
#include "mm2chain_dump.h"

namespace {

struct Record {                       // one recorded mm_chain_dp call
	mm2b_dump_hdr_t h;
	std::vector<mm2b_anchor_t> a, b;
	std::vector<uint64_t> u;
};

struct Group {                        // reads that share their chaining arguments, in dump order
	mm2b_params_t par;
	std::vector<const Record*> reads;
};

bool read_exact(gzFile f, void *dst, size_t bytes)
{
	char *p = (char*)dst;
	while (bytes) {
		const unsigned chunk = (unsigned)std::min<size_t>(bytes, 1u << 30);
		const int got = gzread(f, p, chunk);
		if (got <= 0) return false;
		p += got, bytes -= (size_t)got;
	}
	return true;
}

// gzopen reads plain files as well; returns false on a malformed dump
bool load_dump(const char *path, std::vector<Record> &out)
{
	gzFile f = gzopen(path, "rb");
	if (!f) { fprintf(stderr, "mm2b-replay: cannot open %s\n", path); return false; }
	gzbuffer(f, 1 << 20);
	bool ok = true;
	for (;;) {
		Record r;
		const int got = gzread(f, &r.h, sizeof r.h);
		if (got == 0) break;                                  // clean end of file
		if (got != (int)sizeof r.h || r.h.magic != MM2B_DUMP_MAGIC || r.h.n < 0 || r.h.n_u < 0 || r.h.n_v < 0 || r.h.n_v > r.h.n) {
			fprintf(stderr, "mm2b-replay: %s: bad record header after %zu records\n", path, out.size());
			ok = false;
			break;
		}
		r.a.resize((size_t)r.h.n), r.u.resize((size_t)r.h.n_u), r.b.resize((size_t)r.h.n_v);
		ok = read_exact(f, r.a.data(), r.a.size() * sizeof(mm2b_anchor_t));
		if (ok && (r.h.flags & MM2B_DUMP_HAS_FPV) && r.h.n > 0) ok = gzseek(f, (z_off_t)(12 * r.h.n), SEEK_CUR) >= 0;
		if (ok && !r.u.empty()) ok = read_exact(f, r.u.data(), r.u.size() * 8);
		if (ok && !r.b.empty()) ok = read_exact(f, r.b.data(), r.b.size() * sizeof(mm2b_anchor_t));
		if (!ok) { fprintf(stderr, "mm2b-replay: %s: truncated record %zu\n", path, out.size()); break; }
		out.push_back(std::move(r));
	}
	gzclose(f);
	return ok;
}

bool same_params(const mm2b_params_t &p, const mm2b_dump_hdr_t &h)
{
	return p.max_dist_x == h.max_dist_x && p.max_dist_y == h.max_dist_y && p.bw == h.bw && p.max_skip == h.max_skip && p.max_iter == h.max_iter
	    && p.min_cnt == h.min_cnt && p.min_sc == h.min_sc && p.is_cdna == h.is_cdna && p.n_segs == h.n_segs && p.gap_scale == h.gap_scale;
}

template <class T> struct Pinned {    // mm2b_host_alloc'ed array (pinned: the copies run at PCIe speed)
	T *p = nullptr;
	size_t n = 0;
	bool reserve(size_t want)
	{
		if (want <= n) return true;
		mm2b_host_free(p);
		p = (T*)mm2b_host_alloc(std::max<size_t>(want, 1) * sizeof(T));
		n = p ? want : 0;
		return p != nullptr;
	}
	~Pinned() { mm2b_host_free(p); }
};

struct Totals { long long reads = 0, anchors = 0, chains = 0, chained = 0, calls = 0, bad_reads = 0, heavy = 0; double seconds = 0; };

int usage()
{
	fprintf(stderr, "usage: mm2b-replay [-g dev,dev,..] [-r repeats] [-B reads_per_call] [--check] [--quiet] dump[.gz] ...\n");
	return 2;
}

}  // namespace

int main(int argc, char **argv)
{
	std::vector<int> devs;
	std::vector<const char*> paths;
	int repeats = 1;
	long long reads_per_call = 0;
	bool check = false, quiet = false;
	for (int i = 1; i < argc; ++i) {
		const std::string s = argv[i];
		if (s == "--check") check = true;
		else if (s == "--quiet") quiet = true;
		else if (s == "-r" && i + 1 < argc) repeats = std::max(1, atoi(argv[++i]));
		else if (s == "-B" && i + 1 < argc) reads_per_call = std::max(0LL, atoll(argv[++i]));
		else if (s == "-g" && i + 1 < argc) {
			for (char *tok = strtok(argv[++i], ","); tok; tok = strtok(nullptr, ",")) devs.push_back(atoi(tok));
		} else if (!s.empty() && s[0] == '-') return usage();
		else paths.push_back(argv[i]);
	}
	if (paths.empty()) return usage();

	std::vector<Record> recs;
	for (const char *p : paths) if (!load_dump(p, recs)) return 2;
	std::vector<Group> groups;
	for (const Record &r : recs) {
		Group *g = nullptr;
		for (Group &c : groups) if (same_params(c.par, r.h)) { g = &c; break; }
		if (!g) {
			groups.emplace_back();
			g = &groups.back();
			g->par = mm2b_params_t{r.h.max_dist_x, r.h.max_dist_y, r.h.bw, r.h.max_skip, r.h.max_iter, r.h.min_cnt, r.h.min_sc, r.h.is_cdna, r.h.n_segs, r.h.gap_scale};
		}
		g->reads.push_back(&r);
	}

	if (mm2b_init((int)devs.size(), devs.empty() ? nullptr : devs.data()) != MM2B_OK) {
		fprintf(stderr, "mm2b-replay: %s\n", mm2b_last_error());
		return 3;
	}
	Totals tot;
	int rc = 0;
	{   // (scope: the pinned arrays go back before the backend shuts down)
	Pinned<mm2b_anchor_t> a, b;
	Pinned<uint64_t> u;
	Pinned<int64_t> off, u_off, b_off;
	Pinned<int32_t> n_u, n_v, status;
	for (size_t gi = 0; gi < groups.size() && rc == 0; ++gi) {
		const Group &g = groups[gi];
		const size_t per_call = reads_per_call > 0 ? (size_t)reads_per_call : g.reads.size();
		Totals gt;
		for (size_t first = 0; first < g.reads.size() && rc == 0; first += per_call) {
			const size_t nr = std::min(per_call, g.reads.size() - first);
			size_t na = 0;
			for (size_t r = 0; r < nr; ++r) na += g.reads[first + r]->a.size();
			if (!a.reserve(na) || !b.reserve(na) || !u.reserve(na) || !off.reserve(nr + 1) || !u_off.reserve(nr + 1) || !b_off.reserve(nr + 1)
			    || !n_u.reserve(nr) || !n_v.reserve(nr) || !status.reserve(nr)) {
				fprintf(stderr, "mm2b-replay: pinned allocation failed: %s\n", mm2b_last_error());
				rc = 3;
				break;
			}
			off.p[0] = 0;
			for (size_t r = 0; r < nr; ++r) {
				const Record &rec = *g.reads[first + r];
				if (!rec.a.empty()) memcpy(a.p + off.p[r], rec.a.data(), rec.a.size() * sizeof(mm2b_anchor_t));
				off.p[r + 1] = off.p[r] + (int64_t)rec.a.size();
			}
			mm2b_stats_t st;
			double best = 1e30;
			for (int rep = 0; rep < repeats + (repeats > 1); ++rep) {          // with -r > 1 the first pass is an untimed warm-up
				const auto t0 = std::chrono::steady_clock::now();
				const int err = mm2b_chain_batch(&g.par, (int64_t)nr, off.p, a.p, n_u.p, n_v.p, status.p, u_off.p, b_off.p, u.p, (int64_t)na, b.p, (int64_t)na, &st);
				const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
				if (err != MM2B_OK) { fprintf(stderr, "mm2b-replay: mm2b_chain_batch: %s\n", mm2b_last_error()); rc = 3; break; }
				if (repeats == 1 || rep > 0) best = std::min(best, dt);
			}
			if (rc) break;
			gt.reads += (long long)nr, gt.anchors += (long long)na, gt.chains += st.n_chains, gt.chained += st.n_chained, gt.heavy += st.n_heavy_reads, gt.seconds += best, ++gt.calls;
			if (check) {
				for (size_t r = 0; r < nr; ++r) {
					const Record &rec = *g.reads[first + r];
					const bool u_null = (rec.h.flags & MM2B_DUMP_U_NULL) != 0;
					bool ok = n_u.p[r] == rec.h.n_u && n_v.p[r] == rec.h.n_v && (status.p[r] != MM2B_READ_OK) == u_null;
					ok = ok && (rec.u.empty() || memcmp(u.p + u_off.p[r], rec.u.data(), rec.u.size() * 8) == 0);
					ok = ok && (rec.b.empty() || memcmp(b.p + b_off.p[r], rec.b.data(), rec.b.size() * sizeof(mm2b_anchor_t)) == 0);
					if (!ok) {
						if (gt.bad_reads < 5) fprintf(stderr, "mm2b-replay: group %zu read %zu (n=%lld): n_u %d vs %d, n_v %d vs %d, status %d\n", gi, first + r,
						                              (long long)rec.h.n, n_u.p[r], rec.h.n_u, n_v.p[r], rec.h.n_v, status.p[r]);
						++gt.bad_reads;
					}
				}
			}
		}
		if (!quiet)
			printf("group %zu  max_dist %d/%d bw %d skip %d iter %d min_cnt %d min_sc %d cdna %d segs %d gap_scale %g | reads %lld anchors %lld chains %lld chained %lld | "
			       "%lld call(s) %.3f ms  %.3g reads/s  %.3g anchors/s%s\n", gi, g.par.max_dist_x, g.par.max_dist_y, g.par.bw, g.par.max_skip, g.par.max_iter,
			       g.par.min_cnt, g.par.min_sc, g.par.is_cdna, g.par.n_segs, (double)g.par.gap_scale, gt.reads, gt.anchors, gt.chains, gt.chained, gt.calls,
			       gt.seconds * 1e3, gt.reads / std::max(gt.seconds, 1e-12), gt.anchors / std::max(gt.seconds, 1e-12),
			       check ? (gt.bad_reads ? "  MISMATCH" : "  identical to the recorded results") : "");
		tot.reads += gt.reads, tot.anchors += gt.anchors, tot.chains += gt.chains, tot.chained += gt.chained, tot.calls += gt.calls, tot.seconds += gt.seconds, tot.bad_reads += gt.bad_reads, tot.heavy += gt.heavy;
	}
	}
	if (rc == 0)
		printf("total  devices %d | reads %lld (%lld on the heavy-read kernel) anchors %lld chains %lld chained %lld | %lld call(s) %.3f ms  %.3g reads/s  %.3g anchors/s | mismatching reads %lld%s\n",
		       mm2b_num_devices(), tot.reads, tot.heavy, tot.anchors, tot.chains, tot.chained, tot.calls, tot.seconds * 1e3, tot.reads / std::max(tot.seconds, 1e-12),
		       tot.anchors / std::max(tot.seconds, 1e-12), tot.bad_reads, check ? "" : " (not checked)");
	mm2b_shutdown();
	if (rc == 0 && check && tot.bad_reads) rc = 1;
	return rc;
}
