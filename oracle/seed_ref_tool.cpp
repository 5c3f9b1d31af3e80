// TEST INFRASTRUCTURE — the reference's own seeding, observed.  Built in place from /root/reference by oracle/Makefile into
// oracle/_ref/mm2-seed-ref (never shipped in the product, never on the product path).
//
//   mm2-seed-ref <preset> <ref.fa | ref.mmi> <reads.fa> <seeds.bin> [index.bin] [mid_occ]
//
// For every read of reads.fa it runs the reference's collect_minimizers (map.c:64-78 -> mm_sketch, sketch.c:77) and
// collect_seed_hits (map.c:215-247) — both static in map.c, so this file includes map.c from where it lies — and records
// qlen, the minimizers, the sorted anchors, rep_len and mini_pos.  index.bin, if asked for, is the flat index written by the
// product's own mm2b_index_flatten (host/idx_flatten.cpp), so GPU tests can load an index on a box without /root/reference.
#include "map.c"
#define MM2B_HOST_DECLARES_MM_CHAIN_DP          /* mmpriv.h:65 already declares it (with mm128_t) */
#include "mm2seed_b200.h"
#include <stdio.h>

extern "C" int mm2b_index_flatten(const mm_idx_t *mi, mm2b_index_desc_t *out);

static void put(FILE *fp, const void *p, size_t n) { if (n && fwrite(p, 1, n, fp) != n) { perror("write"); exit(1); } }

int main(int argc, char **argv)
{
	if (argc < 5) { fprintf(stderr, "usage: mm2-seed-ref <preset> <ref> <reads.fa> <seeds.bin> [index.bin] [mid_occ]\n"); return 2; }
	mm_idxopt_t io;
	mm_mapopt_t mo;
	mm_verbose = 1;
	mm_set_opt(0, &io, &mo);
	if (mm_set_opt(argv[1], &io, &mo) < 0) { fprintf(stderr, "unknown preset %s\n", argv[1]); return 2; }
	if (argc > 6) mo.mid_occ = atoi(argv[6]);
	mm_idx_reader_t *rd = mm_idx_reader_open(argv[2], &io, 0);
	if (!rd) { fprintf(stderr, "cannot open %s\n", argv[2]); return 1; }
	mm_idx_t *mi = mm_idx_reader_read(rd, 8);
	if (!mi) { fprintf(stderr, "cannot read the index\n"); return 1; }
	mm_mapopt_update(&mo, mi);
	if (argc > 5 && argv[5][0] && strcmp(argv[5], "-") != 0) {
		mm2b_index_desc_t d;
		if (mm2b_index_flatten(mi, &d) != 0) { fprintf(stderr, "flatten failed\n"); return 1; }
		FILE *fi = fopen(argv[5], "wb");
		const int32_t hdr[6] = {0x5849324d /* "M2IX" */, d.k, d.w, d.is_hpc, d.n_seq, mo.mid_occ};
		put(fi, hdr, sizeof(hdr)), put(fi, &d.n_keys, 8), put(fi, &d.n_pos, 8);
		put(fi, d.keys, (size_t)d.n_keys * 8), put(fi, d.vals, (size_t)d.n_keys * 8), put(fi, d.pos, (size_t)d.n_pos * 8);
		fclose(fi);
	}
	FILE *fo = fopen(argv[4], "wb");
	const int32_t hdr[4] = {0x5332324d /* "M22S" */, mi->k, mi->w, mo.mid_occ};
	put(fo, hdr, sizeof(hdr));
	mm_bseq_file_t *fp = mm_bseq_open(argv[3]);
	if (!fp) { fprintf(stderr, "cannot open %s\n", argv[3]); return 1; }
	void *km = km_init();
	int n = 0;
	mm_bseq1_t *seqs;
	long n_reads = 0;
	while ((seqs = mm_bseq_read(fp, 100000000, 0, &n)) != 0) {
		for (int i = 0; i < n; ++i) {
			mm128_v mv = {0, 0, 0};
			int qlen = seqs[i].l_seq, rep_len = 0, n_mini_pos = 0;
			const char *seq = seqs[i].seq;
			int64_t n_a = 0;
			uint64_t *mini_pos = 0;
			collect_minimizers(km, &mo, mi, 1, &qlen, &seq, &mv);
			mm128_t *a = collect_seed_hits(km, &mo, mo.mid_occ, mi, seqs[i].name, &mv, qlen, &n_a, &rep_len, &n_mini_pos, &mini_pos);
			const int32_t rec[4] = {qlen, (int32_t)mv.n, rep_len, n_mini_pos};
			put(fo, rec, sizeof(rec)), put(fo, &n_a, 8);
			put(fo, mv.a, mv.n * 16), put(fo, a, (size_t)n_a * 16), put(fo, mini_pos, (size_t)n_mini_pos * 8);
			kfree(km, mv.a), kfree(km, a), kfree(km, mini_pos);
			free(seqs[i].seq), free(seqs[i].name);
			if (seqs[i].qual) free(seqs[i].qual);
			if (seqs[i].comment) free(seqs[i].comment);
			++n_reads;
		}
		free(seqs);
	}
	mm_bseq_close(fp);
	fclose(fo);
	km_destroy(km);
	mm_idx_destroy(mi);
	mm_idx_reader_close(rd);
	fprintf(stderr, "[mm2-seed-ref] %ld reads, k=%d w=%d mid_occ=%d\n", n_reads, hdr[1], hdr[2], hdr[3]);
	return 0;
}
