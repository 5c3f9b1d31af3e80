// The reference's radix_sort_128x (ksort.h:116-151, misc.c:155-156) replayed exactly, equal keys included.  Shared by the chaining
// kernels (final order of a read's chains, chain.c:411) and the seeding kernels (anchors of a read, map.c:245).
#pragma once
#include <stdint.h>
#ifndef MM2B_SORT_CHK
#define MM2B_SORT_CHK(cond, code) do { } while (0)
#endif

namespace mm2b {
namespace {

struct W16 { uint64_t x, y; };          // 16-byte record sorted by .x; 8-byte aligned on purpose

// The reference orders chains with radix_sort_128x keyed on .x only (chain.c:411, ksort.h:116-151): insertion sort up to
// 64 elements (stable), otherwise an in-place MSD byte radix sort whose permutation of equal keys is algorithm-specific.
// mm_join_long depends on the resulting adjacency (hit.c:335), so for n > 64 the exact permutation scheme is replayed
// here by one lane (bucket cursors in shared memory, pending sub-ranges in a small worklist kept in `work`).
__device__ void insertion_by_x(W16 *w, int n)
{
	for (int i = 1; i < n; ++i) {
		const W16 key = w[i];
		int j = i;
		for (; j > 0 && key.x < w[j - 1].x; --j) w[j] = w[j - 1];
		w[j] = key;
	}
}

__device__ void flag_sort_by_x_lane0(W16 *w, int n, int *sm /* >= 768 ints */, int3 *work, int work_cap)
{
	int *head = sm, *tail = sm + 256, *cnt = sm + 512;
	int n_work = 0;
	MM2B_SORT_CHK(work_cap >= 1, 0x10);
	work[n_work++] = make_int3(0, n, 56);
	while (n_work > 0) {
		const int3 job = work[--n_work];
		W16 *a = w + job.x;
		const int m = job.y, shift = job.z;
		for (int k = 0; k < 256; ++k) cnt[k] = 0;
		for (int i = 0; i < m; ++i) ++cnt[(int)(a[i].x >> shift & 0xff)];
		bool one_bucket = false;                        // every key shares this byte: the permutation below would move nothing
		for (int k = 0, acc = 0; k < 256; ++k) head[k] = acc, acc += cnt[k], tail[k] = acc, one_bucket |= cnt[k] == m;
		for (int k = one_bucket ? 256 : 0; k < 256;) {
			if (head[k] == tail[k]) { ++k; continue; }
			int l = (int)(a[head[k]].x >> shift & 0xff);
			if (l == k) { ++head[k]; continue; }
			W16 carry = a[head[k]];
			do {
				const W16 out = a[head[l]];
				a[head[l]++] = carry;
				carry = out;
				l = (int)(carry.x >> shift & 0xff);
			} while (l != k);
			a[head[k]++] = carry;
		}
		if (shift) {
			const int next = shift > 8 ? shift - 8 : 0;
			for (int k = 0; k < 256; ++k) {
				const int beg = tail[k] - cnt[k];
				if (cnt[k] > 64) {
					MM2B_SORT_CHK(n_work < work_cap, 0x10);
					if (n_work < work_cap) work[n_work++] = make_int3(job.x + beg, cnt[k], next);
					else insertion_by_x(a + beg, cnt[k]);   // unreachable: work_cap >= n/65 + 1 pending ranges always fit
				} else if (cnt[k] > 1) insertion_by_x(a + beg, cnt[k]);
			}
		}
	}
}

}  // namespace
}  // namespace mm2b
