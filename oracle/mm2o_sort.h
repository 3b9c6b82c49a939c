/* mm2o_sort.h -- ORACLE (test infrastructure only).
 *
 * Restates minimap2 v2.26 ksort.h `KRADIX_SORT_INIT` (radix_sort_128x,
 * radix_sort_64 instantiated in misc.c) and the heap/ksmall helpers used by
 * seed.c and index.c.  The radix sort is an in-place MSD "American flag" sort
 * with insertion sort below RS_MIN_SIZE elements; it is NOT stable and the
 * order it leaves equal keys in feeds the chaining DP,
 * so it is restated operation for operation.
 * Reached from /root/reference/src/lib.rs:482,587 via mm_map.
 */
#ifndef MM2O_SORT_H
#define MM2O_SORT_H

#include <stdint.h>
#include <stddef.h>
#include <algorithm>
#include "mm2o.h"

#define RS_MIN_SIZE 64
#define RS_MAX_BITS 8

template<typename T, typename KeyFn>
static inline void rs_insertsort(T *beg, T *end, KeyFn key)
{
	T *i;
	for (i = beg + 1; i < end; ++i)
		if (key(*i) < key(*(i - 1))) {
			T *j, tmp = *i;
			for (j = i; j > beg && key(tmp) < key(*(j - 1)); --j)
				*j = *(j - 1);
			*j = tmp;
		}
}

template<typename T, typename KeyFn>
static void rs_sort(T *beg, T *end, int n_bits, int s, KeyFn key)
{
	struct bucket_t { T *b, *e; };
	T *i;
	int size = 1 << n_bits, m = size - 1;
	bucket_t *k, b[1 << RS_MAX_BITS], *be = b + size;
	for (k = b; k != be; ++k) k->b = k->e = beg;
	for (i = beg; i != end; ++i) ++b[key(*i) >> s & m].e;
	for (k = b + 1; k != be; ++k)
		k->e += (k - 1)->e - beg, k->b = (k - 1)->e;
	for (k = b; k != be;) {
		if (k->b != k->e) {
			bucket_t *l;
			if ((l = b + (key(*k->b) >> s & m)) != k) {
				T tmp = *k->b, swap;
				do {
					swap = tmp; tmp = *l->b; *l->b++ = swap;
					l = b + (key(tmp) >> s & m);
				} while (l != k);
				*k->b++ = tmp;
			} else ++k->b;
		} else ++k;
	}
	for (b->b = beg, k = b + 1; k != be; ++k) k->b = (k - 1)->e;
	if (s) {
		s = s > n_bits ? s - n_bits : 0;
		for (k = b; k != be; ++k)
			if (k->e - k->b > RS_MIN_SIZE) rs_sort(k->b, k->e, n_bits, s, key);
			else if (k->e - k->b > 1) rs_insertsort(k->b, k->e, key);
	}
}

struct mm2o_key128 { uint64_t operator()(const mm128_t &a) const { return a.x; } };
struct mm2o_key64  { uint64_t operator()(const uint64_t &a) const { return a; } };

static inline void radix_sort_128x(mm128_t *beg, mm128_t *end)
{
	if (end - beg <= RS_MIN_SIZE) rs_insertsort(beg, end, mm2o_key128());
	else rs_sort(beg, end, RS_MAX_BITS, (8 - 1) * RS_MAX_BITS, mm2o_key128());
}

static inline void radix_sort_64(uint64_t *beg, uint64_t *end)
{
	if (end - beg <= RS_MIN_SIZE) rs_insertsort(beg, end, mm2o_key64());
	else rs_sort(beg, end, RS_MAX_BITS, (8 - 1) * RS_MAX_BITS, mm2o_key64());
}

/* ksort.h: ks_heapdown / ks_heapmake for uint64_t (max-heap on '<') */
static inline void ks_heapdown_uint64_t(size_t i, size_t n, uint64_t l[])
{
	size_t k = i;
	uint64_t tmp = l[i];
	while ((k = (k << 1) + 1) < n) {
		if (k != n - 1 && l[k] < l[k + 1]) ++k;
		if (l[k] < tmp) break;
		l[i] = l[k]; i = k;
	}
	l[i] = tmp;
}

static inline void ks_heapmake_uint64_t(size_t lsize, uint64_t l[])
{
	size_t i;
	for (i = (lsize >> 1) - 1; i != (size_t)(-1); --i)
		ks_heapdown_uint64_t(i, lsize, l);
}

/* ksort.h: ks_ksmall -- k-th smallest; the value is algorithm-independent */
static inline uint32_t ks_ksmall_uint32_t(size_t n, uint32_t arr[], size_t kk)
{
	std::nth_element(arr, arr + kk, arr + n);
	return arr[kk];
}

#endif
