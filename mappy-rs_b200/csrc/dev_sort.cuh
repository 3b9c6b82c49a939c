/* dev_sort.cuh -- exact device replay of minimap2's radix_sort_128x.
 *
 * ksort.h KRADIX_SORT_INIT (v2.26) is an in-place MSD "American flag" sort with
 * insertion sort below 65 elements.  It is not stable, and the order in which
 * it leaves equal keys decides chaining tie-breaks (chain ends with equal f in
 * mg_chain_backtrack, equal anchor positions, equal region keys), so wherever
 * equal keys can occur the device must reproduce that permutation, not just a
 * sorted order.  The permutation step is a pointer chase with no parallel form;
 * it is replayed by ONE lane per read while the rest of the chunk's reads run
 * on other warps.  Passes over key bytes that are identical in every element
 * move nothing and are skipped (the result is identical by construction).
 * Records are (x, y) pairs in separate arrays; the key is x.
 */
#ifndef MMG_DEV_SORT_CUH
#define MMG_DEV_SORT_CUH
#include "dev_common.cuh"

template<typename YT>
__device__ __forceinline__ void dev_insertsort_t(uint64_t *x, YT *y, int beg, int end)
{
	for (int i = beg + 1; i < end; ++i)
		if (x[i] < x[i - 1]) {
			uint64_t tx = x[i];
			YT ty = y[i];
			int j;
			for (j = i; j > beg && tx < x[j - 1]; --j) x[j] = x[j - 1], y[j] = y[j - 1];
			x[j] = tx, y[j] = ty;
		}
}

/* bkt: 512 ints of scratch (bucket begin/end); stk: 3 ints per pending range, at least 3*(n/65+1) ints.
 * LEAF_SORT = false stops before the insertion sorts of the leaf ranges (<= 64 elements): those are stable,
 * so the caller may finish with ANY stable sort of the whole array by key (the leaves are disjoint, contiguous
 * and already in key order relative to each other) - sort.cu does that with the whole CTA.
 * A pass in which every element of the range has the same key byte moves nothing; it is detected after the
 * counting loop and skipped.  Bucket loops only span the byte values that occur. */
template<typename YT, bool LEAF_SORT>
static __device__ void dev_radix_sort_t(uint64_t *x, YT *y, int n, int *bkt, int *stk)
{
	if (n <= 64) { if (LEAF_SORT) dev_insertsort_t(x, y, 0, n); return; }
	int *bb = bkt, *be = bkt + 256;
	uint64_t diff = 0;
	for (int i = 1; i < n; ++i) diff |= x[i] ^ x[0];
	if (diff == 0) return;
	int s0 = ((63 - __clzll((long long)diff)) >> 3) << 3; /* highest key byte that differs */
	int sp = 0;
	for (int k = 0; k < 256; ++k) be[k] = 0;
	stk[0] = 0, stk[1] = n, stk[2] = s0, sp = 1;
	while (sp > 0) {
		--sp;
		const int beg = stk[3 * sp], end = stk[3 * sp + 1];
		int s = stk[3 * sp + 2];
		int kmin, kmax;
		for (;;) { /* be[] is all zero here */
			kmin = 255, kmax = 0;
			for (int i = beg; i < end; ++i) {
				int b = (int)((x[i] >> s) & 255);
				++be[b];
				kmin = b < kmin ? b : kmin, kmax = b > kmax ? b : kmax;
			}
			if (kmin != kmax || s == 0) break;
			be[kmin] = 0, s -= 8;           /* one bucket holds the whole range: nothing moves, next byte */
		}
		{
			int pos = beg;
			for (int k = kmin; k <= kmax; ++k) { int cnt = be[k]; bb[k] = pos; pos += cnt; be[k] = pos; }
		}
		for (int k = kmin; k <= kmax;) {
			if (bb[k] != be[k]) {
				int l = (int)((x[bb[k]] >> s) & 255);
				if (l != k) {
					uint64_t tx = x[bb[k]];
					YT ty = y[bb[k]];
					do {
						uint64_t sx = tx;
						YT sy = ty;
						int q = bb[l]++;
						tx = x[q], ty = y[q];
						x[q] = sx, y[q] = sy;
						l = (int)((tx >> s) & 255);
					} while (l != k);
					x[bb[k]] = tx, y[bb[k]] = ty;
					++bb[k];
				} else ++bb[k];
			} else ++k;
		}
		{
			const int s2 = s > 8 ? s - 8 : 0;
			int start = beg;
			for (int k = kmin; k <= kmax; ++k) {
				int e = be[k], sz = e - start;
				be[k] = 0;
				if (s) {
					if (sz > 64) stk[3 * sp] = start, stk[3 * sp + 1] = e, stk[3 * sp + 2] = s2, ++sp;
					else if (LEAF_SORT && sz > 1) dev_insertsort_t(x, y, start, e);
				}
				start = e;
			}
		}
	}
}

/* The same sort called by a WHOLE WARP (all 32 lanes, converged): the counting loops, the bucket prefix sums and
 * the insertion sorts of the leaf ranges (one leaf per lane) run in parallel; only the cycle-leader permutation of
 * a pass, which has no parallel form, stays on lane 0.  Ranges are disjoint, so the order in which they are
 * processed does not change the result.  bkt must be shared memory (atomics), 512 ints. */
template<typename YT, bool LEAF_SORT = true>
static __device__ void dev_radix_sort_warp(uint64_t *x, YT *y, int n, int *bkt, int *stk)
{
	const int lane = mmg_lane();
	const uint32_t lt = mmg_lanemask_lt();
	if (n <= 64) {
		if (LEAF_SORT && lane == 0) dev_insertsort_t(x, y, 0, n);
		__syncwarp();
		return;
	}
	int *bb = bkt, *be = bkt + 256;
	uint64_t diff = 0;
	{
		const uint64_t x0 = x[0];
		for (int i = 1 + lane; i < n; i += 32) diff |= x[i] ^ x0;
		const uint32_t lo = __reduce_or_sync(MMG_FULL, (uint32_t)diff), hi = __reduce_or_sync(MMG_FULL, (uint32_t)(diff >> 32));
		diff = (uint64_t)hi << 32 | lo;
	}
	if (diff == 0) return;
	const int s0 = ((63 - __clzll((long long)diff)) >> 3) << 3;
	for (int k = lane; k < 256; k += 32) be[k] = 0;
	if (lane == 0) stk[0] = 0, stk[1] = n, stk[2] = s0;
	int sp = 1;
	__syncwarp();
	while (sp > 0) {
		--sp;
		const int beg = stk[3 * sp], end = stk[3 * sp + 1];
		int s = stk[3 * sp + 2];
		int kmin, kmax;
		for (;;) { /* be[] is all zero here */
			int mn = 255, mx = 0;
			for (int i = beg + lane; i < end; i += 32) {
				int b = (int)((x[i] >> s) & 255);
				atomicAdd(&be[b], 1);
				mn = b < mn ? b : mn, mx = b > mx ? b : mx;
			}
			kmin = __reduce_min_sync(MMG_FULL, mn), kmax = __reduce_max_sync(MMG_FULL, mx);
			__syncwarp();
			if (kmin != kmax || s == 0) break;
			if (lane == 0) be[kmin] = 0;    /* one bucket holds the whole range: nothing moves, next byte */
			__syncwarp();
			s -= 8;
		}
		{
			int carry = beg;
			for (int k0 = kmin; k0 <= kmax; k0 += 32) {
				const int k = k0 + lane, cnt = k <= kmax ? be[k] : 0;
				int tot, ex = mmg_warp_excl_scan(cnt, &tot);
				if (k <= kmax) bb[k] = carry + ex, be[k] = carry + ex + cnt;
				carry += tot;
			}
		}
		__syncwarp();
		if (lane == 0) {
			for (int k = kmin; k <= kmax;) {
				if (bb[k] != be[k]) {
					int l = (int)((x[bb[k]] >> s) & 255);
					if (l != k) {
						uint64_t tx = x[bb[k]];
						YT ty = y[bb[k]];
						do {
							uint64_t sx = tx;
							YT sy = ty;
							int q = bb[l]++;
							tx = x[q], ty = y[q];
							x[q] = sx, y[q] = sy;
							l = (int)((tx >> s) & 255);
						} while (l != k);
						x[bb[k]] = tx, y[bb[k]] = ty;
						++bb[k];
					} else ++bb[k];
				} else ++k;
			}
		}
		__syncwarp();
		if (s) {
			const int s2 = s > 8 ? s - 8 : 0;
			for (int k0 = kmin; k0 <= kmax; k0 += 32) {
				const int k = k0 + lane;
				int start = 0, e = 0, sz = 0;
				if (k <= kmax) e = be[k], start = k == kmin ? beg : be[k - 1], sz = e - start;
				const bool push = sz > 64;
				const uint32_t pm = __ballot_sync(MMG_FULL, push);
				if (push) { const int q = sp + __popc(pm & lt); stk[3 * q] = start, stk[3 * q + 1] = e, stk[3 * q + 2] = s2; }
				sp += __popc(pm);
				if (LEAF_SORT && sz > 1 && sz <= 64) dev_insertsort_t(x, y, start, e);
			}
		}
		__syncwarp();
		for (int k = kmin + lane; k <= kmax; k += 32) be[k] = 0;
		__syncwarp();
	}
}

/* The radix passes (no leaf sorts) called by a WHOLE CTA: the pending ranges sit on a shared stack and every warp takes
 * one at a time, so that after the first pass (one range, one warp) the independent sub-ranges are permuted by
 * different warps at the same time.  bkt_all: 512 ints of shared memory per warp; ctl: 3 ints {lock, stack size,
 * ranges not finished}; stack: 3 ints per range, at least n / 65 + 1 ranges.
 * (The SIMT emulator of the CPU test-suite runs its fibers cooperatively and cannot spin on a lock: there warp 0
 * does all ranges through dev_radix_sort_warp.) */
template<typename YT>
static __device__ void dev_radix_passes_cta(uint64_t *x, YT *y, int n, int *bkt_all, int *ctl, int *stack, int *stk_global)
{
	const int lane = mmg_lane(), wib = threadIdx.x >> 5;
	const uint32_t lt = mmg_lanemask_lt();
	if (n <= 64) return;
#ifdef MMG_EMU
	if (wib == 0) dev_radix_sort_warp<YT, false>(x, y, n, bkt_all, stk_global);
	__syncthreads();
	return;
#else
	(void)stk_global;
	int *bb = bkt_all + wib * 512, *be = bb + 256;
	for (int k = lane; k < 256; k += 32) be[k] = 0;
	if (threadIdx.x == 0) ctl[0] = 0, ctl[1] = 1, ctl[2] = 1, stack[0] = 0, stack[1] = n, stack[2] = 56;
	__syncthreads();
	for (;;) {
		int beg = 0, end = 0, s = 0, got = 0;
		if (lane == 0) {
			for (;;) {
				if (*(volatile int*)&ctl[2] == 0) { got = -1; break; }
				if (*(volatile int*)&ctl[1] > 0 && atomicCAS(&ctl[0], 0, 1) == 0) {
					int sp = *(volatile int*)&ctl[1];
					if (sp > 0) {
						--sp;
						beg = *(volatile int*)&stack[3 * sp], end = *(volatile int*)&stack[3 * sp + 1], s = *(volatile int*)&stack[3 * sp + 2];
						*(volatile int*)&ctl[1] = sp;
						got = 1;
					}
					__threadfence_block();
					atomicExch(&ctl[0], 0);
					if (got) break;
				}
				__nanosleep(128); /* nothing to take yet: leave the issue slots to the warps that are permuting */
			}
		}
		got = __shfl_sync(MMG_FULL, got, 0);
		if (got < 0) break;
		beg = __shfl_sync(MMG_FULL, beg, 0), end = __shfl_sync(MMG_FULL, end, 0), s = __shfl_sync(MMG_FULL, s, 0);
		int kmin, kmax;
		for (;;) { /* be[] is all zero here */
			int mn = 255, mx = 0;
			for (int i = beg + lane; i < end; i += 32) {
				int b = (int)((x[i] >> s) & 255);
				atomicAdd(&be[b], 1);
				mn = b < mn ? b : mn, mx = b > mx ? b : mx;
			}
			kmin = __reduce_min_sync(MMG_FULL, mn), kmax = __reduce_max_sync(MMG_FULL, mx);
			__syncwarp();
			if (kmin != kmax || s == 0) break;
			if (lane == 0) be[kmin] = 0;
			__syncwarp();
			s -= 8;
		}
		{
			int carry = beg;
			for (int k0 = kmin; k0 <= kmax; k0 += 32) {
				const int k = k0 + lane, cnt = k <= kmax ? be[k] : 0;
				int tot, ex = mmg_warp_excl_scan(cnt, &tot);
				if (k <= kmax) bb[k] = carry + ex, be[k] = carry + ex + cnt;
				carry += tot;
			}
		}
		__syncwarp();
		if (lane == 0) {
			for (int k = kmin; k <= kmax;) {
				if (bb[k] != be[k]) {
					int l = (int)((x[bb[k]] >> s) & 255);
					if (l != k) {
						uint64_t tx = x[bb[k]];
						YT ty = y[bb[k]];
						do {
							uint64_t sx = tx;
							YT sy = ty;
							int q = bb[l]++;
							tx = x[q], ty = y[q];
							x[q] = sx, y[q] = sy;
							l = (int)((tx >> s) & 255);
						} while (l != k);
						x[bb[k]] = tx, y[bb[k]] = ty;
						++bb[k];
					} else ++bb[k];
				} else ++k;
			}
		}
		__syncwarp();
		if (s) {
			const int s2 = s > 8 ? s - 8 : 0;
			for (int k0 = kmin; k0 <= kmax; k0 += 32) {
				const int k = k0 + lane;
				int start = 0, e = 0, sz = 0;
				if (k <= kmax) e = be[k], start = k == kmin ? beg : be[k - 1], sz = e - start;
				const bool push = sz > 64;
				const uint32_t pm = __ballot_sync(MMG_FULL, push);
				if (pm) {
					int base = 0;
					if (lane == 0) {
						while (atomicCAS(&ctl[0], 0, 1) != 0) {}
						base = *(volatile int*)&ctl[1];
					}
					base = __shfl_sync(MMG_FULL, base, 0);
					if (push) { const int q = base + __popc(pm & lt); stack[3 * q] = start, stack[3 * q + 1] = e, stack[3 * q + 2] = s2; }
					__syncwarp();
					if (lane == 0) {
						__threadfence_block();
						*(volatile int*)&ctl[1] = base + __popc(pm);
						atomicAdd(&ctl[2], __popc(pm));
						__threadfence_block();
						atomicExch(&ctl[0], 0);
					}
				}
			}
		}
		__syncwarp();
		for (int k = kmin + lane; k <= kmax; k += 32) be[k] = 0;
		__syncwarp();
		if (lane == 0) { __threadfence_block(); atomicSub(&ctl[2], 1); }
	}
	__syncthreads();
#endif
}

static __device__ void dev_radix_sort_128x(uint64_t *x, uint64_t *y, int n, int *bkt, int *stk)
{
	dev_radix_sort_t<uint64_t, true>(x, y, n, bkt, stk);
}

#endif
