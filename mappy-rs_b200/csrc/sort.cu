/* sort.cu -- per-read segmented sort of anchors by target position
 * (north-star (c), second half): one CTA per read, keys staged in shared memory.
 *
 * Replaces radix_sort_128x(a, a + n_a) at the end of map.c collect_seed_hits
 * (mm_map path, /root/reference/src/lib.rs:482,587; ksort.h, minimap2 v2.26).
 *
 * A read's anchors (typically 10^2..10^4 16-byte records) are sorted where they
 * fit: (key, source index) pairs go to shared memory, a bitonic network orders
 * them by (x, index), and the payload y is gathered once on the way out, so HBM
 * sees one read and one write of every record.  Reads whose anchors exceed the
 * shared-memory tile run the same network in a global scratch slice.
 *
 * Equal keys: upstream's sort is unstable above 64 elements, and the order it
 * leaves equal x in reaches the chaining DP.  Sorting by (x, index) equals
 * upstream's insertion sort for n <= 64; for larger n the kernel detects equal
 * neighbours (rare: the same minimizer twice in a read hitting one target
 * position) and only then replays upstream's permutation exactly on one thread
 * (dev_sort.cuh).
 * Bound: shared-memory bandwidth; HBM traffic 32 B/anchor - see DESIGN.md.
 */
#include "dev_common.cuh"
#include "dev_sort.cuh"
#include "stages.h"

__device__ __forceinline__ bool key_gt(uint64_t xa, uint32_t ia, uint64_t xb, uint32_t ib)
{
	return xa > xb || (xa == xb && ia > ib);
}

/* ELEMS = shared-memory tile (records).  The SMALL instantiation keeps many CTAs resident per SM
 * (the per-read chain of dependent global loads is latency, not bandwidth) and defers reads that
 * do not fit its tile to big_list; the BIG instantiation (LIST = true) takes its reads from there. */
template<int ELEMS, bool LIST>
__global__ void __launch_bounds__(SORT_THREADS)
sort_kernel(ChunkDev c, uint32_t r0, uint32_t r1, uint32_t *work, uint32_t *big_list, uint32_t *n_big, int small_max)
{
	MMG_DYN_SMEM(smem_raw);
	__shared__ uint32_t s_item;
	__shared__ int s_tie;
	__shared__ int s_bkt[512];
	uint64_t *sx = (uint64_t*)smem_raw;
	uint32_t *si = (uint32_t*)(sx + ELEMS);
	const int tid = threadIdx.x, nt = blockDim.x;
	const uint32_t n_items = LIST ? *n_big : r1 - r0;

	for (;;) {
		if (tid == 0) s_item = atomicAdd(work, 1u), s_tie = 0;
		__syncthreads();
		if (s_item >= n_items) break;
		const uint32_t r = LIST ? big_list[s_item] : mmg_read_of(c, r0 + s_item);
		const int n = (int)c.n_a[r];
		if (!LIST && n > small_max) { /* uniform over the CTA */
			if (tid == 0) big_list[atomicAdd(n_big, 1u)] = r;
			__syncthreads();
			continue;
		}
		const uint64_t ab = c.a_off[r] - c.a_off0;
		const uint64_t *ax = c.ax + ab, *ay = c.ay + ab;
		uint64_t *bx = c.bx + ab, *by = c.by + ab;
		if (n > 1) {
			int m = 1;
			while (m < n) m <<= 1;
			uint64_t *kx;
			uint32_t *ki;
			const bool in_smem = m <= ELEMS;
			if (in_smem) kx = sx, ki = si;
			else kx = c.zx + 2 * ab, ki = (uint32_t*)(c.zy + 2 * ab); /* global tile: m < 2n */
			const uint64_t *srcx = ax, *srcy = ay;
			for (int pass = 0; pass < 2; ++pass) {
				for (int i = tid; i < m; i += nt) kx[i] = i < n ? srcx[i] : MMG_INF64, ki[i] = (uint32_t)i;
				__syncthreads();
				for (int k = 2; k <= m; k <<= 1) {
					for (int j = k >> 1; j > 0; j >>= 1) {
						for (int t = tid; t < (m >> 1); t += nt) {
							int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)); /* index with bit j clear */
							int p = i | j;
							bool up = (i & k) == 0;
							uint64_t xa = kx[i], xb = kx[p];
							uint32_t ia = ki[i], ib = ki[p];
							if (key_gt(xa, ia, xb, ib) == up) kx[i] = xb, ki[i] = ib, kx[p] = xa, ki[p] = ia;
						}
						__syncthreads();
					}
				}
				int tie = 0;
				for (int i = tid; i < n; i += nt) {
					uint64_t x = kx[i];
					bx[i] = x, by[i] = srcy[ki[i]];
					if (i + 1 < n && kx[i + 1] == x) tie = 1;
				}
				if (pass == 0 && tie && n > 64) s_tie = 1;
				__syncthreads();
				if (pass == 1 || !s_tie) break;
				/* Equal keys in a read above upstream's insertion-sort size: the order upstream's unstable
				 * radix sort leaves them in is replayed.  One thread runs the radix passes over (key, source
				 * index) in the tile; the insertion sorts of the leaf ranges are stable, so they are replaced
				 * by a second run of the network on (key, position after the radix passes). */
				if (!in_smem) { /* the global tile is the scratch the second pass would need: serial replay */
					for (int i = tid; i < n; i += nt) bx[i] = ax[i], by[i] = ay[i];
					__syncthreads();
					if (tid == 0) {
						dev_radix_sort_128x(bx, by, n, s_bkt, c.f + ab);
						c.flags[r] |= 1u;
					}
					break;
				}
				for (int i = tid; i < n; i += nt) kx[i] = ax[i], ki[i] = (uint32_t)i;
				__syncthreads();
				if (tid == 0) {
					dev_radix_sort_t<uint32_t, false>(kx, ki, n, s_bkt, c.f + ab);
					c.flags[r] |= 1u;
				}
				__syncthreads();
				uint64_t *zx = c.zx + 2 * ab, *zy = c.zy + 2 * ab;
				for (int i = tid; i < n; i += nt) zx[i] = kx[i], zy[i] = ay[ki[i]];
				__syncthreads();
				srcx = zx, srcy = zy;
			}
		} else if (n == 1) {
			if (tid == 0) bx[0] = ax[0], by[0] = ay[0];
		}
		__syncthreads();
	}
}

/* ---- large reads: LSD radix sort of the whole read by one CTA ------------------------------------------
 * A record is key<<28 | source index, with key = strand:1 | linear target coordinate:35 (seq_off[rid] + pos), which
 * orders exactly like x = strand<<63 | rid<<32 | pos.  8-bit digits; a pass is a warp-level multi-split: warp w
 * owns a contiguous slice, __match_any_sync groups the lanes of a step by digit, the group leader bumps the
 * warp's private counter, and a lane's rank is its position inside the group - stable, no atomics.  Passes whose
 * digit is the same in every record are skipped.  The records ping-pong in the read's 2n-word global scratch
 * (L1/L2 resident), so any read length works and shared memory only holds the 8 x 256 counters.
 * O(n) work per pass instead of the O(n log^2 n) of the bitonic network the small reads use. */
#define RSORT_WARPS 8
#define RSORT_IDX_BITS 28
#define RSORT_IDX_MASK 0x0fffffffULL

__device__ __forceinline__ uint64_t rsort_record(uint64_t x, uint32_t i, const uint64_t *seq_off)
{
	const uint64_t lin = seq_off[(uint32_t)(x >> 32) & 0x7fffffffu] + (uint32_t)x;
	return ((x >> 63) << 35 | lin) << RSORT_IDX_BITS | i;
}

/* sorts the n keys srcx[] (stable); afterwards A[0..n) holds the records in order.  Returns the buffer holding them. */
/* lanes of the warp that hold the same 8-bit digit (and the same validity): one ballot per varying bit, cheaper than
 * __match_any_sync, which the compiler expands into a loop */
__device__ __forceinline__ uint32_t rsort_peers(uint32_t d, bool valid, uint32_t vary)
{
	uint32_t peers = __ballot_sync(MMG_FULL, valid);
	if (!valid) peers = ~peers;
	while (vary) { /* only the bits of the digit that differ somewhere in the read (uniform over the CTA) */
		const int b = __ffs((int)vary) - 1;
		vary &= vary - 1;
		const bool bit = (d >> b) & 1u;
		const uint32_t bal = __ballot_sync(MMG_FULL, bit);
		peers &= bit ? bal : ~bal;
	}
	return peers;
}

static __device__ uint64_t *rsort_read(const uint64_t *srcx, int n, const uint64_t *seq_off, uint64_t *A, uint64_t *B, uint32_t *cnt, uint32_t *s_red)
{
	const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
	const uint32_t lt = mmg_lanemask_lt();
	uint64_t diff = 0;
	{
		const uint64_t r0 = rsort_record(srcx[0], 0, seq_off);
		for (int i = tid; i < n; i += RSORT_WARPS * 32) {
			const uint64_t rec = rsort_record(srcx[i], (uint32_t)i, seq_off);
			A[i] = rec;
			diff |= rec ^ r0;
		}
		uint32_t lo = __reduce_or_sync(MMG_FULL, (uint32_t)diff), hi = __reduce_or_sync(MMG_FULL, (uint32_t)(diff >> 32));
		if (lane == 0) s_red[w] = lo, s_red[RSORT_WARPS + w] = hi;
		__syncthreads();
		lo = 0, hi = 0;
		for (int q = 0; q < RSORT_WARPS; ++q) lo |= s_red[q], hi |= s_red[RSORT_WARPS + q];
		diff = (uint64_t)hi << 32 | lo;
		__syncthreads();
	}
	const int chunk = ((n + RSORT_WARPS - 1) / RSORT_WARPS + 31) & ~31;
	const int beg = w * chunk < n ? w * chunk : n, end = beg + chunk < n ? beg + chunk : n;
	for (int shift = RSORT_IDX_BITS; shift < 64; shift += 8) {
		const uint32_t vary = (uint32_t)(diff >> shift) & 0xffu;
		if (vary == 0) continue;
		for (int j = tid; j < RSORT_WARPS * 256; j += RSORT_WARPS * 32) cnt[j] = 0;
		__syncthreads();
		for (int b0 = beg; b0 < end; b0 += 32) {
			const int i = b0 + lane;
			const bool valid = i < end;
			const uint32_t d = valid ? (uint32_t)(A[i] >> shift) & 0xffu : 0u;
			const uint32_t peers = rsort_peers(d, valid, vary);
			if (valid && (peers & lt) == 0) cnt[w * 256 + d] += (uint32_t)__popc(peers);
			__syncwarp();
		}
		__syncthreads();
		{ /* exclusive scan in (digit, warp) order: thread t owns digit t */
			uint32_t loc[RSORT_WARPS], tot = 0;
			for (int q = 0; q < RSORT_WARPS; ++q) loc[q] = tot, tot += cnt[q * 256 + tid];
			uint32_t x = tot;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				uint32_t y = __shfl_up_sync(MMG_FULL, x, o);
				if (lane >= o) x += y;
			}
			if (lane == 31) s_red[w] = x;
			__syncthreads();
			uint32_t pre = 0;
			for (int q = 0; q < w; ++q) pre += s_red[q];
			const uint32_t base = pre + x - tot;
			for (int q = 0; q < RSORT_WARPS; ++q) cnt[q * 256 + tid] = base + loc[q];
		}
		__syncthreads();
		for (int b0 = beg; b0 < end; b0 += 32) {
			const int i = b0 + lane;
			const bool valid = i < end;
			const uint64_t rec = valid ? A[i] : 0;
			const uint32_t d = valid ? (uint32_t)(rec >> shift) & 0xffu : 0u;
			const uint32_t peers = rsort_peers(d, valid, vary);
			const uint32_t base = valid ? cnt[w * 256 + d] : 0;
			__syncwarp();
			if (valid && (peers & lt) == 0) cnt[w * 256 + d] = base + (uint32_t)__popc(peers);
			if (valid) B[base + (uint32_t)__popc(peers & lt)] = rec;
			__syncwarp();
		}
		__syncthreads();
		uint64_t *tmp = A; A = B, B = tmp;
	}
	return A;
}

/* TIE = false: reads of big_list; a read with equal keys (above upstream's insertion-sort size) is appended to
 * tie_list.  TIE = true: reads of tie_list: upstream's unstable radix passes are replayed by one warp over
 * (key, source index) staged in shared memory (the cycle-leader permutation is a chain of dependent accesses: the
 * latency of shared memory instead of L2), the stable leaf insertion sorts are replaced by a second radix sort on
 * (key, position after those passes) - see the tie path of sort_kernel. */
#define RSORT_TIE_TILE 12288      /* records staged in shared memory by the equal-key replay (144 KB + 27 KB of per-warp buckets: one CTA per SM; these reads are rare) */
template<bool TIE>
__global__ void __launch_bounds__(RSORT_WARPS * 32)
radix_sort_kernel(ChunkDev c, const uint64_t *seq_off, uint32_t *work, const uint32_t *list, const uint32_t *n_list, uint32_t *tie_list, uint32_t *n_tie)
{
	MMG_DYN_SMEM(smem_raw);
	__shared__ uint32_t s_item;
	__shared__ int s_tie;
	__shared__ int s_bkt[TIE ? RSORT_WARPS * 512 : 512];
	__shared__ int s_ctl[4], s_stack[TIE ? 3 * (RSORT_TIE_TILE / 65 + 2) : 3];
	__shared__ uint32_t s_cnt[RSORT_WARPS * 256];
	__shared__ uint32_t s_red[2 * RSORT_WARPS];
	const int tid = threadIdx.x, nt = blockDim.x;
	const uint32_t n_items = *n_list;
	for (;;) {
		if (tid == 0) s_item = atomicAdd(work, 1u), s_tie = 0;
		__syncthreads();
		if (s_item >= n_items) break;
		const uint32_t r = list[s_item];
		const int n = (int)c.n_a[r];
		const uint64_t ab = c.a_off[r] - c.a_off0;
		const uint64_t *ax = c.ax + ab, *ay = c.ay + ab;
		uint64_t *bx = c.bx + ab, *by = c.by + ab;
		uint64_t *zx = c.zx + 2 * ab, *zy = c.zy + 2 * ab;
		const uint64_t *srcx = ax, *srcy = ay;
		if (TIE) {
			uint64_t *kx = n <= RSORT_TIE_TILE ? (uint64_t*)smem_raw : zx;
			uint32_t *ki = n <= RSORT_TIE_TILE ? (uint32_t*)((uint64_t*)smem_raw + RSORT_TIE_TILE) : (uint32_t*)(zx + n);
			for (int i = tid; i < n; i += nt) kx[i] = ax[i], ki[i] = (uint32_t)i;
			__syncthreads();
			if (n <= RSORT_TIE_TILE) dev_radix_passes_cta<uint32_t>(kx, ki, n, s_bkt, s_ctl, s_stack, c.f + ab); /* every warp */
			else if (tid < 32) dev_radix_sort_warp<uint32_t, false>(kx, ki, n, s_bkt, c.f + ab);
			__syncthreads();
			for (int i = tid; i < n; i += nt) zy[i] = kx[i], zy[n + i] = ay[ki[i]];
			if (tid == 0) c.flags[r] |= 1u;
			__syncthreads();
			srcx = zy, srcy = zy + n;
		}
		if (n > 1) {
			/* the records ping-pong in the read's global scratch (L2 resident); staging them in shared memory was measured
			 * slower: it costs 3 of the 8 resident CTAs per SM */
			const uint64_t *S = rsort_read(srcx, n, seq_off, zx, zx + n, s_cnt, s_red);
			int tie = 0;
			for (int i = tid; i < n; i += nt) {
				const uint64_t rec = S[i];
				const uint32_t j = (uint32_t)(rec & RSORT_IDX_MASK);
				bx[i] = srcx[j], by[i] = srcy[j];
				if (i + 1 < n && (S[i + 1] >> RSORT_IDX_BITS) == (rec >> RSORT_IDX_BITS)) tie = 1;
			}
			if (!TIE && tie && n > 64) s_tie = 1;
			__syncthreads();
			if (!TIE && s_tie && tid == 0) tie_list[atomicAdd(n_tie, 1u)] = r;
		}
		if (n == 1 && tid == 0) bx[0] = ax[0], by[0] = ay[0];
		__syncthreads();
	}
}

static int g_sort_small_max = 512;
void mmg_sort_set_small_max(int v) { g_sort_small_max = v < 0 ? 0 : v > SORT_SMALL_ELEMS ? SORT_SMALL_ELEMS : v; }

int launch_sort(const ChunkDev &c, const DevIndex &di, uint32_t r0, uint32_t r1, int n_sms, cudaStream_t st, uint32_t *work)
{
	/* work[0]: read counter of the small pass, work[1]: length of big_list, work[2]: list counter of the radix pass,
	 * work[3]: length of tie_list, work[4]: list counter of the tie pass */
	const size_t smem_small = (size_t)SORT_SMALL_ELEMS * 12;
	int grid = n_sms * 8, need = (int)(r1 - r0);
	if (grid > need) grid = need;
	if (grid < 1) grid = 1;
	MMG_LAUNCH((sort_kernel<SORT_SMALL_ELEMS, false>), grid, SORT_THREADS, smem_small, st, c, r0, r1, work, c.big_list, work + 1, g_sort_small_max);
	MMG_LAUNCH((radix_sort_kernel<false>), grid, RSORT_WARPS * 32, 0, st, c, di.seq_off, work + 2, (const uint32_t*)c.big_list, (const uint32_t*)(work + 1), c.tie_list, work + 3);
	const size_t smem_tie = (size_t)RSORT_TIE_TILE * 12;
	static unsigned char attr_done[64];
	if (mmg_once_per_device(attr_done)) { cudaFuncSetAttribute(radix_sort_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tie); }
	grid = n_sms * 2;
	if (grid > need) grid = need;
	if (grid < 1) grid = 1;
	MMG_LAUNCH((radix_sort_kernel<true>), grid, RSORT_WARPS * 32, smem_tie, st, c, di.seq_off, work + 4, (const uint32_t*)c.tie_list, (const uint32_t*)(work + 3), c.tie_list, work + 3);
	return 0;
}
