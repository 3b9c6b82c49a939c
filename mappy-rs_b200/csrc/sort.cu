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
sort_kernel(ChunkDev c, uint32_t r0, uint32_t r1, uint32_t *work, uint32_t *big_list, uint32_t *n_big)
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
		const uint32_t r = LIST ? big_list[s_item] : r0 + s_item;
		const int n = (int)c.n_a[r];
		if (!LIST && n > ELEMS) { /* uniform over the CTA */
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

int launch_sort(const ChunkDev &c, uint32_t r0, uint32_t r1, int n_sms, cudaStream_t st, uint32_t *work)
{
	/* work[0]: read counter of the small pass, work[1]: length of big_list, work[2]: list counter of the big pass */
	const size_t smem_big = (size_t)SORT_SMEM_ELEMS * 12, smem_small = (size_t)SORT_SMALL_ELEMS * 12;
	static bool attr_done = false;
	if (!attr_done) { cudaFuncSetAttribute(sort_kernel<SORT_SMEM_ELEMS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_big); attr_done = true; }
	int grid = n_sms * 8, need = (int)(r1 - r0);
	if (grid > need) grid = need;
	if (grid < 1) grid = 1;
	MMG_LAUNCH((sort_kernel<SORT_SMALL_ELEMS, false>), grid, SORT_THREADS, smem_small, st, c, r0, r1, work, c.big_list, work + 1);
	grid = n_sms * 2;
	if (grid > need) grid = need;
	if (grid < 1) grid = 1;
	MMG_LAUNCH((sort_kernel<SORT_SMEM_ELEMS, true>), grid, SORT_THREADS, smem_big, st, c, r0, r1, work + 2, c.big_list, work + 1);
	return 0;
}
