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

__global__ void __launch_bounds__(SORT_THREADS)
sort_kernel(ChunkDev c, uint32_t r0, uint32_t r1, uint32_t *work)
{
	MMG_DYN_SMEM(smem_raw);
	__shared__ uint32_t s_item;
	__shared__ int s_tie;
	__shared__ int s_bkt[512];
	uint64_t *sx = (uint64_t*)smem_raw;
	uint32_t *si = (uint32_t*)(sx + SORT_SMEM_ELEMS);
	const int tid = threadIdx.x, nt = blockDim.x;

	for (;;) {
		if (tid == 0) s_item = atomicAdd(work, 1u), s_tie = 0;
		__syncthreads();
		const uint32_t r = r0 + s_item;
		if (r >= r1) break;
		const int n = (int)c.n_a[r];
		const uint64_t ab = c.a_off[r] - c.a_off0;
		const uint64_t *ax = c.ax + ab, *ay = c.ay + ab;
		uint64_t *bx = c.bx + ab, *by = c.by + ab;
		if (n > 1) {
			int m = 1;
			while (m < n) m <<= 1;
			uint64_t *kx;
			uint32_t *ki;
			if (m <= SORT_SMEM_ELEMS) kx = sx, ki = si;
			else kx = c.zx + 2 * ab, ki = (uint32_t*)(c.zy + 2 * ab); /* global tile: m < 2n */
			for (int i = tid; i < m; i += nt) kx[i] = i < n ? ax[i] : MMG_INF64, ki[i] = (uint32_t)i;
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
				bx[i] = x, by[i] = ay[ki[i]];
				if (i + 1 < n && kx[i + 1] == x) tie = 1;
			}
			if (tie && n > 64) s_tie = 1;
			__syncthreads();
			if (s_tie) { /* replay upstream's unstable permutation from the unsorted input */
				for (int i = tid; i < n; i += nt) bx[i] = ax[i], by[i] = ay[i];
				__syncthreads();
				if (tid == 0) {
					dev_radix_sort_128x(bx, by, n, s_bkt, c.f + ab);
					c.flags[r] |= 1u;
				}
			}
		} else if (n == 1) {
			if (tid == 0) bx[0] = ax[0], by[0] = ay[0];
		}
		__syncthreads();
	}
}

int launch_sort(const ChunkDev &c, uint32_t r0, uint32_t r1, int n_sms, cudaStream_t st, uint32_t *work)
{
	size_t smem = (size_t)SORT_SMEM_ELEMS * 12;
	static bool attr_done = false;
	if (!attr_done) { cudaFuncSetAttribute(sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_done = true; }
	int grid = n_sms * 2, need = (int)(r1 - r0);
	if (grid > need) grid = need;
	if (grid < 1) grid = 1;
	MMG_LAUNCH(sort_kernel, grid, SORT_THREADS, smem, st, c, r0, r1, work);
	return 0;
}
