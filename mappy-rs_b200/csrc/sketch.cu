/* sketch.cu -- batched (w,k)-minimizer sketch, one warp per read (north-star (b)).
 *
 * Replaces mm_sketch() + hash64() reached through collect_minimizers() in
 * mm_map (/root/reference/src/lib.rs:482,587; upstream sketch.c / map.c,
 * minimap2 v2.26).  Output is bit-identical to mm_sketch(): records
 * x = hash<<8|span, y = pos<<1|strand in emission order.
 *
 * Formulation.  mm_sketch is a serial state machine (ring buffer of the last w
 * k-mers, running minimum, l = run length of usable bases).  Here every step
 * handles 32 bases in parallel:
 *   1. bases -> 2-bit codes; usable bases are compacted with a ballot and their
 *      k-mers are cut out of a 128-bit shift register assembled with two
 *      __reduce_or_sync (no per-base dependency chain);
 *   2. symmetric k-mers (fwd == rev) are dropped BEFORE the window, exactly as
 *      upstream's `continue` does; what remains, plus every ambiguous base, is
 *      the event stream; l is the distance to the last ambiguous event;
 *   3. events go to a shared-memory ring; lane t then owns event e_base+t and
 *      finds P(e) = min over the last w events (ties to the newest) by a
 *      w-step scan, gets P(e-1) from its neighbour with one shuffle, and emits
 *      records under upstream's three rules (first full window, new minimum,
 *      minimum left the window) including the duplicate-key rules;
 *   4. an exclusive warp scan of the per-lane record counts gives the output
 *      slots, so the output order equals the serial one.
 * Bound: integer issue (~60 ops/base) - see DESIGN.md.
 */
#include "dev_common.cuh"
#include "stages.h"

__device__ __forceinline__ uint64_t dev_hash64(uint64_t key, uint64_t mask)
{
	key = (~key + (key << 21)) & mask;
	key = key ^ key >> 24;
	key = ((key + (key << 3)) + (key << 8)) & mask;
	key = key ^ key >> 14;
	key = ((key + (key << 2)) + (key << 4)) & mask;
	key = key ^ key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}

__device__ __forceinline__ int dev_nt4(unsigned c)
{
	unsigned u = c & 0xdfu;
	return u == 'A' ? 0 : u == 'C' ? 1 : u == 'G' ? 2 : (u == 'T' || u == 'U') ? 3 : 4;
}

/* reverse complement of the k 2-bit bases in the low 2k bits of v */
__device__ __forceinline__ uint64_t dev_revcomp(uint64_t v, int k)
{
	uint64_t r = __brevll(v);
	r = ((r & 0xaaaaaaaaaaaaaaaaULL) >> 1) | ((r & 0x5555555555555555ULL) << 1);
	r = ~r;
	return r >> (64 - 2 * k);
}

template<int RING>
__global__ void __launch_bounds__(SKETCH_WARPS * 32)
sketch_kernel(ChunkDev c, int w, int k, uint32_t *work)
{
	MMG_DYN_SMEM(smem_raw);
	const int lane = mmg_lane(), wib = threadIdx.x >> 5;
	const uint32_t RM = RING - 1;
	uint64_t *xr = (uint64_t*)smem_raw + (size_t)wib * RING;
	uint32_t *yr = (uint32_t*)((uint64_t*)smem_raw + (size_t)SKETCH_WARPS * RING) + (size_t)wib * RING;
	uint32_t *lr = (uint32_t*)((uint64_t*)smem_raw + (size_t)SKETCH_WARPS * RING) + (size_t)SKETCH_WARPS * RING + (size_t)wib * RING;
	const uint64_t mask = (1ULL << 2 * k) - 1;
	const uint32_t lt = mmg_lanemask_lt();
	unsigned long long tot_mz = 0, tot_bases = 0;

	for (;;) {
		uint32_t r = mmg_next_item(work);
		if (r >= c.n_reads) break;
		const uint64_t base = c.off[r] - c.off0;
		const int len = (int)(c.off[r + 1] - c.off[r]);
		const char *s = c.seq + c.off[r];
		uint64_t *ox = c.mz_x + base;
		uint32_t *oy = c.mz_y + base;

		for (int j = lane; j < RING; j += 32) xr[j] = MMG_INF64;
		__syncwarp();

		uint64_t prev = 0;          /* last 32 usable bases, newest in the low bits */
		int nvalid = 0;             /* usable bases so far */
		int e_base = 0;             /* events so far */
		int lastN = -1;             /* event index of the last ambiguous base */
		int pidx = -1; uint64_t px = MMG_INF64; /* selection after the last event */
		int n_out = 0;

		for (int pos0 = 0; pos0 < len; pos0 += 32) {
			const int i = pos0 + lane;
			const bool inb = i < len;
			const int cc = inb ? dev_nt4((unsigned char)s[i]) : 4;
			const bool valid = inb && cc < 4, isN = inb && cc == 4;
			const uint32_t vmask = __ballot_sync(MMG_FULL, valid);
			const int nv = __popc(vmask), ci = __popc(vmask & lt);
			/* pack this step's usable bases, newest lowest */
			uint64_t contrib = valid ? (uint64_t)cc << (2 * (nv - 1 - ci)) : 0;
			uint32_t lo = __reduce_or_sync(MMG_FULL, (uint32_t)contrib);
			uint32_t hi = __reduce_or_sync(MMG_FULL, (uint32_t)(contrib >> 32));
			const uint64_t cur = (uint64_t)hi << 32 | lo;
			const int sh = 2 * nv;
			const uint64_t Wlo = sh == 64 ? cur : sh == 0 ? prev : (prev << sh) | cur;
			const uint64_t Whi = sh == 64 ? prev : sh == 0 ? 0 : prev >> (64 - sh);
			const int s2 = valid ? 2 * (nv - 1 - ci) : 0;
			const uint64_t fwd = ((Wlo >> s2) | (s2 ? Whi << (64 - s2) : 0)) & mask;
			const uint64_t rev = dev_revcomp(fwd, k);
			/* a k-mer assembled from fewer than k usable bases can never equal its
			 * (zero-filled) reverse upstream, so the symmetric test starts at k bases */
			const bool pal = valid && (nvalid + ci + 1 >= k) && fwd == rev;
			const bool isev = isN || (valid && !pal);
			const uint32_t emask = __ballot_sync(MMG_FULL, isev), nmask = __ballot_sync(MMG_FULL, isN);
			const int ei = e_base + __popc(emask & lt);
			int lastN_e = lastN;
			{
				uint32_t nb = nmask & lt;
				if (nb) { int ln = 31 - __clz((int)nb); lastN_e = e_base + __popc(emask & ((1u << ln) - 1u)); }
			}
			const int l = isN ? 0 : ei - lastN_e;
			if (isev) {
				uint64_t x = MMG_INF64;
				uint32_t y = 0;
				if (!isN && l >= k) {
					int z = fwd < rev ? 0 : 1;
					x = dev_hash64(z ? rev : fwd, mask) << 8 | (uint64_t)k;
					y = (uint32_t)i << 1 | (uint32_t)z;
				}
				xr[ei & RM] = x, yr[ei & RM] = y, lr[ei & RM] = (uint32_t)l;
			}
			prev = Wlo;
			nvalid += nv;
			if (nmask) { int ln = 31 - __clz((int)nmask); lastN = e_base + __popc(emask & ((1u << ln) - 1u)); }
			const int n_ev = __popc(emask);
			__syncwarp();

			/* ---- window minimum + emission: lane t owns event e_base + t ---- */
			const bool act = lane < n_ev;
			const int e = e_base + lane;
			uint64_t xe = MMG_INF64, bxv = MMG_INF64;
			int le = 0, bj = e;
			if (act) {
				xe = xr[e & RM], le = (int)lr[e & RM];
				for (int j = e - w + 1; j <= e; ++j) {      /* P(e): newest among equal keys */
					uint64_t vx = xr[j & RM];
					if (bxv >= vx) bxv = vx, bj = j;
				}
			}
			uint64_t pmx = __shfl_up_sync(MMG_FULL, bxv, 1);
			int pmi = __shfl_up_sync(MMG_FULL, bj, 1);
			if (lane == 0) pmx = px, pmi = pidx;
			const bool caseA = act && le == w + k - 1 && pmx != MMG_INF64;
			const bool caseB = act && xe <= pmx;
			const bool caseC = act && !caseB && pmi == e - w;
			const bool emitB = caseB && le >= w + k && pmx != MMG_INF64;
			const bool emitC = caseC && le >= w + k - 1;
			const bool dupC = emitC && bxv != MMG_INF64;
			int cnt = (emitB || emitC) ? 1 : 0;
			if (caseA) for (int j = e - w + 1; j < e; ++j) cnt += (xr[j & RM] == pmx && j != pmi);
			if (dupC) for (int j = e - w + 1; j <= e; ++j) cnt += (xr[j & RM] == bxv && j != bj);
			int tot, o = mmg_warp_excl_scan(cnt, &tot);
			if (cnt) {
				o += n_out;
				if (caseA) for (int j = e - w + 1; j < e; ++j)
					if (xr[j & RM] == pmx && j != pmi) ox[o] = xr[j & RM], oy[o] = yr[j & RM], ++o;
				if (emitB || emitC) ox[o] = xr[pmi & RM], oy[o] = yr[pmi & RM], ++o;
				if (dupC) for (int j = e - w + 1; j <= e; ++j)
					if (xr[j & RM] == bxv && j != bj) ox[o] = xr[j & RM], oy[o] = yr[j & RM], ++o;
			}
			n_out += tot;
			if (n_ev > 0) {
				px = __shfl_sync(MMG_FULL, bxv, n_ev - 1);
				pidx = __shfl_sync(MMG_FULL, bj, n_ev - 1);
			}
			e_base += n_ev;
			__syncwarp();
		}
		if (lane == 0) {
			if (px != MMG_INF64) ox[n_out] = xr[pidx & RM], oy[n_out] = yr[pidx & RM], ++n_out;
			c.n_mz[r] = (uint32_t)n_out;
		}
		n_out = __shfl_sync(MMG_FULL, n_out, 0);
		tot_mz += n_out, tot_bases += len;
		__syncwarp();
	}
	if (lane == 0 && tot_bases) {
		atomicAdd(&c.stats[0], tot_bases);
		atomicAdd(&c.stats[1], tot_mz);
	}
}

int launch_sketch(const ChunkDev &c, const DevIndex &di, int n_sms, cudaStream_t st, uint32_t *work)
{
	const int w = di.w, k = di.k;
	int grid = n_sms * 8;
	int need = ((int)c.n_reads + SKETCH_WARPS - 1) / SKETCH_WARPS;
	if (grid > need) grid = need;
	if (grid < 1) grid = 1;
	if (w <= 32) {
		size_t smem = (size_t)SKETCH_WARPS * 64 * 16;
		MMG_LAUNCH(sketch_kernel<64>, grid, SKETCH_WARPS * 32, smem, st, c, w, k, work);
	} else {
		size_t smem = (size_t)SKETCH_WARPS * 512 * 16;
		static bool attr_done = false;
		if (!attr_done) { cudaFuncSetAttribute(sketch_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_done = true; }
		MMG_LAUNCH(sketch_kernel<512>, grid, SKETCH_WARPS * 32, smem, st, c, w, k, work);
	}
	return 0;
}
