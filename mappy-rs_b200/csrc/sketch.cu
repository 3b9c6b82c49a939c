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

#define SK_LEVELS 4   /* level rings of the doubling window minimum: w <= 32 */

template<typename KT> struct SkKey;
/* k <= 15: a k-mer and its hash fit 30 bits, every step is 32-bit arithmetic (hash64 restricted to 2k <= 30
 * bits only ever reads the low 2k bits of its intermediates, so the 32-bit evaluation is bit-identical) */
template<> struct SkKey<uint32_t> {
	static __device__ __forceinline__ uint32_t inf() { return 0xffffffffu; }
	static __device__ __forceinline__ uint32_t hash(uint32_t key, uint32_t mask)
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
	/* bits [s2, s2 + 32) of the 128-bit register (Whi:Wlo), s2 <= 62 */
	static __device__ __forceinline__ uint32_t cut(uint64_t Wlo, uint64_t Whi, int s2, uint32_t mask)
	{
		const uint32_t w0 = (uint32_t)Wlo, w1 = (uint32_t)(Wlo >> 32), w2 = (uint32_t)Whi;
		const bool up = s2 >= 32;
		return __funnelshift_r(up ? w1 : w0, up ? w2 : w1, (uint32_t)s2 & 31u) & mask;
	}
	static __device__ __forceinline__ uint32_t revcomp(uint32_t v, int k)
	{
		uint32_t r = __brev(v);
		r = ((r & 0xaaaaaaaau) >> 1) | ((r & 0x55555555u) << 1);
		return ~r >> (32 - 2 * k);
	}
};
template<> struct SkKey<uint64_t> {
	static __device__ __forceinline__ uint64_t inf() { return MMG_INF64; }
	static __device__ __forceinline__ uint64_t hash(uint64_t key, uint64_t mask)
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
	static __device__ __forceinline__ uint64_t cut(uint64_t Wlo, uint64_t Whi, int s2, uint64_t mask)
	{
		return ((Wlo >> s2) | (s2 ? Whi << (64 - s2) : 0)) & mask;
	}
	static __device__ __forceinline__ uint64_t revcomp(uint64_t v, int k)
	{
		uint64_t r = __brevll(v);
		r = ((r & 0xaaaaaaaaaaaaaaaaULL) >> 1) | ((r & 0x5555555555555555ULL) << 1);
		return ~r >> (64 - 2 * k);
	}
};

/* A,C,G,T/U (either case) -> 0..3, anything else 4 */
__device__ __forceinline__ int dev_nt4(unsigned c)
{
	const unsigned d = (c & 0xdfu) - 'A';
	const bool ok = d < 21u && ((0x180045u >> d) & 1u);
	unsigned code = (c >> 1) & 3u;
	code ^= code >> 1;
	return ok ? (int)code : 4;
}

/* The ring holds, per event, the hash (KT; all-ones = no usable k-mer), y = pos<<1|strand and l.
 * Read mode: read r is c.seq[c.off[r] .. c.off[r+1]), records go to the read's base offset.
 * Segment mode (c.seg_beg != 0; index construction): segment r is c.seq[seg_beg[r] .. +seg_len[r]), records go to
 * slot r * c.seg_cap. */
template<int RING, typename KT>
__global__ void __launch_bounds__(SKETCH_WARPS * 32)
sketch_kernel(ChunkDev c, int w, int k, uint32_t *work)
{
	MMG_DYN_SMEM(smem_raw);
	typedef SkKey<KT> K;
	const int lane = mmg_lane(), wib = threadIdx.x >> 5;
	const uint32_t RM = RING - 1;
	KT *xr = (KT*)smem_raw + (size_t)wib * RING;
	uint32_t *yr = (uint32_t*)((KT*)smem_raw + (size_t)SKETCH_WARPS * RING) + (size_t)wib * RING;
	uint32_t *lr = (uint32_t*)((KT*)smem_raw + (size_t)SKETCH_WARPS * RING) + (size_t)SKETCH_WARPS * RING + (size_t)wib * RING;
	/* level rings of the doubling window minimum (RING == 64, w <= 32: levels 1..4), after the three base rings */
	unsigned char *lvl_base = smem_raw + (size_t)SKETCH_WARPS * RING * (sizeof(KT) + 8) + (size_t)wib * SK_LEVELS * RING * (sizeof(KT) + 4);
	KT *lvl_v = (KT*)lvl_base;
	uint32_t *lvl_m = (uint32_t*)(lvl_base + (size_t)SK_LEVELS * RING * sizeof(KT));
	const KT mask = (KT)(((uint64_t)1 << 2 * k) - 1), INF = K::inf();
	const uint32_t lt = mmg_lanemask_lt();
	unsigned long long tot_mz = 0, tot_bases = 0;

	for (;;) {
		uint32_t r = mmg_next_item(work);
		if (r >= c.n_reads) break;
		r = mmg_read_of(c, r);
		uint64_t base;
		int len;
		const char *s;
		if (c.seg_beg) base = (uint64_t)r * c.seg_cap, len = (int)c.seg_len[r], s = c.seq + c.seg_beg[r];
		else base = c.off[r] - c.off0, len = (int)(c.off[r + 1] - c.off[r]), s = c.seq + c.off[r];
		uint64_t *ox = c.mz_x + base;
		uint32_t *oy = c.mz_y + base;

		for (int j = lane; j < RING; j += 32) xr[j] = INF;
		if (RING == 64) for (int j = lane; j < SK_LEVELS * RING; j += 32) lvl_v[j] = INF, lvl_m[j] = 1u << 8;
		__syncwarp();

		uint64_t prev = 0;          /* last 32 usable bases, newest in the low bits */
		int nvalid = 0;             /* usable bases so far */
		int e_base = 0;             /* events so far */
		int lastN = -1;             /* event index of the last ambiguous base */
		int pidx = -1; KT px = INF; /* selection after the last event */
		int n_out = 0;
		unsigned ch = lane < len ? (unsigned char)s[lane] : 0u;

		for (int pos0 = 0; pos0 < len; pos0 += 32) {
			const int i = pos0 + lane;
			const bool inb = i < len;
			const unsigned ch_next = i + 32 < len ? (unsigned char)s[i + 32] : 0u; /* in flight while this step computes */
			const int cc = inb ? dev_nt4(ch) : 4;
			ch = ch_next;
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
			const KT fwd = K::cut(Wlo, Whi, s2, mask);
			const KT rev = K::revcomp(fwd, k);
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
				KT x = INF;
				uint32_t y = 0;
				if (!isN && l >= k) {
					int z = fwd < rev ? 0 : 1;
					x = K::hash(z ? rev : fwd, mask);
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
			KT xe = INF, bxv = INF;
			int le = 0, bj = e, neq = 0;
			if (act) xe = xr[e & RM], le = (int)lr[e & RM];
			if (RING == 64) {
				/* P(e) by doubling: A_k[e] summarises the 2^k events ending at e as (minimum, age of its newest copy, number
				 * of copies); A_k[e] = A_(k-1)[e] (+) A_(k-1)[e - 2^(k-1)].  The window of w events is the disjoint union of
				 * one piece per set bit of w, newest first: A_K[e] from registers, the others from the level rings (events of
				 * earlier steps left theirs there).  log2(w) combines instead of w. */
				KT cv = xe;
				uint32_t cm = 1u << 8; /* age 0, one copy */
				const int K = 31 - __clz(w);     /* floor(log2(w)) */
				for (int q = 1; q <= K; ++q) {
					const int half = 1 << (q - 1);
					KT *lv = lvl_v + (size_t)(q > 1 ? q - 2 : 0) * RING;   /* ring of level q-1 (levels >= 1 sit at index level-1) */
					uint32_t *lm = lvl_m + (size_t)(q > 1 ? q - 2 : 0) * RING;
					if (q > 1 && act) lv[e & RM] = cv, lm[e & RM] = cm; /* level q-1 of this event (level 0 is xr itself) */
					__syncwarp();
					if (act) {
						const KT ov = q == 1 ? xr[(e - half) & RM] : lv[(e - half) & RM];
						const uint32_t om = q == 1 ? (1u << 8) : lm[(e - half) & RM];
						if (ov < cv) cv = ov, cm = om + (uint32_t)half;       /* older piece wins: its ages shift by the newer piece's size */
						else if (ov == cv) cm += om & 0xffffff00u;          /* tie: newest copy stays, counts add */
					}
				}
				/* the pieces of the lower set bits of w lie behind the 2^K newest events */
				int off = 1 << K;
				for (int q = K - 1; q >= 0; --q) {
					if (!((w >> q) & 1)) continue;
					if (act) {
						const KT ov = q == 0 ? xr[(e - off) & RM] : (lvl_v + (size_t)(q - 1) * RING)[(e - off) & RM];
						const uint32_t om = q == 0 ? (1u << 8) : (lvl_m + (size_t)(q - 1) * RING)[(e - off) & RM];
						if (ov < cv) cv = ov, cm = om + (uint32_t)off;
						else if (ov == cv) cm += om & 0xffffff00u;
					}
					off += 1 << q;
				}
				if (act) bxv = cv, bj = e - (int)(cm & 0xffu), neq = (int)(cm >> 8);
			} else if (act) {
				for (int j = e - w + 1; j <= e; ++j) {      /* P(e): newest among equal keys; neq = copies of it in the window */
					const KT vx = xr[j & RM];
					const bool less = vx < bxv, same = vx == bxv;
					neq = less ? 1 : neq + (same ? 1 : 0);
					if (less || same) bxv = vx, bj = j;
				}
			}
			KT pmx = __shfl_up_sync(MMG_FULL, bxv, 1);
			int pmi = __shfl_up_sync(MMG_FULL, bj, 1);
			if (lane == 0) pmx = px, pmi = pidx;
			const bool caseA = act && le == w + k - 1 && pmx != INF;
			const bool caseB = act && xe <= pmx;
			const bool caseC = act && !caseB && pmi == e - w;
			const bool emitB = caseB && le >= w + k && pmx != INF;
			const bool emitC = caseC && le >= w + k - 1;
			const bool dupC = emitC && bxv != INF && neq > 1;
			int cnt = (emitB || emitC) ? 1 : 0;
			if (caseA) for (int j = e - w + 1; j < e; ++j) cnt += (xr[j & RM] == pmx && j != pmi);
			if (dupC) cnt += neq - 1;
			int tot, o = mmg_warp_excl_scan(cnt, &tot);
			if (cnt) {
				o += n_out;
				if (caseA) for (int j = e - w + 1; j < e; ++j)
					if (xr[j & RM] == pmx && j != pmi) ox[o] = (uint64_t)xr[j & RM] << 8 | (uint64_t)k, oy[o] = yr[j & RM], ++o;
				if (emitB || emitC) ox[o] = (uint64_t)xr[pmi & RM] << 8 | (uint64_t)k, oy[o] = yr[pmi & RM], ++o;
				if (dupC) for (int j = e - w + 1; j <= e; ++j)
					if (xr[j & RM] == bxv && j != bj) ox[o] = (uint64_t)xr[j & RM] << 8 | (uint64_t)k, oy[o] = yr[j & RM], ++o;
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
			if (px != INF) ox[n_out] = (uint64_t)xr[pidx & RM] << 8 | (uint64_t)k, oy[n_out] = yr[pidx & RM], ++n_out;
			c.n_mz[r] = (uint32_t)n_out;
		}
		n_out = __shfl_sync(MMG_FULL, n_out, 0);
		tot_mz += n_out, tot_bases += len;
		__syncwarp();
	}
	if (lane == 0 && tot_bases && c.stats) {
		atomicAdd(&c.stats[0], tot_bases);
		atomicAdd(&c.stats[1], tot_mz);
	}
}

template<int RING, typename KT>
static void launch_sketch_t(const ChunkDev &c, int w, int k, int grid, cudaStream_t st, uint32_t *work)
{
	size_t smem = (size_t)SKETCH_WARPS * RING * (sizeof(KT) + 8) + (RING == 64 ? (size_t)SKETCH_WARPS * SK_LEVELS * RING * (sizeof(KT) + 4) : 0);
	if (smem > 48 * 1024) {
		static unsigned char attr_done[64];
		if (mmg_once_per_device(attr_done)) { cudaFuncSetAttribute(sketch_kernel<RING, KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); }
	}
	MMG_LAUNCH((sketch_kernel<RING, KT>), grid, SKETCH_WARPS * 32, smem, st, c, w, k, work);
}

int launch_sketch(const ChunkDev &c, const DevIndex &di, int n_sms, cudaStream_t st, uint32_t *work)
{
	const int w = di.w, k = di.k;
	int grid = n_sms * 8;
	int need = ((int)c.n_reads + SKETCH_WARPS - 1) / SKETCH_WARPS;
	if (grid > need) grid = need;
	if (grid < 1) grid = 1;
	if (w <= 32) {
		if (k <= 15) launch_sketch_t<64, uint32_t>(c, w, k, grid, st, work);
		else launch_sketch_t<64, uint64_t>(c, w, k, grid, st, work);
	} else {
		if (k <= 15) launch_sketch_t<512, uint32_t>(c, w, k, grid, st, work);
		else launch_sketch_t<512, uint64_t>(c, w, k, grid, st, work);
	}
	return 0;
}
