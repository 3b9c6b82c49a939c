/* chain.cu -- anchor chaining: warp-cooperative DP, backtrack, chain compaction
 * (north-star (d)), one warp per read.
 *
 * Replaces lchain.c mm_lchain_dp (comput_sc, mg_log2), mg_chain_backtrack,
 * mg_chain_bk_end and compact_a of minimap2 v2.26 on the mm_map path
 * (/root/reference/src/lib.rs:482,587).  Bit-exact, including the max_skip
 * early exit, the t[] marks, the `max_ii` shortcut and upstream's unstable sort
 * of chain ends.
 *
 * DP.  Anchor i scans its predecessors j = i-1 .. st in steps of 32 lanes
 * (descending j = ascending lane).  Upstream's loop carries three serial pieces
 * of state; each has a warp form that gives the same result:
 *   - max_f/max_j (strict improvement, first best wins): exclusive prefix max
 *     over the lanes tells every lane whether it WOULD have improved;
 *   - t[] marks (t[p[j]] = i): marks only ever point to smaller j, i.e. to
 *     later lanes or later steps, so "all lanes store, __syncwarp, all lanes
 *     load" shows every lane exactly the marks of its predecessors; marks
 *     stored by lanes past the break point are never read again for this i;
 *   - n_skip (saturating counter, break above max_skip): replayed over the
 *     ballot bits of improving / marked lanes (closed form when nothing is marked).
 * f/p/t live in HBM slices of the read (L1/L2 resident while the read is active).
 * Bound: INT32 issue + L1 latency of the i -> i+1 dependency - see DESIGN.md.
 */
#include <stdlib.h>
#include "dev_common.cuh"
#include "dev_sort.cuh"
#include "dev_chain.cuh"
#include "stages.h"

__device__ __forceinline__ float dev_mg_log2(float x) /* lchain.c mg_log2; x >= 2 */
{
	uint32_t zi = __float_as_uint(x);
	float log_2 = (float)(((zi >> 23) & 255u) - 128u);
	zi &= ~(255u << 23);
	zi += 127u << 23;
	float zf = __uint_as_float(zi);
	float t = __fadd_rn(__fmul_rn(-0.34484843f, zf), 2.02466578f);
	t = __fsub_rn(__fmul_rn(t, zf), 0.67487759f);
	return __fadd_rn(log_2, t);
}

__device__ __forceinline__ int32_t dev_comput_sc(uint64_t aix, uint64_t aiy, uint64_t ajx, uint64_t ajy,
                                                 int32_t max_dist_x, int32_t max_dist_y, int32_t bw, float pen_gap, float pen_skip)
{
	int32_t dq = (int32_t)aiy - (int32_t)ajy, dr, dd, dg, q_span, sc;
	if (dq <= 0 || dq > max_dist_x) return INT32_MIN_;
	dr = (int32_t)(aix - ajx);
	if (dr == 0 || dq > max_dist_y) return INT32_MIN_;
	dd = dr > dq ? dr - dq : dq - dr;
	if (dd > bw) return INT32_MIN_;
	dg = dr < dq ? dr : dq;
	q_span = (int32_t)(ajy >> 32 & 0xff);
	sc = q_span < dg ? q_span : dg;
	if (dd || dg > q_span) {
		float lin_pen = __fadd_rn(__fmul_rn(pen_gap, (float)dd), __fmul_rn(pen_skip, (float)dg));
		float log_pen = dd >= 1 ? dev_mg_log2((float)(dd + 1)) : 0.0f;
		sc -= (int)__fadd_rn(lin_pen, __fmul_rn(.5f, log_pen));
	}
	return sc;
}

/* is anchor a (a <= b in sort order) outside b's window: another strand/contig, or more than d behind on the target?
 * Same strand and contig means equal high words, and then the low words (target positions) are ordered like the keys. */
__device__ __forceinline__ bool chain_out_of_range(uint64_t a, uint64_t b, uint32_t d)
{
	return (uint32_t)(a >> 32) != (uint32_t)(b >> 32) || (uint32_t)b - (uint32_t)a > d;
}

__global__ void __launch_bounds__(CHAIN_WARPS * 32)
chain_dp_kernel(ChunkDev c, DevOpt o, uint32_t r0, uint32_t r1, uint32_t *work)
{
	const int lane = mmg_lane();
	const uint32_t lt = mmg_lanemask_lt();
	unsigned long long tot_iter = 0, tot_anchor = 0;
	for (;;) {
		uint32_t r = r0 + mmg_next_item(work);
		if (r >= r1) break;
		r = mmg_read_of(c, r);
		const int n = (int)c.n_a[r];
		const uint64_t ab = c.a_off[r] - c.a_off0;
		const int qlen = (int)(c.off[r + 1] - c.off[r]);
		const uint64_t *ax = c.bx + ab, *ay = c.by + ab;
		int32_t *f = c.f + ab, *p = c.p + ab, *t = c.t + ab;
		/* map.c mm_map_frag: chaining gaps */
		int32_t max_dist_y = o.max_gap, max_dist_x;
		if (o.max_gap_ref > 0) max_dist_x = o.max_gap_ref;
		else if (o.max_frag_len > 0) { max_dist_x = o.max_frag_len - qlen; if (max_dist_x < o.max_gap) max_dist_x = o.max_gap; }
		else max_dist_x = o.max_gap;
		if (max_dist_x < o.bw) max_dist_x = o.bw;
		if (max_dist_y < o.bw) max_dist_y = o.bw;
		const int32_t bw = o.bw, max_skip = o.max_chain_skip, max_iter = o.max_chain_iter;
		const float pen_gap = o.chn_pen_gap, pen_skip = o.chn_pen_skip;

		int st = 0, max_ii = -1;
		uint64_t mii_x = 0;
		int32_t mii_f = 0;
		bool try_bulk = true; /* only after an anchor that had no predecessor: dense chains never pay for the test */
		for (int i = 0; i < n; ++i) {
			if (try_bulk) { /* Anchors without any predecessor in range (the previous anchor is on another strand/contig or
			   * more than max_dist_x behind - most index hits on a large reference) are settled 32 at a time:
			   * f = span, p = -1, and the loop state afterwards is what the scalar path leaves: st at the
			   * anchor itself, and the anchor as the `max_ii` candidate. */
				const int j = i + lane;
				uint64_t xj = 0, xp = 0;
				if (j < n) { xj = ax[j]; if (j > 0) xp = ax[j - 1]; }
				const bool iso = j < n && (j == 0 || chain_out_of_range(xp, xj, (uint32_t)max_dist_x));
				const uint32_t im = __ballot_sync(MMG_FULL, iso);
				const int run = im == MMG_FULL ? 32 : __ffs((int)~im) - 1;
				if (run > 0) {
					int32_t sp = 0;
					if (lane < run) { sp = (int32_t)(ay[j] >> 32 & 0xff); f[j] = sp, p[j] = -1, t[j] = 0; }
					const int last = run - 1;
					st = i + last, max_ii = i + last;
					mii_x = __shfl_sync(MMG_FULL, xj, last), mii_f = __shfl_sync(MMG_FULL, sp, last);
					i += last;
					__syncwarp();
					continue;
				}
				try_bulk = false;
			}
			const uint64_t aix = ax[i], aiy = ay[i];
			/* advance st: first j that shares the target strand and is within max_dist_x (usually st itself still is) */
			bool st_out;
			st_out = st < i && chain_out_of_range(ax[st], aix, (uint32_t)max_dist_x);
			while (st_out) {
				int j = st + lane;
				bool out = j < i && chain_out_of_range(ax[j], aix, (uint32_t)max_dist_x);
				uint32_t m = __ballot_sync(MMG_FULL, out);
				int lead = m == MMG_FULL ? 32 : __ffs((int)~m) - 1;
				st += lead;
				if (lead < 32) break;
			}
			if (i - st > max_iter) st = i - max_iter;
			try_bulk = st == i;
			int32_t max_f = (int32_t)(aiy >> 32 & 0xff), n_skip = 0;
			int max_j = -1, end_j = st - 1;
			bool broke = false;
			for (int jb = i - 1; jb >= st && !broke; jb -= 32) {
				const int j = jb - lane;
				const bool inr = j >= st;
				int32_t sc = INT32_MIN_, pj = -1;
				if (inr) {
					sc = dev_comput_sc(aix, aiy, ax[j], ay[j], max_dist_x, max_dist_y, bw, pen_gap, pen_skip);
					if (sc != INT32_MIN_) { sc += f[j]; pj = p[j]; }
				}
				const bool ok = sc != INT32_MIN_;
				if (ok && pj >= 0) t[pj] = i;
				__syncwarp();
				const bool marked = ok && t[j] == i;
				/* would lane improve max_f?  Only lanes up to the first lane that holds the step's maximum can:
				 * if that is lane 0 (the usual case in a colinear chain: the nearest predecessor scores best), or
				 * the maximum does not beat the carried max_f, no scan is needed; otherwise an exclusive prefix
				 * max over earlier lanes and the carried max_f decides. */
				const int32_t mxs = __reduce_max_sync(MMG_FULL, sc);
				const uint32_t top = __ballot_sync(MMG_FULL, sc == mxs);
				bool improve;
				if (mxs <= max_f) improve = false;
				else if (top & 1u) improve = lane == 0;
				else {
					int32_t incl = sc;
#pragma unroll
					for (int d = 1; d < 32; d <<= 1) {
						int32_t v = __shfl_up_sync(MMG_FULL, incl, d);
						if (lane >= d && v > incl) incl = v;
					}
					int32_t excl = __shfl_up_sync(MMG_FULL, incl, 1);
					if (lane == 0 || excl < max_f) excl = max_f;
					improve = ok && sc > excl;
				}
				const bool mk = marked && !improve;
				const uint32_t I = __ballot_sync(MMG_FULL, improve), M = __ballot_sync(MMG_FULL, mk);
				int brk = -1;
				if (M == 0) { n_skip -= __popc(I); if (n_skip < 0) n_skip = 0; }
				else if (I < (M & (0u - M))) {
					/* every improving lane precedes every marked lane: n_skip drops first, then only counts up */
					int32_t n0 = n_skip - __popc(I);
					if (n0 < 0) n0 = 0;
					const int need = max_skip - n0 + 1; /* the marked lane at which n_skip exceeds max_skip */
					if (__popc(M) >= need) {
						const uint32_t at = __ballot_sync(MMG_FULL, mk && __popc(M & (lt | (1u << lane))) == need);
						brk = __ffs((int)at) - 1;
					} else n_skip = n0 + __popc(M);
				} else {
					/* every lane is a map x -> max(x + sa, sc): improving = (-1, 0), marked = (+1, -inf),
					 * otherwise the identity; maps of this form are closed under composition, so an
					 * inclusive warp scan gives n_skip after every lane */
					int32_t sa = improve ? -1 : mk ? 1 : 0, sc2 = improve ? 0 : -(1 << 28);
#pragma unroll
					for (int d = 1; d < 32; d <<= 1) {
						int32_t pa = __shfl_up_sync(MMG_FULL, sa, d), pc = __shfl_up_sync(MMG_FULL, sc2, d);
						if (lane >= d) { pc += sa; sc2 = pc > sc2 ? pc : sc2; sa += pa; }
					}
					int32_t after = n_skip + sa;
					if (after < sc2) after = sc2;
					const uint32_t over = __ballot_sync(MMG_FULL, mk && after > max_skip);
					if (over) brk = __ffs((int)over) - 1;
					else n_skip = __shfl_sync(MMG_FULL, after, 31);
				}
				/* best score among the lanes before the break; without a break that is the step's maximum from above */
				int32_t mx = mxs;
				uint32_t at_mx = top;
				if (brk >= 0) {
					const int32_t c2 = lane < brk ? sc : INT32_MIN_;
					mx = __reduce_max_sync(MMG_FULL, c2);
					at_mx = __ballot_sync(MMG_FULL, c2 == mx);
				}
				if (mx > max_f) {
					max_f = mx;
					max_j = jb - (__ffs((int)at_mx) - 1);
				}
				if (brk >= 0) end_j = jb - brk, broke = true;
			}
			tot_iter += broke ? i - end_j : i - st;
			/* lchain.c: the best-scoring anchor in range is tried even if the loop stopped before it */
			if (max_ii < 0 || aix - mii_x > (uint64_t)(int64_t)max_dist_x) {
				int32_t best = INT32_MIN_;
				int bestj = -1;
				for (int jb = i - 1; jb >= st; jb -= 32) {
					int j = jb - lane;
					if (j >= st) { int32_t v = f[j]; if (v > best) best = v, bestj = j; }
				}
				int32_t mx = __reduce_max_sync(MMG_FULL, best);
				int cj = best == mx ? bestj : -1;
				max_ii = __reduce_max_sync(MMG_FULL, cj);
				if (max_ii >= 0) mii_x = ax[max_ii], mii_f = f[max_ii];
			}
			if (max_ii >= 0 && max_ii < end_j) {
				int32_t tmp = dev_comput_sc(aix, aiy, mii_x, ay[max_ii], max_dist_x, max_dist_y, bw, pen_gap, pen_skip);
				if (tmp != INT32_MIN_ && max_f < tmp + mii_f) max_f = tmp + mii_f, max_j = max_ii;
			}
			if (lane == 0) f[i] = max_f, p[i] = max_j, t[i] = 0;
			if (max_ii < 0 || (aix - mii_x <= (uint64_t)(int64_t)max_dist_x && mii_f < max_f))
				max_ii = i, mii_x = aix, mii_f = max_f;
			__syncwarp();
		}
		tot_anchor += n;
	}
	if (lane == 0 && tot_anchor) {
		atomicAdd(&c.stats[4], tot_anchor);
		atomicAdd(&c.stats[5], tot_iter);
	}
}

/* ---- the same DP, several reads per warp ------------------------------------------------------------------
 * After the isolated-anchor filter an anchor has ~11 predecessors in range, so a 32-lane step of chain_dp_kernel is
 * a third full and the ~200 warp instructions an anchor costs are spent on 11 cells.  Here a warp is cut into
 * 32 / G lane groups; every group owns a read and walks its anchors with G-lane steps, and the groups run in
 * lockstep through the phases of one anchor each (settle isolated anchors / advance st / predecessor steps /
 * refresh max_ii / finalise), so one instruction stream serves 32 / G anchors.  The serial pieces of upstream's
 * loop are resolved exactly as in chain_dp_kernel, always with the general forms (prefix maximum for "would this
 * lane have improved max_f", composition of the maps x -> max(x + a, c) for n_skip) over G lanes. */
template<int G> __device__ __forceinline__ int32_t grp_max(int32_t v)
{
#pragma unroll
	for (int d = G >> 1; d; d >>= 1) { const int32_t o = __shfl_xor_sync(MMG_FULL, v, d, G); v = o > v ? o : v; }
	return v;
}
template<int G> __device__ __forceinline__ long long grp_max64(long long v)
{
#pragma unroll
	for (int d = G >> 1; d; d >>= 1) { const long long o = __shfl_xor_sync(MMG_FULL, v, d, G); v = o > v ? o : v; }
	return v;
}

template<int G>
__global__ void __launch_bounds__(CHAIN_WARPS * 32)
chain_dp_group_kernel(ChunkDev c, DevOpt o, uint32_t r0, uint32_t r1, uint32_t *work)
{
	const int lane = mmg_lane(), gl = lane & (G - 1), gbase = lane & ~(G - 1);
	const uint32_t gmask = G == 32 ? 0xffffffffu : ((1u << G) - 1u);
	const int32_t bw = o.bw, max_skip = o.max_chain_skip, max_iter = o.max_chain_iter;
	const float pen_gap = o.chn_pen_gap, pen_skip = o.chn_pen_skip;
	unsigned long long tot_iter = 0, tot_anchor = 0;
	/* the read of my group */
	bool done = false;
	int n = 0, i = 0, st = 0, max_ii = -1;
	uint64_t mii_x = 0;
	int32_t mii_f = 0, max_dist_x = 0, max_dist_y = 0;
	bool try_bulk = true;
	const uint64_t *ax = 0, *ay = 0;
	int32_t *f = 0, *p = 0, *t = 0;
	for (;;) {
		/* ---- groups that finished their read take the next one (all collectives stay warp-wide and converged) ---- */
		for (;;) {
			const bool want = !done && i >= n;
			if (!__any_sync(MMG_FULL, want)) break;
			uint32_t r = 0;
			if (want && gl == 0) r = r0 + atomicAdd(work, 1u);
			r = __shfl_sync(MMG_FULL, r, 0, G);
			if (!want) continue;
			if (r >= r1) { done = true; continue; }
			r = mmg_read_of(c, r);
			n = (int)c.n_a[r];
			const uint64_t ab = c.a_off[r] - c.a_off0;
			const int qlen = (int)(c.off[r + 1] - c.off[r]);
			ax = c.bx + ab, ay = c.by + ab, f = c.f + ab, p = c.p + ab, t = c.t + ab;
			max_dist_y = o.max_gap;                       /* map.c mm_map_frag: chaining gaps */
			if (o.max_gap_ref > 0) max_dist_x = o.max_gap_ref;
			else if (o.max_frag_len > 0) { max_dist_x = o.max_frag_len - qlen; if (max_dist_x < o.max_gap) max_dist_x = o.max_gap; }
			else max_dist_x = o.max_gap;
			if (max_dist_x < bw) max_dist_x = bw;
			if (max_dist_y < bw) max_dist_y = bw;
			i = 0, st = 0, max_ii = -1, mii_x = 0, mii_f = 0, try_bulk = true;
			if (gl == 0) tot_anchor += (unsigned long long)n;
		}
		if (__all_sync(MMG_FULL, done)) break;
		bool act = !done;                                 /* my group works on anchor i in this round */
		/* ---- anchors without any predecessor in range, G at a time (see chain_dp_kernel) ---- */
		{
			const bool tb = act && try_bulk;
			const int j = i + gl;
			uint64_t xj = 0, xp = 0;
			if (tb && j < n) { xj = ax[j]; if (j > 0) xp = ax[j - 1]; }
			const bool iso = tb && j < n && (j == 0 || chain_out_of_range(xp, xj, (uint32_t)max_dist_x));
			const uint32_t im = (__ballot_sync(MMG_FULL, iso) >> gbase) & gmask;
			const int run = tb ? (im == gmask ? G : __ffs((int)~im) - 1) : 0;
			int32_t sp = 0;
			if (gl < run) { sp = (int32_t)(ay[j] >> 32 & 0xff); f[j] = sp, p[j] = -1, t[j] = 0; }
			const int last = run > 0 ? run - 1 : 0;
			const uint64_t lx = __shfl_sync(MMG_FULL, xj, last, G);
			const int32_t lf = __shfl_sync(MMG_FULL, sp, last, G);
			if (run > 0) {
				st = i + last, max_ii = i + last, mii_x = lx, mii_f = lf;
				i += run;
				act = false;                                  /* this round is spent */
			} else if (tb) try_bulk = false;
		}
		uint64_t aix = 0, aiy = 0;
		if (act) aix = ax[i], aiy = ay[i];
		/* ---- advance st: first j that shares the target strand and is within max_dist_x ---- */
		{
			bool st_out = act && st < i && chain_out_of_range(ax[st], aix, (uint32_t)max_dist_x);
			while (__any_sync(MMG_FULL, st_out)) {
				const int j = st + gl;
				const bool out = st_out && j < i && chain_out_of_range(ax[j], aix, (uint32_t)max_dist_x);
				const uint32_t m = (__ballot_sync(MMG_FULL, out) >> gbase) & gmask;
				if (st_out) {
					const int lead = m == gmask ? G : __ffs((int)~m) - 1;
					st += lead;
					if (lead < G) st_out = false;
				}
			}
			if (act) {
				if (i - st > max_iter) st = i - max_iter;
				try_bulk = st == i;
			}
		}
		/* ---- predecessor steps ---- */
		int32_t max_f = (int32_t)(aiy >> 32 & 0xff), n_skip = 0;
		int max_j = -1, end_j = st - 1, jb = i - 1;
		bool broke = false;
		for (;;) {
			const bool ga = act && !broke && jb >= st;
			if (!__any_sync(MMG_FULL, ga)) break;
			const int j = jb - gl;
			const bool inr = ga && j >= st;
			int32_t sc = INT32_MIN_, pj = -1;
			if (inr) {
				sc = dev_comput_sc(aix, aiy, ax[j], ay[j], max_dist_x, max_dist_y, bw, pen_gap, pen_skip);
				if (sc != INT32_MIN_) { sc += f[j]; pj = p[j]; }
			}
			const bool ok = sc != INT32_MIN_;
			if (ok && pj >= 0) t[pj] = i;
			__syncwarp();
			const bool marked = ok && t[j] == i;
			/* would this lane have improved max_f?  exclusive prefix maximum over the earlier lanes and the carried max_f */
			int32_t incl = sc;
#pragma unroll
			for (int d = 1; d < G; d <<= 1) {
				const int32_t v = __shfl_up_sync(MMG_FULL, incl, d, G);
				if (gl >= d && v > incl) incl = v;
			}
			int32_t excl = __shfl_up_sync(MMG_FULL, incl, 1, G);
			if (gl == 0 || excl < max_f) excl = max_f;
			const bool improve = ok && sc > excl;
			const bool mk = marked && !improve;
			/* n_skip after every lane: composition of x -> max(x + sa, sc2) (improving: (-1, 0); marked: (+1, -inf)) */
			int32_t sa = improve ? -1 : mk ? 1 : 0, sc2 = improve ? 0 : -(1 << 28);
#pragma unroll
			for (int d = 1; d < G; d <<= 1) {
				const int32_t pa = __shfl_up_sync(MMG_FULL, sa, d, G);
				int32_t pc = __shfl_up_sync(MMG_FULL, sc2, d, G);
				if (gl >= d) { pc += sa; sc2 = pc > sc2 ? pc : sc2; sa += pa; }
			}
			int32_t after = n_skip + sa;
			if (after < sc2) after = sc2;
			const uint32_t over = (__ballot_sync(MMG_FULL, mk && after > max_skip) >> gbase) & gmask;
			const int brk = over ? __ffs((int)over) - 1 : -1;
			const int32_t after_last = __shfl_sync(MMG_FULL, after, G - 1, G);
			/* best score among the lanes before the break (all lanes without one), first such lane */
			const int32_t c2 = (brk < 0 || gl < brk) ? sc : INT32_MIN_;
			const long long key = grp_max64<G>(((long long)c2 << 8) | (long long)(G - 1 - gl));
			if (ga) {
				const int32_t mx = (int32_t)(key >> 8);
				if (brk < 0) n_skip = after_last;
				if (mx > max_f) max_f = mx, max_j = jb - (G - 1 - (int)(key & 0xff));
				if (brk >= 0) end_j = jb - brk, broke = true;
				jb -= G;
			}
		}
		if (act && gl == 0) tot_iter += (unsigned long long)(broke ? i - end_j : i - st);
		/* ---- lchain.c: the best-scoring anchor in range is tried even if the loop stopped before it ---- */
		{
			bool need = act && (max_ii < 0 || aix - mii_x > (uint64_t)(int64_t)max_dist_x);
			if (__any_sync(MMG_FULL, need)) {
				long long best = -1;                          /* (f + 2^31) << 32 | j : the largest f, then the largest j */
				int jb2 = i - 1;
				for (;;) {
					const bool gn = need && jb2 >= st;
					if (!__any_sync(MMG_FULL, gn)) break;
					const int j = jb2 - gl;
					if (gn && j >= st) { const long long k = ((long long)f[j] + 2147483648LL) << 32 | (long long)(uint32_t)j; if (k > best) best = k; }
					if (gn) jb2 -= G;
				}
				best = grp_max64<G>(best);
				if (need) {
					/* chain_dp_kernel: the lane-wise maxima keep the LARGEST j among equal f within a lane column and the
					 * reduction takes the largest j among the lanes that hold the maximum: the largest j with maximal f */
					max_ii = best < 0 ? -1 : (int)(uint32_t)best;
					if (max_ii >= 0) mii_x = ax[max_ii], mii_f = f[max_ii];
				}
			}
		}
		if (act) {
			if (max_ii >= 0 && max_ii < end_j) {
				const int32_t tmp = dev_comput_sc(aix, aiy, mii_x, ay[max_ii], max_dist_x, max_dist_y, bw, pen_gap, pen_skip);
				if (tmp != INT32_MIN_ && max_f < tmp + mii_f) max_f = tmp + mii_f, max_j = max_ii;
			}
			if (gl == 0) f[i] = max_f, p[i] = max_j, t[i] = 0;
			if (max_ii < 0 || (aix - mii_x <= (uint64_t)(int64_t)max_dist_x && mii_f < max_f))
				max_ii = i, mii_x = aix, mii_f = max_f;
			++i;
		}
		__syncwarp();
	}
#pragma unroll
	for (int d = 16; d; d >>= 1) tot_iter += __shfl_xor_sync(MMG_FULL, tot_iter, d), tot_anchor += __shfl_xor_sync(MMG_FULL, tot_anchor, d);
	if (lane == 0 && tot_anchor) {
		atomicAdd(&c.stats[4], tot_anchor);
		atomicAdd(&c.stats[5], tot_iter);
	}
}

__global__ void __launch_bounds__(CHAIN_WARPS * 32)
backtrack_kernel(ChunkDev c, DevOpt o, uint32_t r0, uint32_t r1, uint32_t *work)
{
	__shared__ int s_bkt[CHAIN_WARPS][512];
	const int lane = mmg_lane(), wib = threadIdx.x >> 5;
	for (;;) {
		uint32_t r = r0 + mmg_next_item(work);
		if (r >= r1) break;
		r = mmg_read_of(c, r);
		const int n = (int)c.n_a[r];
		const uint64_t ab = c.a_off[r] - c.a_off0;
		int n_u = 0, n_v = 0;
		if (n > 0)
			dev_backtrack_compact(n, c.bx + ab, c.by + ab, c.f + ab, c.p + ab, c.t + ab, c.v + ab,
			                      c.zx + 2 * ab, c.zy + 2 * ab, c.cx + ab, c.cy + ab, c.u + ab,
			                      o.min_cnt, o.min_chain_score, o.bw, s_bkt[wib], &n_u, &n_v);
		if (lane == 0) c.n_u[r] = (uint32_t)n_u, c.n_v[r] = (uint32_t)n_v;
		__syncwarp();
	}
}

int launch_chain(const ChunkDev &c, const DevOpt &o, uint32_t r0, uint32_t r1, int n_sms, cudaStream_t st, uint32_t *work)
{
	int grid = n_sms * 12, need = ((int)(r1 - r0) + CHAIN_WARPS - 1) / CHAIN_WARPS;
	if (grid > need) grid = need;
	if (grid < 1) grid = 1;
	static int groups = -1;   /* lanes per read: 32 = chain_dp_kernel (one read per warp); 8 / 16 = chain_dp_group_kernel */
	if (groups < 0) { const char *e = getenv("MMG_CHAIN_LANES"); groups = e ? atoi(e) : 32; }   /* B200, configs[2]: 32 lanes 32.8 ms, 16 lanes 40.2 ms, 8 lanes 42.9 ms per step */
	if (groups == 8) { need = ((int)(r1 - r0) + CHAIN_WARPS * 4 - 1) / (CHAIN_WARPS * 4); if (grid > need) grid = need < 1 ? 1 : need; MMG_LAUNCH(chain_dp_group_kernel<8>, grid, CHAIN_WARPS * 32, 0, st, c, o, r0, r1, work); }
	else if (groups == 16) { need = ((int)(r1 - r0) + CHAIN_WARPS * 2 - 1) / (CHAIN_WARPS * 2); if (grid > need) grid = need < 1 ? 1 : need; MMG_LAUNCH(chain_dp_group_kernel<16>, grid, CHAIN_WARPS * 32, 0, st, c, o, r0, r1, work); }
	else MMG_LAUNCH(chain_dp_kernel, grid, CHAIN_WARPS * 32, 0, st, c, o, r0, r1, work);
	return 0;
}

int launch_backtrack(const ChunkDev &c, const DevOpt &o, uint32_t r0, uint32_t r1, int n_sms, cudaStream_t st, uint32_t *work)
{
	int grid = n_sms * 12, need = ((int)(r1 - r0) + CHAIN_WARPS - 1) / CHAIN_WARPS;
	if (grid > need) grid = need;
	if (grid < 1) grid = 1;
	MMG_LAUNCH(backtrack_kernel, grid, CHAIN_WARPS * 32, 0, st, c, o, r0, r1, work);
	return 0;
}
