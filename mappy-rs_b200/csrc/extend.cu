/* extend.cu -- base-level alignment on the device (north-star (f)).
 *
 * Replaces, on the mm_map path (/root/reference/src/lib.rs:482,587; mappy-rs
 * always sets MM_F_CIGAR, src/lib.rs:339; minimap2 v2.26): align.c
 * mm_align_skeleton / mm_align1 (mm_fix_bad_ends, mm_filter_bad_seeds(_alt),
 * extension windows, gap filling, z-drop handling, mm_split_reg),
 * mm_align_pair -> ksw2_extd2_sse.c ksw_extd2_sse, ksw2.h ksw_backtrack /
 * ksw_apply_zdrop, mm_test_zdrop, mm_append_cigar, mm_fix_cigar,
 * mm_update_extra and hit.c mm_squeeze_a.
 *
 * Three kernels per round:
 *   ext_prep   (warp per read)  decides, for every region, the left extension,
 *              the gap-fill segments between kept anchors and the right
 *              extension.  In mm_align1 these depend only on anchor positions,
 *              never on an earlier DP result, so they become independent JOBS.
 *   ext_dp     (warp per job)   banded dual-affine DP on anti-diagonals: the
 *              cells of one anti-diagonal are independent, lanes stride over
 *              them; difference arrays u,v,x,y,x2,y2 and the sequences live in
 *              shared memory (or a global slice for very long jobs), one
 *              traceback byte per cell goes to HBM.  Upstream evaluates whole
 *              16-lane SSE blocks, so the kernel evaluates the same padded
 *              cells: band-edge cells then read the same neighbours upstream does.
 *   ext_stitch (warp per read)  walks the jobs of a region in order, applies
 *              upstream's z-drop / split rules, concatenates CIGARs, fixes them
 *              and recomputes coordinates, mlen, blen, dp_max.
 * A region split by a z-drop creates a new region that is aligned in the next
 * round.  Bound: INT32 issue (DP cells), traceback bytes to HBM - see DESIGN.md.
 */
#include <stdio.h>
#include <stdlib.h>
#include "dev_common.cuh"
#include "dev_sort.cuh"
#include "dev_regs.cuh"
#include "stages.h"
#include "extend.h"

#define KSW_NEG_INF (-0x40000000)
#define EZ_RIGHT 0x02
#define EZ_APPROX_MAX 0x08
#define EZ_EXTZ_ONLY 0x40
#define EZ_REV_CIGAR 0x80
#define SEED_IGNORE (1ULL << 41)
#define SEED_TANDEM (1ULL << 42)
#define SEED_LONG_JOIN (1ULL << 40)

/* ---------- sequence access ---------- */
__device__ __forceinline__ int ext_nt4(unsigned c)
{
	unsigned u = c & 0xdfu;
	return u == 'A' ? 0 : u == 'C' ? 1 : u == 'G' ? 2 : (u == 'T' || u == 'U') ? 3 : 4;
}
/* base p of qseq0[rev] (align.c mm_align_skeleton: forward codes / reverse complement) */
__device__ __forceinline__ int ext_qbase(const char *seq, int qlen, int rev, int p)
{
	if (!rev) return ext_nt4((unsigned char)seq[p]);
	int c = ext_nt4((unsigned char)seq[qlen - 1 - p]);
	return c < 4 ? 3 - c : 4;
}
__device__ __forceinline__ int ext_tbase(const DevIndex &di, uint64_t off, int pos)
{ /* index.c mm_idx_getseq: 4-bit packed */
	uint64_t i = off + (uint64_t)pos;
	return (int)(di.S[i >> 3] >> ((i & 7) << 2) & 0xf);
}

/* ---------- serial region helpers (lane 0) ---------- */
__device__ void ext_reg_set_coor(DevReg *r, int32_t qlen, const uint64_t *ax, const uint64_t *ay)
{ /* hit.c mm_reg_set_coor + mm_cal_fuzzy_len */
	const int k = r->as, cnt = r->cnt;
	int32_t q_span = (int32_t)(ay[k] >> 32 & 0xff);
	uint32_t rev = (uint32_t)(ax[k] >> 63);
	REG_SET(*r, 10, 1, rev);
	r->rid = (int32_t)(ax[k] << 1 >> 33);
	r->rs = (int32_t)ax[k] + 1 > q_span ? (int32_t)ax[k] + 1 - q_span : 0;
	r->re = (int32_t)ax[k + cnt - 1] + 1;
	if (!rev) {
		r->qs = (int32_t)ay[k] + 1 - q_span;
		r->qe = (int32_t)ay[k + cnt - 1] + 1;
	} else {
		r->qs = qlen - ((int32_t)ay[k + cnt - 1] + 1);
		r->qe = qlen - ((int32_t)ay[k] + 1 - q_span);
	}
	r->mlen = r->blen = 0;
	if (cnt <= 0) return;
	r->mlen = r->blen = q_span;
	for (int i = k + 1; i < k + cnt; ++i) {
		int span = (int)(ay[i] >> 32 & 0xff);
		int tl = (int32_t)ax[i] - (int32_t)ax[i - 1];
		int ql = (int32_t)ay[i] - (int32_t)ay[i - 1];
		r->blen += tl > ql ? tl : ql;
		r->mlen += tl > span && ql > span ? span : tl < ql ? tl : ql;
	}
}

__device__ void ext_fix_bad_ends(const DevReg *r, const uint64_t *ax, const uint64_t *ay, int bw, int min_match, int32_t *as, int32_t *cnt)
{ /* align.c mm_fix_bad_ends */
	int32_t i, l, m;
	*as = r->as, *cnt = r->cnt;
	if (r->cnt < 3) return;
	m = l = (int32_t)(ay[r->as] >> 32 & 0xff);
	for (i = r->as + 1; i < r->as + r->cnt - 1; ++i) {
		int32_t lq, lr, mn, mx;
		int32_t q_span = (int32_t)(ay[i] >> 32 & 0xff);
		if (ay[i] & SEED_LONG_JOIN) break;
		lr = (int32_t)ax[i] - (int32_t)ax[i - 1];
		lq = (int32_t)ay[i] - (int32_t)ay[i - 1];
		mn = lr < lq ? lr : lq;
		mx = lr > lq ? lr : lq;
		if (mx - mn > l >> 1) *as = i;
		l += mn;
		m += mn < q_span ? mn : q_span;
		if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r->mlen >> 1) break;
	}
	*cnt = r->as + r->cnt - *as;
	m = l = (int32_t)(ay[r->as + r->cnt - 1] >> 32 & 0xff);
	for (i = r->as + r->cnt - 2; i > *as; --i) {
		int32_t lq, lr, mn, mx;
		int32_t q_span = (int32_t)(ay[i + 1] >> 32 & 0xff);
		if (ay[i + 1] & SEED_LONG_JOIN) break;
		lr = (int32_t)ax[i + 1] - (int32_t)ax[i];
		lq = (int32_t)ay[i + 1] - (int32_t)ay[i];
		mn = lr < lq ? lr : lq;
		mx = lr > lq ? lr : lq;
		if (mx - mn > l >> 1) *cnt = i + 1 - *as;
		l += mn;
		m += mn < q_span ? mn : q_span;
		if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r->mlen >> 1) break;
	}
}

#define GAPOF(i) (((int32_t)ay[as1 + (i)] - (int32_t)ay[as1 + (i) - 1]) - ((int32_t)ax[as1 + (i)] - (int32_t)ax[as1 + (i) - 1]))

__device__ void ext_filter_bad_seeds(int as1, int cnt1, const uint64_t *ax, uint64_t *ay, int min_gap, int diff_thres, int max_ext_len, int max_ext_cnt, int *K)
{ /* align.c mm_filter_bad_seeds; K[cnt1] scratch */
	int max_st, max_en, n, i, k, mx;
	for (i = 1, n = 0; i < cnt1; ++i) {
		int gap = GAPOF(i);
		if (gap < -min_gap || gap > min_gap) K[n++] = i;
	}
	if (n == 0) return;
	mx = 0, max_st = max_en = -1;
	for (k = 0;; ++k) {
		int gap, l, n_ins = 0, n_del = 0, qs, rs, max_diff = 0, max_diff_l = -1;
		if (k == n || k >= max_en) {
			if (max_en > 0)
				for (i = K[max_st]; i < K[max_en]; ++i) ay[as1 + i] |= SEED_IGNORE;
			mx = 0, max_st = max_en = -1;
			if (k == n) break;
		}
		i = K[k];
		gap = GAPOF(i);
		if (gap > 0) n_ins += gap;
		else n_del += -gap;
		qs = (int32_t)ay[as1 + i - 1];
		rs = (int32_t)ax[as1 + i - 1];
		for (l = k + 1; l < n && l <= k + max_ext_cnt; ++l) {
			int j = K[l], diff;
			if ((int32_t)ay[as1 + j] - qs > max_ext_len || (int32_t)ax[as1 + j] - rs > max_ext_len) break;
			gap = GAPOF(j);
			if (gap > 0) n_ins += gap;
			else n_del += -gap;
			diff = n_ins + n_del - (n_ins > n_del ? n_ins - n_del : n_del - n_ins);
			if (max_diff < diff) max_diff = diff, max_diff_l = l;
		}
		if (max_diff > diff_thres && max_diff > mx) mx = max_diff, max_st = k, max_en = max_diff_l;
	}
}

__device__ void ext_filter_bad_seeds_alt(int as1, int cnt1, const uint64_t *ax, uint64_t *ay, int min_gap, int max_ext, int *K)
{ /* align.c mm_filter_bad_seeds_alt */
	int n, i, k;
	for (i = 1, n = 0; i < cnt1; ++i) {
		int gap = GAPOF(i);
		if (gap < -min_gap || gap > min_gap) K[n++] = i;
	}
	for (k = 0; k < n;) {
		int n_ins = 0, n_del = 0, l, gap;
		i = K[k];
		gap = GAPOF(i);
		if (gap > 0) n_ins += gap;
		else n_del += -gap;
		for (l = k + 1; l < n; ++l) {
			int j = K[l], diff;
			if ((int32_t)ay[as1 + j] - (int32_t)ay[as1 + i] > max_ext) break;
			gap = GAPOF(j);
			if (gap > 0) n_ins += gap;
			else n_del += -gap;
			diff = n_ins + n_del - (n_ins > n_del ? n_ins - n_del : n_del - n_ins);
			if (diff > min_gap) break;
		}
		if (l < n) {
			int j = K[l];
			for (i = K[k]; i < j; ++i) ay[as1 + i] |= SEED_IGNORE;
			k = l + 1;
		} else ++k;
	}
}

/* The extension windows of align.c mm_align1 (non-SR, non-splice branch). */
__device__ void ext_windows(const DevOpt &o, const DevIndex &di, const DevReg *r, ExtReg *x, int qlen, int n_a, const uint64_t *ax, const uint64_t *ay)
{
	const int as1 = x->as1, cnt1 = x->cnt1;
	const int32_t rid = (int32_t)(ax[r->as] << 1 >> 33), tlen = (int32_t)di.seq_len[rid];
	int32_t i, l, rs0, re0, qs0, qe0, rs, re, qs, qe, rs1, qs1, re1, qe1;
	rs = (int32_t)ax[as1] + 1, qs = (int32_t)ay[as1] + 1;                               /* mm_adjust_minier (no HPC) */
	re = (int32_t)ax[as1 + cnt1 - 1] + 1, qe = (int32_t)ay[as1 + cnt1 - 1] + 1;
	rs0 = (int32_t)ax[r->as] + 1 - (int32_t)(ay[r->as] >> 32 & 0xff);
	qs0 = (int32_t)ay[r->as] + 1 - (int32_t)(ay[r->as] >> 32 & 0xff);
	if (rs0 < 0) rs0 = 0;
	rs1 = qs1 = 0;
	for (i = r->as - 1, l = 0; i >= 0 && ax[i] >> 32 == ax[r->as] >> 32; --i) { /* inspect nearby seeds */
		int32_t xx = (int32_t)ax[i] + 1 - (int32_t)(ay[i] >> 32 & 0xff);
		int32_t yy = (int32_t)ay[i] + 1 - (int32_t)(ay[i] >> 32 & 0xff);
		if (xx < rs0 && yy < qs0) {
			if (++l > o.min_cnt) {
				l = rs0 - xx > qs0 - yy ? rs0 - xx : qs0 - yy;
				rs1 = rs0 - l, qs1 = qs0 - l;
				if (rs1 < 0) rs1 = 0;
				break;
			}
		}
	}
	if (qs > 0 && rs > 0) {
		l = qs < o.max_gap ? qs : o.max_gap;
		qs1 = qs1 > qs - l ? qs1 : qs - l;
		qs0 = qs0 < qs1 ? qs0 : qs1;
		l += l * o.a > o.q ? (l * o.a - o.q) / o.e : 0;
		l = l < o.max_gap ? l : o.max_gap;
		l = l < rs ? l : rs;
		rs1 = rs1 > rs - l ? rs1 : rs - l;
		rs0 = rs0 < rs1 ? rs0 : rs1;
		rs0 = rs0 < rs ? rs0 : rs;
	} else rs0 = rs, qs0 = qs;
	re0 = (int32_t)ax[r->as + r->cnt - 1] + 1;
	qe0 = (int32_t)ay[r->as + r->cnt - 1] + 1;
	re1 = tlen, qe1 = qlen;
	for (i = r->as + r->cnt, l = 0; i < n_a && ax[i] >> 32 == ax[r->as] >> 32; ++i) {
		int32_t xx = (int32_t)ax[i] + 1;
		int32_t yy = (int32_t)ay[i] + 1;
		if (xx > re0 && yy > qe0) {
			if (++l > o.min_cnt) {
				l = xx - re0 > yy - qe0 ? xx - re0 : yy - qe0;
				re1 = re0 + l, qe1 = qe0 + l;
				break;
			}
		}
	}
	if (qe < qlen && re < tlen) {
		l = qlen - qe < o.max_gap ? qlen - qe : o.max_gap;
		qe1 = qe1 < qe + l ? qe1 : qe + l;
		qe0 = qe0 > qe1 ? qe0 : qe1;
		l += l * o.a > o.q ? (l * o.a - o.q) / o.e : 0;
		l = l < o.max_gap ? l : o.max_gap;
		l = l < tlen - re ? l : tlen - re;
		re1 = re1 < re + l ? re1 : re + l;
		re0 = re0 > re1 ? re0 : re1;
	} else re0 = re, qe0 = qe;
	x->rs = rs, x->qs = qs, x->re = re, x->qe = qe;
	x->rs0 = rs0, x->qs0 = qs0, x->re0 = re0, x->qe0 = qe0;
}

/* Walk the gap-fill segments of a region exactly like the loop in mm_align1.
 * f(i, rs, qs, re, qe) is called for every segment that gets aligned; returning
 * false stops the walk (z-drop).  Returns the (re, qe) of the last anchor seen. */
template<typename F>
__device__ __forceinline__ void ext_walk_fills(const DevOpt &o, const ExtReg *x, const uint64_t *ax, const uint64_t *ay, int32_t *re_, int32_t *qe_, F f)
{
	const int as1 = x->as1, cnt1 = x->cnt1;
	int32_t rs = x->rs, qs = x->qs, re = x->re, qe = x->qe;
	for (int i = 1; i < cnt1; ++i) {
		if ((ay[as1 + i] & (SEED_IGNORE | SEED_TANDEM)) && i != cnt1 - 1) continue;
		re = (int32_t)ax[as1 + i] + 1, qe = (int32_t)ay[as1 + i] + 1;
		if (i == cnt1 - 1 || (ay[as1 + i] & SEED_LONG_JOIN) || (qe - qs >= o.min_ksw_len && re - rs >= o.min_ksw_len)) {
			if (!f(i, rs, qs, re, qe)) break;
			rs = re, qs = qe;
		}
	}
	*re_ = re, *qe_ = qe;
}

/* ksw2_ll_sse.c ksw_ll_i16 on the whole warp: gmax plus the (qe, te) upstream reports.  The striped SSE
 * code computes plain Smith-Waterman scores over the query padded to 8*slen columns (padding scores 0);
 * te is the last target row whose maximum reaches gmax, qe the column of that row with H == gmax that
 * comes last in the striped layout (index (p % slen) * 8 + p / slen).  scr: 7*(tl+1) int32 + tl uint64. */
template<typename QF, typename TF>
__device__ int ext_ll_i16(int ql, int tl, QF qf, TF tf, const DevOpt &o, int32_t *scr, int *qe_, int *te_)
{
	const int lane = mmg_lane();
	const int a_ = o.a < 0 ? -o.a : o.a, b_ = o.b > 0 ? -o.b : o.b, amb = o.sc_ambi > 0 ? -o.sc_ambi : o.sc_ambi;
	const int gapoe = o.q + o.e, gape = o.e, slen = (ql + 7) / 8, QL = slen * 8, n = tl + 1;
	int32_t *H0 = scr, *H1 = scr + n, *H2 = scr + 2 * n, *E0 = scr + 3 * n, *E1 = scr + 4 * n, *F0 = scr + 5 * n, *F1 = scr + 6 * n;
	unsigned long long *rb = (unsigned long long*)(scr + 7 * n + (n & 1));
	for (int i = lane; i < tl; i += 32) rb[i] = 0;
	__syncwarp();
	for (int d = 0; d < QL + tl - 1; ++d) {
		int ilo = d - (QL - 1) > 0 ? d - (QL - 1) : 0, ihi = d < tl - 1 ? d : tl - 1;
		for (int i = ilo + lane; i <= ihi; i += 32) {
			const int p = d - i;
			int s = 0;
			if (p < ql) {
				const int cq = qf(p), ct = tf(i);
				s = (ct == 4 || cq == 4) ? amb : ct == cq ? a_ : b_;
			}
			const int diag = (i > 0 && p > 0) ? H2[i - 1] : 0;
			const int up = i > 0 ? H1[i - 1] : 0, eup = i > 0 ? E1[i - 1] : 0;
			const int left = p > 0 ? H1[i] : 0, fl = p > 0 ? F1[i] : 0;
			int e = eup - gape > up - gapoe ? eup - gape : up - gapoe;
			int f = fl - gape > left - gapoe ? fl - gape : left - gapoe;
			if (e < 0 || i == 0) e = 0;
			if (f < 0 || p == 0) f = 0;
			int h = diag + s;
			h = h > e ? h : e;
			h = h > f ? h : f;
			h = h > 0 ? h : 0;
			H0[i] = h, E0[i] = e, F0[i] = f;
			unsigned long long key = (unsigned long long)(uint32_t)h << 32 | (uint32_t)((p % slen) * 8 + p / slen);
			if (key > rb[i]) rb[i] = key;
		}
		__syncwarp();
		int32_t *t = H2; H2 = H1, H1 = H0, H0 = t;
		t = E1, E1 = E0, E0 = t;
		t = F1, F1 = F0, F0 = t;
	}
	unsigned long long best = 0; /* (gmax, row) with the LAST row winning ties */
	for (int i = lane; i < tl; i += 32) {
		unsigned long long k = (rb[i] >> 32) << 32 | (uint32_t)i;
		if (k > best) best = k;
	}
#pragma unroll
	for (int d = 16; d; d >>= 1) {
		unsigned long long ob = __shfl_xor_sync(MMG_FULL, best, d);
		if (ob > best) best = ob;
	}
	const int gmax = (int)(best >> 32), te = (int)(uint32_t)best;
	const uint32_t idx = (uint32_t)rb[te];
	*te_ = te, *qe_ = (int)(idx / 8 + idx % 8 * slen);
	return gmax;
}

/* ---------- prep: squeeze anchors, windows, job lists ---------- */
__global__ void __launch_bounds__(CHAIN_WARPS * 32)
ext_prep_kernel(ChunkDev c, DevIndex di, DevOpt o, ExtBufs xb, uint32_t r0, uint32_t r1, int round, uint32_t *work)
{
	const int lane = mmg_lane();
	for (;;) {
		uint32_t r = r0 + mmg_next_item(work);
		if (r >= r1) break;
		r = mmg_read_of(c, r);
		const int n_regs = (int)c.n_regs[r];
		if (n_regs == 0) continue;
		const uint64_t ab = c.a_off[r] - c.a_off0, rb = xb.xr_off[r];
		const int qlen = (int)(c.off[r + 1] - c.off[r]);
		DevReg *regs = c.regs + rb;
		ExtReg *xr = xb.xregs + rb;
		uint64_t *ax = c.cx + ab, *ay = c.cy + ab;       /* squeezed anchors live in cx/cy */
		if (round == 0) {
			/* hit.c mm_squeeze_a: anchors of surviving regions, in order of r->as */
			const uint64_t *sx = c.bx + ab, *sy = c.by + ab;
			if (lane == 0) {
				uint64_t *aux = c.zx + 2 * ab;
				for (int i = 0; i < n_regs; ++i) aux[i] = (uint64_t)(uint32_t)regs[i].as << 32 | (uint32_t)i;
				for (int i = 1; i < n_regs; ++i) { /* values are distinct: any sort gives upstream's array */
					uint64_t tv = aux[i];
					int j = i;
					for (; j > 0 && tv < aux[j - 1]; --j) aux[j] = aux[j - 1];
					aux[j] = tv;
				}
			}
			__syncwarp();
			int as = 0;
			for (int i = 0; i < n_regs; ++i) {
				DevReg *g = &regs[(uint32_t)(c.zx + 2 * ab)[i]];
				const int src = g->as, cnt = g->cnt;
				for (int j = lane; j < cnt; j += 32) ax[as + j] = sx[src + j], ay[as + j] = sy[src + j];
				__syncwarp();
				if (lane == 0) g->as = as;
				as += cnt;
			}
			if (lane == 0) xb.n_sq[r] = (uint32_t)as;
			__syncwarp();
		}
		const int n_a = (int)xb.n_sq[r];
		/* align.c mm_align1_inv, first half: a region split off by an inversion-like z-drop and its left
		 * neighbour may flank an inversion; score the gap between them on the opposite strand */
		for (int i = 1; round > 0 && i < n_regs; ++i) {
			if (xr[i].state != EXT_INV_PENDING) continue;
			const DevReg r1 = regs[i - 1], r2 = regs[i];
			const int r1rev = (int)REG_REV(r1);
			const int ql = r1rev ? r1.qs - r2.qe : r2.qs - r1.qe, tl = r2.rs - r1.re;
			bool ok = (r1.bits >> 8 & 1u) && (r2.bits >> 9 & 1u);
			if (r1.id != r1.parent && r1.parent != PARENT_TMP_PRI) ok = false;
			if (r2.id != r2.parent && r2.parent != PARENT_TMP_PRI) ok = false;
			if (r1.rid != r2.rid || r1rev != (int)REG_REV(r2)) ok = false;
			if (ql < o.min_chain_score || ql > o.max_gap || tl < o.min_chain_score || tl > o.max_gap) ok = false;
			if (ok && (uint64_t)(7 * (tl + 1) + 2) * 4 + (uint64_t)tl * 8 > xb.big_per_warp) { ok = false; if (lane == 0) atomicOr(&c.flags[r], 0x20000000u), atomicOr(c.err, 0x20000000u); }
			int qe = -1, te = -1;
			const int rev_inv = r1rev ? 0 : 1, p0 = r1rev ? r2.qe : qlen - r2.qs;
			if (ok) {
				const char *seq = c.seq + c.off[r];
				const uint64_t toff = di.seq_off[r1.rid];
				int32_t *scr = (int32_t*)(xb.big + (size_t)(blockIdx.x * CHAIN_WARPS + (threadIdx.x >> 5)) * xb.big_per_warp);
				const int rs2 = r2.rs;
				int score = ext_ll_i16(ql, tl,
					[&](int p) { return ext_qbase(seq, qlen, rev_inv, p0 + ql - 1 - p); },   /* mm_seq_rev(ql, qseq) */
					[&](int t) { return ext_tbase(di, toff, rs2 - 1 - t); },                  /* mm_seq_rev(tl, tseq) */
					o, scr, &qe, &te);
				if (score < o.min_dp_max) ok = false;
			}
			__syncwarp();   /* every lane has read xr[i].state before lane 0 changes it */
			if (lane == 0) {
				ExtReg *x = &xr[i];
				x->n_jobs = 0;
				if (ok) {
					const int q_off = ql - (qe + 1), t_off = tl - (te + 1);
					const uint32_t j0 = atomicAdd(xb.n_jobs, 1u);
					if ((uint64_t)j0 + 1 > xb.cap_jobs) atomicOr(&c.flags[r], 0x40000000u), atomicOr(c.err, 0x40000000u), x->state = EXT_DONE;
					else if (p0 + q_off < 0 || r1.re + t_off < 0) x->state = EXT_DONE; /* qe/te landed in the SSE padding: upstream reads before its buffers here */
					else {
						ExtJob *jb = &xb.jobs[j0];
						jb->read = r, jb->reg = (uint32_t)i, jb->kind = EXT_INV, jb->rev = (uint8_t)rev_inv, jb->rid = r1.rid;
						jb->qs = p0 + q_off, jb->qe = p0 + ql, jb->rs = r1.re + t_off, jb->re = r2.rs;
						jb->w = (int)(o.bw * 1.5), jb->zdrop = o.zdrop, jb->end_bonus = -1, jb->flag = EZ_EXTZ_ONLY;
						jb->n_cigar = 0, jb->zdropped = 0, jb->reach_end = 0, jb->zdrop_code = 0, jb->pad[0] = 0;
						const int jq = ql - q_off, jt = tl - t_off;
						int n_col = jq < jt ? jq : jt;
						n_col = ((n_col < jb->w + 1 ? n_col : jb->w + 1) + 15) / 16 + 1;
						jb->tb_size = (uint64_t)(jq + jt - 1) * (uint64_t)n_col * 16 + 16;
						jb->cg_size = (uint32_t)(jq + jt + 2);
						x->job0 = j0, x->n_jobs = 1, x->qs0 = q_off, x->rs0 = t_off;
					}
				} else x->state = EXT_DONE;
			}
			__syncwarp();
		}
		if (lane == 0) {
			int *K = (int*)(c.t + ab);
			for (int i = 0; i < n_regs; ++i) {
				DevReg *g = &regs[i];
				ExtReg *x = &xr[i];
				if (round == 0) x->state = EXT_PENDING, x->job0 = 0, x->n_jobs = 0;
				if (x->state != EXT_PENDING) continue;
				x->n_jobs = 0;
				if (g->cnt == 0) { x->state = EXT_DONE; continue; }
				if (!(o.flag & 0x10000000LL)) ext_fix_bad_ends(g, ax, ay, o.bw, o.min_chain_score * 2, &x->as1, &x->cnt1);
				else x->as1 = g->as, x->cnt1 = g->cnt;
				ext_filter_bad_seeds(x->as1, x->cnt1, ax, ay, 10, 40, o.max_gap >> 1, 10, K);
				ext_filter_bad_seeds_alt(x->as1, x->cnt1, ax, ay, 30, o.max_gap >> 1, K);
				ext_windows(o, di, g, x, qlen, n_a, ax, ay);
				/* count jobs, reserve a contiguous range, then fill it */
				int nj = 0;
				int32_t re_l, qe_l;
				if (x->qs > 0 && x->rs > 0) ++nj;
				ext_walk_fills(o, x, ax, ay, &re_l, &qe_l, [&](int, int32_t, int32_t, int32_t, int32_t) { ++nj; return true; });
				if (qe_l < x->qe0 && re_l < x->re0) ++nj;
				const uint32_t j0 = atomicAdd(xb.n_jobs, (uint32_t)nj);
				x->job0 = j0, x->n_jobs = nj;
				if ((uint64_t)j0 + nj > xb.cap_jobs) { atomicOr(&c.flags[r], 0x40000000u), atomicOr(c.err, 0x40000000u); x->n_jobs = 0; continue; }
				const int rid = g->rid, rev = (int)REG_REV(*g), bw = (int)(o.bw * 1.5 + 1.);
				int bw_long = (int)(o.bw_long * 1.5 + 1.);
				if (bw_long < bw) bw_long = bw;
				uint32_t jn = j0;
				auto put = [&](int kind, int32_t qs, int32_t qe, int32_t rs, int32_t re, int w, int zdrop, int end_bonus, int flag) {
					ExtJob *jb = &xb.jobs[jn++];
					jb->read = r, jb->reg = (uint32_t)i, jb->kind = (uint8_t)kind, jb->rev = (uint8_t)rev, jb->rid = rid;
					jb->qs = qs, jb->qe = qe, jb->rs = rs, jb->re = re, jb->w = w, jb->zdrop = zdrop, jb->end_bonus = end_bonus, jb->flag = flag;
					jb->n_cigar = 0, jb->zdropped = 0, jb->reach_end = 0, jb->zdrop_code = 0, jb->pad[0] = 0;
					const int ql = qe - qs, tl = re - rs;
					int w2 = w < 0 ? (tl > ql ? tl : ql) : w;
					int n_col = ql < tl ? ql : tl;
					n_col = ((n_col < w2 + 1 ? n_col : w2 + 1) + 15) / 16 + 1;
					jb->tb_size = (uint64_t)(ql + tl - 1) * (uint64_t)n_col * 16 + 16;
					jb->cg_size = (uint32_t)(ql + tl + 2);
				};
				if (x->qs > 0 && x->rs > 0)
					put(EXT_LEFT, x->qs0, x->qs, x->rs0, x->rs, bw, (g->bits >> 14 & 1u) ? o.zdrop_inv : o.zdrop, o.end_bonus, EZ_EXTZ_ONLY | EZ_RIGHT | EZ_REV_CIGAR);
				ext_walk_fills(o, x, ax, ay, &re_l, &qe_l, [&](int, int32_t rs, int32_t qs, int32_t re, int32_t qe) {
					put(EXT_FILL, qs, qe, rs, re, bw_long, o.zdrop, -1, EZ_APPROX_MAX);
					return true;
				});
				if (qe_l < x->qe0 && re_l < x->re0)
					put(EXT_RIGHT, qe_l, x->qe0, re_l, x->re0, bw, o.zdrop, o.end_bonus, EZ_EXTZ_ONLY);
			}
		}
		__syncwarp();
	}
}

/* number of jobs queued so far, clamped to the arena (prep flags the overflow); read by scan, DP and stitch */
__device__ __forceinline__ uint32_t ext_n_jobs(const ExtBufs &xb)
{
	const uint32_t n = *xb.n_jobs;
	return (uint64_t)n > xb.cap_jobs ? (uint32_t)xb.cap_jobs : n;
}
/* true when the CIGAR slices of the queued jobs fit the arena (else DP and stitch do nothing and the host reports it) */
__device__ __forceinline__ bool ext_cigar_fits(const ExtBufs &xb) { return xb.cg_base[1] <= xb.cap_cg; }

/* Exclusive scan of the per-job cigar sizes of jobs [j0, *n_jobs).  It must be the ordered scan: ext_stitch builds a
 * region's CIGAR in place over the contiguous slices of the region's consecutive jobs.  Chained over tiles of 8192 jobs:
 * block b takes tiles b, b + grid, ...; a tile adds its sum to the inclusive prefix its predecessor publishes in
 * `chain` (value << 1 | ready; zeroed by the host).  Every block is resident (one per SM) and takes its tiles in
 * increasing order, so the tile a block waits for is always owned by a block that is not waiting on a later one. */
#define JOB_SCAN_PER 8
#define JOB_SCAN_TILE (1024 * JOB_SCAN_PER)
__global__ void __launch_bounds__(1024)
ext_job_scan_kernel(ExtBufs xb, uint32_t j0, unsigned long long *chain)
{
	__shared__ unsigned long long s_b[32];
	__shared__ unsigned long long s_cb;
	const int lane = mmg_lane(), wib = threadIdx.x >> 5;
	const uint32_t j1 = ext_n_jobs(xb);
	if (j1 <= j0) { if (blockIdx.x == 0 && threadIdx.x == 0) xb.cg_base[1] = xb.cg_base[0]; return; }
	const uint32_t n_tiles = (j1 - j0 + JOB_SCAN_TILE - 1) / JOB_SCAN_TILE;
	for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
		const uint64_t i = (uint64_t)j0 + (uint64_t)tile * JOB_SCAN_TILE + (uint64_t)threadIdx.x * JOB_SCAN_PER;
		uint32_t sz[JOB_SCAN_PER];
		unsigned long long vb = 0;
#pragma unroll
		for (int k = 0; k < JOB_SCAN_PER; ++k) sz[k] = i + k < j1 ? xb.jobs[i + k].cg_size : 0u, vb += sz[k];
		unsigned long long xbv = vb;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			unsigned long long yb = __shfl_up_sync(MMG_FULL, xbv, d);
			if (lane >= d) xbv += yb;
		}
		if (lane == 31) s_b[wib] = xbv;
		__syncthreads();
		if (wib == 0) {
			unsigned long long wb = s_b[lane], sb = wb;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				unsigned long long yb = __shfl_up_sync(MMG_FULL, sb, d);
				if (lane >= d) sb += yb;
			}
			s_b[lane] = sb - wb;
			if (lane == 31) { /* sb = the tile's sum */
				unsigned long long before;
				if (tile == 0) before = xb.cg_base[0];
				else {
					volatile unsigned long long *prev = chain + (tile - 1);
					unsigned long long v;
					while (!((v = *prev) & 1ull)) {
#ifndef MMG_EMU
						__nanosleep(64);
#endif
					}
					before = v >> 1;
				}
				s_cb = before;
				__threadfence();
				*(volatile unsigned long long*)(chain + tile) = (before + sb) << 1 | 1ull;
				if (tile == n_tiles - 1) xb.cg_base[1] = before + sb;
			}
		}
		__syncthreads();
		unsigned long long at = s_cb + s_b[wib] + xbv - vb;
#pragma unroll
		for (int k = 0; k < JOB_SCAN_PER; ++k)
			if (i + k < j1) xb.jobs[i + k].tb_off = 0, xb.jobs[i + k].cg_off = at, at += sz[k];
		__syncthreads();
	}
}

/* ---------- the DP ---------- */
struct DpMem {
	int32_t *H;
	uint32_t *ga, *gb, *gc; /* the six difference arrays, 16-bit fields biased by 128, one 16-byte record per four columns:
	                         * ga = (u01, u23, y01, y23), gb = (x01, v01, x23, v23), gc = (x2_01, x2_23, y2_01, y2_23) */
	uint8_t *s;             /* match scores, bytes biased by 128 */
	uint8_t *sf;            /* target, then (contiguous) the reversed query, as in upstream's single kcalloc block */
	int T16, flat_sz;       /* flat_sz = bytes addressable from sf (T16 + Q16 + 16) */
};
#define EXT_DP_BYTES(T16, Q16) ((size_t)18 * (T16) + (Q16) + 16)   /* H 4, ga/gb/gc 4 each, s 1, target 1 per column */

/* scalar views of the packed arrays (boundary cells, the H bookkeeping): index of column t's 16-bit field */
enum { DP_U, DP_Y, DP_X, DP_V, DP_X2, DP_Y2 };
template<int A> __device__ __forceinline__ uint16_t *dp_field(const DpMem &m, int t)
{
	uint32_t *g = A == DP_U || A == DP_Y ? m.ga : A == DP_X || A == DP_V ? m.gb : m.gc;
	const int lo = A == DP_X || A == DP_V ? (t & 2) * 2 + (t & 1) : (t & 3);     /* gb interleaves x and v per pair */
	const int off = A == DP_Y || A == DP_Y2 ? 4 : A == DP_V ? 2 : 0;
	return (uint16_t*)g + 2 * (t & ~3) + lo + off;
}
template<int A> __device__ __forceinline__ int dp_get(const DpMem &m, int t) { return (int)*dp_field<A>(m, t) - 128; }
template<int A> __device__ __forceinline__ void dp_set(const DpMem &m, int t, int v) { *dp_field<A>(m, t) = (uint16_t)(v + 128); }

__device__ __forceinline__ bool ez_apply_zdrop(int32_t *ez_max, int *ez_max_t, int *ez_max_q, int32_t H, int r, int t, int zdrop, int e)
{ /* ksw2.h ksw_apply_zdrop with is_rot = 1 */
	if (H > *ez_max) {
		*ez_max = H, *ez_max_t = t, *ez_max_q = r - t;
	} else if (t >= *ez_max_t && r - t >= *ez_max_q) {
		int tl = t - *ez_max_t, ql = (r - t) - *ez_max_q, l;
		l = tl > ql ? tl - ql : ql - tl;
		if (zdrop >= 0 && *ez_max - H > zdrop + l * e) return true;
	}
	return false;
}

__device__ __forceinline__ void ext_dpmem_set(DpMem &m, unsigned char *base, int T16, int Q16)
{
	m.T16 = T16, m.flat_sz = T16 + Q16 + 16;
	m.H = (int32_t*)base;
	m.ga = (uint32_t*)(base + (size_t)4 * T16), m.gb = m.ga + T16, m.gc = m.gb + T16;
	m.s = (uint8_t*)(m.gc + T16);
	m.sf = m.s + T16;
}

/* ---- two cells per 32-bit word (16-bit fields) ----
 * The recurrence of one cell is ~45 scalar integer operations.  sm_100a has single-instruction 16x2 max / add-max
 * (VIMNMX3.U16x2, VIADDMNMX.S16x2), so the core loop keeps two cells in one register: every quantity is held in a
 * 16-bit field with a bias that keeps it non-negative, plain 32-bit adds then act on both fields at once (no carry
 * can cross a field).  The arg-max direction rides in the low three bits of the compared values (value*8 + priority:
 * "first maximum wins" for left-aligned gaps, "last maximum wins" for right-aligned ones, as the > / >= comparisons
 * of ksw2_extd2_sse.c resolve ties), and the four "gap continues" bits are read off one carry bit per term. */
#ifdef MMG_EMU
static inline uint32_t dp_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
	const uint64_t v = (uint64_t)b << 32 | a;
	uint32_t r = 0;
	for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xff) << (8 * i);
	return r;
}
static inline uint32_t dp_f2(uint32_t lo, uint32_t hi) { return (lo & 0xffffu) | hi << 16; }
static inline uint32_t dp_max3u(uint32_t a, uint32_t b, uint32_t c)
{
	auto mx = [](uint32_t x, uint32_t y) { return x > y ? x : y; };
	return dp_f2(mx(mx(a & 0xffff, b & 0xffff), c & 0xffff), mx(mx(a >> 16, b >> 16), c >> 16));
}
static inline uint32_t dp_minu(uint32_t a, uint32_t b) { return dp_f2((a & 0xffff) < (b & 0xffff) ? a : b, (a >> 16) < (b >> 16) ? a >> 16 : b >> 16); }
static inline uint32_t dp_addmaxs(uint32_t a, uint32_t b, uint32_t c)
{
	auto f = [](uint32_t x, uint32_t y, uint32_t z) { int16_t s = (int16_t)(uint16_t)(x + y), m = (int16_t)(uint16_t)z; return (uint32_t)(uint16_t)(s > m ? s : m); };
	return dp_f2(f(a & 0xffff, b & 0xffff, c & 0xffff), f(a >> 16, b >> 16, c >> 16));
}
#else
__device__ __forceinline__ uint32_t dp_prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
__device__ __forceinline__ uint32_t dp_max3u(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }
__device__ __forceinline__ uint32_t dp_minu(uint32_t a, uint32_t b) { return __vminu2(a, b); }
__device__ __forceinline__ uint32_t dp_addmaxs(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }
#endif
struct alignas(16) DpQuad { uint32_t x, y, z, w; };
struct alignas(8) DpPair { uint32_t x, y; };
#define DP_W(c) ((uint32_t)(c) * 0x00010001u)          /* the same 16-bit value in both fields */
#define DP_LD32(p) (*(const uint32_t*)(p))
#define DP_ST32(p, v) (*(uint32_t*)(p) = (v))

/* constants of one pass (all uniform) */
struct DpK {
	uint32_t p_s, p_a, p_b, p_a2, p_b2;     /* value*8 + priority addends (the score term also carries its extra bias) */
	uint32_t d_flip;                         /* direction = d_flip - priority (left) or priority (right) */
	uint32_t zcap;                           /* sc_mch + 256 */
	uint32_t fa, fb, fa2, fb2;               /* addends that put "term > 0" (left) / "term >= 0" (right) in bits 12..15 */
	uint32_t ra, rb, ra2, rb2;               /* re-bias of the clamped terms (signed 16-bit fields) */
	uint32_t floor1, floor2;                 /* 128 - qe, 128 - qe2 */
	bool right;
};

/* One 16x2 word: two cells.  Inputs are biased by 128 (S, U, Y, Y2 of the cell; Xs, Vs, X2s of its left neighbour).
 * Outputs: the six new array values (fields biased by 128) and the traceback byte. */
__device__ __forceinline__ void dp_cell2(const DpK &k, uint32_t S, uint32_t U, uint32_t Y, uint32_t Y2, uint32_t Xs, uint32_t Vs, uint32_t X2s,
                                         uint32_t &nu, uint32_t &nv, uint32_t &nx, uint32_t &ny, uint32_t &nx2, uint32_t &ny2, uint32_t &nd)
{
	const uint32_t A = Xs + Vs, B = Y + U, A2 = X2s + Vs, B2 = Y2 + U;                 /* term + 256 */
	const uint32_t M = dp_max3u(dp_max3u(S * 8 + k.p_s, A * 8 + k.p_a, B * 8 + k.p_b), A2 * 8 + k.p_a2, B2 * 8 + k.p_b2);
	const uint32_t pr = M & 0x00070007u;
	const uint32_t z = dp_minu((M >> 3) & 0x1fff1fffu, k.zcap);                        /* z + 256 */
	nd = k.right ? pr : k.d_flip - pr;
	nu = z - Vs, nv = z - U;                                                           /* (z - v[t-1]) + 128, (z - u[t]) + 128 */
	const uint32_t Fa = A - z + k.fa, Fb = B - z + k.fb, Fa2 = A2 - z + k.fa2, Fb2 = B2 - z + k.fb2;
	nx = dp_addmaxs(Fa, k.ra, k.floor1), ny = dp_addmaxs(Fb, k.rb, k.floor1);
	nx2 = dp_addmaxs(Fa2, k.ra2, k.floor2), ny2 = dp_addmaxs(Fb2, k.rb2, k.floor2);
	const uint32_t g1 = (Fb2 & 0x80008000u) | (Fa2 & 0x7fff7fffu), g2 = (Fb & 0x20002000u) | (Fa & 0xdfffdfffu);
	const uint32_t g = (g1 & 0xc000c000u) | (g2 & 0x3fff3fffu);
	nd |= (g >> 9) & 0x00780078u;
}

/* ---- backtrack (ksw_backtrack, is_rot = 1) from cell (i, j); returns the number of CIGAR operations ----
 * The walk is serial, and one traceback byte per step straight from global memory costs a full L2 round trip
 * per CIGAR base.  The warp therefore fetches a window of 32 anti-diagonals x 8 target columns below the
 * current cell (lane = diagonal), and every lane replays the same walk out of registers (one shuffle per
 * step) until it leaves the window: at least 8 and typically 16 steps per memory round trip.  A byte with
 * bit 7 set stands for upstream's force_state (cell outside the band of its diagonal). */
__device__ __noinline__ int ext_backtrack_warp(const uint8_t *tb, int n_col, int qlen, int tlen, int w, int i, int j, bool rev_cigar, uint32_t *cigar)
{
	const int lane = mmg_lane();
	int n_cigar = 0;
	int state = 0;
	uint32_t cur_op = 0xf, cur_len = 0;      /* the open CIGAR run (all lanes agree; lane 0 writes) */
	auto push = [&](uint32_t op, int len) {
		if (op != cur_op) {
			if (cur_len) { if (lane == 0) cigar[n_cigar] = cur_len << 4 | cur_op; ++n_cigar; }
			cur_op = op, cur_len = 0;
		}
		cur_len += (uint32_t)len;
	};
	while (i >= 0 && j >= 0) {
		const int r0 = i + j, i0 = i;
		unsigned long long win = 0;
		{
			const int r = r0 - lane;
			if (r >= 0) {
				int st = 0, en = tlen - 1;
				if (st < r - qlen + 1) st = r - qlen + 1;
				if (en > r) en = r;
				if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
				if (en > (r + w) >> 1) en = (r + w) >> 1;
				st = st / 16 * 16, en = (en + 16) / 16 * 16 - 1;  /* off[r], off_end[r] */
				const uint8_t *row = tb + (size_t)r * n_col - st;
#pragma unroll
				for (int k = 0; k < 8; ++k) {
					const int ii = i0 - k;
					unsigned long long v = 0;
					if (ii >= 0) v = ii < st ? 0x82u : ii > en ? 0x81u : row[ii];
					win |= v << (8 * k);
				}
			}
		}
		while (i >= 0 && j >= 0) {
			const int dr = r0 - (i + j), di = i0 - i;
			if (dr >= 32 || di >= 8) break;
			const uint32_t tmp = (uint32_t)(__shfl_sync(MMG_FULL, win, dr) >> (8 * di)) & 0xff;
			if (tmp & 0x80) state = (int)(tmp & 3);
			else {
				if (state == 0) state = tmp & 7;
				else if (!(tmp >> (state + 2) & 1)) state = 0;
				if (state == 0) state = tmp & 7;
			}
			if (state == 0) push(0, 1), --i, --j;
			else if (state == 1 || state == 3) push(2, 1), --i;
			else push(1, 1), --j;
		}
	}
	if (i >= 0) push(2, i + 1);
	if (j >= 0) push(1, j + 1);
	push(0xf, 0);                              /* flush the open run */
	__syncwarp();
	if (!rev_cigar)
		for (int k = lane; k < n_cigar >> 1; k += 32) { uint32_t t = cigar[k]; cigar[k] = cigar[n_cigar - 1 - k], cigar[n_cigar - 1 - k] = t; }
	return n_cigar;
}

/* One pass of ksw_extd2_sse over (qlen x tlen).  All lanes of the warp call it.  SMEM = true: the job's arrays are
 * the warp's shared-memory slice; the pointers are derived from the shared array inside this function so that the
 * compiler emits shared-memory loads/stores with 32-bit addresses instead of generic ones. */
template<bool SMEM>
__device__ __noinline__ void ext_dp_pass_t(unsigned char *gbase, int T16, int Q16, const DevOpt &o, int qlen, int tlen, int w, int zdrop, int end_bonus, int flag,
                                           uint8_t *tb, uint32_t *cigar, ExtJob *res, unsigned long long *n_cell)
{
	DpMem m;
	MMG_DYN_SMEM(dp_smem);
	if (SMEM) ext_dpmem_set(m, dp_smem + (size_t)(threadIdx.x >> 5) * EXT_SMEM_PER_WARP, T16, Q16);
	else ext_dpmem_set(m, gbase, T16, Q16);
	const int lane = mmg_lane();
	int q = o.q, e = o.e, q2 = o.q2, e2 = o.e2;
	if (q2 + e2 < q + e) { int t = q; q = q2, q2 = t, t = e, e = e2, e2 = t; }
	const int qe = q + e, qe2 = q2 + e2;
	const int a_ = o.a < 0 ? -o.a : o.a, b_ = o.b > 0 ? -o.b : o.b, amb = o.sc_ambi > 0 ? -o.sc_ambi : o.sc_ambi;
	const int sc_mch = a_, sc_mis = b_, sc_N = amb == 0 ? -e2 : amb;
	const bool approx_max = (flag & EZ_APPROX_MAX) != 0, right = (flag & EZ_RIGHT) != 0;
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	int n_col_ = qlen < tlen ? qlen : tlen;
	n_col_ = ((n_col_ < w + 1 ? n_col_ : w + 1) + 15) / 16 + 1;
	const int n_col = n_col_ * 16;
	int long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
	if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
	const int long_diff = long_thres * (e - e2) - (q2 - q) - e2;
	/* constants of the packed core (dp_cell2): every term is carried with bias 256, its test constant puts
	 * "term - z + gap_open > 0" (left-aligned; ">= 0" right-aligned) into bit 12, 13, 14 or 15 of its field */
	/* (kept in the warp's slot of shared memory behind the DP slices: the pass has no registers to spare) */
	DpK &dk = *(DpK*)(dp_smem + (size_t)EXT_DP_WARPS * EXT_SMEM_PER_WARP + (size_t)(threadIdx.x >> 5) * EXT_DPK_BYTES);
	__syncwarp();
	if (lane == 0) {
		const int ge = right ? 0 : 1;
		dk.right = right;
		dk.p_s = DP_W(128 * 8 + (right ? 0 : 4)), dk.p_a = DP_W(right ? 1 : 3), dk.p_b = DP_W(2), dk.p_a2 = DP_W(right ? 3 : 1), dk.p_b2 = DP_W(right ? 4 : 0);
		dk.d_flip = DP_W(4), dk.zcap = DP_W(sc_mch + 256);
		dk.fa = DP_W(q + 4096 - ge), dk.fb = DP_W(q + 8192 - ge), dk.fa2 = DP_W(q2 + 16384 - ge), dk.fb2 = DP_W(q2 + 32768 - ge);
		dk.ra = DP_W((uint16_t)(128 - qe - (4096 - ge))), dk.rb = DP_W((uint16_t)(128 - qe - (8192 - ge)));
		dk.ra2 = DP_W((uint16_t)(128 - qe2 - (16384 - ge))), dk.rb2 = DP_W((uint16_t)(128 - qe2 - (32768 - ge)));
		dk.floor1 = DP_W((uint16_t)(128 - qe)), dk.floor2 = DP_W((uint16_t)(128 - qe2));
	}
	__syncwarp();
	const uint32_t sc_b4 = (uint32_t)(sc_mch + 128) * 0x01010101u, scN_b4 = (uint32_t)(sc_N + 128) * 0x01010101u;
	/* ksw_reset_extz */
	int32_t ez_max = 0, ez_score = KSW_NEG_INF, ez_mqe = KSW_NEG_INF, ez_mte = KSW_NEG_INF;
	int ez_max_q = -1, ez_max_t = -1, ez_mqe_t = -1, ez_mte_q = -1, zdropped = 0, reach_end = 0;
	/* NB: mm_align_pair never calls the kernel with scores that make -min_sc > 2*(q+e) on this path */
	{
		const uint32_t g1 = DP_W(128 - q - e), g2 = DP_W(128 - q2 - e2);
		for (int g = lane; g < T16 >> 2; g += 32) {
			uint32_t *a = m.ga + 4 * g, *b = m.gb + 4 * g, *c = m.gc + 4 * g;
			a[0] = a[1] = a[2] = a[3] = g1, b[0] = b[1] = b[2] = b[3] = g1;
			c[0] = c[1] = c[2] = c[3] = g2;
			DP_ST32(m.s + 4 * g, 0x80808080u);
			if (!approx_max) m.H[4 * g] = m.H[4 * g + 1] = m.H[4 * g + 2] = m.H[4 * g + 3] = KSW_NEG_INF;
		}
	}
	__syncwarp();
	int32_t H0 = 0;
	int last_H0_t = 0, last_st = -1, last_en = -1;
	unsigned long long cells = 0;
	const int n_diag = qlen + tlen - 1;
	for (int r = 0; r < n_diag; ++r) {
		int st = 0, en = tlen - 1, st0, en0;
		if (st < r - qlen + 1) st = r - qlen + 1;
		if (en > r) en = r;
		if (st < (r - w + 1) >> 1) st = (r - w + 1) >> 1;
		if (en > (r + w) >> 1) en = (r + w) >> 1;
		if (st > en) { zdropped = 1; break; }
		st0 = st, en0 = en;
		st = st / 16 * 16, en = (en + 16) / 16 * 16 - 1;
		const int bnd = r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2;
		/* the left neighbour of column st: upstream takes x[st-1], v[st-1], x2[st-1] only when that column was
		 * computed on the previous diagonal (the sweep reads them with everything else), else the boundary values */
		const bool edge_mem = st > 0 && st - 1 >= last_st && st - 1 <= last_en;
		const uint32_t x1 = (uint32_t)(128 - q - e) << 16, x21 = (uint32_t)(128 - q2 - e2) << 16, v1 = (uint32_t)(128 + (st > 0 ? -q - e : bnd)) << 16;
		__syncwarp();
		if (en >= r && lane == 0) dp_set<DP_Y>(m, r, -q - e), dp_set<DP_Y2>(m, r, -q2 - e2), dp_set<DP_U>(m, r, bnd);
		__syncwarp();
		/* One sweep per 128 columns (four per lane, as two 16x2 words), from high t to low t.
		 * Scores: upstream stores 16 lanes at a time starting at st0 (lanes past en0 included; bytes below st0 keep
		 * their old value, which the 16-aligned core does read): one aligned word of the target, the query bytes
		 * through a funnel of two aligned words, byte-parallel compare, merged into the lane's own word of s[].
		 * Core: every load of a sweep precedes its stores, so a cell reads its left neighbour's OLD x/v/x2. */
		const int lim = st0 + ((en0 - st0) / 16 + 1) * 16 < T16 ? st0 + ((en0 - st0) / 16 + 1) * 16 : T16;
		const int qoff = T16 + (qlen - 1 - r);               /* qrr = qr + (qlen-1-r), qr = sf + T16 */
		const uint32_t qsel = 0x3210u + 0x1111u * (uint32_t)(qoff & 3);
		const int top = en > lim - 1 ? en : lim - 1;
		uint8_t *pr = tb + (size_t)r * n_col - st;
		for (int base = st + ((top - st) >> 7 << 7); base >= st; base -= 128) {
			const int t = base + 4 * lane;
			const bool act = t <= en, sact = t < lim && t + 4 > st0;
			uint32_t sw = 0, A0, A1, A2, A3, B0, B1, B2, B3, C0, C1, C2, C3, xp, vp, x2p;   /* only read where act */
			if (sact) {
				const uint32_t tw = DP_LD32(m.sf + t);
				const int qa = (qoff + t) & ~3;
				const uint32_t qlo = DP_LD32(m.sf + qa), qhi = qa + 4 < m.flat_sz ? DP_LD32(m.sf + qa + 4) : 0u;
				const uint32_t qw = dp_prmt(qlo, qhi, qsel);
				const uint32_t nz = (((tw ^ qw) + 0x7f7f7f7fu) >> 7) & 0x01010101u;            /* 1 where the bases differ */
				const uint32_t isn = (((tw | qw) >> 2) & 0x01010101u) * 0xffu;                  /* 0xff where either is N */
				sw = sc_b4 - nz * (uint32_t)(sc_mch - sc_mis);
				sw = (sw & ~isn) | (scN_b4 & isn);
				const int lo = st0 - t, hi = lim - t;
				if (lo > 0 || hi < 4) {
					const uint32_t vm = (hi >= 4 ? 0xffffffffu : (1u << 8 * hi) - 1) & ~(lo <= 0 ? 0u : (1u << 8 * lo) - 1);
					sw = (sw & vm) | (DP_LD32(m.s + t) & ~vm);
				}
				DP_ST32(m.s + t, sw);
			} else if (act) sw = DP_LD32(m.s + t);
			if (act) {
				const DpQuad a = *(const DpQuad*)(m.ga + t), b = *(const DpQuad*)(m.gb + t), c = *(const DpQuad*)(m.gc + t);
				A0 = a.x, A1 = a.y, A2 = a.z, A3 = a.w, B0 = b.x, B1 = b.y, B2 = b.z, B3 = b.w, C0 = c.x, C1 = c.y, C2 = c.z, C3 = c.w;
				if (t == st && !edge_mem) xp = x1, vp = v1, x2p = x21;
				else {
					const DpPair pb = *(const DpPair*)(m.gb + t - 2);
					xp = pb.x, vp = pb.y, x2p = m.gc[t - 3];
				}
			}
			__syncwarp();
			if (act) {
				uint32_t nu0, nv0, nx0, ny0, nx20, ny20, nd0, nu1, nv1, nx1, ny1, nx21, ny21, nd1;
				dp_cell2(dk, dp_prmt(sw, 0, 0x4140), A0, A2, C2, __funnelshift_l(xp, B0, 16), __funnelshift_l(vp, B1, 16), __funnelshift_l(x2p, C0, 16),
				         nu0, nv0, nx0, ny0, nx20, ny20, nd0);
				dp_cell2(dk, dp_prmt(sw, 0, 0x4342), A1, A3, C3, __funnelshift_l(B0, B2, 16), __funnelshift_l(B1, B3, 16), __funnelshift_l(C0, C1, 16),
				         nu1, nv1, nx1, ny1, nx21, ny21, nd1);
				DpQuad a, b, c;
				a.x = nu0, a.y = nu1, a.z = ny0, a.w = ny1, b.x = nx0, b.y = nv0, b.z = nx1, b.w = nv1, c.x = nx20, c.y = nx21, c.z = ny20, c.w = ny21;
				*(DpQuad*)(m.ga + t) = a, *(DpQuad*)(m.gb + t) = b, *(DpQuad*)(m.gc + t) = c;
				DP_ST32(pr + t, dp_prmt(nd0, nd1, 0x6420));
			}
			__syncwarp();
		}
		cells += (unsigned long long)(en0 - st0 + 1);
		if (!approx_max) {
			int32_t max_H, max_t;
			if (r > 0) {
				const int en1 = st0 + (en0 - st0) / 4 * 4;
				int32_t hen = en0 > 0 ? m.H[en0 - 1] + dp_get<DP_U>(m, en0) : m.H[en0] + dp_get<DP_V>(m, en0);
				__syncwarp();
				/* H[t] += v[t] for t in [st0, en0); upstream's 4-lane SSE max followed by a scalar tail */
				long long kg = -1, kr = -1; /* (value, tie) keys; larger wins */
				for (int t = st0 + lane; t < en0; t += 32) {
					int32_t h = m.H[t] + dp_get<DP_V>(m, t);
					m.H[t] = h;
					if (t < en1) { /* group region: value, then SSE lane (t-st0)&3 ascending, then t ascending */
						long long k = ((long long)h - KSW_NEG_INF) << 28 | (long long)(3 - ((t - st0) & 3)) << 26 | (long long)(0x3ffffff - t);
						if (k > kg) kg = k;
					} else { /* scalar tail: value, then t ascending */
						long long k = ((long long)h - KSW_NEG_INF) << 28 | (long long)(0x3ffffff - t);
						if (k > kr) kr = k;
					}
				}
				if (lane == 0) m.H[en0] = hen;
#pragma unroll
				for (int d = 16; d; d >>= 1) {
					long long og = __shfl_xor_sync(MMG_FULL, kg, d), orr = __shfl_xor_sync(MMG_FULL, kr, d);
					if (og > kg) kg = og;
					if (orr > kr) kr = orr;
				}
				max_H = hen, max_t = en0;
				if (kg >= 0) {
					int32_t vg = (int32_t)((kg >> 28) + KSW_NEG_INF);
					if (vg > max_H) max_H = vg, max_t = 0x3ffffff - (int)(kg & 0x3ffffffLL);
				}
				if (kr >= 0) {
					int32_t vr = (int32_t)((kr >> 28) + KSW_NEG_INF);
					if (vr > max_H) max_H = vr, max_t = 0x3ffffff - (int)(kr & 0x3ffffffLL);
				}
				__syncwarp();
			} else {
				max_H = dp_get<DP_V>(m, 0) - qe, max_t = 0;
				__syncwarp();
				if (lane == 0) m.H[0] = max_H;
				__syncwarp();
			}
			const int32_t Hen0 = m.H[en0], Hst0 = m.H[st0];
			if (en0 == tlen - 1 && Hen0 > ez_mte) ez_mte = Hen0, ez_mte_q = r - en0;
			if (r - st0 == qlen - 1 && Hst0 > ez_mqe) ez_mqe = Hst0, ez_mqe_t = st0;
			if (ez_apply_zdrop(&ez_max, &ez_max_t, &ez_max_q, max_H, r, max_t, zdrop, e2)) { zdropped = 1; break; }
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez_score = m.H[tlen - 1];
		} else {
			if (r > 0) {
				if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
					int32_t d0 = dp_get<DP_V>(m, last_H0_t), d1 = dp_get<DP_U>(m, last_H0_t + 1);
					if (d0 > d1) H0 += d0;
					else H0 += d1, ++last_H0_t;
				} else if (last_H0_t >= st0 && last_H0_t <= en0) {
					H0 += dp_get<DP_V>(m, last_H0_t);
				} else {
					++last_H0_t, H0 += dp_get<DP_U>(m, last_H0_t);
				}
			} else H0 = dp_get<DP_V>(m, 0) - qe, last_H0_t = 0;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez_score = H0;
		}
		last_st = st, last_en = en;
	}
	__syncwarp();
	int n_cigar = 0;
	{
		int i = -1, j = -1;
		if (!zdropped && !(flag & EZ_EXTZ_ONLY)) i = tlen - 1, j = qlen - 1;
		else if (!zdropped && (flag & EZ_EXTZ_ONLY) && ez_mqe + end_bonus > ez_max) reach_end = 1, i = ez_mqe_t, j = qlen - 1;
		else if (ez_max_t >= 0 && ez_max_q >= 0) i = ez_max_t, j = ez_max_q;
		if (i >= 0 || j >= 0 || reach_end) n_cigar = ext_backtrack_warp(tb, n_col, qlen, tlen, w, i, j, (flag & EZ_REV_CIGAR) != 0, cigar);
	}
	if (lane == 0) {
		res->max = ez_max, res->max_q = ez_max_q, res->max_t = ez_max_t, res->mqe = ez_mqe, res->mqe_t = ez_mqe_t, res->score = ez_score;
		res->zdropped = (uint8_t)zdropped, res->reach_end = (uint8_t)reach_end, res->n_cigar = n_cigar;
		*n_cell += cells;
	}
	__syncwarp();
}

__device__ __forceinline__ void ext_dp_pass(bool in_smem, unsigned char *gbase, int T16, int Q16, const DevOpt &o, int qlen, int tlen, int w, int zdrop, int end_bonus,
                                            int flag, uint8_t *tb, uint32_t *cigar, ExtJob *res, unsigned long long *n_cell)
{
	if (in_smem) ext_dp_pass_t<true>(gbase, T16, Q16, o, qlen, tlen, w, zdrop, end_bonus, flag, tb, cigar, res, n_cell);
	else ext_dp_pass_t<false>(gbase, T16, Q16, o, qlen, tlen, w, zdrop, end_bonus, flag, tb, cigar, res, n_cell);
}

/* align.c mm_test_zdrop without the inversion test (lane 0); returns max_zdrop and the most-dropped region */
__device__ int ext_test_zdrop(const DpMem &m, const DevOpt &o, int qlen, int n_cigar, const uint32_t *cigar, int pos[2][2])
{
	const int a_ = o.a < 0 ? -o.a : o.a, b_ = o.b > 0 ? -o.b : o.b, amb = o.sc_ambi > 0 ? -o.sc_ambi : o.sc_ambi;
	const uint8_t *tseq = m.sf, *qr = m.sf + m.T16; /* qseq[j] = qr[qlen-1-j] */
	int32_t score = 0, mx = -2147483647 - 1, max_i = -1, max_j = -1, i = 0, j = 0, max_zdrop = 0;
	pos[0][0] = pos[0][1] = pos[1][0] = pos[1][1] = -1;
	auto upd = [&](int32_t sc, int ii, int jj) {
		if (sc < mx) {
			int li = ii - max_i, lj = jj - max_j;
			int diff = li > lj ? li - lj : lj - li;
			int z = mx - sc - diff * o.e;
			if (z > max_zdrop) {
				max_zdrop = z;
				pos[0][0] = max_i, pos[0][1] = ii;
				pos[1][0] = max_j, pos[1][1] = jj;
			}
		} else mx = sc, max_i = ii, max_j = jj;
	};
	for (int k = 0; k < n_cigar; ++k) {
		const int op = cigar[k] & 0xf, len = (int)(cigar[k] >> 4);
		if (op == 0) {
			for (int l = 0; l < len; ++l) {
				const int ct = tseq[i + l], cq = qr[qlen - 1 - (j + l)];
				score += (ct == 4 || cq == 4) ? amb : ct == cq ? a_ : b_;
				upd(score, i + l, j + l);
			}
			i += len, j += len;
		} else if (op == 1 || op == 2 || op == 3) {
			score -= o.q + o.e * len;
			if (op == 1) j += len;
			else i += len;
			upd(score, i, j);
		}
	}
	return max_zdrop;
}

/* Local-alignment score of the reverse complement of the most-dropped query stretch against its
 * target stretch (ksw_ll_i16 inside mm_test_zdrop; only the score is used).  Anti-diagonal
 * Smith-Waterman over the whole warp; `scr` is int32 scratch (the finished traceback slice). */
__device__ int ext_inv_score(const DpMem &m, const DevOpt &o, int qlen, int pos[2][2], int32_t *scr)
{
	const int lane = mmg_lane();
	const int q_len = pos[1][1] - pos[1][0], t_len = pos[0][1] - pos[0][0];
	if (q_len <= 0 || t_len <= 0) return 0;
	const int a_ = o.a < 0 ? -o.a : o.a, b_ = o.b > 0 ? -o.b : o.b, amb = o.sc_ambi > 0 ? -o.sc_ambi : o.sc_ambi;
	const int gapoe = o.q + o.e, gape = o.e;
	const uint8_t *tseq = m.sf + pos[0][0], *qr = m.sf + m.T16;
	const int n = t_len + 1;
	int32_t *H0 = scr, *H1 = scr + n, *H2 = scr + 2 * n, *E0 = scr + 3 * n, *E1 = scr + 4 * n, *F0 = scr + 5 * n, *F1 = scr + 6 * n;
	int gmax = 0;
	/* query base j of qseq2: c = qseq[pos[1][1] - j - 1], complemented; qseq[x] = qr[qlen-1-x] */
	for (int d = 0; d < q_len + t_len - 1; ++d) {
		int ilo = d - (q_len - 1) > 0 ? d - (q_len - 1) : 0, ihi = d < t_len - 1 ? d : t_len - 1;
		for (int i = ilo + lane; i <= ihi; i += 32) {
			const int j = d - i;
			int cq = qr[qlen - 1 - (pos[1][1] - j - 1)];
			cq = cq >= 4 ? 4 : 3 - cq;
			const int ct = tseq[i];
			const int s = (ct == 4 || cq == 4) ? amb : ct == cq ? a_ : b_;
			const int diag = (i > 0 && j > 0) ? H2[i - 1] : 0;
			const int up = i > 0 ? H1[i - 1] : 0, eup = i > 0 ? E1[i - 1] : 0;
			const int left = j > 0 ? H1[i] : 0, fl = j > 0 ? F1[i] : 0;
			int e = eup - gape > up - gapoe ? eup - gape : up - gapoe;
			int f = fl - gape > left - gapoe ? fl - gape : left - gapoe;
			if (e < 0) e = 0;
			if (f < 0) f = 0;
			if (i == 0) e = 0;
			if (j == 0) f = 0;
			int h = diag + s;
			h = h > e ? h : e;
			h = h > f ? h : f;
			h = h > 0 ? h : 0;
			H0[i] = h, E0[i] = e, F0[i] = f;
			gmax = gmax > h ? gmax : h;
		}
		__syncwarp();
		int32_t *t = H2; H2 = H1, H1 = H0, H0 = t;
		t = E1, E1 = E0, E0 = t;
		t = F1, F1 = F0, F0 = t;
	}
	gmax = __reduce_max_sync(MMG_FULL, gmax);
	return gmax > 32767 ? 32767 : gmax;
}

#include "extend_fill.inc"

__global__ void __launch_bounds__(EXT_DP_WARPS * 32, 4)
ext_dp_kernel(ChunkDev c, DevIndex di, DevOpt o, ExtBufs xb, uint32_t j0, uint64_t tb_slice, int pass, int last_pass, uint32_t *work)
{
	MMG_DYN_SMEM(smem_raw);
	const int lane = mmg_lane(), wib = threadIdx.x >> 5;
	unsigned char *my_smem = smem_raw + (size_t)wib * EXT_SMEM_PER_WARP;
	unsigned long long n_cell = 0;
	const uint32_t j1 = ext_n_jobs(xb);
	if (!ext_cigar_fits(xb)) return;
	/* the traceback of a job lives only while the job runs: every resident warp owns one slice of the arena */
	uint8_t *tb = xb.tb + (size_t)(blockIdx.x * EXT_DP_WARPS + wib) * tb_slice;
	/* pass 0 walks the jobs of this round, 32 per claim, skipping the ones ext_fill_kernel finished; a job whose
	 * traceback does not fit this pass' slice goes on a list for the next pass (fewer warps, larger slices) */
	uint32_t seg0 = 0, n_list = 0;
	for (int k = 0; k + 1 < pass; ++k) seg0 += xb.ovf_n[k];
	if (pass > 0) n_list = xb.ovf_n[pass - 1];
	const uint32_t seg1 = pass > 0 ? seg0 + n_list : 0;
	uint32_t todo = 0, base = 0;
	for (;;) {
		uint32_t ji;
		if (pass == 0) {
			while (!todo) {
				base = j0 + 32u * mmg_next_item(work);
				if (base >= j1) break;
				const uint32_t jl = base + (uint32_t)lane;
				todo = __ballot_sync(MMG_FULL, jl < j1 && !xb.jobs[jl].pad[0]);
			}
			if (!todo) break;
			ji = base + (uint32_t)(__ffs((int)todo) - 1);
			todo &= todo - 1;
		} else {
			const uint32_t k = mmg_next_item(work);
			if (k >= n_list) break;
			ji = xb.ovf[seg0 + k];
		}
		ExtJob *jb = &xb.jobs[ji];
		if (jb->tb_size > tb_slice) {
			if (lane == 0) {
				if (!last_pass) xb.ovf[seg1 + atomicAdd(&xb.ovf_n[pass], 1u)] = ji;
				else {
					atomicOr(&c.flags[jb->read], 0x10000000u), atomicOr(c.err, 0x10000000u);
					jb->n_cigar = 0, jb->zdropped = 1, jb->max = 0, jb->max_q = jb->max_t = -1, jb->zdrop_code = 0, jb->reach_end = 0, jb->pad[0] = 1;
				}
			}
			__syncwarp();
			continue;
		}
		const int qlen = jb->qe - jb->qs, tlen = jb->re - jb->rs;
		const int flag = jb->flag;
		const uint32_t r = jb->read;
		const char *seq = c.seq + c.off[r];
		const int rlen = (int)(c.off[r + 1] - c.off[r]);
		uint32_t *cigar = xb.jcigar + jb->cg_off;
		if (lane == 0) jb->pad[0] = 1;
		if (qlen <= 0 || tlen <= 0) { /* ksw_extd2_sse returns an empty ez */
			if (lane == 0) {
				jb->max = 0, jb->max_q = jb->max_t = jb->mqe_t = -1, jb->mqe = jb->score = KSW_NEG_INF;
				jb->zdropped = 0, jb->reach_end = 0, jb->n_cigar = 0, jb->zdrop_code = 0;
			}
			__syncwarp();
			continue;
		}
		const int T16 = (tlen + 15) / 16 * 16, Q16 = (qlen + 15) / 16 * 16;
		const size_t need = EXT_DP_BYTES(T16, Q16);
		DpMem m;
		unsigned char *base = need <= EXT_SMEM_PER_WARP ? my_smem : xb.big + (size_t)(blockIdx.x * EXT_DP_WARPS + wib) * xb.big_per_warp;
		if (need > EXT_SMEM_PER_WARP && need > xb.big_per_warp) { /* longer than the per-warp global slice: reported by the host */
			if (lane == 0) atomicOr(&c.flags[r], 0x20000000u), atomicOr(c.err, 0x20000000u), jb->n_cigar = 0, jb->zdropped = 1, jb->max = 0, jb->max_q = jb->max_t = -1, jb->zdrop_code = 0, jb->reach_end = 0;
			__syncwarp();
			continue;
		}
		const bool in_smem = need <= EXT_SMEM_PER_WARP;
		ext_dpmem_set(m, base, T16, Q16);
		/* load the sequences: target codes, reversed query codes, zero padding */
		const uint64_t toff = di.seq_off[jb->rid];
		const bool is_left = jb->kind == EXT_LEFT;
		for (int i = lane; i < m.flat_sz; i += 32) {
			int v = 0;
			if (i < tlen) v = is_left ? ext_tbase(di, toff, jb->re - 1 - i) : ext_tbase(di, toff, jb->rs + i);
			else if (i >= T16 && i - T16 < qlen) {
				const int j = qlen - 1 - (i - T16);                 /* qr[t] = query[qlen-1-t] */
				v = is_left ? ext_qbase(seq, rlen, jb->rev, jb->qe - 1 - j) : ext_qbase(seq, rlen, jb->rev, jb->qs + j);
			}
			m.sf[i] = (uint8_t)v;
		}
		__syncwarp();
		if ((int64_t)tlen * qlen > o.max_sw_mat && o.max_sw_mat > 0) { /* mm_align_pair: treated as z-dropped */
			if (lane == 0) jb->max = 0, jb->max_q = jb->max_t = jb->mqe_t = -1, jb->mqe = jb->score = KSW_NEG_INF, jb->zdropped = 1, jb->reach_end = 0, jb->n_cigar = 0, jb->zdrop_code = 0;
			__syncwarp();
			continue;
		}
		ext_dp_pass(in_smem, base, T16, Q16, o, qlen, tlen, jb->w, jb->zdrop, jb->end_bonus, flag, tb, cigar, jb, &n_cell);
		if (jb->kind == EXT_FILL) {
			/* mm_test_zdrop on the first-pass CIGAR; a second, exact pass when the score drops too much */
			int code = 0, pos[2][2];
			if (lane == 0) {
				int max_zdrop = ext_test_zdrop(m, o, qlen, jb->n_cigar, cigar, pos);
				const int q_len = pos[1][1] - pos[1][0], t_len = pos[0][1] - pos[0][0];
				if (max_zdrop > o.zdrop_inv && q_len < o.max_gap && t_len < o.max_gap) code = 3; /* inversion test needed */
				else code = max_zdrop > o.zdrop ? 1 : 0;
			}
			code = __shfl_sync(MMG_FULL, code, 0);
			if (code == 3) {
				int sc = 0;
				pos[0][0] = __shfl_sync(MMG_FULL, pos[0][0], 0), pos[0][1] = __shfl_sync(MMG_FULL, pos[0][1], 0);
				pos[1][0] = __shfl_sync(MMG_FULL, pos[1][0], 0), pos[1][1] = __shfl_sync(MMG_FULL, pos[1][1], 0);
				sc = ext_inv_score(m, o, qlen, pos, (int32_t*)tb);
				if (sc >= o.min_chain_score * o.a && sc >= o.min_dp_max) code = 2;
				else {
					int mz = 0;
					if (lane == 0) mz = ext_test_zdrop(m, o, qlen, jb->n_cigar, cigar, pos);
					mz = __shfl_sync(MMG_FULL, mz, 0);
					code = mz > o.zdrop ? 1 : 0;
				}
			}
			if (code != 0) ext_dp_pass(in_smem, base, T16, Q16, o, qlen, tlen, jb->w, code == 2 ? o.zdrop_inv : o.zdrop, -1, 0, tb, cigar, jb, &n_cell);
			if (lane == 0) jb->zdrop_code = (uint8_t)code;
		}
		__syncwarp();
	}
	if (lane == 0 && n_cell) atomicAdd(&c.stats[7], n_cell);
}

#include "extend_stitch.inc"
