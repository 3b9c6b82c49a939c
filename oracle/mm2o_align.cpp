/* mm2o_align.cpp -- ORACLE (test infrastructure only).
 * Restates minimap2 v2.26 align.c (ksw_gen_simple_mat, mm_fix_cigar,
 * mm_update_extra, mm_append_cigar, mm_align_pair, mm_test_zdrop,
 * mm_fix_bad_ends, mm_filter_bad_seeds(_alt), mm_adjust_minier, mm_align1,
 * mm_align1_inv, mm_align_skeleton), hit.c mm_squeeze_a, and format.c
 * write_cs_core / write_MD_core (mm_gen_cs, mm_gen_MD) for the single-segment,
 * non-splice, non-SR path.  Reached from /root/reference/src/lib.rs:482,587
 * because mappy-rs always sets MM_F_CIGAR (src/lib.rs:339); `cs` is requested
 * by every map_batch worker (src/lib.rs:589).
 * Pinned by the reference only through `map_one` (r_st == 0, r_en == 400;
 * src/lib.rs:1094-1106); everything else here is parity unpinned.
 */
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <assert.h>
#include <string>
#include <vector>
#include "mm2o.h"
#include "mm2o_sort.h"
#include "mm2o_ksw2.h"

static inline float mg_log2(float x) // NB: this doesn't work when x<2 (mmpriv.h)
{
	union { float f; uint32_t i; } z = { x };
	float log_2 = ((z.i >> 23) & 255) - 128;
	z.i &= ~(255 << 23);
	z.i += 127 << 23;
	log_2 += (-0.34484843f * z.f + 2.02466578f) * z.f - 0.67487759f;
	return log_2;
}

static void ksw_gen_simple_mat(int m, int8_t *mat, int8_t a, int8_t b, int8_t sc_ambi)
{
	int i, j;
	a = a < 0 ? -a : a;
	b = b > 0 ? -b : b;
	sc_ambi = sc_ambi > 0 ? -sc_ambi : sc_ambi;
	for (i = 0; i < m - 1; ++i) {
		for (j = 0; j < m - 1; ++j)
			mat[i * m + j] = i == j ? a : b;
		mat[i * m + m - 1] = sc_ambi;
	}
	for (j = 0; j < m; ++j)
		mat[(m - 1) * m + j] = sc_ambi;
}

static inline void mm_seq_rev(uint32_t len, uint8_t *seq)
{
	uint32_t i;
	uint8_t t;
	for (i = 0; i < len >> 1; ++i)
		t = seq[i], seq[i] = seq[len - 1 - i], seq[len - 1 - i] = t;
}

static inline void update_max_zdrop(int32_t score, int i, int j, int32_t *max, int *max_i, int *max_j, int e, int *max_zdrop, int pos[2][2])
{
	if (score < *max) {
		int li = i - *max_i;
		int lj = j - *max_j;
		int diff = li > lj ? li - lj : lj - li;
		int z = *max - score - diff * e;
		if (z > *max_zdrop) {
			*max_zdrop = z;
			pos[0][0] = *max_i, pos[0][1] = i;
			pos[1][0] = *max_j, pos[1][1] = j;
		}
	} else *max = score, *max_i = i, *max_j = j;
}

static int mm_test_zdrop(const mm_mapopt_t *opt, const uint8_t *qseq, const uint8_t *tseq, uint32_t n_cigar, const uint32_t *cigar, const int8_t *mat)
{
	uint32_t k;
	int32_t score = 0, max = INT32_MIN, max_i = -1, max_j = -1, i = 0, j = 0, max_zdrop = 0;
	int pos[2][2] = {{-1, -1}, {-1, -1}}, q_len, t_len;

	// find the score and the region where score drops most along diagonal
	for (k = 0, score = 0; k < n_cigar; ++k) {
		uint32_t l, op = cigar[k] & 0xf, len = cigar[k] >> 4;
		if (op == MM_CIGAR_MATCH) {
			for (l = 0; l < len; ++l) {
				score += mat[tseq[i + l] * 5 + qseq[j + l]];
				update_max_zdrop(score, i + l, j + l, &max, &max_i, &max_j, opt->e, &max_zdrop, pos);
			}
			i += len, j += len;
		} else if (op == MM_CIGAR_INS || op == MM_CIGAR_DEL || op == MM_CIGAR_N_SKIP) {
			score -= opt->q + opt->e * len;
			if (op == MM_CIGAR_INS) j += len;
			else i += len;
			update_max_zdrop(score, i, j, &max, &max_i, &max_j, opt->e, &max_zdrop, pos);
		}
	}

	// test if there is an inversion in the most dropped region
	q_len = pos[1][1] - pos[1][0], t_len = pos[0][1] - pos[0][0];
	if (!(opt->flag & (MM_F_SPLICE | MM_F_SR | MM_F_FOR_ONLY | MM_F_REV_ONLY)) && max_zdrop > opt->zdrop_inv && q_len < opt->max_gap && t_len < opt->max_gap) {
		std::vector<uint8_t> qseq2(q_len > 0 ? q_len : 1);
		int q_off, t_off;
		for (i = 0; i < q_len; ++i) {
			int c = qseq[pos[1][1] - i - 1];
			qseq2[i] = c >= 4 ? 4 : 3 - c;
		}
		score = ksw_ll_i16(q_len, qseq2.data(), 5, mat, t_len, tseq + pos[0][0], opt->q, opt->e, &q_off, &t_off);
		if (score >= opt->min_chain_score * opt->a && score >= opt->min_dp_max)
			return 2; // there is a potential inversion
	}
	return max_zdrop > opt->zdrop ? 1 : 0;
}

static void mm_fix_cigar(mm_reg1_t *r, const uint8_t *qseq, const uint8_t *tseq, int *qshift, int *tshift)
{
	mm_extra_t *p = r->p;
	int32_t toff = 0, qoff = 0, to_shrink = 0;
	uint32_t k;
	std::vector<uint32_t> &cigar = p->cigar;
	uint32_t n_cigar = (uint32_t)cigar.size();
	*qshift = *tshift = 0;
	if (n_cigar <= 1) return;
	for (k = 0; k < n_cigar; ++k) { // indel left alignment
		uint32_t op = cigar[k] & 0xf, len = cigar[k] >> 4;
		if (len == 0) to_shrink = 1;
		if (op == MM_CIGAR_MATCH) {
			toff += len, qoff += len;
		} else if (op == MM_CIGAR_INS || op == MM_CIGAR_DEL) {
			if (k > 0 && k < n_cigar - 1 && (cigar[k - 1] & 0xf) == 0 && (cigar[k + 1] & 0xf) == 0) {
				int l, prev_len = cigar[k - 1] >> 4;
				if (op == MM_CIGAR_INS) {
					for (l = 0; l < prev_len; ++l)
						if (qseq[qoff - 1 - l] != qseq[qoff + len - 1 - l])
							break;
				} else {
					for (l = 0; l < prev_len; ++l)
						if (tseq[toff - 1 - l] != tseq[toff + len - 1 - l])
							break;
				}
				if (l > 0)
					cigar[k - 1] -= l << 4, cigar[k + 1] += l << 4, qoff -= l, toff -= l;
				if (l == prev_len) to_shrink = 1;
			}
			if (op == MM_CIGAR_INS) qoff += len;
			else toff += len;
		} else if (op == MM_CIGAR_N_SKIP) {
			toff += len;
		}
	}
	assert(qoff == r->qe - r->qs && toff == r->re - r->rs);
	for (k = 0; k + 2 < n_cigar; ++k) { // fix CIGAR like 5I6D7I
		if ((cigar[k] & 0xf) > 0 && (cigar[k] & 0xf) + (cigar[k + 1] & 0xf) == 3) {
			uint32_t l, s[3] = {0, 0, 0};
			for (l = k; l < n_cigar; ++l) { // count number of adjacent I and D
				uint32_t op = cigar[l] & 0xf;
				if (op == MM_CIGAR_INS || op == MM_CIGAR_DEL || cigar[l] >> 4 == 0)
					s[op] += cigar[l] >> 4;
				else break;
			}
			if (s[1] > 0 && s[2] > 0 && l - k > 2) { // turn to a single I and a single D
				cigar[k]     = s[1] << 4 | MM_CIGAR_INS;
				cigar[k + 1] = s[2] << 4 | MM_CIGAR_DEL;
				for (k += 2; k < l; ++k)
					cigar[k] &= 0xf;
				to_shrink = 1;
			}
			k = l;
		}
	}
	if (to_shrink) { // squeeze out zero-length operations
		int32_t l = 0;
		for (k = 0; k < n_cigar; ++k) // squeeze out zero-length operations
			if (cigar[k] >> 4 != 0)
				cigar[l++] = cigar[k];
		n_cigar = l;
		for (k = l = 0; k < n_cigar; ++k) // merge two adjacent operations if they are the same
			if (k == n_cigar - 1 || (cigar[k] & 0xf) != (cigar[k + 1] & 0xf))
				cigar[l++] = cigar[k];
			else cigar[k + 1] += cigar[k] >> 4 << 4; // add length to the next CIGAR operator
		n_cigar = l;
	}
	if ((cigar[0] & 0xf) == MM_CIGAR_INS || (cigar[0] & 0xf) == MM_CIGAR_DEL) { // get rid of leading I or D
		int32_t l = cigar[0] >> 4;
		if ((cigar[0] & 0xf) == MM_CIGAR_INS) {
			if (r->rev) r->qe -= l;
			else r->qs += l;
			*qshift = l;
		} else r->rs += l, *tshift = l;
		--n_cigar;
		memmove(cigar.data(), cigar.data() + 1, n_cigar * 4);
	}
	cigar.resize(n_cigar);
}

static void mm_update_extra(mm_reg1_t *r, const uint8_t *qseq, const uint8_t *tseq, const int8_t *mat, int8_t q, int8_t e, int log_gap)
{
	uint32_t k, l;
	int32_t qshift, tshift, toff = 0, qoff = 0;
	double s = 0.0, max = 0.0;
	mm_extra_t *p = r->p;
	if (p == 0) return;
	mm_fix_cigar(r, qseq, tseq, &qshift, &tshift);
	qseq += qshift, tseq += tshift; // qseq and tseq may be shifted due to the removal of leading I/D
	r->blen = r->mlen = 0;
	for (k = 0; k < p->cigar.size(); ++k) {
		uint32_t op = p->cigar[k] & 0xf, len = p->cigar[k] >> 4;
		if (op == MM_CIGAR_MATCH) {
			int n_ambi = 0, n_diff = 0;
			for (l = 0; l < len; ++l) {
				int cq = qseq[qoff + l], ct = tseq[toff + l];
				if (ct > 3 || cq > 3) ++n_ambi;
				else if (ct != cq) ++n_diff;
				s += mat[ct * 5 + cq];
				if (s < 0) s = 0;
				else max = max > s ? max : s;
			}
			r->blen += len - n_ambi, r->mlen += len - (n_ambi + n_diff), p->n_ambi += n_ambi;
			toff += len, qoff += len;
		} else if (op == MM_CIGAR_INS) {
			int n_ambi = 0;
			for (l = 0; l < len; ++l)
				if (qseq[qoff + l] > 3) ++n_ambi;
			r->blen += len - n_ambi, p->n_ambi += n_ambi;
			if (log_gap) s -= q + (double)e * mg_log2(1.0 + len);
			else s -= q + e;
			if (s < 0) s = 0;
			qoff += len;
		} else if (op == MM_CIGAR_DEL) {
			int n_ambi = 0;
			for (l = 0; l < len; ++l)
				if (tseq[toff + l] > 3) ++n_ambi;
			r->blen += len - n_ambi, p->n_ambi += n_ambi;
			if (log_gap) s -= q + (double)e * mg_log2(1.0 + len);
			else s -= q + e;
			if (s < 0) s = 0;
			toff += len;
		} else if (op == MM_CIGAR_N_SKIP) {
			toff += len;
		}
	}
	p->dp_max = (int32_t)(max + .499);
	assert(qoff == r->qe - r->qs && toff == r->re - r->rs);
}

static void mm_append_cigar(mm_reg1_t *r, const std::vector<uint32_t> &cigar)
{
	if (cigar.empty()) return;
	if (r->p == 0) {
		r->p = new mm_extra_t();
		r->p->capacity = 0, r->p->dp_score = r->p->dp_max = r->p->dp_max2 = 0, r->p->n_ambi = r->p->trans_strand = 0;
	}
	std::vector<uint32_t> &c = r->p->cigar;
	if (!c.empty() && (c.back() & 0xf) == (cigar[0] & 0xf)) { // same CIGAR op at the boundary
		c.back() += (cigar[0] >> 4) << 4;
		c.insert(c.end(), cigar.begin() + 1, cigar.end());
	} else c.insert(c.end(), cigar.begin(), cigar.end());
}

static void mm_align_pair(const mm_mapopt_t *opt, int qlen, const uint8_t *qseq, int tlen, const uint8_t *tseq, const int8_t *mat, int w, int end_bonus, int zdrop, int flag, ksw_extz_t *ez, mm2o_stats_t *st)
{
	if (opt->transition != 0 && opt->b != opt->transition)
		flag |= KSW_EZ_GENERIC_SC;
	{ static FILE *jl = getenv("MM2O_JOBLOG") ? fopen(getenv("MM2O_JOBLOG"), "w") : 0; if (jl) fprintf(jl, "%d %d %d %d\n", qlen, tlen, w, flag); }
	if (opt->max_sw_mat > 0 && (int64_t)tlen * qlen > opt->max_sw_mat) {
		ksw_reset_extz(ez);
		ez->zdropped = 1;
	} else if (opt->q == opt->q2 && opt->e == opt->e2) {
		/* single affine gap: upstream's separate kernel (a 4-tuple `scoring`, /root/reference/src/lib.rs:369-376) */
		ksw_extz2(qlen, qseq, tlen, tseq, 5, mat, opt->q, opt->e, w, zdrop, end_bonus, flag, ez, st ? &st->n_cell : 0);
	} else
		ksw_extd2(qlen, qseq, tlen, tseq, 5, mat, opt->q, opt->e, opt->q2, opt->e2, w, zdrop, end_bonus, flag, ez, st ? &st->n_cell : 0);
}

static void mm_fix_bad_ends(const mm_reg1_t *r, const mm128_t *a, int bw, int min_match, int32_t *as, int32_t *cnt)
{
	int32_t i, l, m;
	*as = r->as, *cnt = r->cnt;
	if (r->cnt < 3) return;
	m = l = a[r->as].y >> 32 & 0xff;
	for (i = r->as + 1; i < r->as + r->cnt - 1; ++i) {
		int32_t lq, lr, min, max;
		int32_t q_span = a[i].y >> 32 & 0xff;
		if (a[i].y & MM_SEED_LONG_JOIN) break;
		lr = (int32_t)a[i].x - (int32_t)a[i - 1].x;
		lq = (int32_t)a[i].y - (int32_t)a[i - 1].y;
		min = lr < lq ? lr : lq;
		max = lr > lq ? lr : lq;
		if (max - min > l >> 1) *as = i;
		l += min;
		m += min < q_span ? min : q_span;
		if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r->mlen >> 1) break;
	}
	*cnt = r->as + r->cnt - *as;
	m = l = a[r->as + r->cnt - 1].y >> 32 & 0xff;
	for (i = r->as + r->cnt - 2; i > *as; --i) {
		int32_t lq, lr, min, max;
		int32_t q_span = a[i + 1].y >> 32 & 0xff;
		if (a[i + 1].y & MM_SEED_LONG_JOIN) break;
		lr = (int32_t)a[i + 1].x - (int32_t)a[i].x;
		lq = (int32_t)a[i + 1].y - (int32_t)a[i].y;
		min = lr < lq ? lr : lq;
		max = lr > lq ? lr : lq;
		if (max - min > l >> 1) *cnt = i + 1 - *as;
		l += min;
		m += min < q_span ? min : q_span;
		if (l >= bw << 1 || (m >= min_match && m >= bw) || m >= r->mlen >> 1) break;
	}
}

static void mm_filter_bad_seeds(int as1, int cnt1, mm128_t *a, int min_gap, int diff_thres, int max_ext_len, int max_ext_cnt)
{
	int max_st, max_en, n, i, k, max;
	std::vector<int> K(cnt1 > 0 ? cnt1 : 1);
	for (i = 1, n = 0; i < cnt1; ++i) { // collect all gaps
		int gap = ((int32_t)a[as1 + i].y - (int32_t)a[as1 + i - 1].y) - ((int32_t)a[as1 + i].x - (int32_t)a[as1 + i - 1].x);
		if (gap < -min_gap || gap > min_gap)
			K[n++] = i;
	}
	if (n == 0) return;
	max = 0, max_st = max_en = -1;
	for (k = 0;; ++k) { // traverse each gap
		int gap, l, n_ins = 0, n_del = 0, qs, rs, max_diff = 0, max_diff_l = -1;
		if (k == n || k >= max_en) {
			if (max_en > 0)
				for (i = K[max_st]; i < K[max_en]; ++i)
					a[as1 + i].y |= MM_SEED_IGNORE;
			max = 0, max_st = max_en = -1;
			if (k == n) break;
		}
		i = K[k];
		gap = ((int32_t)a[as1 + i].y - (int32_t)a[as1 + i - 1].y) - (int32_t)(a[as1 + i].x - a[as1 + i - 1].x);
		if (gap > 0) n_ins += gap;
		else n_del += -gap;
		qs = (int32_t)a[as1 + i - 1].y;
		rs = (int32_t)a[as1 + i - 1].x;
		for (l = k + 1; l < n && l <= k + max_ext_cnt; ++l) {
			int j = K[l], diff;
			if ((int32_t)a[as1 + j].y - qs > max_ext_len || (int32_t)a[as1 + j].x - rs > max_ext_len) break;
			gap = ((int32_t)a[as1 + j].y - (int32_t)a[as1 + j - 1].y) - (int32_t)(a[as1 + j].x - a[as1 + j - 1].x);
			if (gap > 0) n_ins += gap;
			else n_del += -gap;
			diff = n_ins + n_del - abs(n_ins - n_del);
			if (max_diff < diff)
				max_diff = diff, max_diff_l = l;
		}
		if (max_diff > diff_thres && max_diff > max)
			max = max_diff, max_st = k, max_en = max_diff_l;
	}
}

static void mm_filter_bad_seeds_alt(int as1, int cnt1, mm128_t *a, int min_gap, int max_ext)
{
	int n, i, k;
	std::vector<int> K(cnt1 > 0 ? cnt1 : 1);
	for (i = 1, n = 0; i < cnt1; ++i) { // collect all gaps
		int gap = ((int32_t)a[as1 + i].y - (int32_t)a[as1 + i - 1].y) - ((int32_t)a[as1 + i].x - (int32_t)a[as1 + i - 1].x);
		if (gap < -min_gap || gap > min_gap)
			K[n++] = i;
	}
	for (k = 0; k < n;) { // traverse each gap
		int n_ins = 0, n_del = 0, l, gap;
		i = K[k];
		gap = ((int32_t)a[as1 + i].y - (int32_t)a[as1 + i - 1].y) - (int32_t)(a[as1 + i].x - a[as1 + i - 1].x);
		if (gap > 0) n_ins += gap;
		else n_del += -gap;
		for (l = k + 1; l < n; ++l) {
			int j = K[l], diff;
			if ((int32_t)a[as1 + j].y - (int32_t)a[as1 + i].y > max_ext) break;
			gap = ((int32_t)a[as1 + j].y - (int32_t)a[as1 + j - 1].y) - (int32_t)(a[as1 + j].x - a[as1 + j - 1].x);
			if (gap > 0) n_ins += gap;
			else n_del += -gap;
			diff = n_ins + n_del - abs(n_ins - n_del);
			if (diff > min_gap)
				break;
		}
		if (l < n) {
			int j = K[l];
			for (i = K[k]; i < j; ++i)
				a[as1 + i].y |= MM_SEED_IGNORE;
			k = l + 1;
		} else ++k;
	}
}

static inline void mm_adjust_minier(const mm_idx_t *mi, mm128_t *a, int32_t *r, int32_t *q)
{ /* non-HPC index: the k-mer end is the anchor coordinate */
	*r = (int32_t)a->x + 1;
	*q = (int32_t)a->y + 1;
}

static void mm_align1(const mm_mapopt_t *opt, const mm_idx_t *mi, int qlen, uint8_t *qseq0[2], mm_reg1_t *r, mm_reg1_t *r2, int n_a, mm128_t *a, ksw_extz_t *ez, mm2o_stats_t *st)
{
	int32_t rid = a[r->as].x << 1 >> 33, rev = a[r->as].x >> 63, as1, cnt1;
	uint8_t *qseq;
	int32_t i, l, bw, bw_long, dropped = 0, extra_flag = 0, rs0, re0, qs0, qe0;
	int32_t rs, re, qs, qe;
	int32_t rs1, qs1, re1, qe1;
	int8_t mat[25];

	r2->cnt = 0;
	if (r->cnt == 0) return;
	ksw_gen_simple_mat(5, mat, opt->a, opt->b, opt->sc_ambi);
	bw = (int)(opt->bw * 1.5 + 1.);
	bw_long = (int)(opt->bw_long * 1.5 + 1.);
	if (bw_long < bw) bw_long = bw;

	if (!(opt->flag & MM_F_NO_END_FLT))
		mm_fix_bad_ends(r, a, opt->bw, opt->min_chain_score * 2, &as1, &cnt1);
	else as1 = r->as, cnt1 = r->cnt;
	mm_filter_bad_seeds(as1, cnt1, a, 10, 40, opt->max_gap >> 1, 10);
	mm_filter_bad_seeds_alt(as1, cnt1, a, 30, opt->max_gap >> 1);
	mm_adjust_minier(mi, &a[as1], &rs, &qs);
	mm_adjust_minier(mi, &a[as1 + cnt1 - 1], &re, &qe);
	assert(cnt1 > 0);

	/* Look for the start and end of regions to perform DP. */
	// compute rs0 and qs0
	rs0 = (int32_t)a[r->as].x + 1 - (int32_t)(a[r->as].y >> 32 & 0xff);
	qs0 = (int32_t)a[r->as].y + 1 - (int32_t)(a[r->as].y >> 32 & 0xff);
	if (rs0 < 0) rs0 = 0; // this may happen when HPC is in use
	assert(qs0 >= 0); // this should never happen, or it is logic error
	rs1 = qs1 = 0;
	for (i = r->as - 1, l = 0; i >= 0 && a[i].x >> 32 == a[r->as].x >> 32; --i) { // inspect nearby seeds
		int32_t x = (int32_t)a[i].x + 1 - (int32_t)(a[i].y >> 32 & 0xff);
		int32_t y = (int32_t)a[i].y + 1 - (int32_t)(a[i].y >> 32 & 0xff);
		if (x < rs0 && y < qs0) {
			if (++l > opt->min_cnt) {
				l = rs0 - x > qs0 - y ? rs0 - x : qs0 - y;
				rs1 = rs0 - l, qs1 = qs0 - l;
				if (rs1 < 0) rs1 = 0; // not strictly necessary; better have this guard for explicit
				break;
			}
		}
	}
	if (qs > 0 && rs > 0) {
		l = qs < opt->max_gap ? qs : opt->max_gap;
		qs1 = qs1 > qs - l ? qs1 : qs - l;
		qs0 = qs0 < qs1 ? qs0 : qs1; // at least include qs0
		l += l * opt->a > opt->q ? (l * opt->a - opt->q) / opt->e : 0;
		l = l < opt->max_gap ? l : opt->max_gap;
		l = l < rs ? l : rs;
		rs1 = rs1 > rs - l ? rs1 : rs - l;
		rs0 = rs0 < rs1 ? rs0 : rs1;
		rs0 = rs0 < rs ? rs0 : rs;
	} else rs0 = rs, qs0 = qs;
	// compute re0 and qe0
	re0 = (int32_t)a[r->as + r->cnt - 1].x + 1;
	qe0 = (int32_t)a[r->as + r->cnt - 1].y + 1;
	re1 = mi->seq[rid].len, qe1 = qlen;
	for (i = r->as + r->cnt, l = 0; i < n_a && a[i].x >> 32 == a[r->as].x >> 32; ++i) { // inspect nearby seeds
		int32_t x = (int32_t)a[i].x + 1;
		int32_t y = (int32_t)a[i].y + 1;
		if (x > re0 && y > qe0) {
			if (++l > opt->min_cnt) {
				l = x - re0 > y - qe0 ? x - re0 : y - qe0;
				re1 = re0 + l, qe1 = qe0 + l;
				break;
			}
		}
	}
	if (qe < qlen && re < (int32_t)mi->seq[rid].len) {
		l = qlen - qe < opt->max_gap ? qlen - qe : opt->max_gap;
		qe1 = qe1 < qe + l ? qe1 : qe + l;
		qe0 = qe0 > qe1 ? qe0 : qe1; // at least include qe0
		l += l * opt->a > opt->q ? (l * opt->a - opt->q) / opt->e : 0;
		l = l < opt->max_gap ? l : opt->max_gap;
		l = l < (int32_t)mi->seq[rid].len - re ? l : mi->seq[rid].len - re;
		re1 = re1 < re + l ? re1 : re + l;
		re0 = re0 > re1 ? re0 : re1;
	} else re0 = re, qe0 = qe;
	if (a[r->as].y & MM_SEED_SELF) {
		int max_ext = r->qs > r->rs ? r->qs - r->rs : r->rs - r->qs;
		if (r->rs - rs0 > max_ext) rs0 = r->rs - max_ext;
		if (r->qs - qs0 > max_ext) qs0 = r->qs - max_ext;
		max_ext = r->qe > r->re ? r->qe - r->re : r->re - r->qe;
		if (re0 - r->re > max_ext) re0 = r->re + max_ext;
		if (qe0 - r->qe > max_ext) qe0 = r->qe + max_ext;
	}

	assert(re0 > rs0);
	std::vector<uint8_t> tseq_v(re0 - rs0 + 16);
	uint8_t *tseq = tseq_v.data();

	if (qs > 0 && rs > 0) { // left extension; probably the condition can be changed to "qs > qs0 && rs > rs0"
		qseq = &qseq0[rev][qs0];
		mm_idx_getseq(mi, rid, rs0, rs, tseq);
		mm_seq_rev(qs - qs0, qseq);
		mm_seq_rev(rs - rs0, tseq);
		mm_align_pair(opt, qs - qs0, qseq, rs - rs0, tseq, mat, bw, opt->end_bonus, r->split_inv ? opt->zdrop_inv : opt->zdrop, extra_flag | KSW_EZ_EXTZ_ONLY | KSW_EZ_RIGHT | KSW_EZ_REV_CIGAR, ez, st);
		if (!ez->cigar.empty()) {
			mm_append_cigar(r, ez->cigar);
			r->p->dp_score += ez->max;
		}
		rs1 = rs - (ez->reach_end ? ez->mqe_t + 1 : ez->max_t + 1);
		qs1 = qs - (ez->reach_end ? qs - qs0 : ez->max_q + 1);
		mm_seq_rev(qs - qs0, qseq);
	} else rs1 = rs, qs1 = qs;
	re1 = rs, qe1 = qs;
	assert(qs1 >= 0 && rs1 >= 0);

	for (i = 1; i < cnt1; ++i) { // gap filling
		if ((a[as1 + i].y & (MM_SEED_IGNORE | MM_SEED_TANDEM)) && i != cnt1 - 1) continue;
		mm_adjust_minier(mi, &a[as1 + i], &re, &qe);
		re1 = re, qe1 = qe;
		if (i == cnt1 - 1 || (a[as1 + i].y & MM_SEED_LONG_JOIN) || (qe - qs >= opt->min_ksw_len && re - rs >= opt->min_ksw_len)) {
			int j, bw1 = bw_long, zdrop_code;
			if (a[as1 + i].y & MM_SEED_LONG_JOIN)
				bw1 = qe - qs > re - rs ? qe - qs : re - rs;
			// perform normal gapped alignment
			qseq = &qseq0[rev][qs];
			mm_idx_getseq(mi, rid, rs, re, tseq);
			mm_align_pair(opt, qe - qs, qseq, re - rs, tseq, mat, bw1, -1, opt->zdrop, extra_flag | KSW_EZ_APPROX_MAX, ez, st); // first pass: with approximate Z-drop
			// test Z-drop and inversion Z-drop
			if ((zdrop_code = mm_test_zdrop(opt, qseq, tseq, (uint32_t)ez->cigar.size(), ez->cigar.data(), mat)) != 0)
				mm_align_pair(opt, qe - qs, qseq, re - rs, tseq, mat, bw1, -1, zdrop_code == 2 ? opt->zdrop_inv : opt->zdrop, extra_flag, ez, st); // second pass: lift approximate
			// update CIGAR
			if (!ez->cigar.empty())
				mm_append_cigar(r, ez->cigar);
			if (ez->zdropped) { // truncated by Z-drop; TODO: sometimes Z-drop kicks in because the next seed placement is wrong. This can be fixed in principle.
				if (!r->p) {
					assert(ez->cigar.empty());
					r->p = new mm_extra_t();
					r->p->capacity = 0, r->p->dp_score = r->p->dp_max = r->p->dp_max2 = 0, r->p->n_ambi = r->p->trans_strand = 0;
				}
				for (j = i - 1; j >= 0; --j)
					if ((int32_t)a[as1 + j].x <= rs + ez->max_t)
						break;
				dropped = 1;
				if (j < 0) j = 0;
				r->p->dp_score += ez->max;
				re1 = rs + (ez->max_t + 1);
				qe1 = qs + (ez->max_q + 1);
				if (cnt1 - (j + 1) >= opt->min_cnt) {
					mm_split_reg(r, r2, as1 + j + 1 - r->as, qlen, a, !!(opt->flag & MM_F_QSTRAND));
					if (zdrop_code == 2) r2->split_inv = 1;
				}
				break;
			} else r->p->dp_score += ez->score;
			rs = re, qs = qe;
		}
	}

	if (!dropped && qe < qe0 && re < re0) { // right extension
		qseq = &qseq0[rev][qe];
		mm_idx_getseq(mi, rid, re, re0, tseq);
		mm_align_pair(opt, qe0 - qe, qseq, re0 - re, tseq, mat, bw, opt->end_bonus, opt->zdrop, extra_flag | KSW_EZ_EXTZ_ONLY, ez, st);
		if (!ez->cigar.empty()) {
			mm_append_cigar(r, ez->cigar);
			r->p->dp_score += ez->max;
		}
		re1 = re + (ez->reach_end ? ez->mqe_t + 1 : ez->max_t + 1);
		qe1 = qe + (ez->reach_end ? qe0 - qe : ez->max_q + 1);
	}
	assert(qe1 <= qlen);

	r->rs = rs1, r->re = re1;
	if (rev) r->qs = qlen - qe1, r->qe = qlen - qs1;
	else r->qs = qs1, r->qe = qe1;

	assert(re1 - rs1 <= re0 - rs0);
	if (r->p) {
		mm_idx_getseq(mi, rid, rs1, re1, tseq);
		mm_update_extra(r, &qseq0[r->rev][qs1], tseq, mat, opt->q, opt->e, !(opt->flag & MM_F_SPLICE));
	}
}

static int mm_align1_inv(const mm_mapopt_t *opt, const mm_idx_t *mi, int qlen, uint8_t *qseq0[2], const mm_reg1_t *r1, const mm_reg1_t *r2, mm_reg1_t *r_inv, ksw_extz_t *ez, mm2o_stats_t *st)
{ // NB: this doesn't work with the qstrand mode
	int tl, ql, score, ret = 0, q_off, t_off;
	uint8_t *qseq;
	int8_t mat[25];

	memset((void*)r_inv, 0, sizeof(mm_reg1_t));
	if (!(r1->split & 1) || !(r2->split & 2)) return 0;
	if (r1->id != r1->parent && r1->parent != MM_PARENT_TMP_PRI) return 0;
	if (r2->id != r2->parent && r2->parent != MM_PARENT_TMP_PRI) return 0;
	if (r1->rid != r2->rid || r1->rev != r2->rev) return 0;
	ql = r1->rev ? r1->qs - r2->qe : r2->qs - r1->qe;
	tl = r2->rs - r1->re;
	if (ql < opt->min_chain_score || ql > opt->max_gap) return 0;
	if (tl < opt->min_chain_score || tl > opt->max_gap) return 0;

	ksw_gen_simple_mat(5, mat, opt->a, opt->b, opt->sc_ambi);
	std::vector<uint8_t> tseq_v(tl + 16);
	uint8_t *tseq = tseq_v.data();
	mm_idx_getseq(mi, r1->rid, r1->re, r2->rs, tseq);
	qseq = r1->rev ? &qseq0[0][r2->qe] : &qseq0[1][qlen - r2->qs];

	mm_seq_rev(ql, qseq);
	mm_seq_rev(tl, tseq);
	score = ksw_ll_i16(ql, qseq, 5, mat, tl, tseq, opt->q, opt->e, &q_off, &t_off);
	mm_seq_rev(ql, qseq);
	mm_seq_rev(tl, tseq);
	if (score < opt->min_dp_max) return 0;
	q_off = ql - (q_off + 1), t_off = tl - (t_off + 1);
	mm_align_pair(opt, ql - q_off, qseq + q_off, tl - t_off, tseq + t_off, mat, (int)(opt->bw * 1.5), -1, opt->zdrop, KSW_EZ_EXTZ_ONLY, ez, st);
	if (ez->cigar.empty()) return 0; // should never be here
	mm_append_cigar(r_inv, ez->cigar);
	r_inv->p->dp_score = ez->max;
	r_inv->id = -1;
	r_inv->parent = MM_PARENT_UNSET;
	r_inv->inv = 1;
	r_inv->rev = !r1->rev;
	r_inv->rid = r1->rid;
	r_inv->div = -1.0f;
	if (r_inv->rev == 0) {
		r_inv->qs = r2->qe + q_off;
		r_inv->qe = r_inv->qs + ez->max_q + 1;
	} else {
		r_inv->qe = r2->qs - q_off;
		r_inv->qs = r_inv->qe - (ez->max_q + 1);
	}
	r_inv->rs = r1->re + t_off;
	r_inv->re = r_inv->rs + ez->max_t + 1;
	mm_update_extra(r_inv, &qseq[q_off], &tseq[t_off], mat, opt->q, opt->e, !(opt->flag & MM_F_SPLICE));
	ret = 1;
	return ret;
}

/* hit.c: mm_squeeze_a */
static int mm_squeeze_a(int n_regs, mm_reg1_t *regs, mm128_t *a)
{ // squeeze out regions in a[] that are not referenced by regs[]
	int i, as = 0;
	std::vector<uint64_t> aux(n_regs > 0 ? n_regs : 1);
	for (i = 0; i < n_regs; ++i)
		aux[i] = (uint64_t)regs[i].as << 32 | i;
	radix_sort_64(aux.data(), aux.data() + n_regs);
	for (i = 0; i < n_regs; ++i) {
		mm_reg1_t *r = &regs[(int32_t)aux[i]];
		if (r->as != as) {
			memmove(&a[as], &a[r->as], r->cnt * 16);
			r->as = as;
		}
		as += r->cnt;
	}
	return as;
}

static inline mm_reg1_t *mm_insert_reg(const mm_reg1_t *r, int i, int *n_regs, mm_reg1_t *regs)
{
	regs = (mm_reg1_t*)realloc((void*)regs, (*n_regs + 1) * sizeof(mm_reg1_t));
	if (i + 1 != *n_regs)
		memmove((void*)&regs[i + 2], (void*)&regs[i + 1], sizeof(mm_reg1_t) * (*n_regs - i - 1));
	regs[i + 1] = *r;
	++*n_regs;
	return regs;
}

mm_reg1_t *mm_align_skeleton(const mm_mapopt_t *opt, const mm_idx_t *mi, int qlen, const char *qstr, int *n_regs_, mm_reg1_t *regs, mm128_t *a, mm2o_stats_t *st)
{
	int32_t i, n_regs = *n_regs_, n_a;
	uint8_t *qseq0[2];
	ksw_extz_t ez;

	// encode the query sequence
	std::vector<uint8_t> qbuf(qlen * 2 + 16);
	qseq0[0] = qbuf.data();
	qseq0[1] = qseq0[0] + qlen;
	for (i = 0; i < qlen; ++i) {
		qseq0[0][i] = seq_nt4_table[(uint8_t)qstr[i]];
		qseq0[1][qlen - 1 - i] = qseq0[0][i] < 4 ? 3 - qseq0[0][i] : 4;
	}

	// align through seed hits
	n_a = mm_squeeze_a(n_regs, regs, a);
	ksw_reset_extz(&ez);
	for (i = 0; i < n_regs; ++i) {
		mm_reg1_t r2;
		memset((void*)&r2, 0, sizeof(r2));
		mm_align1(opt, mi, qlen, qseq0, &regs[i], &r2, n_a, a, &ez, st);
		if (r2.cnt > 0) regs = mm_insert_reg(&r2, i, &n_regs, regs);
		if (i > 0 && regs[i].split_inv && !(opt->flag & MM_F_NO_INV)) {
			if (mm_align1_inv(opt, mi, qlen, qseq0, &regs[i - 1], &regs[i], &r2, &ez, st)) {
				regs = mm_insert_reg(&r2, i, &n_regs, regs);
				++i; // skip the inserted INV alignment
			}
		}
	}
	*n_regs_ = n_regs;
	mm_filter_regs(opt, qlen, n_regs_, regs);
	mm_hit_sort(n_regs_, regs, opt->alt_drop);
	return regs;
}

/********** format.c: cs / MD **********/

static void get_aligned_seqs(const mm_idx_t *mi, const mm_reg1_t *r, const char *seq, std::vector<uint8_t> &tseq, std::vector<uint8_t> &qseq)
{
	int i, q_len = r->qe - r->qs, t_len = r->re - r->rs;
	tseq.assign(t_len + 1, 0), qseq.assign(q_len + 1, 0);
	mm_idx_getseq(mi, r->rid, r->rs, r->re, tseq.data());
	if (!r->rev) {
		for (i = r->qs; i < r->qe; ++i)
			qseq[i - r->qs] = seq_nt4_table[(uint8_t)seq[i]];
	} else {
		for (i = r->qs; i < r->qe; ++i) {
			uint8_t c = seq_nt4_table[(uint8_t)seq[i]];
			qseq[r->qe - i - 1] = c >= 4 ? 4 : 3 - c;
		}
	}
}

std::string mm_gen_cs(const mm_idx_t *mi, const mm_reg1_t *r, const char *seq, int no_iden)
{
	std::string s;
	if (r->p == 0) return s;
	std::vector<uint8_t> tseq, qseq;
	get_aligned_seqs(mi, r, seq, tseq, qseq);
	int q_off = 0, t_off = 0;
	for (size_t i = 0; i < r->p->cigar.size(); ++i) {
		int j, op = r->p->cigar[i] & 0xf, len = r->p->cigar[i] >> 4;
		if (op == MM_CIGAR_MATCH || op == MM_CIGAR_EQ_MATCH || op == MM_CIGAR_X_MISMATCH) {
			std::string tmp;
			for (j = 0; j < len; ++j) {
				if (qseq[q_off + j] != tseq[t_off + j]) {
					if (!tmp.empty()) {
						if (!no_iden) s += "=" + tmp;
						else s += ":" + std::to_string(tmp.size());
						tmp.clear();
					}
					s += '*';
					s += "acgtn"[tseq[t_off + j]];
					s += "acgtn"[qseq[q_off + j]];
				} else tmp += "ACGTN"[qseq[q_off + j]];
			}
			if (!tmp.empty()) {
				if (!no_iden) s += "=" + tmp;
				else s += ":" + std::to_string(tmp.size());
			}
			q_off += len, t_off += len;
		} else if (op == MM_CIGAR_INS) {
			s += '+';
			for (j = 0; j < len; ++j) s += "acgtn"[qseq[q_off + j]];
			q_off += len;
		} else if (op == MM_CIGAR_DEL) {
			s += '-';
			for (j = 0; j < len; ++j) s += "acgtn"[tseq[t_off + j]];
			t_off += len;
		}
	}
	return s;
}

std::string mm_gen_MD(const mm_idx_t *mi, const mm_reg1_t *r, const char *seq)
{
	std::string s;
	if (r->p == 0) return s;
	std::vector<uint8_t> tseq, qseq;
	get_aligned_seqs(mi, r, seq, tseq, qseq);
	int q_off = 0, t_off = 0, l_MD = 0;
	for (size_t i = 0; i < r->p->cigar.size(); ++i) {
		int j, op = r->p->cigar[i] & 0xf, len = r->p->cigar[i] >> 4;
		if (op == MM_CIGAR_MATCH || op == MM_CIGAR_EQ_MATCH || op == MM_CIGAR_X_MISMATCH) {
			for (j = 0; j < len; ++j) {
				if (qseq[q_off + j] != tseq[t_off + j]) {
					s += std::to_string(l_MD);
					s += "ACGTN"[tseq[t_off + j]];
					l_MD = 0;
				} else ++l_MD;
			}
			q_off += len, t_off += len;
		} else if (op == MM_CIGAR_INS) {
			q_off += len;
		} else if (op == MM_CIGAR_DEL) {
			s += std::to_string(l_MD);
			s += '^';
			for (j = 0; j < len; ++j) s += "ACGTN"[tseq[t_off + j]];
			l_MD = 0;
			t_off += len;
		} else if (op == MM_CIGAR_N_SKIP) {
			t_off += len;
		}
	}
	if (l_MD > 0) s += std::to_string(l_MD);
	return s;
}
