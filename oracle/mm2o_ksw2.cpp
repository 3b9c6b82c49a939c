/* mm2o_ksw2.cpp -- ORACLE (test infrastructure only).
 * Restates minimap2 v2.26 ksw2.h (ksw_reset_extz, ksw_apply_zdrop, ksw_push_cigar,
 * ksw_backtrack), ksw2_extd2_sse.c (ksw_extd2_sse: dual affine gap, the kernel
 * mm_align_pair picks when q != q2 || e != e2, i.e. map-ont and map-hifi) and
 * ksw2_ll_sse.c (ksw_ll_qinit + ksw_ll_i16, the local score used by the
 * inversion test).  Reached from /root/reference/src/lib.rs:482,587 via
 * mm_map -> align_regs -> mm_align_skeleton because mappy-rs forces MM_F_CIGAR
 * (src/lib.rs:339).
 *
 * ksw_extd2_sse is restated at the granularity of its 16-lane SSE blocks with
 * wrapping int8 arithmetic: upstream evaluates whole blocks, so cells just
 * outside the band hold well-defined (if meaningless) values that band-edge
 * cells read on later anti-diagonals.  A "clean" scalar DP would differ there.
 * Memory beyond the kcalloc'ed block is taken as zero.
 * Parity unpinned by the reference except `map_one` (src/lib.rs:1094-1106).
 */
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include <vector>
#include <emmintrin.h>
#include <smmintrin.h>
#include "mm2o_ksw2.h"

/* 1 = run the byte-by-byte restatement of the SSE blocks instead of the intrinsics (tests compare the two) */
int mm2o_ksw_force_scalar = 0;

void ksw_reset_extz(ksw_extz_t *ez)
{
	ez->max_q = ez->max_t = ez->mqe_t = ez->mte_q = -1;
	ez->max = 0, ez->score = ez->mqe = ez->mte = KSW_NEG_INF;
	ez->cigar.clear(), ez->zdropped = 0, ez->reach_end = 0;
}

static inline int ksw_apply_zdrop(ksw_extz_t *ez, int is_rot, int32_t H, int a, int b, int zdrop, int8_t e)
{
	int r, t;
	if (is_rot) r = a, t = b;
	else r = a + b, t = a;
	if (H > (int32_t)ez->max) {
		ez->max = H, ez->max_t = t, ez->max_q = r - t;
	} else if (t >= ez->max_t && r - t >= ez->max_q) {
		int tl = t - ez->max_t, ql = (r - t) - ez->max_q, l;
		l = tl > ql ? tl - ql : ql - tl;
		if (zdrop >= 0 && (int32_t)ez->max - H > zdrop + l * e) {
			ez->zdropped = 1;
			return 1;
		}
	}
	return 0;
}

static inline void ksw_push_cigar(std::vector<uint32_t> &cigar, uint32_t op, int len)
{
	if (cigar.empty() || op != (cigar.back() & 0xf)) cigar.push_back((uint32_t)len << 4 | op);
	else cigar.back() += (uint32_t)len << 4;
}

/* ksw2.h: ksw_backtrack with is_rot = 1, min_intron_len = 0 */
static void ksw_backtrack_rot(int is_rev, const uint8_t *p, const int *off, const int *off_end, int n_col, int i0, int j0, std::vector<uint32_t> &cigar)
{
	int i = i0, j = j0, r, state = 0;
	uint32_t tmp;
	cigar.clear();
	while (i >= 0 && j >= 0) { // at the beginning of the loop, _state_ tells us which state to check
		int force_state = -1;
		r = i + j;
		if (i < off[r]) force_state = 2;
		if (off_end && i > off_end[r]) force_state = 1;
		tmp = force_state < 0 ? p[(size_t)r * n_col + i - off[r]] : 0;
		if (state == 0) state = tmp & 7; // if requesting the H state, find state one maximizes it.
		else if (!(tmp >> (state + 2) & 1)) state = 0; // if requesting other states, _state_ stays the same if it is a continuation; otherwise, set to H
		if (state == 0) state = tmp & 7;
		if (force_state >= 0) state = force_state;
		if (state == 0) ksw_push_cigar(cigar, 0, 1), --i, --j; // match
		else if (state == 1 || state == 3) ksw_push_cigar(cigar, 2, 1), --i; // deletion
		else ksw_push_cigar(cigar, 1, 1), --j; // insertion
	}
	if (i >= 0) ksw_push_cigar(cigar, 2, i + 1); // first deletion
	if (j >= 0) ksw_push_cigar(cigar, 1, j + 1); // first insertion
	if (!is_rev)
		for (size_t k = 0; k < cigar.size() >> 1; ++k) // reverse CIGAR
			tmp = cigar[k], cigar[k] = cigar[cigar.size() - 1 - k], cigar[cigar.size() - 1 - k] = tmp;
}

#define I8(v) ((int8_t)(v))

/* The 16-lane blocks of ksw_extd2_sse are evaluated with the same SSE2/SSE4.1 instructions upstream uses
 * (what SIMDe maps to natively on x86-64, Cargo.toml:24) unless mm2o_ksw_force_scalar is set, in which case the
 * blocks are walked byte by byte with wrapping int8 arithmetic: the two must agree bit for bit
 * (tests/test_oracle_fixtures.py::test_ksw_sse_blocks_equal_scalar_restatement). */
void ksw_extd2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
               int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez, uint64_t *n_cell)
{
	const bool use_sse = !mm2o_ksw_force_scalar;
	int r, t, qe = q + e, n_col_, tlen_, qlen_, last_st, last_en, wl, wr, max_sc, min_sc, long_thres, long_diff;
	int with_cigar = !(flag & KSW_EZ_SCORE_ONLY), approx_max = !!(flag & KSW_EZ_APPROX_MAX);
	int32_t H0 = 0, last_H0_t = 0;
	int8_t sc_mch, sc_mis, sc_N, m1;

	ksw_reset_extz(ez);
	if (m <= 1 || qlen <= 0 || tlen <= 0) return;

	if (q2 + e2 < q + e) t = q, q = q2, q2 = t, t = e, e = e2, e2 = t; // make sure q+e no larger than q2+e2
	qe = q + e;
	const int qe2 = q2 + e2;
	sc_mch = mat[0], sc_mis = mat[1];
	sc_N = mat[m * m - 1] == 0 ? -e2 : mat[m * m - 1];
	m1 = m - 1; // wildcard

	if (w < 0) w = tlen > qlen ? tlen : qlen;
	wl = wr = w;
	tlen_ = (tlen + 15) / 16;
	n_col_ = qlen < tlen ? qlen : tlen;
	n_col_ = ((n_col_ < w + 1 ? n_col_ : w + 1) + 15) / 16 + 1;
	qlen_ = (qlen + 15) / 16;
	for (t = 1, max_sc = mat[0], min_sc = mat[1]; t < m * m; ++t) {
		max_sc = max_sc > mat[t] ? max_sc : mat[t];
		min_sc = min_sc < mat[t] ? min_sc : mat[t];
	}
	if (-min_sc > 2 * (q + e)) return; // otherwise, we won't see any mismatches

	long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
	if (q2 + e2 + long_thres * e2 > q + e + long_thres * e)
		++long_thres;
	long_diff = long_thres * (e - e2) - (q2 - q) - e2;

	/* one zero-initialised block holding u,v,x,y,x2,y2,s,sf,qr back to back (kcalloc) */
	const size_t T16 = (size_t)tlen_ * 16, mem_sz = ((size_t)tlen_ * 8 + qlen_ + 1) * 16;
	std::vector<int8_t> mem(mem_sz + 64, 0);
	int8_t *u = mem.data(), *v = u + T16, *x = v + T16, *y = x + T16, *x2 = y + T16, *y2 = x2 + T16, *s = y2 + T16;
	uint8_t *sf = (uint8_t*)(s + T16), *qr = sf + T16;
	memset(u, -q - e, T16);
	memset(v, -q - e, T16);
	memset(x, -q - e, T16);
	memset(y, -q - e, T16);
	memset(x2, -q2 - e2, T16);
	memset(y2, -q2 - e2, T16);
	std::vector<int32_t> H;
	if (!approx_max) H.assign(T16, KSW_NEG_INF);
	std::vector<uint8_t> p;
	std::vector<int> off, off_end;
	const int n_col = n_col_ * 16;
	if (with_cigar) {
		p.assign(((size_t)(qlen + tlen - 1) * n_col_ + 1) * 16, 0);
		off.assign(qlen + tlen - 1, 0);
		off_end.assign(qlen + tlen - 1, 0);
	}

	for (t = 0; t < qlen; ++t) qr[t] = query[qlen - 1 - t];
	memcpy(sf, target, tlen);
	const uint8_t *mem_end = (const uint8_t*)mem.data() + mem_sz;

	for (r = 0, last_st = last_en = -1; r < qlen + tlen - 1; ++r) {
		int st = 0, en = tlen - 1, st0, en0;
		int8_t x1, x21, v1;
		const uint8_t *qrr = qr + (qlen - 1 - r);
		// find the boundaries
		if (st < r - qlen + 1) st = r - qlen + 1;
		if (en > r) en = r;
		if (st < (r - wr + 1) >> 1) st = (r - wr + 1) >> 1; // take the ceil
		if (en > (r + wl) >> 1) en = (r + wl) >> 1; // take the floor
		if (st > en) {
			ez->zdropped = 1;
			break;
		}
		st0 = st, en0 = en;
		st = st / 16 * 16, en = (en + 16) / 16 * 16 - 1;
		// set boundary conditions
		if (st > 0) {
			if (st - 1 >= last_st && st - 1 <= last_en) {
				x1 = x[st - 1], x21 = x2[st - 1], v1 = v[st - 1]; // (r-1,s-1) calculated in the last round
			} else {
				x1 = -q - e, x21 = -q2 - e2;
				v1 = -q - e;
			}
		} else {
			x1 = -q - e, x21 = -q2 - e2;
			v1 = r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2;
		}
		if (en >= r) {
			y[r] = -q - e, y2[r] = -q2 - e2;
			u[r] = r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2;
		}
		// loop fission: set scores first (16 lanes at a time from st0; lanes past en0 are written too)
		if (use_sse && !(flag & KSW_EZ_GENERIC_SC)) {
			const __m128i m1_ = _mm_set1_epi8(m1), sc_mch_ = _mm_set1_epi8(sc_mch), sc_mis_ = _mm_set1_epi8(sc_mis), sc_N_ = _mm_set1_epi8(sc_N);
			for (t = st0; t <= en0; t += 16) {
				__m128i sq = _mm_loadu_si128((const __m128i*)&sf[t]), st_ = _mm_loadu_si128((const __m128i*)&qrr[t]);
				__m128i mask = _mm_or_si128(_mm_cmpeq_epi8(sq, m1_), _mm_cmpeq_epi8(st_, m1_));
				__m128i tmp = _mm_cmpeq_epi8(sq, st_);
				tmp = _mm_blendv_epi8(sc_mis_, sc_mch_, tmp);
				tmp = _mm_blendv_epi8(tmp, sc_N_, mask);
				_mm_storeu_si128((__m128i*)((int8_t*)s + t), tmp);
			}
		} else
		for (t = st0; t <= en0; t += 16) {
			for (int k = 0; k < 16; ++k) {
				const uint8_t *psq = &sf[t + k], *pst = &qrr[t + k];
				uint8_t sq = psq < mem_end ? *psq : 0, sq2 = pst < mem_end ? *pst : 0;
				int8_t tmp;
				if (!(flag & KSW_EZ_GENERIC_SC)) {
					int mask = (sq == (uint8_t)m1) || (sq2 == (uint8_t)m1);
					tmp = sq == sq2 ? sc_mch : sc_mis;
					if (mask) tmp = sc_N;
				} else tmp = mat[(sq < m ? sq : m - 1) * m + (sq2 < m ? sq2 : m - 1)];
				if ((uint8_t*)(s + t + k) < (uint8_t*)sf) s[t + k] = tmp; // s is followed by sf in memory; st0+16k never overruns it upstream
				else s[t + k] = tmp;
			}
		}
		// core loop
		if (with_cigar) {
			off[r] = st, off_end[r] = en;
			uint8_t *pr = p.data() + (size_t)r * n_col - st;
			int8_t px = x1, px2 = x21, pv = v1;
			if (use_sse) {
				const __m128i q_ = _mm_set1_epi8(q), q2_ = _mm_set1_epi8(q2), qe_ = _mm_set1_epi8(qe), qe2_ = _mm_set1_epi8(qe2), zero_ = _mm_setzero_si128();
				const __m128i sc_mch_ = _mm_set1_epi8(sc_mch);
				__m128i x1_ = _mm_cvtsi32_si128((uint8_t)x1), x21_ = _mm_cvtsi32_si128((uint8_t)x21), v1_ = _mm_cvtsi32_si128((uint8_t)v1);
				const bool right = (flag & KSW_EZ_RIGHT) != 0;
				for (t = st; t <= en; t += 16) {
					__m128i d, z, a, b, a2, b2, xt1, x2t1, vt1, ut, tmp;
					z = _mm_loadu_si128((const __m128i*)(s + t));
					xt1 = _mm_loadu_si128((const __m128i*)(x + t));
					tmp = _mm_srli_si128(xt1, 15);
					xt1 = _mm_or_si128(_mm_slli_si128(xt1, 1), x1_);
					x1_ = tmp;
					vt1 = _mm_loadu_si128((const __m128i*)(v + t));
					tmp = _mm_srli_si128(vt1, 15);
					vt1 = _mm_or_si128(_mm_slli_si128(vt1, 1), v1_);
					v1_ = tmp;
					a = _mm_add_epi8(xt1, vt1);
					ut = _mm_loadu_si128((const __m128i*)(u + t));
					b = _mm_add_epi8(_mm_loadu_si128((const __m128i*)(y + t)), ut);
					x2t1 = _mm_loadu_si128((const __m128i*)(x2 + t));
					tmp = _mm_srli_si128(x2t1, 15);
					x2t1 = _mm_or_si128(_mm_slli_si128(x2t1, 1), x21_);
					x21_ = tmp;
					a2 = _mm_add_epi8(x2t1, vt1);
					b2 = _mm_add_epi8(_mm_loadu_si128((const __m128i*)(y2 + t)), ut);
					if (!right) {
						d = _mm_and_si128(_mm_cmpgt_epi8(a, z), _mm_set1_epi8(1));
						z = _mm_max_epi8(z, a);
						d = _mm_blendv_epi8(d, _mm_set1_epi8(2), _mm_cmpgt_epi8(b, z));
						z = _mm_max_epi8(z, b);
						d = _mm_blendv_epi8(d, _mm_set1_epi8(3), _mm_cmpgt_epi8(a2, z));
						z = _mm_max_epi8(z, a2);
						d = _mm_blendv_epi8(d, _mm_set1_epi8(4), _mm_cmpgt_epi8(b2, z));
						z = _mm_max_epi8(z, b2);
					} else {
						d = _mm_andnot_si128(_mm_cmpgt_epi8(z, a), _mm_set1_epi8(1));
						z = _mm_max_epi8(z, a);
						d = _mm_blendv_epi8(_mm_set1_epi8(2), d, _mm_cmpgt_epi8(z, b));
						z = _mm_max_epi8(z, b);
						d = _mm_blendv_epi8(_mm_set1_epi8(3), d, _mm_cmpgt_epi8(z, a2));
						z = _mm_max_epi8(z, a2);
						d = _mm_blendv_epi8(_mm_set1_epi8(4), d, _mm_cmpgt_epi8(z, b2));
						z = _mm_max_epi8(z, b2);
					}
					z = _mm_min_epi8(z, sc_mch_);
					_mm_storeu_si128((__m128i*)(u + t), _mm_sub_epi8(z, vt1));
					_mm_storeu_si128((__m128i*)(v + t), _mm_sub_epi8(z, ut));
					tmp = _mm_sub_epi8(z, q_);
					a = _mm_sub_epi8(a, tmp), b = _mm_sub_epi8(b, tmp);
					tmp = _mm_sub_epi8(z, q2_);
					a2 = _mm_sub_epi8(a2, tmp), b2 = _mm_sub_epi8(b2, tmp);
					if (!right) {
						tmp = _mm_cmpgt_epi8(a, zero_);
						_mm_storeu_si128((__m128i*)(x + t), _mm_sub_epi8(_mm_and_si128(tmp, a), qe_));
						d = _mm_or_si128(d, _mm_and_si128(tmp, _mm_set1_epi8(0x08)));
						tmp = _mm_cmpgt_epi8(b, zero_);
						_mm_storeu_si128((__m128i*)(y + t), _mm_sub_epi8(_mm_and_si128(tmp, b), qe_));
						d = _mm_or_si128(d, _mm_and_si128(tmp, _mm_set1_epi8(0x10)));
						tmp = _mm_cmpgt_epi8(a2, zero_);
						_mm_storeu_si128((__m128i*)(x2 + t), _mm_sub_epi8(_mm_and_si128(tmp, a2), qe2_));
						d = _mm_or_si128(d, _mm_and_si128(tmp, _mm_set1_epi8(0x20)));
						tmp = _mm_cmpgt_epi8(b2, zero_);
						_mm_storeu_si128((__m128i*)(y2 + t), _mm_sub_epi8(_mm_and_si128(tmp, b2), qe2_));
						d = _mm_or_si128(d, _mm_and_si128(tmp, _mm_set1_epi8(0x40)));
					} else {
						tmp = _mm_cmpgt_epi8(zero_, a);
						_mm_storeu_si128((__m128i*)(x + t), _mm_sub_epi8(_mm_andnot_si128(tmp, a), qe_));
						d = _mm_or_si128(d, _mm_andnot_si128(tmp, _mm_set1_epi8(0x08)));
						tmp = _mm_cmpgt_epi8(zero_, b);
						_mm_storeu_si128((__m128i*)(y + t), _mm_sub_epi8(_mm_andnot_si128(tmp, b), qe_));
						d = _mm_or_si128(d, _mm_andnot_si128(tmp, _mm_set1_epi8(0x10)));
						tmp = _mm_cmpgt_epi8(zero_, a2);
						_mm_storeu_si128((__m128i*)(x2 + t), _mm_sub_epi8(_mm_andnot_si128(tmp, a2), qe2_));
						d = _mm_or_si128(d, _mm_andnot_si128(tmp, _mm_set1_epi8(0x20)));
						tmp = _mm_cmpgt_epi8(zero_, b2);
						_mm_storeu_si128((__m128i*)(y2 + t), _mm_sub_epi8(_mm_andnot_si128(tmp, b2), qe2_));
						d = _mm_or_si128(d, _mm_andnot_si128(tmp, _mm_set1_epi8(0x40)));
					}
					_mm_storeu_si128((__m128i*)(pr + t), d);
				}
			} else
			for (t = st; t <= en; ++t) {
				int8_t z, a, b, a2, b2, xt1, x2t1, vt1, ut, tmp;
				uint8_t d;
				z = s[t];
				xt1 = px, px = x[t];     // xt1 <- x[r-1][t-1]
				vt1 = pv, pv = v[t];     // vt1 <- v[r-1][t-1]
				x2t1 = px2, px2 = x2[t];
				a = I8(xt1 + vt1);
				ut = u[t];
				b = I8(y[t] + ut);
				a2 = I8(x2t1 + vt1);
				b2 = I8(y2[t] + ut);
				if (!(flag & KSW_EZ_RIGHT)) { // gap left-alignment
					d = a > z ? 1 : 0;
					z = z > a ? z : a;
					d = b > z ? 2 : d;
					z = z > b ? z : b;
					d = a2 > z ? 3 : d;
					z = z > a2 ? z : a2;
					d = b2 > z ? 4 : d;
					z = z > b2 ? z : b2;
					z = z < sc_mch ? z : sc_mch;
					u[t] = I8(z - vt1);
					v[t] = I8(z - ut);
					tmp = I8(z - q);
					a = I8(a - tmp), b = I8(b - tmp);
					tmp = I8(z - q2);
					a2 = I8(a2 - tmp), b2 = I8(b2 - tmp);
					x[t] = I8((a > 0 ? a : 0) - qe);
					d |= a > 0 ? 0x08 : 0;
					y[t] = I8((b > 0 ? b : 0) - qe);
					d |= b > 0 ? 0x10 : 0;
					x2[t] = I8((a2 > 0 ? a2 : 0) - qe2);
					d |= a2 > 0 ? 0x20 : 0;
					y2[t] = I8((b2 > 0 ? b2 : 0) - qe2);
					d |= b2 > 0 ? 0x40 : 0;
				} else { // gap right-alignment
					d = z > a ? 0 : 1;
					z = z > a ? z : a;
					d = z > b ? d : 2;
					z = z > b ? z : b;
					d = z > a2 ? d : 3;
					z = z > a2 ? z : a2;
					d = z > b2 ? d : 4;
					z = z > b2 ? z : b2;
					z = z < sc_mch ? z : sc_mch;
					u[t] = I8(z - vt1);
					v[t] = I8(z - ut);
					tmp = I8(z - q);
					a = I8(a - tmp), b = I8(b - tmp);
					tmp = I8(z - q2);
					a2 = I8(a2 - tmp), b2 = I8(b2 - tmp);
					x[t] = I8((0 > a ? 0 : a) - qe);
					d |= 0 > a ? 0 : 0x08;
					y[t] = I8((0 > b ? 0 : b) - qe);
					d |= 0 > b ? 0 : 0x10;
					x2[t] = I8((0 > a2 ? 0 : a2) - qe2);
					d |= 0 > a2 ? 0 : 0x20;
					y2[t] = I8((0 > b2 ? 0 : b2) - qe2);
					d |= 0 > b2 ? 0 : 0x40;
				}
				pr[t] = d;
			}
		} else { // score only
			int8_t px = x1, px2 = x21, pv = v1;
			for (t = st; t <= en; ++t) {
				int8_t z, a, b, a2, b2, xt1, x2t1, vt1, ut, tmp;
				z = s[t];
				xt1 = px, px = x[t];
				vt1 = pv, pv = v[t];
				x2t1 = px2, px2 = x2[t];
				a = I8(xt1 + vt1);
				ut = u[t];
				b = I8(y[t] + ut);
				a2 = I8(x2t1 + vt1);
				b2 = I8(y2[t] + ut);
				z = z > a ? z : a;
				z = z > b ? z : b;
				z = z > a2 ? z : a2;
				z = z > b2 ? z : b2;
				z = z < sc_mch ? z : sc_mch;
				u[t] = I8(z - vt1);
				v[t] = I8(z - ut);
				tmp = I8(z - q);
				a = I8(a - tmp), b = I8(b - tmp);
				tmp = I8(z - q2);
				a2 = I8(a2 - tmp), b2 = I8(b2 - tmp);
				x[t] = I8((a > 0 ? a : 0) - qe);
				y[t] = I8((b > 0 ? b : 0) - qe);
				x2[t] = I8((a2 > 0 ? a2 : 0) - qe2);
				y2[t] = I8((b2 > 0 ? b2 : 0) - qe2);
			}
		}
		if (n_cell) *n_cell += en0 - st0 + 1;
		if (!approx_max) { // find the exact max with a 32-bit score array
			int32_t max_H, max_t;
			// compute H[], max_H and max_t
			if (r > 0) {
				int32_t HH[4], tt[4], en1 = st0 + (en0 - st0) / 4 * 4, i;
				max_H = H[en0] = en0 > 0 ? H[en0 - 1] + u[en0] : H[en0] + v[en0]; // special casing the last element
				max_t = en0;
				for (i = 0; i < 4; ++i) HH[i] = max_H, tt[i] = max_t;
				if (use_sse) {
					__m128i max_H_ = _mm_set1_epi32(max_H), max_t_ = _mm_set1_epi32(max_t);
					for (t = st0; t < en1; t += 4) {
						__m128i H1 = _mm_loadu_si128((const __m128i*)&H[t]);
						int32_t v4;
						memcpy(&v4, v + t, 4);
						H1 = _mm_add_epi32(H1, _mm_cvtepi8_epi32(_mm_cvtsi32_si128(v4)));
						_mm_storeu_si128((__m128i*)&H[t], H1);
						const __m128i tmp = _mm_cmpgt_epi32(H1, max_H_);
						max_H_ = _mm_blendv_epi8(max_H_, H1, tmp);
						max_t_ = _mm_blendv_epi8(max_t_, _mm_set1_epi32(t), tmp);
					}
					_mm_storeu_si128((__m128i*)HH, max_H_);
					_mm_storeu_si128((__m128i*)tt, max_t_);
				} else
				for (t = st0; t < en1; t += 4) { // this implements: H[t]+=v8[t]-qe; if(H[t]>max_H) max_H=H[t],max_t=t;
					for (i = 0; i < 4; ++i) {
						H[t + i] += (int32_t)v[t + i];
						if (H[t + i] > HH[i]) HH[i] = H[t + i], tt[i] = t;
					}
				}
				for (i = 0; i < 4; ++i)
					if (max_H < HH[i]) max_H = HH[i], max_t = tt[i] + i;
				for (; t < en0; ++t) { // for the rest of values that haven't been computed with SSE
					H[t] += (int32_t)v[t];
					if (H[t] > max_H)
						max_H = H[t], max_t = t;
				}
			} else H[0] = v[0] - qe, max_H = H[0], max_t = 0; // special casing r==0
			// update ez
			if (en0 == tlen - 1 && H[en0] > ez->mte)
				ez->mte = H[en0], ez->mte_q = r - en0;
			if (r - st0 == qlen - 1 && H[st0] > ez->mqe)
				ez->mqe = H[st0], ez->mqe_t = st0;
			if (ksw_apply_zdrop(ez, 1, max_H, r, max_t, zdrop, e2)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1)
				ez->score = H[tlen - 1];
		} else { // find approximate max; Z-drop might be inaccurate, too.
			if (r > 0) {
				if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
					int32_t d0 = v[last_H0_t];
					int32_t d1 = u[last_H0_t + 1];
					if (d0 > d1) H0 += d0;
					else H0 += d1, ++last_H0_t;
				} else if (last_H0_t >= st0 && last_H0_t <= en0) {
					H0 += v[last_H0_t];
				} else {
					++last_H0_t, H0 += u[last_H0_t];
				}
			} else H0 = v[0] - qe, last_H0_t = 0;
			if ((flag & KSW_EZ_APPROX_DROP) && ksw_apply_zdrop(ez, 1, H0, r, last_H0_t, zdrop, e2)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1)
				ez->score = H0;
		}
		last_st = st, last_en = en;
	}
	if (with_cigar) { // backtrack
		int rev_cigar = !!(flag & KSW_EZ_REV_CIGAR);
		if (!ez->zdropped && !(flag & KSW_EZ_EXTZ_ONLY)) {
			ksw_backtrack_rot(rev_cigar, p.data(), off.data(), off_end.data(), n_col, tlen - 1, qlen - 1, ez->cigar);
		} else if (!ez->zdropped && (flag & KSW_EZ_EXTZ_ONLY) && ez->mqe + end_bonus > (int)ez->max) {
			ez->reach_end = 1;
			ksw_backtrack_rot(rev_cigar, p.data(), off.data(), off_end.data(), n_col, ez->mqe_t, qlen - 1, ez->cigar);
		} else if (ez->max_t >= 0 && ez->max_q >= 0) {
			ksw_backtrack_rot(rev_cigar, p.data(), off.data(), off_end.data(), n_col, ez->max_t, ez->max_q, ez->cigar);
		}
	}
}

/* ksw2_extz2_sse.c: ksw_extz2_sse -- the single-affine kernel mm_align_pair dispatches to when q == q2 and
 * e == e2, which is what a 4-tuple `scoring` gives (/root/reference/src/lib.rs:369-376 sets q2 = q, e2 = e).
 * Upstream keeps u, v, x, y SHIFTED by q+e (z by 2(q+e)) so that every value is a non-negative byte and unsigned
 * max/min apply; restated here with the same intrinsics.  The device runs its dual-gap kernel with equal gap
 * pairs for this case; tests/test_emu_parity.py and tests/test_gpu_parity.py check that the two agree. */
void ksw_extz2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m, const int8_t *mat,
               int8_t q, int8_t e, int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez, uint64_t *n_cell)
{
	int r, t, qe = q + e, n_col_, tlen_, qlen_, last_st, last_en, wl, wr, max_sc, min_sc;
	const int with_cigar = !(flag & KSW_EZ_SCORE_ONLY), approx_max = !!(flag & KSW_EZ_APPROX_MAX);
	int32_t H0 = 0, last_H0_t = 0;

	ksw_reset_extz(ez);
	if (m <= 0 || qlen <= 0 || tlen <= 0) return;
	const __m128i zero_ = _mm_set1_epi8(0), q_ = _mm_set1_epi8(q), qe2_ = _mm_set1_epi8((q + e) * 2);
	const __m128i flag1_ = _mm_set1_epi8(1), flag2_ = _mm_set1_epi8(2), flag8_ = _mm_set1_epi8(0x08), flag16_ = _mm_set1_epi8(0x10);
	const __m128i sc_mch_ = _mm_set1_epi8(mat[0]), sc_mis_ = _mm_set1_epi8(mat[1]);
	const __m128i sc_N_ = mat[m * m - 1] == 0 ? _mm_set1_epi8(-e) : _mm_set1_epi8(mat[m * m - 1]);
	const __m128i m1_ = _mm_set1_epi8(m - 1), max_sc_ = _mm_set1_epi8(mat[0] + (q + e) * 2);

	if (w < 0) w = tlen > qlen ? tlen : qlen;
	wl = wr = w;
	tlen_ = (tlen + 15) / 16;
	n_col_ = qlen < tlen ? qlen : tlen;
	n_col_ = ((n_col_ < w + 1 ? n_col_ : w + 1) + 15) / 16 + 1;
	qlen_ = (qlen + 15) / 16;
	for (t = 1, max_sc = mat[0], min_sc = mat[1]; t < m * m; ++t) {
		max_sc = max_sc > mat[t] ? max_sc : mat[t];
		min_sc = min_sc < mat[t] ? min_sc : mat[t];
	}
	if (-min_sc > 2 * (q + e)) return; // otherwise, we won't see any mismatches

	/* kcalloc(tlen_ * 6 + qlen_ + 1, 16): u, v, x, y, s, sf, qr back to back, zero = the shifted form of -q-e */
	const size_t T16 = (size_t)tlen_ * 16, mem_sz = ((size_t)tlen_ * 6 + qlen_ + 1) * 16;
	std::vector<uint8_t> mem(mem_sz + 64, 0);
	uint8_t *u8 = mem.data(), *v8 = u8 + T16, *x8 = v8 + T16, *y8 = x8 + T16, *s8 = y8 + T16, *sf = s8 + T16, *qr = sf + T16;
	std::vector<int32_t> H;
	if (!approx_max) H.assign(T16, KSW_NEG_INF);
	std::vector<uint8_t> p;
	std::vector<int> off, off_end;
	const int n_col = n_col_ * 16;
	if (with_cigar) {
		p.assign(((size_t)(qlen + tlen - 1) * n_col_ + 1) * 16, 0);
		off.assign(qlen + tlen - 1, 0);
		off_end.assign(qlen + tlen - 1, 0);
	}
	for (t = 0; t < qlen; ++t) qr[t] = query[qlen - 1 - t];
	memcpy(sf, target, tlen);

	for (r = 0, last_st = last_en = -1; r < qlen + tlen - 1; ++r) {
		int st = 0, en = tlen - 1, st0, en0;
		uint8_t x1, v1;
		const uint8_t *qrr = qr + (qlen - 1 - r);
		if (st < r - qlen + 1) st = r - qlen + 1;
		if (en > r) en = r;
		if (st < (r - wr + 1) >> 1) st = (r - wr + 1) >> 1; // take the ceil
		if (en > (r + wl) >> 1) en = (r + wl) >> 1; // take the floor
		if (st > en) {
			ez->zdropped = 1;
			break;
		}
		st0 = st, en0 = en;
		st = st / 16 * 16, en = (en + 16) / 16 * 16 - 1;
		// set boundary conditions
		if (st > 0) {
			if (st - 1 >= last_st && st - 1 <= last_en)
				x1 = x8[st - 1], v1 = v8[st - 1]; // (r-1,s-1) calculated in the last round
			else x1 = v1 = 0; // not calculated; set to zeros
		} else x1 = 0, v1 = r ? q : 0;
		if (en >= r) y8[r] = 0, u8[r] = r ? q : 0;
		// loop fission: set scores first
		for (t = st0; t <= en0; t += 16) {
			__m128i sq = _mm_loadu_si128((const __m128i*)&sf[t]), st_ = _mm_loadu_si128((const __m128i*)&qrr[t]);
			__m128i mask = _mm_or_si128(_mm_cmpeq_epi8(sq, m1_), _mm_cmpeq_epi8(st_, m1_));
			__m128i tmp = _mm_cmpeq_epi8(sq, st_);
			tmp = _mm_blendv_epi8(sc_mis_, sc_mch_, tmp);
			tmp = _mm_blendv_epi8(tmp, sc_N_, mask);
			_mm_storeu_si128((__m128i*)(s8 + t), tmp);
		}
		// core loop
		__m128i x1_ = _mm_cvtsi32_si128(x1), v1_ = _mm_cvtsi32_si128(v1);
		uint8_t *pr = with_cigar ? p.data() + (size_t)r * n_col - st : 0;
		if (with_cigar) off[r] = st, off_end[r] = en;
		const bool right = (flag & KSW_EZ_RIGHT) != 0;
		for (t = st; t <= en; t += 16) {
			__m128i d = zero_, z, a, b, xt1, vt1, ut, tmp;
			z = _mm_add_epi8(_mm_loadu_si128((const __m128i*)(s8 + t)), qe2_);
			xt1 = _mm_loadu_si128((const __m128i*)(x8 + t));
			tmp = _mm_srli_si128(xt1, 15);
			xt1 = _mm_or_si128(_mm_slli_si128(xt1, 1), x1_);
			x1_ = tmp;
			vt1 = _mm_loadu_si128((const __m128i*)(v8 + t));
			tmp = _mm_srli_si128(vt1, 15);
			vt1 = _mm_or_si128(_mm_slli_si128(vt1, 1), v1_);
			v1_ = tmp;
			a = _mm_add_epi8(xt1, vt1);
			ut = _mm_loadu_si128((const __m128i*)(u8 + t));
			b = _mm_add_epi8(_mm_loadu_si128((const __m128i*)(y8 + t)), ut);
			if (!with_cigar) {
				z = _mm_max_epi8(z, a);
			} else if (!right) {
				d = _mm_and_si128(_mm_cmpgt_epi8(a, z), flag1_);       // d = a > z? 1 : 0
				z = _mm_max_epi8(z, a);
				tmp = _mm_cmpgt_epi8(b, z);
				d = _mm_blendv_epi8(d, flag2_, tmp);                   // d = b > z? 2 : d
			} else {
				d = _mm_andnot_si128(_mm_cmpgt_epi8(z, a), flag1_);    // d = z > a? 0 : 1
				z = _mm_max_epi8(z, a);
				tmp = _mm_cmpgt_epi8(z, b);
				d = _mm_blendv_epi8(flag2_, d, tmp);                   // d = z > b? d : 2
			}
			z = _mm_max_epu8(z, b);
			z = _mm_min_epu8(z, max_sc_);
			_mm_storeu_si128((__m128i*)(u8 + t), _mm_sub_epi8(z, vt1));
			_mm_storeu_si128((__m128i*)(v8 + t), _mm_sub_epi8(z, ut));
			z = _mm_sub_epi8(z, q_);
			a = _mm_sub_epi8(a, z);
			b = _mm_sub_epi8(b, z);
			if (!with_cigar || !right) {
				tmp = _mm_cmpgt_epi8(a, zero_);
				_mm_storeu_si128((__m128i*)(x8 + t), _mm_and_si128(tmp, a));
				d = _mm_or_si128(d, _mm_and_si128(tmp, flag8_));       // d = a > 0? 0x08 : 0
				tmp = _mm_cmpgt_epi8(b, zero_);
				_mm_storeu_si128((__m128i*)(y8 + t), _mm_and_si128(tmp, b));
				d = _mm_or_si128(d, _mm_and_si128(tmp, flag16_));      // d = b > 0? 0x10 : 0
			} else {
				tmp = _mm_cmpgt_epi8(zero_, a);
				_mm_storeu_si128((__m128i*)(x8 + t), _mm_andnot_si128(tmp, a));
				d = _mm_or_si128(d, _mm_andnot_si128(tmp, flag8_));    // d = 0 > a? 0 : 0x08
				tmp = _mm_cmpgt_epi8(zero_, b);
				_mm_storeu_si128((__m128i*)(y8 + t), _mm_andnot_si128(tmp, b));
				d = _mm_or_si128(d, _mm_andnot_si128(tmp, flag16_));   // d = 0 > b? 0 : 0x10
			}
			if (with_cigar) _mm_storeu_si128((__m128i*)(pr + t), d);
		}
		if (n_cell) *n_cell += en0 - st0 + 1;
		if (!approx_max) { // find the exact max with a 32-bit score array
			int32_t max_H, max_t;
			if (r > 0) {
				int32_t HH[4], tt[4], en1 = st0 + (en0 - st0) / 4 * 4, i;
				max_H = H[en0] = en0 > 0 ? H[en0 - 1] + u8[en0] - qe : H[en0] + v8[en0] - qe; // special casing the last element
				max_t = en0;
				__m128i max_H_ = _mm_set1_epi32(max_H), max_t_ = _mm_set1_epi32(max_t);
				const __m128i qe_ = _mm_set1_epi32(q + e);
				for (t = st0; t < en1; t += 4) { // this implements: H[t]+=v8[t]-qe; if(H[t]>max_H) max_H=H[t],max_t=t;
					__m128i H1 = _mm_loadu_si128((const __m128i*)&H[t]);
					__m128i t_ = _mm_setr_epi32(v8[t], v8[t + 1], v8[t + 2], v8[t + 3]);
					H1 = _mm_add_epi32(H1, t_);
					H1 = _mm_sub_epi32(H1, qe_);
					_mm_storeu_si128((__m128i*)&H[t], H1);
					t_ = _mm_set1_epi32(t);
					const __m128i tmp = _mm_cmpgt_epi32(H1, max_H_);
					max_H_ = _mm_blendv_epi8(max_H_, H1, tmp);
					max_t_ = _mm_blendv_epi8(max_t_, t_, tmp);
				}
				_mm_storeu_si128((__m128i*)HH, max_H_);
				_mm_storeu_si128((__m128i*)tt, max_t_);
				for (i = 0; i < 4; ++i)
					if (max_H < HH[i]) max_H = HH[i], max_t = tt[i] + i;
				for (; t < en0; ++t) { // for the rest of values that haven't been computed with SSE
					H[t] += (int32_t)v8[t] - qe;
					if (H[t] > max_H)
						max_H = H[t], max_t = t;
				}
			} else H[0] = v8[0] - qe - qe, max_H = H[0], max_t = 0; // special casing r==0
			// update ez
			if (en0 == tlen - 1 && H[en0] > ez->mte)
				ez->mte = H[en0], ez->mte_q = r - en;
			if (r - st0 == qlen - 1 && H[st0] > ez->mqe)
				ez->mqe = H[st0], ez->mqe_t = st0;
			if (ksw_apply_zdrop(ez, 1, max_H, r, max_t, zdrop, e)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1)
				ez->score = H[tlen - 1];
		} else { // find approximate max; Z-drop might be inaccurate, too.
			if (r > 0) {
				if (last_H0_t >= st0 && last_H0_t <= en0 && last_H0_t + 1 >= st0 && last_H0_t + 1 <= en0) {
					int32_t d0 = v8[last_H0_t] - qe;
					int32_t d1 = u8[last_H0_t + 1] - qe;
					if (d0 > d1) H0 += d0;
					else H0 += d1, ++last_H0_t;
				} else if (last_H0_t >= st0 && last_H0_t <= en0) {
					H0 += v8[last_H0_t] - qe;
				} else {
					++last_H0_t, H0 += u8[last_H0_t] - qe;
				}
				if ((flag & KSW_EZ_APPROX_DROP) && ksw_apply_zdrop(ez, 1, H0, r, last_H0_t, zdrop, e)) break;
			} else H0 = v8[0] - qe - qe, last_H0_t = 0;
			if (r == qlen + tlen - 2 && en0 == tlen - 1)
				ez->score = H0;
		}
		last_st = st, last_en = en;
	}
	if (with_cigar) { // backtrack
		int rev_cigar = !!(flag & KSW_EZ_REV_CIGAR);
		if (!ez->zdropped && !(flag & KSW_EZ_EXTZ_ONLY)) {
			ksw_backtrack_rot(rev_cigar, p.data(), off.data(), off_end.data(), n_col, tlen - 1, qlen - 1, ez->cigar);
		} else if (!ez->zdropped && (flag & KSW_EZ_EXTZ_ONLY) && ez->mqe + end_bonus > (int)ez->max) {
			ez->reach_end = 1;
			ksw_backtrack_rot(rev_cigar, p.data(), off.data(), off_end.data(), n_col, ez->mqe_t, qlen - 1, ez->cigar);
		} else if (ez->max_t >= 0 && ez->max_q >= 0) {
			ksw_backtrack_rot(rev_cigar, p.data(), off.data(), off_end.data(), n_col, ez->max_t, ez->max_q, ez->cigar);
		}
	}
}

/* ksw2_ll_sse.c: ksw_ll_qinit(size=2) + ksw_ll_i16 -- striped (Farrar) local alignment score,
 * restated lane for lane (8 x int16) because qe/te tie-breaking follows the striped layout. */
int ksw_ll_i16(int qlen, const uint8_t *query, int m, const int8_t *mat, int tlen, const uint8_t *target, int gapo, int gape, int *qe, int *te)
{
	const int p = 8, slen = (qlen + p - 1) / p;
	int i, gmax = 0, qlen8 = slen * 8;
	/* query profile: qp[a][j][k] = mat[a][query[k*slen + j]] (0 beyond qlen) */
	std::vector<int16_t> qp((size_t)m * slen * 8), H0v((size_t)slen * 8, 0), H1v((size_t)slen * 8, 0), E((size_t)slen * 8, 0), Hmax((size_t)slen * 8, 0);
	for (int a = 0; a < m; ++a)
		for (int j = 0; j < slen; ++j)
			for (int k = 0; k < 8; ++k) {
				int pos = k * slen + j;
				qp[((size_t)a * slen + j) * 8 + k] = pos >= qlen ? 0 : mat[a * m + query[pos]];
			}
	int16_t *H0 = H0v.data(), *H1 = H1v.data();
	const int gapoe = gapo + gape;
	auto adds = [](int a, int b) { int s = a + b; return (int16_t)(s > 32767 ? 32767 : s < -32768 ? -32768 : s); };
	auto subsu = [](int16_t a, int b) { int s = (int)(uint16_t)a - b; return (int16_t)(uint16_t)(s < 0 ? 0 : s); };
	*qe = *te = -1;
	for (i = 0; i < tlen; ++i) {
		int j, k, imax;
		int16_t e[8], h[8], f[8], mx[8];
		const int16_t *S = &qp[(size_t)target[i] * slen * 8];
		for (k = 0; k < 8; ++k) f[k] = 0, mx[k] = 0;
		for (k = 7; k > 0; --k) h[k] = H0[(size_t)(slen - 1) * 8 + k - 1]; // h = H0[slen-1] shifted by one lane
		h[0] = 0;
		for (j = 0; j < slen; ++j) {
			for (k = 0; k < 8; ++k) {
				int16_t hh = adds(h[k], S[(size_t)j * 8 + k]);
				e[k] = E[(size_t)j * 8 + k];
				hh = hh > e[k] ? hh : e[k];
				hh = hh > f[k] ? hh : f[k];
				mx[k] = mx[k] > hh ? mx[k] : hh;
				H1[(size_t)j * 8 + k] = hh;
				hh = subsu(hh, gapoe);
				e[k] = subsu(e[k], gape);
				e[k] = e[k] > hh ? e[k] : hh;
				E[(size_t)j * 8 + k] = e[k];
				f[k] = subsu(f[k], gape);
				f[k] = f[k] > hh ? f[k] : hh;
				h[k] = H0[(size_t)j * 8 + k];
			}
		}
		for (k = 0; k < 8; ++k) {
			int16_t carry[8];
			for (int l = 7; l > 0; --l) carry[l] = f[l - 1];
			carry[0] = 0;
			memcpy(f, carry, sizeof(f));
			int done = 0;
			for (j = 0; j < slen; ++j) {
				int any = 0;
				for (int l = 0; l < 8; ++l) {
					int16_t hh = H1[(size_t)j * 8 + l];
					hh = hh > f[l] ? hh : f[l];
					H1[(size_t)j * 8 + l] = hh;
					hh = subsu(hh, gapoe);
					f[l] = subsu(f[l], gape);
					if (f[l] > hh) any = 1;
				}
				if (!any) { done = 1; break; }
			}
			if (done) break;
		}
		imax = 0;
		for (k = 0; k < 8; ++k) imax = imax > mx[k] ? imax : mx[k];
		if (imax >= gmax) {
			gmax = imax; *te = i;
			memcpy(Hmax.data(), H1, (size_t)slen * 8 * sizeof(int16_t));
		}
		int16_t *tmp = H1; H1 = H0; H0 = tmp;
	}
	for (i = 0; i < qlen8; ++i)
		if ((int)(uint16_t)Hmax[i] == gmax) *qe = i / 8 + i % 8 * slen;
	return gmax;
}
