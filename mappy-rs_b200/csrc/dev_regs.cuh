/* dev_regs.cuh -- device functions over a read's region records (DevReg).
 * Each function names the minimap2 v2.26 routine it replaces; all are reached
 * from mm_map (/root/reference/src/lib.rs:482,587).  Serial functions are run
 * by lane 0 of the read's warp. */
#ifndef MMG_DEV_REGS_CUH
#define MMG_DEV_REGS_CUH
#include "dev_common.cuh"
#include "dev_sort.cuh"

#define REG_SET(r, shift, width, v) ((r).bits = ((r).bits & ~(((1u << (width)) - 1u) << (shift))) | ((uint32_t)(v) << (shift)))
#define PARENT_UNSET   (-1)
#define PARENT_TMP_PRI (-2)

__device__ __forceinline__ uint32_t dev_wang_hash(uint32_t key) /* khash.h __ac_Wang_hash */
{
	key += ~(key << 15);
	key ^= (key >> 10);
	key += (key << 3);
	key ^= (key >> 6);
	key += ~(key << 11);
	key ^= (key >> 16);
	return key;
}

__device__ __forceinline__ uint64_t dev_hash64u(uint64_t key) /* hit.c hash64 (unmasked) */
{
	key = (~key + (key << 21));
	key = key ^ key >> 24;
	key = ((key + (key << 3)) + (key << 8));
	key = key ^ key >> 14;
	key = ((key + (key << 2)) + (key << 4));
	key = key ^ key >> 28;
	key = (key + (key << 31));
	return key;
}

/* ---- glibc 2.39 logf (sysdeps/ieee754/flt-32/e_logf.c), bit-exact -------- */
__device__ __forceinline__ float dev_logf(float x)
{
	const double T_invc[16] = {
		0x1.661ec79f8f3bep+0, 0x1.571ed4aaf883dp+0, 0x1.49539f0f010bp+0, 0x1.3c995b0b80385p+0,
		0x1.30d190c8864a5p+0, 0x1.25e227b0b8eap+0, 0x1.1bb4a4a1a343fp+0, 0x1.12358f08ae5bap+0,
		0x1.0953f419900a7p+0, 0x1p+0, 0x1.e608cfd9a47acp-1, 0x1.ca4b31f026aap-1,
		0x1.b2036576afce6p-1, 0x1.9c2d163a1aa2dp-1, 0x1.886e6037841edp-1, 0x1.767dcf5534862p-1 };
	const double T_logc[16] = {
		-0x1.57bf7808caadep-2, -0x1.2bef0a7c06ddbp-2, -0x1.01eae7f513a67p-2, -0x1.b31d8a68224e9p-3,
		-0x1.6574f0ac07758p-3, -0x1.1aa2bc79c81p-3, -0x1.a4e76ce8c0e5ep-4, -0x1.1973c5a611cccp-4,
		-0x1.252f438e10c1ep-5, 0x0p+0, 0x1.aa5aa5df25984p-5, 0x1.c5e53aa362eb4p-4,
		0x1.526e57720db08p-3, 0x1.bc2860d22477p-3, 0x1.1058bc8a07ee1p-2, 0x1.4043057b6ee09p-2 };
	const double Ln2 = 0x1.62e42fefa39efp-1, A0 = -0x1.00ea348b88334p-2, A1 = 0x1.5575b0be00b6ap-2, A2 = -0x1.ffffef20a4123p-2;
	uint32_t ix = __float_as_uint(x);
	if (ix == 0x3f800000u) return 0.0f;
	if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) { /* x < 0x1p-126 or inf or nan: not reachable from mapq inputs */
		if (ix * 2 == 0) return -__int_as_float(0x7f800000);
		if (ix == 0x7f800000u) return x;
		if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return __int_as_float(0x7fc00000);
		ix = __float_as_uint(__fmul_rn(x, 0x1p23f));
		ix -= 23u << 23;
	}
	uint32_t tmp = ix - 0x3f330000u;
	int i = (int)((tmp >> 19) % 16u);
	int k = (int32_t)tmp >> 23;
	uint32_t iz = ix - (tmp & 0xff800000u);
	double z = (double)__uint_as_float(iz);
	double r = __dsub_rn(__dmul_rn(z, T_invc[i]), 1.0);
	double y0 = __dadd_rn(T_logc[i], __dmul_rn((double)k, Ln2));
	double r2 = __dmul_rn(r, r);
	double y = __dadd_rn(__dmul_rn(A1, r), A2);
	y = __dadd_rn(__dmul_rn(A0, r2), y);
	y = __dadd_rn(__dmul_rn(y, r2), __dadd_rn(y0, r));
	return (float)y;
}

/* hit.c: mm_reg_set_coor + mm_cal_fuzzy_len; the two length sums run on the whole warp */
__device__ __forceinline__ void dev_reg_set_coor(DevReg *r, int32_t qlen, const uint64_t *ax, const uint64_t *ay)
{
	const int lane = mmg_lane();
	const int k = r->as, cnt = r->cnt;
	int32_t blen = 0, mlen = 0;
	for (int i = k + 1 + lane; i < k + cnt; i += 32) {
		int span = (int)(ay[i] >> 32 & 0xff);
		int tl = (int32_t)ax[i] - (int32_t)ax[i - 1];
		int ql = (int32_t)ay[i] - (int32_t)ay[i - 1];
		blen += tl > ql ? tl : ql;
		mlen += tl > span && ql > span ? span : tl < ql ? tl : ql;
	}
	blen = __reduce_add_sync(MMG_FULL, blen);
	mlen = __reduce_add_sync(MMG_FULL, mlen);
	if (lane == 0) {
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
		if (cnt <= 0) r->mlen = r->blen = 0;
		else r->mlen = q_span + mlen, r->blen = q_span + blen;
	}
	__syncwarp();
}

/* hit.c: mm_gen_regs (whole warp; sort replay on lane 0) */
static __device__ void dev_gen_regs(uint32_t hash, int qlen, int n_u, const uint64_t *u, const uint64_t *ax, const uint64_t *ay, DevReg *regs,
                             uint64_t *zx, uint64_t *zy, int *bkt, int *stk)
{
	const int lane = mmg_lane();
	if (lane == 0) {
		int k = 0;
		for (int i = 0; i < n_u; ++i) {
			uint32_t h = (uint32_t)dev_hash64u((dev_hash64u(ax[k]) + dev_hash64u(ay[k])) ^ hash);
			zx[i] = u[i] ^ h;
			zy[i] = (uint64_t)k << 32 | (uint32_t)u[i];
			k += (int)(uint32_t)u[i];
		}
		dev_radix_sort_128x(zx, zy, n_u, bkt, stk);
		for (int i = 0; i < n_u >> 1; ++i) { /* larger score first */
			uint64_t tx = zx[i], ty = zy[i];
			zx[i] = zx[n_u - 1 - i], zy[i] = zy[n_u - 1 - i];
			zx[n_u - 1 - i] = tx, zy[n_u - 1 - i] = ty;
		}
		for (int i = 0; i < n_u; ++i) {
			DevReg g;
			g.id = i, g.parent = PARENT_UNSET;
			g.score = g.score0 = (int32_t)(zx[i] >> 32);
			g.hash = (uint32_t)zx[i];
			g.cnt = (int32_t)(uint32_t)zy[i];
			g.as = (int32_t)(zy[i] >> 32);
			g.div = -1.0f;
			g.rid = g.qs = g.qe = g.rs = g.re = 0;
			g.subsc = 0, g.mlen = g.blen = 0, g.n_sub = 0, g.bits = 0;
			g.dp_score = g.dp_max = g.dp_max2 = g.n_ambi = 0;
			g.n_cigar = 0, g.pad = 0, g.cigar_off = 0;
			regs[i] = g;
		}
	}
	__syncwarp();
	for (int i = 0; i < n_u; ++i) dev_reg_set_coor(&regs[i], qlen, ax, ay);
}

__device__ __forceinline__ int dev_alt_score(int score, float alt_diff_frac)
{
	if (score < 0) return score;
	score = (int)(score * (1.0 - alt_diff_frac) + .499);
	return score > 0 ? score : 1;
}

/* hit.c: mm_set_parent.  cov[n], w[n] scratch.  No ALT contigs on this path (mi->n_alt == 0). */
static __device__ void dev_set_parent(float mask_level, int mask_len, int n, DevReg *r, int sub_diff, float alt_diff_frac, uint64_t *cov, int *w)
{
	if (n <= 0) return;
	for (int i = 0; i < n; ++i) r[i].id = i;
	w[0] = 0, r[0].parent = 0;
	int k = 1;
	for (int i = 1; i < n; ++i) {
		DevReg *ri = &r[i];
		int si = ri->qs, ei = ri->qe, n_cov = 0, uncov_len = 0, j;
		for (j = 0; j < k; ++j) { /* overlapping primaries */
			DevReg *rp = &r[w[j]];
			int sj = rp->qs, ej = rp->qe;
			if (ej <= si || sj >= ei) continue;
			if (sj < si) sj = si;
			if (ej > ei) ej = ei;
			cov[n_cov++] = (uint64_t)sj << 32 | (uint32_t)ej;
		}
		bool is_new = false;
		if (n_cov == 0) is_new = true;
		else {
			int x = si;
			for (int a = 1; a < n_cov; ++a) { /* radix_sort_64: values only, any sort gives the same array */
				uint64_t tv = cov[a];
				int b = a;
				for (; b > 0 && tv < cov[b - 1]; --b) cov[b] = cov[b - 1];
				cov[b] = tv;
			}
			for (int a = 0; a < n_cov; ++a) {
				if ((int)(cov[a] >> 32) > x) uncov_len += (int)(cov[a] >> 32) - x;
				x = (int32_t)cov[a] > x ? (int32_t)cov[a] : x;
			}
			if (ei > x) uncov_len += ei - x;
			for (j = 0; j < k; ++j) {
				DevReg *rp = &r[w[j]];
				int sj = rp->qs, ej = rp->qe, mn, mx, ol;
				if (ej <= si || sj >= ei) continue;
				mn = ej - sj < ei - si ? ej - sj : ei - si;
				mx = ej - sj > ei - si ? ej - sj : ei - si;
				ol = si < sj ? (ei < sj ? 0 : ei < ej ? ei - sj : ej - sj) : (ej < si ? 0 : ej < ei ? ej - si : ei - si);
				if (__fsub_rn(__fdiv_rn((float)ol, (float)mn), __fdiv_rn((float)uncov_len, (float)mx)) > mask_level && uncov_len <= mask_len) {
					int cnt_sub = 0, sci = ri->score;
					ri->parent = rp->parent;
					rp->subsc = rp->subsc > sci ? rp->subsc : sci;
					if (ri->cnt >= rp->cnt) cnt_sub = 1;
					if (REG_HASP(*rp) && REG_HASP(*ri) && (rp->rid != ri->rid || rp->rs != ri->rs || rp->re != ri->re || ol != mn)) {
						sci = ri->dp_max;
						rp->dp_max2 = rp->dp_max2 > sci ? rp->dp_max2 : sci;
						if (rp->dp_max - ri->dp_max <= sub_diff) cnt_sub = 1;
					}
					if (cnt_sub) ++rp->n_sub;
					break;
				}
			}
			if (j == k) is_new = true;
		}
		if (is_new) w[k++] = i, ri->parent = i, ri->n_sub = 0;
	}
	(void)alt_diff_frac;
}

/* hit.c: mm_set_sam_pri */
__device__ __forceinline__ void dev_set_sam_pri(int n, DevReg *r)
{
	int n_pri = 0;
	for (int i = 0; i < n; ++i)
		if (r[i].id == r[i].parent) { ++n_pri; REG_SET(r[i], 12, 1, n_pri == 1); }
		else REG_SET(r[i], 12, 1, 0);
}

/* hit.c: mm_sync_regs.  tmp: scratch of max_id+1 ints */
static __device__ void dev_sync_regs(int n_regs, DevReg *regs, int *tmp)
{
	int max_id = -1;
	if (n_regs <= 0) return;
	for (int i = 0; i < n_regs; ++i) max_id = max_id > regs[i].id ? max_id : regs[i].id;
	int n_tmp = max_id + 1;
	for (int i = 0; i < n_tmp; ++i) tmp[i] = -1;
	for (int i = 0; i < n_regs; ++i) if (regs[i].id >= 0) tmp[regs[i].id] = i;
	for (int i = 0; i < n_regs; ++i) {
		DevReg *r = &regs[i];
		r->id = i;
		if (r->parent == PARENT_TMP_PRI) r->parent = i;
		else if (r->parent >= 0 && tmp[r->parent] >= 0) r->parent = tmp[r->parent];
		else r->parent = PARENT_UNSET;
	}
	dev_set_sam_pri(n_regs, regs);
}

/* hit.c: mm_select_sub -- in place, reading r[p] exactly as upstream does (a parent
 * slot may already hold a later record once earlier records were dropped) */
static __device__ void dev_select_sub(float pri_ratio, int min_diff, int best_n, int check_strand, int min_strand_sc, int *n_, DevReg *r, int *tmp)
{
	if (pri_ratio > 0.0f && *n_ > 0) {
		int i, k, n = *n_, n_2nd = 0;
		for (i = k = 0; i < n; ++i) {
			int p = r[i].parent;
			if (p == i || REG_INV(r[i])) {
				r[k++] = r[i];
			} else if (((float)r[i].score >= __fmul_rn((float)r[p].score, pri_ratio) || r[i].score + min_diff >= r[p].score) && n_2nd < best_n) {
				if (!(r[i].qs == r[p].qs && r[i].qe == r[p].qe && r[i].rid == r[p].rid && r[i].rs == r[p].rs && r[i].re == r[p].re))
					r[k++] = r[i], ++n_2nd;
			} else if (check_strand && n_2nd < best_n && r[i].score > min_strand_sc && REG_REV(r[p]) != REG_REV(r[i])) {
				REG_SET(r[i], 13, 1, 1);
				r[k++] = r[i], ++n_2nd;
			}
		}
		if (k != n) dev_sync_regs(k, r, tmp);
		*n_ = k;
	}
}

/* hit.c: mm_filter_strand_retained */
static __device__ int dev_filter_strand_retained(int n_regs, DevReg *r)
{
	int i, k;
	for (i = k = 0; i < n_regs; ++i) {
		int p = r[i].parent;
		if (!REG_SRET(r[i]) || r[i].div < __fmul_rn(r[p].div, 5.0f) || r[i].div < 0.01f) {
			if (k < i) r[k++] = r[i];
			else ++k;
		}
	}
	return k;
}

/* esterr.c: mm_est_err.  mini_pos[i] = span<<32 | q_pos>>1 is rebuilt from the kept seeds */
__device__ __forceinline__ int32_t dev_for_qpos(int32_t qlen, uint64_t x, uint64_t y)
{
	int32_t v = (int32_t)y, q_span = (int32_t)(y >> 32 & 0xff);
	if (x >> 63) v = qlen - 1 - (v + 1 - q_span);
	return v;
}

/* Called by the whole warp.  Upstream walks the read's minimizers j and the region's anchors k together and
 * advances k when anchor k sits at minimizer j; if anchor k is at none of the remaining minimizers the walk ends
 * with k stuck.  Query positions increase along both lists, so anchor k matches exactly when its position occurs
 * among the minimizers: 32 anchors are looked up per step by binary search, and the walk ends at the first one
 * that is absent - the same n_match and the same last matched minimizer `en`. */
static __device__ void dev_est_err(const DevIndex &di, int qlen, int n_regs, DevReg *regs, const uint64_t *ax, const uint64_t *ay, int n, const uint32_t *sq, const uint32_t *sm)
{
	if (n == 0) return;
	const int lane = mmg_lane();
	unsigned sum_k = 0;
	for (int i = lane; i < n; i += 32) sum_k += sm[i] >> 8 & 0xff;
	sum_k = __reduce_add_sync(MMG_FULL, sum_k);
	const float avg_k = __fdiv_rn((float)(uint64_t)sum_k, (float)n);
	for (int i = 0; i < n_regs; ++i) {
		DevReg *r = &regs[i];
		const bool rev = REG_REV(*r);
		const int cnt = r->cnt, as = r->as;
		if (lane == 0) r->div = -1.0f;
		if (cnt == 0) continue;
		int32_t st = -1;
		{
			int a0 = rev ? as + cnt - 1 : as;
			int32_t x = dev_for_qpos(qlen, ax[a0], ay[a0]), L = 0, R = n - 1;
			while (L <= R) {
				int32_t m = (int32_t)(((uint64_t)L + R) >> 1);
				int32_t y = (int32_t)(sq[m] >> 1);
				if (y < x) L = m + 1;
				else if (y > x) R = m - 1;
				else { st = m; break; }
			}
		}
		if (st < 0) continue;
		int32_t en = st, n_match = 1;
		for (int k0 = 1; k0 < cnt; k0 += 32) {
			const int k = k0 + lane;
			int32_t found = -1;
			if (k < cnt) {
				const int a1 = rev ? as + cnt - 1 - k : as + k;
				const int32_t x = dev_for_qpos(qlen, ax[a1], ay[a1]);
				int32_t L = st + 1, R = n - 1;
				while (L <= R) {
					int32_t m = (int32_t)(((uint64_t)L + R) >> 1);
					int32_t y = (int32_t)(sq[m] >> 1);
					if (y < x) L = m + 1;
					else if (y > x) R = m - 1;
					else { found = m; break; }
				}
			}
			const uint32_t miss = __ballot_sync(MMG_FULL, k < cnt && found < 0);
			const uint32_t have = __ballot_sync(MMG_FULL, k < cnt);
			const int good = miss ? __ffs((int)miss) - 1 : __popc(have); /* leading anchors of this step that matched */
			if (good > 0) {
				n_match += good;
				en = __shfl_sync(MMG_FULL, found, good - 1);
			}
			if (miss) break;
		}
		if (lane == 0) {
			const int32_t l_ref = (int32_t)di.seq_len[r->rid];
			int32_t n_tot = en - st + 1;
			if ((float)r->qs > avg_k && (float)r->rs > avg_k) ++n_tot;
			if ((float)(qlen - r->qs) > avg_k && (float)(l_ref - r->re) > avg_k) ++n_tot;
			r->div = n_match >= n_tot ? 0.0f : (float)(1.0 - pow((double)n_match / n_tot, 1.0 / (double)avg_k));
		}
	}
	__syncwarp();
}

/* hit.c: mm_set_inv_mapq -- an inversion takes the smaller mapq of its two neighbours on the target */
static __device__ void dev_set_inv_mapq(int n_regs, DevReg *regs, uint64_t *zx, uint64_t *zy, int *bkt, int *stk)
{
	int i, n_aux;
	if (n_regs < 3) return;
	for (i = 0; i < n_regs; ++i) if (REG_INV(regs[i])) break;
	if (i == n_regs) return;
	for (i = n_aux = 0; i < n_regs; ++i)
		if (regs[i].parent == i || regs[i].parent < 0)
			zy[n_aux] = (uint64_t)i, zx[n_aux++] = (uint64_t)(uint32_t)regs[i].rid << 32 | (uint32_t)regs[i].rs;
	dev_radix_sort_128x(zx, zy, n_aux, bkt, stk);
	for (i = 1; i < n_aux - 1; ++i) {
		DevReg *inv = &regs[(int)zy[i]];
		if (REG_INV(*inv)) {
			const DevReg *l = &regs[(int)zy[i - 1]], *r = &regs[(int)zy[i + 1]];
			uint32_t mq = REG_MAPQ(*l) < REG_MAPQ(*r) ? REG_MAPQ(*l) : REG_MAPQ(*r);
			REG_SET(*inv, 0, 8, mq);
		}
	}
}

/* hit.c: mm_set_mapq (is_sr = 0).  Inversion regions only arise after alignment. */
static __device__ void dev_set_mapq(int n_regs, DevReg *regs, int min_chain_sc, int match_sc, int rep_len)
{
	const float q_coef = 40.0f;
	int64_t sum_sc = 0;
	if (n_regs == 0) return;
	for (int i = 0; i < n_regs; ++i)
		if (regs[i].parent == regs[i].id) sum_sc += regs[i].score;
	const float uniq_ratio = __fdiv_rn((float)sum_sc, (float)(sum_sc + rep_len));
	for (int i = 0; i < n_regs; ++i) {
		DevReg *r = &regs[i];
		uint32_t mq = 0;
		if (REG_INV(*r)) mq = 0;
		else if (r->parent == r->id) {
			int mapq, subsc;
			float pen_s1 = __fmul_rn(r->score > 100 ? 1.0f : __fmul_rn(0.01f, (float)r->score), uniq_ratio);
			float pen_cm = r->cnt > 10 ? 1.0f : __fmul_rn(0.1f, (float)r->cnt);
			pen_cm = pen_s1 < pen_cm ? pen_s1 : pen_cm;
			subsc = r->subsc > min_chain_sc ? r->subsc : min_chain_sc;
			if (REG_HASP(*r) && r->dp_max2 > 0 && r->dp_max > 0) {
				float identity = __fdiv_rn((float)r->mlen, (float)r->blen);
				float x = __fdiv_rn(__fdiv_rn(__fmul_rn((float)r->dp_max2, (float)subsc), (float)r->dp_max), (float)r->score0);
				float v = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(identity, pen_cm), q_coef), __fsub_rn(1.0f, __fmul_rn(x, x))),
				                    dev_logf(__fdiv_rn((float)r->dp_max, (float)match_sc)));
				mapq = (int)v;
				{
					float a = __fmul_rn(__fmul_rn(__fmul_rn(6.02f, identity), identity), (float)(r->dp_max - r->dp_max2));
					int mapq_alt = (int)__fadd_rn(__fdiv_rn(a, (float)match_sc), .499f);
					mapq = mapq < mapq_alt ? mapq : mapq_alt;
				}
			} else {
				float x = __fdiv_rn((float)subsc, (float)r->score0);
				if (REG_HASP(*r)) {
					float identity = __fdiv_rn((float)r->mlen, (float)r->blen);
					mapq = (int)__fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(identity, pen_cm), q_coef), __fsub_rn(1.0f, x)),
					                      dev_logf(__fdiv_rn((float)r->dp_max, (float)match_sc)));
				} else {
					mapq = (int)__fmul_rn(__fmul_rn(__fmul_rn(pen_cm, q_coef), __fsub_rn(1.0f, x)), dev_logf((float)r->score));
				}
			}
			mapq -= (int)__fadd_rn(__fmul_rn(4.343f, dev_logf((float)(r->n_sub + 1))), .499f);
			mapq = mapq > 0 ? mapq : 0;
			mq = (uint32_t)(mapq < 60 ? mapq : 60);
			if (REG_HASP(*r) && r->dp_max > r->dp_max2 && mq == 0) mq = 1;
		} else mq = 0;
		REG_SET(*r, 0, 8, mq);
	}
}

#endif
