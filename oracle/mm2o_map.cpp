/* mm2o_map.cpp -- ORACLE (test infrastructure only).
 * Restates minimap2 v2.26 seed.c (mm_seed_mz_flt, mm_seed_collect_all,
 * mm_seed_select, mm_collect_matches) and map.c (collect_minimizers,
 * collect_seed_hits, chain_post, align_regs, mm_map_frag, mm_map) for the
 * single-segment long-read path.
 * Reference call sites: /root/reference/src/lib.rs:482-488 (`Aligner.map`) and
 * :587-593 (worker threads of `map_batch`) -> crate minimap2 `Aligner::map` ->
 * `mm_map(idx, len, seq, &n_regs, tbuf, &mapopt, qname=NULL)`.
 */
#include <stdlib.h>
#include <string.h>
#include <assert.h>
#include "mm2o.h"
#include "mm2o_sort.h"

struct mm_seed_t {
	uint32_t n;
	uint32_t q_pos;
	uint32_t q_span:8, flt:1, seg_id:8, is_tandem:1;
	const uint64_t *cr;
};

/* seed.c: mm_seed_mz_flt */
static void mm_seed_mz_flt(mm128_v *mv, int32_t q_occ_max, float q_occ_frac)
{
	size_t i, j, st;
	if ((int64_t)mv->size() <= q_occ_max || q_occ_frac <= 0.0f || q_occ_max <= 0) return;
	std::vector<mm128_t> a(mv->size());
	for (i = 0; i < mv->size(); ++i)
		a[i].x = (*mv)[i].x, a[i].y = i;
	radix_sort_128x(a.data(), a.data() + a.size());
	for (st = 0, i = 1; i <= mv->size(); ++i) {
		if (i == mv->size() || a[i].x != a[st].x) {
			int32_t cnt = i - st;
			if (cnt > q_occ_max && cnt > mv->size() * q_occ_frac)
				for (j = st; j < i; ++j)
					(*mv)[a[j].y].x = 0;
			st = i;
		}
	}
	for (i = j = 0; i < mv->size(); ++i)
		if ((*mv)[i].x != 0)
			(*mv)[j++] = (*mv)[i];
	mv->resize(j);
}

/* seed.c: mm_seed_collect_all */
static mm_seed_t *mm_seed_collect_all(const mm_idx_t *mi, const mm128_v *mv, int32_t *n_m_)
{
	mm_seed_t *m;
	size_t i;
	int32_t k;
	m = (mm_seed_t*)malloc((mv->size() ? mv->size() : 1) * sizeof(mm_seed_t));
	for (i = k = 0; i < mv->size(); ++i) {
		const uint64_t *cr;
		mm_seed_t *q;
		const mm128_t *p = &(*mv)[i];
		uint32_t q_pos = (uint32_t)p->y, q_span = p->x & 0xff;
		int t;
		cr = mm_idx_get(mi, p->x >> 8, &t);
		if (t == 0) continue;
		q = &m[k++];
		q->q_pos = q_pos, q->q_span = q_span, q->cr = cr, q->n = t, q->seg_id = p->y >> 32;
		q->is_tandem = q->flt = 0;
		if (i > 0 && p->x >> 8 == (*mv)[i - 1].x >> 8) q->is_tandem = 1;
		if (i < mv->size() - 1 && p->x >> 8 == (*mv)[i + 1].x >> 8) q->is_tandem = 1;
	}
	*n_m_ = k;
	return m;
}

#define MAX_MAX_HIGH_OCC 128

/* seed.c: mm_seed_select */
static void mm_seed_select(int32_t n, mm_seed_t *a, int len, int max_occ, int max_max_occ, int dist)
{ // for high-occ minimizers, choose up to max_high_occ in each high-occ streak
	int32_t i, last0, m;
	uint64_t b[MAX_MAX_HIGH_OCC]; // this is to avoid a heap allocation

	if (n == 0 || n == 1) return;
	for (i = m = 0; i < n; ++i)
		if (a[i].n > (uint32_t)max_occ) ++m;
	if (m == 0) return; // no high-frequency k-mers; do nothing
	for (i = 0, last0 = -1; i <= n; ++i) {
		if (i == n || a[i].n <= (uint32_t)max_occ) {
			if (i - last0 > 1) {
				int32_t ps = last0 < 0 ? 0 : (uint32_t)a[last0].q_pos >> 1;
				int32_t pe = i == n ? len : (uint32_t)a[i].q_pos >> 1;
				int32_t j, k, st = last0 + 1, en = i;
				int32_t max_high_occ = (int32_t)((double)(pe - ps) / dist + .499);
				if (max_high_occ > 0) {
					if (max_high_occ > MAX_MAX_HIGH_OCC)
						max_high_occ = MAX_MAX_HIGH_OCC;
					for (j = st, k = 0; j < en && k < max_high_occ; ++j, ++k)
						b[k] = (uint64_t)a[j].n << 32 | j;
					ks_heapmake_uint64_t(k, b); // initialize the binomial heap
					for (; j < en; ++j) { // if there are more, choose top max_high_occ
						if (a[j].n < (int32_t)(b[0] >> 32)) { // then update the heap
							b[0] = (uint64_t)a[j].n << 32 | j;
							ks_heapdown_uint64_t(0, k, b);
						}
					}
					for (j = 0; j < k; ++j) a[(uint32_t)b[j]].flt = 1;
				}
				for (j = st; j < en; ++j) a[j].flt ^= 1;
				for (j = st; j < en; ++j)
					if (a[j].n > (uint32_t)max_max_occ)
						a[j].flt = 1;
			}
			last0 = i;
		}
	}
}

/* seed.c: mm_collect_matches */
static mm_seed_t *mm_collect_matches(int *_n_m, int qlen, int max_occ, int max_max_occ, int dist, const mm_idx_t *mi, const mm128_v *mv, int64_t *n_a, int *rep_len, int *n_mini_pos, uint64_t **mini_pos)
{
	int rep_st = 0, rep_en = 0, n_m, n_m0;
	size_t i;
	mm_seed_t *m;
	*n_mini_pos = 0;
	*mini_pos = (uint64_t*)malloc((mv->size() ? mv->size() : 1) * sizeof(uint64_t));
	m = mm_seed_collect_all(mi, mv, &n_m0);
	if (dist > 0 && max_max_occ > max_occ) {
		mm_seed_select(n_m0, m, qlen, max_occ, max_max_occ, dist);
	} else {
		for (i = 0; i < (size_t)n_m0; ++i)
			if (m[i].n > (uint32_t)max_occ)
				m[i].flt = 1;
	}
	for (i = 0, n_m = 0, *rep_len = 0, *n_a = 0; i < (size_t)n_m0; ++i) {
		mm_seed_t *q = &m[i];
		if (q->flt) {
			int en = (q->q_pos >> 1) + 1, st = en - q->q_span;
			if (st > rep_en) {
				*rep_len += rep_en - rep_st;
				rep_st = st, rep_en = en;
			} else rep_en = en;
		} else {
			*n_a += q->n;
			(*mini_pos)[(*n_mini_pos)++] = (uint64_t)q->q_span << 32 | q->q_pos >> 1;
			m[n_m++] = *q;
		}
	}
	*rep_len += rep_en - rep_st;
	*_n_m = n_m;
	return m;
}

/* map.c: collect_seed_hits (skip_seed() reduces to the strand-only filters
 * because mappy-rs passes qname = NULL) */
static mm128_t *collect_seed_hits(const mm_mapopt_t *opt, int max_occ, const mm_idx_t *mi, const mm128_v *mv, int qlen, int64_t *n_a, int *rep_len,
                                  int *n_mini_pos, uint64_t **mini_pos, mm2o_stats_t *st)
{
	int i, n_m;
	mm_seed_t *m;
	mm128_t *a;
	m = mm_collect_matches(&n_m, qlen, max_occ, opt->max_max_occ, opt->occ_dist, mi, mv, n_a, rep_len, n_mini_pos, mini_pos);
	a = (mm128_t*)malloc((*n_a ? *n_a : 1) * sizeof(mm128_t));
	if (st) st->n_seed += n_m, st->n_hit += *n_a;
	for (i = 0, *n_a = 0; i < n_m; ++i) {
		mm_seed_t *q = &m[i];
		const uint64_t *r = q->cr;
		uint32_t k;
		for (k = 0; k < q->n; ++k) {
			int32_t rpos = (uint32_t)r[k] >> 1;
			mm128_t *p;
			if (opt->flag & (MM_F_FOR_ONLY | MM_F_REV_ONLY)) { // map.c: skip_seed
				if ((r[k] & 1) == (q->q_pos & 1)) { // forward strand
					if (opt->flag & MM_F_REV_ONLY) continue;
				} else {
					if (opt->flag & MM_F_FOR_ONLY) continue;
				}
			}
			p = &a[(*n_a)++];
			if ((r[k] & 1) == (q->q_pos & 1)) { // forward strand
				p->x = (r[k] & 0xffffffff00000000ULL) | rpos;
				p->y = (uint64_t)q->q_span << 32 | q->q_pos >> 1;
			} else { // reverse strand (query-strand mode is outside the mappy-rs path)
				p->x = 1ULL << 63 | (r[k] & 0xffffffff00000000ULL) | rpos;
				p->y = (uint64_t)q->q_span << 32 | (qlen - ((q->q_pos >> 1) + 1 - q->q_span) - 1);
			}
			p->y |= (uint64_t)q->seg_id << MM_SEED_SEG_SHIFT;
			if (q->is_tandem) p->y |= MM_SEED_TANDEM;
		}
	}
	free(m);
	radix_sort_128x(a, a + (*n_a));
	return a;
}

/* khash.h: __ac_Wang_hash */
static inline uint32_t ac_Wang_hash(uint32_t key)
{
	key += ~(key << 15);
	key ^= (key >> 10);
	key += (key << 3);
	key ^= (key >> 6);
	key += ~(key << 11);
	key ^= (key >> 16);
	return key;
}

/* khash.h: __ac_X31_hash_string */
static inline uint32_t ac_X31_hash_string(const char *s)
{
	uint32_t h = (uint32_t)*s;
	if (h) for (++s; *s; ++s) h = (h << 5) - h + (uint32_t)*s;
	return h;
}

static void chain_post(const mm_mapopt_t *opt, const mm_idx_t *mi, int *n_regs, mm_reg1_t *regs)
{
	if (!(opt->flag & MM_F_ALL_CHAINS)) { // don't choose primary mapping(s)
		mm_set_parent(opt->mask_level, opt->mask_len, *n_regs, regs, opt->a * 2 + opt->b, opt->flag & MM_F_HARD_MLEVEL, opt->alt_drop);
		mm_select_sub(opt->pri_ratio, mi->k * 2, opt->best_n, 1, opt->max_gap * 0.8, n_regs, regs);
	}
}

static mm_reg1_t *align_regs(const mm_mapopt_t *opt, const mm_idx_t *mi, int qlen, const char *seq, int *n_regs, mm_reg1_t *regs, mm128_t *a, mm2o_stats_t *st)
{
	if (!(opt->flag & MM_F_CIGAR)) return regs;
	regs = mm_align_skeleton(opt, mi, qlen, seq, n_regs, regs, a, st); // this calls mm_filter_regs()
	if (!(opt->flag & MM_F_ALL_CHAINS)) { // don't choose primary mapping(s)
		mm_set_parent(opt->mask_level, opt->mask_len, *n_regs, regs, opt->a * 2 + opt->b, opt->flag & MM_F_HARD_MLEVEL, opt->alt_drop);
		mm_select_sub(opt->pri_ratio, mi->k * 2, opt->best_n, 0, opt->max_gap * 0.8, n_regs, regs);
		mm_set_sam_pri(*n_regs, regs);
	}
	return regs;
}

static void copy_regs(std::vector<mm_reg1_t> &dst, const mm_reg1_t *r, int n)
{
	dst.assign(r, r + n);
	for (int i = 0; i < n; ++i) dst[i].p = 0;
}

/* map.c: mm_map_frag with n_segs == 1, then mm_map */
mm_reg1_t *mm_map(const mm_idx_t *mi, int qlen, const char *seq, int *n_regs, const mm_mapopt_t *opt, const char *qname,
                  mm2o_stats_t *st, mm2o_trace_t *tr)
{
	int rep_len = 0, n_regs0 = 0, n_mini_pos = 0;
	int max_chain_gap_qry, max_chain_gap_ref, is_splice = !!(opt->flag & MM_F_SPLICE), is_sr = !!(opt->flag & MM_F_SR);
	uint32_t hash;
	int64_t n_a = 0;
	uint64_t *u = 0, *mini_pos = 0;
	mm128_t *a;
	mm128_v mv;
	mm_reg1_t *regs0;
	float chn_pen_gap, chn_pen_skip;

	*n_regs = 0;
	if (qlen == 0) return 0;
	if (opt->max_qlen > 0 && qlen > opt->max_qlen) return 0;

	hash  = qname && !(opt->flag & MM_F_NO_HASH_NAME) ? ac_X31_hash_string(qname) : 0;
	hash ^= ac_Wang_hash(qlen) + ac_Wang_hash(opt->seed);
	hash  = ac_Wang_hash(hash);

	// map.c: collect_minimizers (sdust masking is off: sdust_thres == 0)
	mm_sketch(seq, qlen, mi->w, mi->k, 0, mi->flag & MM_I_HPC, &mv);
	if (st) st->n_bases += qlen, st->n_mz += mv.size();
	if (opt->q_occ_frac > 0.0f) mm_seed_mz_flt(&mv, opt->mid_occ, opt->q_occ_frac);
	if (tr) tr->mv = mv;
	a = collect_seed_hits(opt, opt->mid_occ, mi, &mv, qlen, &n_a, &rep_len, &n_mini_pos, &mini_pos, st);
	if (st) st->n_anchor += n_a;
	if (tr) tr->a_sorted.assign(a, a + n_a), tr->rep_len = rep_len;

	// set max chaining gap on the query and the reference sequence
	if (is_sr)
		max_chain_gap_qry = qlen > opt->max_gap ? qlen : opt->max_gap;
	else max_chain_gap_qry = opt->max_gap;
	if (opt->max_gap_ref > 0) {
		max_chain_gap_ref = opt->max_gap_ref; // always honor mm_mapopt_t::max_gap_ref if set
	} else if (opt->max_frag_len > 0) {
		max_chain_gap_ref = opt->max_frag_len - qlen;
		if (max_chain_gap_ref < opt->max_gap) max_chain_gap_ref = opt->max_gap;
	} else max_chain_gap_ref = opt->max_gap;

	chn_pen_gap  = opt->chain_gap_scale * 0.01 * mi->k;
	chn_pen_skip = opt->chain_skip_scale * 0.01 * mi->k;
	if (opt->flag & MM_F_RMQ) {
		a = mm_lchain_rmq(opt->max_gap, opt->rmq_inner_dist, opt->bw, opt->max_chain_skip, opt->rmq_size_cap, opt->min_cnt, opt->min_chain_score,
		                  chn_pen_gap, chn_pen_skip, n_a, a, &n_regs0, &u);
	} else {
		a = mm_lchain_dp(max_chain_gap_ref, max_chain_gap_qry, opt->bw, opt->max_chain_skip, opt->max_chain_iter, opt->min_cnt, opt->min_chain_score,
		                 chn_pen_gap, chn_pen_skip, is_splice, 1, n_a, a, &n_regs0, &u, st ? &st->n_iter : 0);
	}
	if (tr) {
		int64_t na = 0;
		for (int i = 0; i < n_regs0; ++i) na += (int32_t)u[i];
		tr->u_dp.assign(u, u + n_regs0);
		tr->a_dp.assign(a, a + na);
		tr->rechained = 0;
	}

	if (opt->bw_long > opt->bw && (opt->flag & (MM_F_SPLICE | MM_F_SR | MM_F_NO_LJOIN)) == 0 && n_regs0 > 1) { // re-chain/long-join for long sequences
		int32_t st_ = (int32_t)a[0].y, en = (int32_t)a[(int32_t)u[0] - 1].y;
		if (qlen - (en - st_) > opt->rmq_rescue_size || en - st_ > qlen * opt->rmq_rescue_ratio) {
			int32_t i;
			for (i = 0, n_a = 0; i < n_regs0; ++i) n_a += (int32_t)u[i];
			free(u);
			radix_sort_128x(a, a + n_a);
			a = mm_lchain_rmq(opt->max_gap, opt->rmq_inner_dist, opt->bw_long, opt->max_chain_skip, opt->rmq_size_cap, opt->min_cnt, opt->min_chain_score,
			                  chn_pen_gap, chn_pen_skip, n_a, a, &n_regs0, &u);
			if (st) st->n_rechain += 1;
			if (tr) tr->rechained = 1;
		}
	}
	// (the short-read re-chain branch needs opt->max_occ > opt->mid_occ; max_occ is 0 on this path)
	if (tr) {
		int64_t na = 0;
		for (int i = 0; i < n_regs0; ++i) na += (int32_t)u[i];
		tr->u.assign(u, u + n_regs0);
		tr->a.assign(a, a + na);
	}
	if (st) for (int i = 0; i < n_regs0; ++i) st->n_kept += (int32_t)u[i];

	regs0 = mm_gen_regs(hash, qlen, n_regs0, u, a, !!(opt->flag & MM_F_QSTRAND));
	if (tr) copy_regs(tr->regs_gen, regs0, n_regs0);

	chain_post(opt, mi, &n_regs0, regs0);
	if (!is_sr && !(opt->flag & MM_F_QSTRAND)) {
		mm_est_err(mi, qlen, n_regs0, regs0, a, n_mini_pos, mini_pos);
		n_regs0 = mm_filter_strand_retained(n_regs0, regs0);
	}
	if (tr) copy_regs(tr->regs_chain, regs0, n_regs0);

	regs0 = align_regs(opt, mi, qlen, seq, &n_regs0, regs0, a, st);
	mm_set_mapq(n_regs0, regs0, opt->min_chain_score, opt->a, rep_len, is_sr);
	*n_regs = n_regs0;
	if (st) st->n_regs += n_regs0;

	free(a);
	free(u);
	free(mini_pos);
	return regs0;
}

void mm_free_regs(mm_reg1_t *regs, int n)
{
	for (int i = 0; i < n; ++i) if (regs[i].p) delete regs[i].p;
	free(regs);
}
