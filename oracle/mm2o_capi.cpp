/* mm2o_capi.cpp -- ORACLE (test infrastructure only).
 * Flat C entry points over the oracle for ctypes (tests/, bench.py cpu_baseline
 * and --impl reference, __graft_entry__.smoke()).  The batch mapper mirrors the
 * topology of mappy-rs `map_batch` (/root/reference/src/lib.rs:541-636): N worker
 * threads, one read per task, index shared read-only.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <thread>
#include <atomic>
#include <vector>
#include <string>
#include "mm2o.h"

/* identical layout to include/mmg.h: mmg_hit_t (kept separate on purpose: the
 * oracle shares no code with the product) */
struct mm2o_hit_t {
	int32_t rid, rs, re, qs, qe;
	int32_t mlen, blen;
	int32_t score, score0, cnt, subsc, n_sub;
	int32_t parent, id;
	int32_t dp_score, dp_max, dp_max2;
	int32_t nm, n_ambi;
	uint32_t hash;
	float div;
	uint8_t rev, mapq, is_primary, flags; /* flags: 1 sam_pri, 2 inv, 4 strand_retained, 8 split&1, 16 split&2, 32 has_cigar */
	uint32_t n_cigar;
	uint64_t cigar_off;
};

struct mm2o_result_t {
	std::vector<uint64_t> hit_off;   /* n_reads + 1 */
	std::vector<mm2o_hit_t> hits;
	std::vector<uint32_t> cigar;
	std::vector<uint64_t> cs_off;    /* per hit + 1 (empty unless cs requested) */
	std::string cs;
	mm2o_stats_t stats;
};

struct mm2o_aligner_t {
	mm_idx_t *mi;
	mm_idxopt_t io;
	mm_mapopt_t mo;
};

static void fill_hit(const mm_idx_t *mi, const mm_reg1_t *r, mm2o_hit_t *h)
{
	memset(h, 0, sizeof(*h));
	h->rid = r->rid, h->rs = r->rs, h->re = r->re, h->qs = r->qs, h->qe = r->qe;
	h->mlen = r->mlen, h->blen = r->blen;
	h->score = r->score, h->score0 = r->score0, h->cnt = r->cnt, h->subsc = r->subsc, h->n_sub = r->n_sub;
	h->parent = r->parent, h->id = r->id;
	h->hash = r->hash, h->div = r->div;
	h->rev = r->rev, h->mapq = r->mapq, h->is_primary = (r->parent == r->id);
	h->flags = (r->sam_pri ? 1 : 0) | (r->inv ? 2 : 0) | (r->strand_retained ? 4 : 0) | ((r->split & 1) ? 8 : 0) | ((r->split & 2) ? 16 : 0) | (r->p ? 32 : 0);
	if (r->p) {
		h->dp_score = r->p->dp_score, h->dp_max = r->p->dp_max, h->dp_max2 = r->p->dp_max2;
		h->n_ambi = r->p->n_ambi;
		h->nm = r->blen - r->mlen + r->p->n_ambi; /* crate minimap2 0.1.15 Aligner::map */
		h->n_cigar = (uint32_t)r->p->cigar.size();
	}
}

extern "C" {

mm2o_aligner_t *mm2o_open(const char *fn_idx, const char *preset, int is_fasta)
{
	mm2o_aligner_t *al = new mm2o_aligner_t();
	mm_set_opt(0, &al->io, &al->mo);
	if (preset && preset[0] && mm_set_opt(preset, &al->io, &al->mo) < 0) { delete al; return 0; }
	al->mo.flag |= 4; /* src/lib.rs:339 */
	al->mi = is_fasta ? mm_idx_from_fasta(fn_idx, al->io.w, al->io.k, al->io.bucket_bits, al->io.flag) : mm_idx_load(fn_idx);
	if (!al->mi) { delete al; return 0; }
	mm_mapopt_update(&al->mo, al->mi);
	return al;
}

/* build from in-memory sequences (synthetic references of the benchmark) */
mm2o_aligner_t *mm2o_build(const char *preset, int n_seq, const char **names, const char **seqs, const uint32_t *lens)
{
	mm2o_aligner_t *al = new mm2o_aligner_t();
	mm_set_opt(0, &al->io, &al->mo);
	if (preset && preset[0] && mm_set_opt(preset, &al->io, &al->mo) < 0) { delete al; return 0; }
	al->mo.flag |= 4;
	al->mi = mm_idx_build(al->io.w, al->io.k, al->io.bucket_bits, al->io.flag, n_seq, names, seqs, lens);
	mm_mapopt_update(&al->mo, al->mi);
	return al;
}

/* the same with the k / w overrides of the constructor (src/lib.rs:341-346 write idxopt.k / idxopt.w) */
mm2o_aligner_t *mm2o_build_kw(const char *preset, int k, int w, int n_seq, const char **names, const char **seqs, const uint32_t *lens)
{
	mm2o_aligner_t *al = new mm2o_aligner_t();
	mm_set_opt(0, &al->io, &al->mo);
	if (preset && preset[0] && mm_set_opt(preset, &al->io, &al->mo) < 0) { delete al; return 0; }
	al->mo.flag |= 4;
	if (k > 0) al->io.k = (short)k;
	if (w > 0) al->io.w = (short)w;
	al->mi = mm_idx_build(al->io.w, al->io.k, al->io.bucket_bits, al->io.flag, n_seq, names, seqs, lens);
	mm_mapopt_update(&al->mo, al->mi);
	return al;
}

void mm2o_close(mm2o_aligner_t *al) { if (al) { mm_idx_destroy(al->mi); delete al; } }

int mm2o_dump_index(mm2o_aligner_t *al, const char *fn) { return mm_idx_dump(fn, al->mi); }

/* option access by name: the python side never needs the struct layout */
#define OPT_FIELDS(X) X(seed) X(max_qlen) X(bw) X(bw_long) X(max_gap) X(max_gap_ref) X(max_frag_len) X(max_chain_skip) X(max_chain_iter) \
	X(min_cnt) X(min_chain_score) X(rmq_size_cap) X(rmq_inner_dist) X(rmq_rescue_size) X(mask_len) X(best_n) X(a) X(b) X(q) X(e) X(q2) X(e2) \
	X(transition) X(sc_ambi) X(zdrop) X(zdrop_inv) X(end_bonus) X(min_dp_max) X(min_ksw_len) X(anchor_ext_len) X(anchor_ext_shift) \
	X(min_mid_occ) X(max_mid_occ) X(mid_occ) X(max_occ) X(max_max_occ) X(occ_dist)

int mm2o_set_opt_int(mm2o_aligner_t *al, const char *name, int64_t v)
{
	if (strcmp(name, "flag") == 0) { al->mo.flag = v; return 0; }
	if (strcmp(name, "max_sw_mat") == 0) { al->mo.max_sw_mat = v; return 0; }
#define X(f) if (strcmp(name, #f) == 0) { al->mo.f = (int)v; return 0; }
	OPT_FIELDS(X)
#undef X
	return -1;
}

int64_t mm2o_get_opt_int(mm2o_aligner_t *al, const char *name)
{
	if (strcmp(name, "flag") == 0) return al->mo.flag;
	if (strcmp(name, "max_sw_mat") == 0) return al->mo.max_sw_mat;
	if (strcmp(name, "k") == 0) return al->mi->k;
	if (strcmp(name, "w") == 0) return al->mi->w;
	if (strcmp(name, "b") == 0) return al->mi->b;
	if (strcmp(name, "idx_flag") == 0) return al->mi->flag;
	if (strcmp(name, "n_seq") == 0) return al->mi->n_seq;
#define X(f) if (strcmp(name, #f) == 0) return al->mo.f;
	OPT_FIELDS(X)
#undef X
	return INT64_MIN;
}

const char *mm2o_seq_name(mm2o_aligner_t *al, int i) { return al->mi->seq[i].name.c_str(); }
uint32_t mm2o_seq_len(mm2o_aligner_t *al, int i) { return al->mi->seq[i].len; }
int mm2o_name2id(mm2o_aligner_t *al, const char *name) { return mm_idx_name2id(al->mi, name); }
int mm2o_getseq(mm2o_aligner_t *al, uint32_t rid, uint32_t st, uint32_t en, uint8_t *buf) { return mm_idx_getseq(al->mi, rid, st, en, buf); }

/* index content as flat arrays: (minimizer, y) pairs of every occurrence */
uint64_t mm2o_index_entries(mm2o_aligner_t *al, uint64_t *mz, uint64_t *y, uint64_t cap)
{
	uint64_t n = 0;
	const mm_idx_t *mi = al->mi;
	for (int i = 0; i < 1 << mi->b; ++i) {
		const mm_idx_bucket_t *b = &mi->B[i];
		for (size_t k = 0; k < b->keys.size(); ++k) {
			uint64_t minier = (b->keys[k] >> 1) << mi->b | i;
			if (b->keys[k] & 1) {
				if (mz && n < cap) mz[n] = minier, y[n] = b->vals[k];
				++n;
			} else {
				uint32_t cnt = (uint32_t)b->vals[k];
				for (uint32_t j = 0; j < cnt; ++j) {
					if (mz && n < cap) mz[n] = minier, y[n] = b->p[(b->vals[k] >> 32) + j];
					++n;
				}
			}
		}
	}
	return n;
}

/* minimizers of one sequence: returns count, fills up to cap */
uint64_t mm2o_sketch(const char *seq, int len, int w, int k, uint32_t rid, int is_hpc, uint64_t *x, uint64_t *y, uint64_t cap)
{
	mm128_v v;
	if (len > 0) mm_sketch(seq, len, w, k, rid, is_hpc, &v);
	for (size_t i = 0; i < v.size() && i < cap; ++i) x[i] = v[i].x, y[i] = v[i].y;
	return v.size();
}

mm2o_result_t *mm2o_map_batch(mm2o_aligner_t *al, const char *seqs, const uint64_t *offsets, uint32_t n_reads, int n_threads, int want_cs)
{
	mm2o_result_t *res = new mm2o_result_t();
	memset(&res->stats, 0, sizeof(res->stats));
	std::vector<std::vector<mm2o_hit_t> > hits(n_reads);
	std::vector<std::vector<uint32_t> > cig(n_reads);
	std::vector<std::vector<std::string> > css(n_reads);
	std::atomic<uint32_t> next(0);
	if (n_threads < 1) n_threads = 1;
	std::vector<mm2o_stats_t> tst(n_threads);
	auto worker = [&](int tid) {
		mm2o_stats_t st;
		memset(&st, 0, sizeof(st));
		for (;;) {
			uint32_t i = next.fetch_add(1);
			if (i >= n_reads) break;
			int n_regs = 0, qlen = (int)(offsets[i + 1] - offsets[i]);
			std::string s(seqs + offsets[i], qlen); /* mm_map wants a NUL-terminated copy (src/lib.rs:856-860 also copies) */
			mm_reg1_t *regs = mm_map(al->mi, qlen, s.c_str(), &n_regs, &al->mo, 0, &st, 0);
			hits[i].resize(n_regs);
			for (int j = 0; j < n_regs; ++j) {
				fill_hit(al->mi, &regs[j], &hits[i][j]);
				if (regs[j].p) {
					hits[i][j].cigar_off = cig[i].size();
					cig[i].insert(cig[i].end(), regs[j].p->cigar.begin(), regs[j].p->cigar.end());
					if (want_cs) css[i].push_back(want_cs == 2 ? mm_gen_MD(al->mi, &regs[j], s.c_str()) : mm_gen_cs(al->mi, &regs[j], s.c_str(), want_cs == 3 ? 0 : 1)); /* 1 short cs, 2 MD, 3 long cs */
				} else if (want_cs) css[i].push_back(std::string());
			}
			mm_free_regs(regs, n_regs);
		}
		tst[tid] = st;
	};
	if (n_threads == 1) worker(0);
	else {
		std::vector<std::thread> th;
		for (int t = 0; t < n_threads; ++t) th.emplace_back(worker, t);
		for (auto &t : th) t.join();
	}
	res->hit_off.resize(n_reads + 1);
	uint64_t nh = 0, nc = 0;
	for (uint32_t i = 0; i < n_reads; ++i) res->hit_off[i] = nh, nh += hits[i].size(), nc += cig[i].size();
	res->hit_off[n_reads] = nh;
	res->hits.reserve(nh);
	res->cigar.reserve(nc);
	if (want_cs) res->cs_off.push_back(0);
	for (uint32_t i = 0; i < n_reads; ++i) {
		uint64_t base = res->cigar.size();
		for (size_t j = 0; j < hits[i].size(); ++j) {
			mm2o_hit_t h = hits[i][j];
			h.cigar_off += base;
			res->hits.push_back(h);
			if (want_cs) { res->cs += css[i][j]; res->cs_off.push_back(res->cs.size()); }
		}
		res->cigar.insert(res->cigar.end(), cig[i].begin(), cig[i].end());
	}
	for (int t = 0; t < n_threads; ++t) {
		const uint64_t *s = (const uint64_t*)&tst[t];
		uint64_t *d = (uint64_t*)&res->stats;
		for (size_t k = 0; k < sizeof(mm2o_stats_t) / 8; ++k) d[k] += s[k];
	}
	return res;
}

uint64_t mm2o_result_n_hits(mm2o_result_t *r) { return r->hits.size(); }
uint64_t mm2o_result_n_cigar(mm2o_result_t *r) { return r->cigar.size(); }
const uint64_t *mm2o_result_hit_off(mm2o_result_t *r) { return r->hit_off.data(); }
const mm2o_hit_t *mm2o_result_hits(mm2o_result_t *r) { return r->hits.data(); }
const uint32_t *mm2o_result_cigar(mm2o_result_t *r) { return r->cigar.data(); }
const uint64_t *mm2o_result_cs_off(mm2o_result_t *r) { return r->cs_off.data(); }
const char *mm2o_result_cs(mm2o_result_t *r) { return r->cs.c_str(); }
void mm2o_result_stats(mm2o_result_t *r, uint64_t *out) { memcpy(out, &r->stats, sizeof(mm2o_stats_t)); }
extern int mm2o_ksw_force_scalar;
/* 1 = ksw_extd2 walks its 16-lane blocks byte by byte instead of using the SSE intrinsics (self-check of the oracle) */
void mm2o_set_ksw_scalar(int v) { mm2o_ksw_force_scalar = v; }
int mm2o_sizeof_hit(void) { return (int)sizeof(mm2o_hit_t); }
void mm2o_result_free(mm2o_result_t *r) { delete r; }

/* ---- stage traces of one read (differential tests against GPU intermediates) ---- */
struct mm2o_trace_handle_t { mm2o_trace_t tr; std::vector<mm2o_hit_t> regs_gen, regs_chain, regs_final; std::vector<uint32_t> cigar; };

mm2o_trace_handle_t *mm2o_trace(mm2o_aligner_t *al, const char *seq, int qlen)
{
	mm2o_trace_handle_t *h = new mm2o_trace_handle_t();
	std::string s(seq, qlen);
	int n_regs = 0;
	mm_reg1_t *regs = mm_map(al->mi, qlen, s.c_str(), &n_regs, &al->mo, 0, 0, &h->tr);
	h->regs_gen.resize(h->tr.regs_gen.size());
	for (size_t i = 0; i < h->regs_gen.size(); ++i) fill_hit(al->mi, &h->tr.regs_gen[i], &h->regs_gen[i]);
	h->regs_chain.resize(h->tr.regs_chain.size());
	for (size_t i = 0; i < h->regs_chain.size(); ++i) fill_hit(al->mi, &h->tr.regs_chain[i], &h->regs_chain[i]);
	h->regs_final.resize(n_regs);
	for (int i = 0; i < n_regs; ++i) {
		fill_hit(al->mi, &regs[i], &h->regs_final[i]);
		if (regs[i].p) { h->regs_final[i].cigar_off = h->cigar.size(); h->cigar.insert(h->cigar.end(), regs[i].p->cigar.begin(), regs[i].p->cigar.end()); }
	}
	mm_free_regs(regs, n_regs);
	return h;
}

/* which: 0 mv, 1 a_sorted, 2 a_dp, 3 a (after re-chain) -> 128-bit records; 4 u_dp, 5 u -> 64-bit */
uint64_t mm2o_trace_n(mm2o_trace_handle_t *h, int which)
{
	switch (which) {
	case 0: return h->tr.mv.size();
	case 1: return h->tr.a_sorted.size();
	case 2: return h->tr.a_dp.size();
	case 3: return h->tr.a.size();
	case 4: return h->tr.u_dp.size();
	case 5: return h->tr.u.size();
	case 6: return h->regs_gen.size();
	case 7: return h->regs_chain.size();
	case 8: return h->regs_final.size();
	case 9: return h->tr.rechained;
	case 10: return (uint64_t)h->tr.rep_len;
	}
	return 0;
}

const void *mm2o_trace_ptr(mm2o_trace_handle_t *h, int which)
{
	switch (which) {
	case 0: return h->tr.mv.data();
	case 1: return h->tr.a_sorted.data();
	case 2: return h->tr.a_dp.data();
	case 3: return h->tr.a.data();
	case 4: return h->tr.u_dp.data();
	case 5: return h->tr.u.data();
	case 6: return h->regs_gen.data();
	case 7: return h->regs_chain.data();
	case 8: return h->regs_final.data();
	case 11: return h->cigar.data();
	}
	return 0;
}

void mm2o_trace_free(mm2o_trace_handle_t *h) { delete h; }

} // extern "C"
