/* mm2o.h -- ORACLE (test infrastructure only; never linked into the product).
 *
 * CPU restatement of the minimap2 v2.26 `mm_map` path that mappy-rs drives
 * through `Aligner.map` / `map_batch` (/root/reference/src/lib.rs:482-488,
 * 587-593 -> crate minimap2 0.1.15 `Aligner::map` -> `mm_map`).
 *
 * The arithmetic lives in the un-vendored dependency
 * `minimap2-sys 0.1.15+minimap2.2.26` (/root/reference/Cargo.toml:24); neither
 * the crate nor upstream minimap2 is present in this container, so this file
 * restates the published algorithm of minimap2 v2.26 (sketch.c, index.c,
 * seed.c, map.c, lchain.c, krmq.h, hit.c, esterr.c, align.c, ksw2_ext*2_sse.c,
 * options.c, format.c).  Each function names the upstream function it follows
 * and the reference call site that reaches it.
 *
 * PARITY STATUS: pinned by the reference's own fixtures for the .mmi format,
 * hash64, mm_sketch, 4-bit sequence decode and the `map_one` known answer
 * (resources/test/test.{fa,mmi}; src/lib.rs:1040-1106).  Everything else
 * (chaining scores, mapq, CIGARs, secondaries...) is "parity unpinned": no
 * golden vector exists in the reference and no minimap2 binary exists here.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * legs may use this library.
 */
#ifndef MM2O_H
#define MM2O_H

#include <stdint.h>
#include <stddef.h>
#include <vector>
#include <string>

/* ---- flags (minimap.h) ---- */
#define MM_F_NO_DIAG       0x001
#define MM_F_NO_DUAL       0x002
#define MM_F_CIGAR         0x004
#define MM_F_OUT_SAM       0x008
#define MM_F_NO_QUAL       0x010
#define MM_F_OUT_CG        0x020
#define MM_F_OUT_CS        0x040
#define MM_F_SPLICE        0x080
#define MM_F_SPLICE_FOR    0x100
#define MM_F_SPLICE_REV    0x200
#define MM_F_NO_LJOIN      0x400
#define MM_F_OUT_CS_LONG   0x800
#define MM_F_SR            0x1000
#define MM_F_FRAG_MODE     0x2000
#define MM_F_NO_PRINT_2ND  0x4000
#define MM_F_2_IO_THREADS  0x8000
#define MM_F_LONG_CIGAR    0x10000
#define MM_F_INDEPEND_SEG  0x20000
#define MM_F_SPLICE_FLANK  0x40000
#define MM_F_SOFTCLIP      0x80000
#define MM_F_FOR_ONLY      0x100000
#define MM_F_REV_ONLY      0x200000
#define MM_F_HEAP_SORT     0x400000
#define MM_F_ALL_CHAINS    0x800000
#define MM_F_OUT_MD        0x1000000
#define MM_F_COPY_COMMENT  0x2000000
#define MM_F_EQX           0x4000000
#define MM_F_PAF_NO_HIT    0x8000000
#define MM_F_NO_END_FLT    0x10000000
#define MM_F_HARD_MLEVEL   0x20000000
#define MM_F_SAM_HIT_ONLY  0x40000000
#define MM_F_RMQ           0x80000000LL
#define MM_F_QSTRAND       0x100000000LL
#define MM_F_NO_INV        0x200000000LL
#define MM_F_NO_HASH_NAME  0x400000000LL

#define MM_I_HPC     0x1
#define MM_I_NO_SEQ  0x2
#define MM_I_NO_NAME 0x4

#define MM_PARENT_UNSET   (-1)
#define MM_PARENT_TMP_PRI (-2)

#define MM_SEED_LONG_JOIN (1ULL<<40)
#define MM_SEED_IGNORE    (1ULL<<41)
#define MM_SEED_TANDEM    (1ULL<<42)
#define MM_SEED_SELF      (1ULL<<43)
#define MM_SEED_SEG_SHIFT 48
#define MM_SEED_SEG_MASK  (0xffULL<<(MM_SEED_SEG_SHIFT))

#define MM_CIGAR_MATCH    0
#define MM_CIGAR_INS      1
#define MM_CIGAR_DEL      2
#define MM_CIGAR_N_SKIP   3
#define MM_CIGAR_EQ_MATCH 7
#define MM_CIGAR_X_MISMATCH 8

struct mm128_t { uint64_t x, y; };
typedef std::vector<mm128_t> mm128_v;

/* minimap.h: mm_idxopt_t */
struct mm_idxopt_t {
	short k, w, flag, bucket_bits;
	int64_t mini_batch_size;
	uint64_t batch_size;
};

/* minimap.h: mm_mapopt_t (v2.26 field set) */
struct mm_mapopt_t {
	int64_t flag;
	int seed;
	int sdust_thres;
	int max_qlen;
	int bw, bw_long;
	int max_gap, max_gap_ref;
	int max_frag_len;
	int max_chain_skip, max_chain_iter;
	int min_cnt;
	int min_chain_score;
	float chain_gap_scale;
	float chain_skip_scale;
	int rmq_size_cap, rmq_inner_dist;
	int rmq_rescue_size;
	float rmq_rescue_ratio;
	float mask_level;
	int mask_len;
	float pri_ratio;
	int best_n;
	float alt_drop;
	int a, b, q, e, q2, e2;
	int transition;
	int sc_ambi;
	int noncan;
	int junc_bonus;
	int zdrop, zdrop_inv;
	int end_bonus;
	int min_dp_max;
	int min_ksw_len;
	int anchor_ext_len, anchor_ext_shift;
	float max_clip_ratio;
	int rank_min_len;
	float rank_frac;
	int pe_ori, pe_bonus;
	float mid_occ_frac;
	float q_occ_frac;
	int32_t min_mid_occ, max_mid_occ;
	int32_t mid_occ;
	int32_t max_occ, max_max_occ, occ_dist;
	int64_t mini_batch_size;
	int64_t max_sw_mat;
	int64_t cap_kalloc;
};

struct mm_idx_seq_t {
	std::string name;
	uint64_t offset;
	uint32_t len;
	uint32_t is_alt;
};

/* One bucket of the index: a sorted table instead of upstream's khash (the
 * iteration order of khash is implementation-defined and nothing on the
 * mapping path depends on it). */
struct mm_idx_bucket_t {
	std::vector<uint64_t> keys; /* minier>>b<<1 | is_single, sorted ascending */
	std::vector<uint64_t> vals;
	std::vector<uint64_t> p;    /* position runs */
};

struct mm_idx_t {
	int32_t b, w, k, flag;
	uint32_t n_seq;
	int32_t n_alt;
	std::vector<mm_idx_seq_t> seq;
	std::vector<uint32_t> S;    /* 4-bit packed sequence */
	std::vector<mm_idx_bucket_t> B;
};

/* minimap.h: mm_extra_t */
struct mm_extra_t {
	uint32_t capacity;
	int32_t dp_score, dp_max, dp_max2;
	uint32_t n_ambi, trans_strand; /* bit-fields upstream */
	std::vector<uint32_t> cigar;   /* len<<4|op */
};

/* minimap.h: mm_reg1_t */
struct mm_reg1_t {
	int32_t id, cnt, rid, score;
	int32_t qs, qe, rs, re;
	int32_t parent, subsc;
	int32_t as;
	int32_t mlen, blen;
	int32_t n_sub;
	int32_t score0;
	uint32_t mapq, split, rev, inv, sam_pri, proper_frag, pe_thru, seg_split, seg_id, split_inv, is_alt, strand_retained;
	uint32_t hash;
	float div;
	mm_extra_t *p;
};

/* per-read statistics counters used as roofline denominators */
struct mm2o_stats_t {
	uint64_t n_bases, n_mz, n_seed, n_hit, n_anchor, n_iter, n_kept, n_cell, n_regs, n_rechain;
};

/* intermediate stages, kept when the caller asks for them (differential tests) */
struct mm2o_trace_t {
	mm128_v mv;            /* after mm_seed_mz_flt */
	mm128_v a_sorted;      /* anchors after radix_sort_128x */
	std::vector<uint64_t> u_dp; mm128_v a_dp;   /* after mm_lchain_dp */
	std::vector<uint64_t> u;    mm128_v a;      /* after optional re-chain */
	int rechained, rep_len;
	std::vector<mm_reg1_t> regs_gen;   /* after mm_gen_regs */
	std::vector<mm_reg1_t> regs_chain; /* after chain_post + est_err + strand filter */
};

/* ---- options.c ---- */
void mm_idxopt_init(mm_idxopt_t *opt);
void mm_mapopt_init(mm_mapopt_t *opt);
int  mm_set_opt(const char *preset, mm_idxopt_t *io, mm_mapopt_t *mo);
void mm_mapopt_update(mm_mapopt_t *opt, const mm_idx_t *mi);

/* ---- sketch.c ---- */
void mm_sketch(const char *str, int len, int w, int k, uint32_t rid, int is_hpc, mm128_v *p);

/* ---- index.c ---- */
mm_idx_t *mm_idx_load(const char *fn);                       /* .mmi v2 */
mm_idx_t *mm_idx_build(int w, int k, int b, int flag, int n_seq, const char **names, const char **seqs, const uint32_t *lens);
mm_idx_t *mm_idx_from_fasta(const char *fn, int w, int k, int b, int flag);
int       mm_idx_dump(const char *fn, const mm_idx_t *mi);
void      mm_idx_destroy(mm_idx_t *mi);
const uint64_t *mm_idx_get(const mm_idx_t *mi, uint64_t minier, int *n);
int32_t   mm_idx_cal_max_occ(const mm_idx_t *mi, float f);
int       mm_idx_getseq(const mm_idx_t *mi, uint32_t rid, uint32_t st, uint32_t en, uint8_t *seq);
int       mm_idx_getseq_rev(const mm_idx_t *mi, uint32_t rid, uint32_t st, uint32_t en, uint8_t *seq);
int       mm_idx_name2id(const mm_idx_t *mi, const char *name);

/* ---- map.c ---- */
mm_reg1_t *mm_map(const mm_idx_t *mi, int qlen, const char *seq, int *n_regs, const mm_mapopt_t *opt, const char *qname,
                  mm2o_stats_t *st, mm2o_trace_t *tr);
void mm_free_regs(mm_reg1_t *regs, int n);

/* ---- lchain.c ---- */
mm128_t *mm_lchain_dp(int max_dist_x, int max_dist_y, int bw, int max_skip, int max_iter, int min_cnt, int min_sc, float chn_pen_gap, float chn_pen_skip,
                      int is_cdna, int n_seg, int64_t n, mm128_t *a, int *n_u_, uint64_t **_u, uint64_t *n_iter_);
mm128_t *mm_lchain_rmq(int max_dist, int max_dist_inner, int bw, int max_chn_skip, int cap_rmq_size, int min_cnt, int min_sc, float chn_pen_gap, float chn_pen_skip,
                       int64_t n, mm128_t *a, int *n_u_, uint64_t **_u);

/* ---- hit.c / esterr.c ---- */
mm_reg1_t *mm_gen_regs(uint32_t hash, int qlen, int n_u, uint64_t *u, mm128_t *a, int is_qstrand);
void mm_set_parent(float mask_level, int mask_len, int n, mm_reg1_t *r, int sub_diff, int hard_mask_level, float alt_diff_frac);
void mm_select_sub(float pri_ratio, int min_diff, int best_n, int check_strand, int min_strand_sc, int *n_, mm_reg1_t *r);
void mm_sync_regs(int n_regs, mm_reg1_t *regs);
int  mm_set_sam_pri(int n, mm_reg1_t *r);
void mm_hit_sort(int *n_regs, mm_reg1_t *r, float alt_diff_frac);
void mm_filter_regs(const mm_mapopt_t *opt, int qlen, int *n_regs, mm_reg1_t *regs);
int  mm_filter_strand_retained(int n_regs, mm_reg1_t *r);
void mm_set_mapq(int n_regs, mm_reg1_t *regs, int min_chain_sc, int match_sc, int rep_len, int is_sr);
void mm_est_err(const mm_idx_t *mi, int qlen, int n_regs, mm_reg1_t *regs, const mm128_t *a, int32_t n, const uint64_t *mini_pos);
void mm_split_reg(mm_reg1_t *r, mm_reg1_t *r2, int n, int qlen, mm128_t *a, int is_qstrand);
void mm_reg_set_coor(mm_reg1_t *r, int32_t qlen, const mm128_t *a, int is_qstrand);

/* ---- align.c ---- */
mm_reg1_t *mm_align_skeleton(const mm_mapopt_t *opt, const mm_idx_t *mi, int qlen, const char *qstr, int *n_regs_, mm_reg1_t *regs, mm128_t *a, mm2o_stats_t *st);

/* ---- format.c ---- */
std::string mm_gen_cs(const mm_idx_t *mi, const mm_reg1_t *r, const char *seq, int no_iden);
std::string mm_gen_MD(const mm_idx_t *mi, const mm_reg1_t *r, const char *seq);

extern const unsigned char seq_nt4_table[256];

#endif
