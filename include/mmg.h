/* mmg.h -- C ABI of the B200-native minimap2 mapping path behind mappy-rs.
 *
 * This is the drop-in boundary: a thin `extern "C"` library (libmmg.so) whose
 * entry points are what the Rust host of mappy-rs binds INSTEAD of the
 * minimap2-sys FFI it calls today.  Every entry point cites the reference
 * interface it replaces (file:line under /root/reference, i.e. Adoni5/mappy-rs
 * src/lib.rs).  Plain pointers and sizes only; no C++/torch types; every call
 * returns 0 (or a non-negative count) on success and a negative MMG_E* code on
 * failure, with a thread-local message behind mmg_last_error().  The library
 * never calls back into the host language.
 *
 * There is no CPU mapping path behind this ABI: without a CUDA device
 * mmg_aligner_create() fails with MMG_ENODEV.
 */
#ifndef MMG_H
#define MMG_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMG_OK        0
#define MMG_EINVAL   (-1)  /* bad argument / unknown preset */
#define MMG_EIO      (-2)  /* cannot open / parse index or FASTA */
#define MMG_ENODEV   (-3)  /* no CUDA device (there is no CPU fallback) */
#define MMG_ECUDA    (-4)  /* CUDA runtime error (see mmg_last_error) */
#define MMG_ENOMEM   (-5)
#define MMG_ENOSEQ   (-6)  /* index has no sequence (MM_I_NO_SEQ) but CIGAR requested */
#define MMG_EUNSUP   (-7)  /* option combination outside the supported path */

/* mapping flags: same bit values as minimap.h MM_F_* (src/lib.rs:339 ORs 4 = MM_F_CIGAR) */
#define MMG_F_CIGAR      0x004
#define MMG_F_NO_LJOIN   0x400
#define MMG_F_FOR_ONLY   0x100000
#define MMG_F_REV_ONLY   0x200000
#define MMG_F_ALL_CHAINS 0x800000
#define MMG_F_RMQ        0x80000000LL
#define MMG_I_HPC    0x1
#define MMG_I_NO_SEQ 0x2

/* Same field order/types as minimap.h v2.26 `mm_idxopt_t` (what minimap2-sys'
 * bindgen struct exposes and src/lib.rs:340-346 writes into). */
typedef struct {
	short k, w, flag, bucket_bits;
	int64_t mini_batch_size;
	uint64_t batch_size;
} mmg_idxopt_t;

/* Same field order/types as minimap.h v2.26 `mm_mapopt_t` (src/lib.rs:339-385
 * writes flag, min_cnt, min_chain_score, min_dp_max, bw, best_n, max_frag_len,
 * a, b, q, e, q2, e2, sc_ambi). */
typedef struct {
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
	const char *split_prefix;
} mmg_mapopt_t;

/* One mapping of one read: the fields of minimap.h `mm_reg1_t` (+ mm_extra_t)
 * that crate minimap2 0.1.15 `Aligner::map` turns into a `Mapping`
 * (src/lib.rs:489-511, 596-618), plus the chain-level fields the parity tests
 * compare.  64-bit aligned, 112 bytes. */
typedef struct {
	int32_t rid, rs, re, qs, qe;       /* target id/start/end, query start/end */
	int32_t mlen, blen;                /* match_len, block_len */
	int32_t score, score0, cnt, subsc, n_sub;
	int32_t parent, id;
	int32_t dp_score, dp_max, dp_max2; /* 0 unless CIGAR was computed */
	int32_t nm, n_ambi;                /* NM = blen - mlen + n_ambi */
	uint32_t hash;
	float div;
	uint8_t rev, mapq, is_primary, flags; /* flags: 1 sam_pri, 2 inv, 4 strand_retained, 8 split&1, 16 split&2, 32 has_cigar */
	uint32_t n_cigar;
	uint64_t cigar_off;                /* into the batch's cigar pool; ops are len<<4|op */
} mmg_hit_t;

typedef struct mmg_index mmg_index;
typedef struct mmg_aligner mmg_aligner;
typedef struct mmg_batch mmg_batch;

/* ---- options --------------------------------------------------------------
 * replaces mm_set_opt(NULL,..) + mm_set_opt(preset,..)   src/lib.rs:333,336 */
int mmg_set_opt(const char *preset, mmg_idxopt_t *io, mmg_mapopt_t *mo);
/* replaces mm_mapopt_update(&mapopts, idx)               src/lib.rs:414 */
int mmg_mapopt_update(mmg_mapopt_t *mo, const mmg_index *idx);

/* ---- index ----------------------------------------------------------------
 * replaces mm_idx_reader_open / mm_idx_reader_read / mm_idx_reader_close
 * (src/lib.rs:398, 407, 412): `path` is a .mmi v2 file or a FASTA file; only
 * the first index part is read, as the reference does. */
int mmg_index_open(const char *path, const mmg_idxopt_t *io, int n_threads, mmg_index **out);
/* index built from in-memory sequences (benchmark harness; mm_idx_str upstream) */
int mmg_index_build(const mmg_idxopt_t *io, int n_seq, const char *const *names, const char *const *seqs,
                    const uint32_t *lens, int n_threads, mmg_index **out);
/* the same, constructed on CUDA device `device` (index_dev.cu) where a device is present and k is odd: the index
 * then stays resident on that device and mmg_aligner_create() on it uses it in place.  mmg_index_open() on a
 * FASTA and mmg_index_build() use device 0.  (mm_idx_gen upstream; `Aligner("ref.fa")`, src/lib.rs:395-410) */
int mmg_index_build_on(const mmg_idxopt_t *io, int n_seq, const char *const *names, const char *const *seqs,
                       const uint32_t *lens, int n_threads, int device, mmg_index **out);
int mmg_index_dump(const mmg_index *idx, const char *path);     /* mm_idx_dump (fn_idx_out, src/lib.rs:391) */
void mmg_index_destroy(mmg_index *idx);
/* replaces direct reads of mm_idx_t.{k,w,b,flag,n_seq}   src/lib.rs:445,658,663,669,711 */
int mmg_index_info(const mmg_index *idx, int32_t out_k_w_b_flag_nseq[5]);
/* replaces idx.seq[i].name / .len                        src/lib.rs:450,737 */
const char *mmg_index_seq_name(const mmg_index *idx, uint32_t i);
uint32_t mmg_index_seq_len(const mmg_index *idx, uint32_t i);
/* replaces mm_idx_name2id                                src/lib.rs:716 */
int mmg_index_name2id(const mmg_index *idx, const char *name);
/* replaces mm_idx_getseq (codes 0-4)                     src/lib.rs:747 */
int mmg_index_getseq(const mmg_index *idx, uint32_t rid, uint32_t st, uint32_t en, uint8_t *seq);
/* number of (minimizer, position) entries and a flat copy of them (tests) */
uint64_t mmg_index_entries(const mmg_index *idx, uint64_t *minier, uint64_t *pos, uint64_t cap);

/* ---- aligner ---------------------------------------------------------------
 * replaces the minimap2::Aligner value + per-thread mm_tbuf_t (src/lib.rs:419-425,
 * 545): uploads the index to `device` once; it stays GPU-resident. */
int mmg_aligner_create(const mmg_index *idx, const mmg_mapopt_t *mo, int device, mmg_aligner **out);
/* the same across several GPUs of one box (what `enable_threading(n)` + N workers on one shared index are in the
 * reference, src/lib.rs:541-553): the index is uploaded / built once and replicated to the other devices with peer
 * copies, every mmg_map_batch on the returned aligner shards its reads by bases over the devices (one host thread and
 * one set of streams per device, no collective) and gathers the results in read order. */
int mmg_aligner_create_multi(const mmg_index *idx, const mmg_mapopt_t *mo, const int *devices, int n_dev, mmg_aligner **out);
void mmg_aligner_destroy(mmg_aligner *al);
/* tuning knobs of the device pipeline (not mapping semantics):
 * "chunk_bases", "chunk_reads", "anchor_cap", "profile" (1 = per-stage CUDA events) */
int mmg_aligner_set(mmg_aligner *al, const char *key, int64_t value);

/* replaces N x mm_map(idx, len, seq, &n_regs, tbuf, &mapopt, NULL) behind
 * Aligner.map / the map_batch worker threads (src/lib.rs:482-488, 587-593).
 * `bases` is the concatenation of the reads (ASCII), `offsets` has n_reads+1
 * entries.  HOST buffers: host->device copies, all kernels and the device->host
 * copy of the results happen inside this call. */
int mmg_map_batch(mmg_aligner *al, const char *bases, const uint64_t *offsets, uint32_t n_reads, mmg_batch **out);

/* ---- streaming -------------------------------------------------------------
 * replaces the work queue, the worker threads and the result queue behind `enable_threading` / `map_batch` / the
 * result iterator (src/lib.rs:297-309, 541-636, 771-906, 972-991).  mmg_submit copies the reads into page-locked
 * staging owned by the library and returns at once (the host may drop its strings, as src/lib.rs:856-866 does after
 * the push); one worker thread of the aligner maps device-sized batches as soon as enough bases are queued, the
 * producer pauses (20 ms) or calls mmg_flush; mmg_next hands back one read's hits at a time (1 = a result, 0 = none
 * within timeout_ms / nothing outstanding, < 0 = error; timeout_ms < 0 waits).  A result stays valid until
 * mmg_result_release.  submit / flush and next / release may be called from different threads. */
typedef struct {
	uint64_t read_id;          /* first_id + position of the read in its mmg_submit call */
	uint32_t n_hits;
	const mmg_hit_t *hits;     /* n_hits records in the order mm_map returns them */
	const uint32_t *cigar;     /* hits[i].cigar_off indexes this pool */
	void *owner;               /* internal */
} mmg_result_t;
int mmg_submit(mmg_aligner *al, const char *bases, const uint64_t *offsets, uint32_t n_reads, uint64_t first_id);
int mmg_flush(mmg_aligner *al);
int mmg_next(mmg_aligner *al, mmg_result_t *out, int timeout_ms);
void mmg_result_release(mmg_aligner *al, mmg_result_t *r);

/* Page-locked host memory for the `bases` buffer of mmg_map_batch: with it the per-chunk host->device copies are
 * asynchronous and overlap the kernels (a pageable buffer is staged by the driver and blocks the submitting thread).
 * Replaces nothing in the reference (its FFI passes one `*const c_char` per read, src/lib.rs:482-488); it is what the
 * batch host (the Rust work-queue drain) should assemble its reads in. */
int mmg_host_alloc(size_t bytes, void **out);
void mmg_host_free(void *p);

/* The same path split in three so that device time can be measured with the
 * inputs already resident in HBM: upload (H2D), run (kernels only; returns when
 * the device is done), fetch (D2H + host marshalling). */
int mmg_batch_upload(mmg_aligner *al, const char *bases, const uint64_t *offsets, uint32_t n_reads, mmg_batch **out);
int mmg_batch_run(mmg_aligner *al, mmg_batch *b);
int mmg_batch_fetch(mmg_aligner *al, mmg_batch *b);

/* results: hits of read i are hits[hit_off[i] .. hit_off[i+1]) in the order
 * mm_map returns them (src/lib.rs:489-511 iterates in that order) */
uint32_t mmg_batch_n_reads(const mmg_batch *b);
uint64_t mmg_batch_n_hits(const mmg_batch *b);
const uint64_t *mmg_batch_hit_off(const mmg_batch *b);
const mmg_hit_t *mmg_batch_hits(const mmg_batch *b);
uint64_t mmg_batch_n_cigar(const mmg_batch *b);
const uint32_t *mmg_batch_cigar(const mmg_batch *b);
/* replaces mm_gen_cs(.., no_iden) / mm_gen_MD called by crate minimap2 when
 * cs/MD are requested (src/lib.rs:484-485, 589-590): writes a NUL-terminated
 * string for one hit given its CIGAR and the read; returns the length (call
 * with buf = NULL to size) or a negative code */
int mmg_gen_cs(const mmg_index *idx, const mmg_hit_t *hit, const uint32_t *cigar, const char *seq, int qlen, int no_iden, char *buf, size_t cap);
int mmg_gen_md(const mmg_index *idx, const mmg_hit_t *hit, const uint32_t *cigar, const char *seq, int qlen, char *buf, size_t cap);
/* the same for every hit of a batch at once (which: 0 = short cs, 1 = MD); strings are concatenated,
 * str_off has n_hits+1 entries; returns the total length, writes only if it fits cap */
int64_t mmg_gen_tags(const mmg_index *idx, const char *bases, const uint64_t *offsets, uint32_t n_reads, const uint64_t *hit_off,
                     const mmg_hit_t *hits, const uint32_t *cigar_pool, int which, int n_threads, char *buf, uint64_t cap, uint64_t *str_off);
/* replaces freeing reg.p and the mm_reg1_t array (crate minimap2 Aligner::map) */
void mmg_batch_destroy(mmg_batch *b);

/* per-batch counters (roofline denominators):
 * n_bases, n_mz, n_seed, n_hit, n_anchor, n_iter, n_kept, n_cell, n_regs, n_rechain,
 * n_dropped (isolated anchors removed before the sort; counted in n_anchor),
 * n_cell_fill (the cells of n_cell computed by the register-resident gap-fill kernel) */
#define MMG_N_STATS 12
int mmg_batch_stats(const mmg_batch *b, uint64_t out[MMG_N_STATS]);
/* device milliseconds per pipeline stage of the last mmg_batch_run ("profile"=1) and launches */
#define MMG_N_STAGES 12
int mmg_stage_times(const mmg_aligner *al, double ms[MMG_N_STAGES], uint64_t launches[MMG_N_STAGES]);
const char *mmg_stage_name(int stage);
/* device milliseconds of the last mmg_batch_run: CUDA events on the library's stream around all its kernels */
double mmg_last_run_ms(const mmg_aligner *al);

/* intermediate device arrays of one batch, copied to the host (differential
 * tests): which = 0 minimizers (x,y), 1 sorted anchors (x,y), 2 chained anchors
 * (x,y), 3 chains u[]; per-read offsets in `off` (n_reads+1). Returns count. */
int64_t mmg_debug_dump(mmg_aligner *al, mmg_batch *b, int which, uint64_t *x, uint64_t *y, uint64_t cap, uint64_t *off);

/* measurement hook: INT32 lane-operations per second (Gop/s) of dependent IADD3/LOP3 chains on `device`, the
 * denominator of the integer-issue rooflines (not in MEASURED_PEAKS.json) */
int mmg_debug_int32_peak(int device, double *gops);
/* test hook: the device logf used by the mapq computation, over an array */
int mmg_debug_logf(const float *x, float *y, uint64_t n);

const char *mmg_last_error(void);
const char *mmg_version(void);
int mmg_sizeof_hit(void);

#ifdef __cplusplus
}
#endif
#endif
