/* mmg_internal.h -- shared declarations of the B200 mapping library (libmmg.so).
 * Host index (flat layout, ready for upload), device views, pipeline types. */
#ifndef MMG_INTERNAL_H
#define MMG_INTERNAL_H

#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>
#include "../../include/mmg.h"

#define MMG_EMPTY_KEY 0xffffffffffffffffULL

/* Host-side index in the layout that is uploaded verbatim.
 *
 * Upstream keeps 2^b khash tables (index.c: mm_idx_bucket_t); a GPU lookup wants
 * one probe = one 16-byte load, so the buckets are flattened into a single
 * open-addressing table (linear probing from an even slot, load factor <= 0.25, so that a lookup reads the
 * two slots of one 32-byte sector per round trip and almost always ends after the first):
 *   slot.key = minimizer<<1 | is_single   (MMG_EMPTY_KEY when free)
 *   slot.val = position word y            (is_single)
 *            = offset<<32 | count into pos[]  (otherwise; runs sorted ascending)
 * which is the same key/value encoding index.c: worker_post() stores. */
struct mmg_index {
	int32_t k, w, b, flag;
	uint32_t n_seq;
	std::vector<std::string> names;
	std::vector<uint32_t> lens;
	std::vector<uint64_t> offs;        /* n_seq + 1 */
	std::vector<uint32_t> S;           /* 4 bits per base (index.c: mm_seq4_set) */
	uint32_t hbits;                    /* table has 1<<hbits slots */
	std::vector<uint64_t> hkeys, hvals;
	std::vector<uint64_t> pos;
	uint64_t n_keys;
	/* An index built on the device (index_dev.cu) stays resident there: dev_* are the uploaded layout on
	 * device dev_device (-1 = none); the host vectors hkeys/hvals/pos are filled on demand (host_tables). */
	int dev_device;
	void *dev_htab; uint64_t *dev_pos; uint32_t *dev_S; uint64_t *dev_seq_off; uint32_t *dev_seq_len;
	uint64_t n_pos;
	bool host_tables;
	/* occurrence histogram taken by the device build (mm_idx_cal_max_occ): occ_hist[c] keys occur c times
	 * (c < 65536), occ_big = the sorted counts above that */
	std::vector<unsigned long long> occ_hist;
	std::vector<uint32_t> occ_big;
	mmg_index() : k(0), w(0), b(0), flag(0), n_seq(0), hbits(0), n_keys(0), dev_device(-1), dev_htab(0), dev_pos(0), dev_S(0),
	              dev_seq_off(0), dev_seq_len(0), n_pos(0), host_tables(true) {}
};

static inline uint64_t mmg_hash_slot(uint64_t minier, uint32_t hbits)
{
	return ((minier * 0x9E3779B97F4A7C15ULL) >> (64 - hbits)) & ~(uint64_t)1; /* probing starts at an even slot */
}

/* index_host.cpp */
int mmg_host_sketch(const char *seq, int len, int w, int k, uint32_t rid, std::vector<uint64_t> &xs, std::vector<uint64_t> &ys);
const uint64_t *mmg_index_lookup(const mmg_index *idx, uint64_t minier, int *n);
int32_t mmg_index_cal_max_occ(const mmg_index *idx, float f);
void mmg_set_error(const char *fmt, ...);
/* index_dev.cu */
int mmg_index_build_device(int w, int k, int b, int flag, int n_seq, const char *const *names, const char *const *seqs, const uint32_t *lens,
                           int device, mmg_index **out);
int mmg_index_ensure_host(mmg_index *idx);
void mmg_index_free_device(mmg_index *idx);
/* stream_host.cpp */
void mmg_stream_shutdown(mmg_aligner *al);

#endif
