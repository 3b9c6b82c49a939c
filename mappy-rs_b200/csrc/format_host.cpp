/* format_host.cpp -- cs / MD strings for finished hits (host marshalling).
 *
 * Replaces mm_gen_cs(km, &buf, &cap, mi, r, seq, no_iden=1) and mm_gen_MD that crate
 * minimap2 0.1.15 `Aligner::map` calls when cs/MD are requested
 * (/root/reference/src/lib.rs:484-485; the map_batch workers hard-wire cs=true,
 * src/lib.rs:589; upstream format.c write_cs_core / write_MD_core, minimap2 v2.26).
 * Both are pure functions of (CIGAR, reference bases, query bases); they run on
 * the host while results are turned into `Mapping` objects (SURVEY.md 8(f) rank 2
 * lists a device version as a next row).
 */
#include <string.h>
#include <string>
#include <vector>
#include <thread>
#include <atomic>
#include "mmg_internal.h"

static inline int nt4_(unsigned char c)
{
	switch (c) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': case 'U': case 'u': return 3;
	default: return 4;
	}
}

struct Aligned { std::vector<uint8_t> t, q; };

static int aligned_seqs(const mmg_index *idx, const mmg_hit_t *h, const char *seq, int qlen, Aligned &a)
{
	if (h->rid < 0 || (uint32_t)h->rid >= idx->n_seq || idx->S.empty()) return -1;
	if (h->qs < 0 || h->qe > qlen || h->qs > h->qe || h->rs < 0 || h->re > (int32_t)idx->lens[h->rid] || h->rs > h->re) return -1;
	const int tl = h->re - h->rs, ql = h->qe - h->qs;
	a.t.resize(tl + 1), a.q.resize(ql + 1);
	const uint64_t o = idx->offs[h->rid];
	for (int i = 0; i < tl; ++i) { uint64_t p = o + h->rs + i; a.t[i] = idx->S[p >> 3] >> ((p & 7) << 2) & 0xf; }
	if (!h->rev) for (int i = h->qs; i < h->qe; ++i) a.q[i - h->qs] = (uint8_t)nt4_((unsigned char)seq[i]);
	else for (int i = h->qs; i < h->qe; ++i) { int c = nt4_((unsigned char)seq[i]); a.q[h->qe - i - 1] = (uint8_t)(c >= 4 ? 4 : 3 - c); }
	return 0;
}

static void put_int(std::string &s, long v)
{ /* run lengths: a few digits, written without snprintf (one call per run of matches adds up over a batch) */
	char b[24];
	int n = 0;
	unsigned long u = v < 0 ? 0UL - (unsigned long)v : (unsigned long)v;
	do { b[n++] = (char)('0' + u % 10); u /= 10; } while (u);
	if (v < 0) b[n++] = '-';
	while (n) s += b[--n];
}

static int gen_cs(const mmg_index *idx, const mmg_hit_t *h, const uint32_t *cigar, const char *seq, int qlen, int no_iden, std::string &s)
{
	Aligned a;
	s.clear();
	if (aligned_seqs(idx, h, seq, qlen, a) < 0) return -1;
	int q_off = 0, t_off = 0;
	for (uint32_t i = 0; i < h->n_cigar; ++i) {
		const int op = cigar[i] & 0xf, len = (int)(cigar[i] >> 4);
		if (op == 0 || op == 7 || op == 8) {
			int l_tmp = 0;
			for (int j = 0; j < len; ++j) {
				if (a.q[q_off + j] != a.t[t_off + j]) {
					if (l_tmp > 0) {
						if (!no_iden) { s += '='; for (int k = j - l_tmp; k < j; ++k) s += "ACGTN"[a.q[q_off + k]]; }
						else { s += ':'; put_int(s, l_tmp); }
						l_tmp = 0;
					}
					s += '*', s += "acgtn"[a.t[t_off + j]], s += "acgtn"[a.q[q_off + j]];
				} else ++l_tmp;
			}
			if (l_tmp > 0) {
				if (!no_iden) { s += '='; for (int k = len - l_tmp; k < len; ++k) s += "ACGTN"[a.q[q_off + k]]; }
				else { s += ':'; put_int(s, l_tmp); }
			}
			q_off += len, t_off += len;
		} else if (op == 1) {
			s += '+';
			for (int j = 0; j < len; ++j) s += "acgtn"[a.q[q_off + j]];
			q_off += len;
		} else if (op == 2) {
			s += '-';
			for (int j = 0; j < len; ++j) s += "acgtn"[a.t[t_off + j]];
			t_off += len;
		} else if (op == 3) t_off += len;
	}
	return q_off == h->qe - h->qs && t_off == h->re - h->rs ? 0 : -1;
}

static int gen_md(const mmg_index *idx, const mmg_hit_t *h, const uint32_t *cigar, const char *seq, int qlen, std::string &s)
{
	Aligned a;
	s.clear();
	if (aligned_seqs(idx, h, seq, qlen, a) < 0) return -1;
	int q_off = 0, t_off = 0, l_MD = 0;
	for (uint32_t i = 0; i < h->n_cigar; ++i) {
		const int op = cigar[i] & 0xf, len = (int)(cigar[i] >> 4);
		if (op == 0 || op == 7 || op == 8) {
			for (int j = 0; j < len; ++j) {
				if (a.q[q_off + j] != a.t[t_off + j]) { put_int(s, l_MD); s += "ACGTN"[a.t[t_off + j]]; l_MD = 0; }
				else ++l_MD;
			}
			q_off += len, t_off += len;
		} else if (op == 1) q_off += len;
		else if (op == 2) {
			put_int(s, l_MD);
			s += '^';
			for (int j = 0; j < len; ++j) s += "ACGTN"[a.t[t_off + j]];
			l_MD = 0, t_off += len;
		} else if (op == 3) t_off += len;
	}
	if (l_MD > 0) put_int(s, l_MD);
	return 0;
}

extern "C" {

int mmg_gen_cs(const mmg_index *idx, const mmg_hit_t *hit, const uint32_t *cigar, const char *seq, int qlen, int no_iden, char *buf, size_t cap)
{
	std::string s;
	if (gen_cs(idx, hit, cigar, seq, qlen, no_iden, s) < 0) { mmg_set_error("cs: hit does not match the sequence/index"); return MMG_EINVAL; }
	if (buf && cap > s.size()) memcpy(buf, s.c_str(), s.size() + 1);
	return (int)s.size();
}

int mmg_gen_md(const mmg_index *idx, const mmg_hit_t *hit, const uint32_t *cigar, const char *seq, int qlen, char *buf, size_t cap)
{
	std::string s;
	if (gen_md(idx, hit, cigar, seq, qlen, s) < 0) { mmg_set_error("MD: hit does not match the sequence/index"); return MMG_EINVAL; }
	if (buf && cap > s.size()) memcpy(buf, s.c_str(), s.size() + 1);
	return (int)s.size();
}

/* cs (which = 0, short form) or MD (which = 1) of every hit of a batch, concatenated; str_off has n_hits+1
 * entries.  Returns the total length; nothing is written when it exceeds cap (call again with a larger buffer). */
int64_t mmg_gen_tags(const mmg_index *idx, const char *bases, const uint64_t *offsets, uint32_t n_reads, const uint64_t *hit_off,
                     const mmg_hit_t *hits, const uint32_t *cigar_pool, int which, int n_threads, char *buf, uint64_t cap, uint64_t *str_off)
{
	const uint64_t n_hits = hit_off[n_reads];
	std::vector<std::string> out(n_hits);
	std::atomic<uint32_t> next(0);
	std::atomic<int> bad(0);
	if (n_threads < 1) n_threads = 1;
	uint32_t step = n_reads / (4u * (uint32_t)n_threads); /* reads per claim: small batches still use every thread */
	step = step < 1 ? 1 : step > 64 ? 64 : step;
	auto work = [&]() {
		for (;;) {
			uint32_t r = next.fetch_add(step);
			if (r >= n_reads) break;
			for (uint32_t rr = r; rr < r + step && rr < n_reads; ++rr) {
				const char *seq = bases + offsets[rr];
				const int qlen = (int)(offsets[rr + 1] - offsets[rr]);
				for (uint64_t i = hit_off[rr]; i < hit_off[rr + 1]; ++i) {
					const mmg_hit_t *h = &hits[i];
					if (!(h->flags & 32)) continue;
					int rc = which == 0 ? gen_cs(idx, h, cigar_pool + h->cigar_off, seq, qlen, 1, out[i]) : gen_md(idx, h, cigar_pool + h->cigar_off, seq, qlen, out[i]);
					if (rc < 0) bad = 1;
				}
			}
		}
	};
	std::vector<std::thread> th;
	for (int t = 1; t < n_threads; ++t) th.emplace_back(work);
	work();
	for (auto &t : th) t.join();
	if (bad) { mmg_set_error("cs/MD: a hit does not match its sequence/index"); return MMG_EINVAL; }
	uint64_t tot = 0;
	for (uint64_t i = 0; i < n_hits; ++i) { if (str_off) str_off[i] = tot; tot += out[i].size(); }
	if (str_off) str_off[n_hits] = tot;
	if (buf && tot <= cap) { uint64_t o = 0; for (uint64_t i = 0; i < n_hits; ++i) { memcpy(buf + o, out[i].data(), out[i].size()); o += out[i].size(); } }
	return (int64_t)tot;
}

} // extern "C"
