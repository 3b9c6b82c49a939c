/* index_host.cpp -- host side of the GPU-resident minimizer index.
 *
 * Loads a minimap2 `.mmi` v2 file or builds the index from sequences, into the
 * flat open-addressing layout of mmg_internal.h that mmg_aligner_create()
 * uploads once (north-star (a)).  Replaces, for the mappy-rs host,
 * mm_idx_reader_open/read/close (/root/reference/src/lib.rs:398-412),
 * mm_idx_name2id (:716), mm_idx_getseq (:747) and mm_mapopt_update's
 * mm_idx_cal_max_occ (:414).  Format: SURVEY.md appendix B.
 *
 * Index construction is a one-off outside the timed mapping path
 * (moving it to the sketch kernel + a device
 * sort is the next row); minimizers are selected here with the same
 * window-minimum formulation the sketch kernel uses (see sketch.cu).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include <stdarg.h>
#include <algorithm>
#include <thread>
#include <atomic>
#include "mmg_internal.h"

static thread_local char g_err[512] = "";
void mmg_set_error(const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
}
extern "C" const char *mmg_last_error(void) { return g_err; }

static inline int nt4(unsigned char c)
{
	switch (c) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': case 'U': case 'u': return 3;
	default: return 4;
	}
}

static inline uint64_t mix64(uint64_t key, uint64_t mask)
{ /* Thomas Wang's invertible integer hash restricted to 2k bits (sketch.c: hash64) */
	key = (~key + (key << 21)) & mask;
	key = key ^ key >> 24;
	key = ((key + (key << 3)) + (key << 8)) & mask;
	key = key ^ key >> 14;
	key = ((key + (key << 2)) + (key << 4)) & mask;
	key = key ^ key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}

/* (w,k)-minimizers in the window-minimum formulation.
 *
 * An "event" is every base that survives the symmetric-k-mer skip (ambiguous
 * bases are events carrying an infinite key).  After event e the selected
 * minimizer P(e) is the minimum key among the last w events, ties to the newest.
 * Records are written when P changes, exactly as sketch.c: mm_sketch() does,
 * including its duplicate-key rules and the l-counter conditions near the start
 * of a valid stretch.  Output x = hash<<8|span, y = rid<<32|pos<<1|strand. */
int mmg_host_sketch(const char *seq, int len, int w, int k, uint32_t rid, std::vector<uint64_t> &xs, std::vector<uint64_t> &ys)
{
	const uint64_t INF = ~0ULL, mask = (1ULL << 2 * k) - 1;
	const int shift1 = 2 * (k - 1);
	std::vector<uint64_t> rx(w, INF), ry(w, INF);
	uint64_t fwd = 0, rev = 0;
	int64_t e = 0;          /* event index */
	int l = 0;              /* valid, non-symmetric bases since the last ambiguous base */
	int64_t pidx = -1;      /* event index of the current selection, -1 = none/infinite */
	uint64_t px = INF, py = INF;
	if (w < 1 || w > 255 || k < 1 || k > 28) return -1;
	auto emit = [&](uint64_t x, uint64_t y) { xs.push_back(x); ys.push_back(y); };
	for (int i = 0; i < len; ++i) {
		int c = nt4((unsigned char)seq[i]);
		uint64_t ix = INF, iy = INF;
		if (c < 4) {
			fwd = (fwd << 2 | (uint64_t)c) & mask;
			rev = (rev >> 2) | (3ULL ^ (uint64_t)c) << shift1;
			if (fwd == rev) continue;      /* not an event */
			++l;
			if (l >= k) {
				int z = fwd < rev ? 0 : 1;
				ix = mix64(z ? rev : fwd, mask) << 8 | (uint64_t)k;
				iy = (uint64_t)rid << 32 | (uint32_t)i << 1 | (uint64_t)z;
			}
		} else l = 0;
		int slot = (int)(e % w);
		rx[slot] = ix, ry[slot] = iy;
		/* first full window of a stretch: earlier copies of the selected key */
		if (l == w + k - 1 && px != INF)
			for (int64_t j = e - w + 1; j < e; ++j)
				if (j >= 0 && rx[j % w] == px && ry[j % w] != py) emit(rx[j % w], ry[j % w]);
		if (ix <= px) {                          /* new selection at this event */
			if (l >= w + k && px != INF) emit(px, py);
			px = ix, py = iy, pidx = e;
		} else if (pidx == e - w) {               /* selection just left the window */
			if (l >= w + k - 1 && px != INF) emit(px, py);
			px = INF;
			for (int64_t j = e - w + 1; j <= e; ++j) {   /* newest among equal keys wins */
				uint64_t vx = j >= 0 ? rx[j % w] : INF, vy = j >= 0 ? ry[j % w] : INF;
				if (px >= vx) px = vx, py = vy, pidx = j;
			}
			if (l >= w + k - 1 && px != INF)
				for (int64_t j = e - w + 1; j <= e; ++j)
					if (j >= 0 && rx[j % w] == px && ry[j % w] != py) emit(rx[j % w], ry[j % w]);
		}
		++e;
	}
	if (px != INF) emit(px, py);
	return 0;
}

/* ------------------------------------------------------------------------- */

const uint64_t *mmg_index_lookup(const mmg_index *idx, uint64_t minier, int *n)
{
	*n = 0;
	if (mmg_index_ensure_host(const_cast<mmg_index*>(idx))) return 0;
	uint64_t m = ((uint64_t)1 << idx->hbits) - 1, s = mmg_hash_slot(minier, idx->hbits);
	*n = 0;
	for (;; s = (s + 1) & m) {
		uint64_t key = idx->hkeys[s];
		if (key == MMG_EMPTY_KEY) return 0;
		if (key >> 1 == minier) {
			if (key & 1) { *n = 1; return &idx->hvals[s]; }
			*n = (uint32_t)idx->hvals[s];
			return &idx->pos[idx->hvals[s] >> 32];
		}
	}
}

static void table_alloc(mmg_index *idx, uint64_t n_keys)
{
	uint32_t hb = 4;
	while (((uint64_t)1 << hb) < n_keys * 4) ++hb;
	idx->hbits = hb, idx->n_keys = n_keys;
	idx->hkeys.assign((size_t)1 << hb, MMG_EMPTY_KEY);
	idx->hvals.assign((size_t)1 << hb, 0);
}

static inline void table_put(mmg_index *idx, uint64_t minier, int single, uint64_t val)
{
	uint64_t m = ((uint64_t)1 << idx->hbits) - 1, s = mmg_hash_slot(minier, idx->hbits);
	while (idx->hkeys[s] != MMG_EMPTY_KEY) s = (s + 1) & m;
	idx->hkeys[s] = minier << 1 | (uint64_t)(single ? 1 : 0);
	idx->hvals[s] = val;
}

int32_t mmg_index_cal_max_occ(const mmg_index *idx, float f)
{ /* index.c: mm_idx_cal_max_occ -- ((1-f) n)-th smallest occurrence count + 1 */
	if (f <= 0.f) return INT32_MAX;
	if (!idx->occ_hist.empty()) { /* device build: the same order statistic from the occurrence histogram */
		if (idx->n_keys == 0) return INT32_MAX;
		uint64_t kk = (uint32_t)((1. - f) * idx->n_keys), acc = 0;
		for (size_t c = 0; c < idx->occ_hist.size(); ++c) {
			acc += idx->occ_hist[c];
			if (acc > kk) return (int32_t)(c + 1);
		}
		return (int32_t)(idx->occ_big[kk - acc] + 1);
	}
	std::vector<uint32_t> a;
	a.reserve(idx->n_keys);
	for (size_t s = 0; s < idx->hkeys.size(); ++s)
		if (idx->hkeys[s] != MMG_EMPTY_KEY) a.push_back(idx->hkeys[s] & 1 ? 1u : (uint32_t)idx->hvals[s]);
	if (a.empty()) return INT32_MAX;
	size_t kk = (uint32_t)((1. - f) * a.size());
	std::nth_element(a.begin(), a.begin() + kk, a.end());
	return (int32_t)(a[kk] + 1);
}

struct MzRec { uint64_t m, y; };

static mmg_index *build_from_seqs(int w, int k, int b, int flag, int n_seq, const char *const *names, const char *const *seqs, const uint32_t *lens, int n_threads, int device = 0)
{
	if (flag & MMG_I_HPC) { mmg_set_error("homopolymer-compressed indexes (MM_I_HPC, map-pb) are outside the supported path"); return 0; }
	if (!getenv("MMG_HOST_INDEX_BUILD")) { /* device construction whenever a device and the configuration allow it */
		mmg_index *didx = 0;
		int rc = mmg_index_build_device(w < 1 ? 1 : w, k, b, flag, n_seq, names, seqs, lens, device, &didx);
		if (rc == MMG_OK) return didx;
		if (rc != MMG_EUNSUP) return 0;
	}
	mmg_index *idx = new mmg_index();
	idx->k = k, idx->w = w < 1 ? 1 : w, idx->b = b, idx->flag = flag, idx->n_seq = n_seq;
	idx->offs.assign(n_seq + 1, 0);
	for (int i = 0; i < n_seq; ++i) {
		idx->names.push_back(names[i]);
		idx->lens.push_back(lens[i]);
		idx->offs[i + 1] = idx->offs[i] + lens[i];
	}
	uint64_t sum_len = idx->offs[n_seq];
	if (!(flag & MMG_I_NO_SEQ)) {
		idx->S.assign((sum_len + 7) / 8, 0);
		for (int i = 0; i < n_seq; ++i) {
			uint64_t o = idx->offs[i];
			for (uint32_t j = 0; j < lens[i]; ++j) {
				uint64_t p = o + j;
				idx->S[p >> 3] |= (uint32_t)nt4((unsigned char)seqs[i][j]) << ((p & 7) << 2);
			}
		}
	}
	/* sketch contigs in parallel */
	std::vector<std::vector<uint64_t> > xs(n_seq), ys(n_seq);
	std::atomic<int> next(0);
	auto work = [&]() {
		for (;;) {
			int i = next.fetch_add(1);
			if (i >= n_seq) break;
			if (lens[i] > 0) mmg_host_sketch(seqs[i], (int)lens[i], idx->w, idx->k, (uint32_t)i, xs[i], ys[i]);
		}
	};
	if (n_threads < 1) n_threads = 1;
	{
		std::vector<std::thread> th;
		int nt = std::min(n_threads, n_seq);
		for (int t = 1; t < nt; ++t) th.emplace_back(work);
		work();
		for (auto &t : th) t.join();
	}
	size_t n = 0;
	for (int i = 0; i < n_seq; ++i) n += xs[i].size();
	std::vector<MzRec> all;
	all.reserve(n);
	for (int i = 0; i < n_seq; ++i) {
		for (size_t j = 0; j < xs[i].size(); ++j) all.push_back(MzRec{ xs[i][j] >> 8, ys[i][j] });
		std::vector<uint64_t>().swap(xs[i]);
		std::vector<uint64_t>().swap(ys[i]);
	}
	std::sort(all.begin(), all.end(), [](const MzRec &a, const MzRec &c) { return a.m < c.m || (a.m == c.m && a.y < c.y); });
	uint64_t n_keys = 0, n_pos = 0;
	for (size_t i = 0; i < all.size();) {
		size_t j = i + 1;
		while (j < all.size() && all[j].m == all[i].m) ++j;
		++n_keys;
		if (j - i > 1) n_pos += j - i;
		i = j;
	}
	if (n_pos >> 32) { /* the flat table stores pos[] offsets in 32 bits (upstream's are per bucket) */
		mmg_set_error("index has %llu multi-occurrence positions: more than 2^32 are not supported", (unsigned long long)n_pos);
		delete idx;
		return 0;
	}
	table_alloc(idx, n_keys);
	idx->pos.reserve(n_pos);
	for (size_t i = 0; i < all.size();) {
		size_t j = i + 1;
		while (j < all.size() && all[j].m == all[i].m) ++j;
		if (j - i == 1) table_put(idx, all[i].m, 1, all[i].y);
		else {
			uint64_t off = idx->pos.size();
			for (size_t t = i; t < j; ++t) idx->pos.push_back(all[t].y);
			table_put(idx, all[i].m, 0, off << 32 | (uint64_t)(j - i));
		}
		i = j;
	}
	return idx;
}

static mmg_index *load_mmi(FILE *fp)
{
	uint32_t x[5];
	if (fread(x, 4, 5, fp) != 5) return 0;
	mmg_index *idx = new mmg_index();
	idx->w = x[0], idx->k = x[1], idx->b = x[2], idx->n_seq = x[3], idx->flag = x[4];
	idx->offs.assign(1, 0);
	bool ok = true;
	for (uint32_t i = 0; ok && i < idx->n_seq; ++i) {
		uint8_t l;
		uint32_t len;
		std::string name;
		ok = fread(&l, 1, 1, fp) == 1;
		if (ok && l) { name.resize(l); ok = fread(&name[0], 1, l, fp) == l; }
		ok = ok && fread(&len, 4, 1, fp) == 1;
		idx->names.push_back(name);
		idx->lens.push_back(len);
		idx->offs.push_back(idx->offs.back() + len);
	}
	/* buckets: first pass into memory (keys need the bucket id to become full minimizers) */
	std::vector<uint64_t> keys, vals; /* key = minier<<1|single ; val rebased into the global pos[] */
	for (uint32_t i = 0; ok && i < (1u << idx->b); ++i) {
		int32_t n;
		uint32_t size;
		ok = fread(&n, 4, 1, fp) == 1;
		if (!ok) break;
		uint64_t base = idx->pos.size();
		if ((base + (n > 0 ? (uint64_t)n : 0)) >> 32) { /* the flat table stores global pos[] offsets in 32 bits */
			mmg_set_error("index has more than 2^32 multi-occurrence positions: not supported");
			ok = false;
			break;
		}
		if (n > 0) {
			idx->pos.resize(base + n);
			ok = fread(&idx->pos[base], 8, n, fp) == (size_t)n;
		}
		ok = ok && fread(&size, 4, 1, fp) == 1;
		for (uint32_t j = 0; ok && j < size; ++j) {
			uint64_t kv[2];
			ok = fread(kv, 8, 2, fp) == 2;
			uint64_t minier = (kv[0] >> 1) << idx->b | i;
			keys.push_back(minier << 1 | (kv[0] & 1));
			vals.push_back((kv[0] & 1) ? kv[1] : ((kv[1] >> 32) + base) << 32 | (uint32_t)kv[1]);
		}
	}
	if (ok && !(idx->flag & MMG_I_NO_SEQ)) {
		idx->S.resize((idx->offs.back() + 7) / 8);
		ok = idx->S.empty() || fread(idx->S.data(), 4, idx->S.size(), fp) == idx->S.size();
	}
	if (!ok) { delete idx; return 0; }
	table_alloc(idx, keys.size());
	for (size_t i = 0; i < keys.size(); ++i) table_put(idx, keys[i] >> 1, (int)(keys[i] & 1), vals[i]);
	return idx;
}

/* FASTA / FASTQ, plain or gzip-compressed, as mm_idx_reader_open reads it through kseq + zlib
 * (/root/reference/src/lib.rs:398; gzread passes uncompressed files through).  The whole file is one index part:
 * mappy-rs sets batch_size to 2^63 - 1 (src/lib.rs:340). */
static mmg_index *load_fasta(const char *path, const mmg_idxopt_t *io, int n_threads)
{
	gzFile fp = gzopen(path, "rb");
	if (!fp) return 0;
	gzbuffer(fp, 1 << 20);
	std::string buf;
	std::vector<char> tmp(1 << 20);
	int nr;
	while ((nr = gzread(fp, tmp.data(), (unsigned)tmp.size())) > 0) buf.append(tmp.data(), (size_t)nr);
	const bool bad = nr < 0;
	gzclose(fp);
	if (bad) { mmg_set_error("error while reading (decompressing) '%s'", path); return 0; }
	std::vector<std::string> names, seqs;
	size_t qual_left = 0;       /* FASTQ: quality characters still to skip */
	bool in_qual = false;
	for (size_t i = 0; i < buf.size();) {
		size_t e = buf.find('\n', i);
		if (e == std::string::npos) e = buf.size();
		size_t l = e;
		while (l > i && (buf[l - 1] == '\r' || buf[l - 1] == ' ')) --l;
		if (in_qual) { /* kseq: the quality string is as long as the sequence, whatever it contains */
			qual_left = l - i >= qual_left ? 0 : qual_left - (l - i);
			if (qual_left == 0) in_qual = false;
		} else if (l > i && (buf[i] == '>' || buf[i] == '@')) {
			size_t p = i + 1;
			while (p < l && buf[p] != ' ' && buf[p] != '\t') ++p;
			names.push_back(buf.substr(i + 1, p - i - 1));
			seqs.push_back(std::string());
		} else if (l > i && buf[i] == '+' && !seqs.empty()) {
			qual_left = seqs.back().size(), in_qual = qual_left > 0;
		} else if (l > i && !seqs.empty()) seqs.back().append(buf, i, l - i);
		i = e + 1;
	}
	if (names.empty()) return 0;
	std::vector<const char*> np, sp;
	std::vector<uint32_t> ln;
	for (size_t j = 0; j < names.size(); ++j) np.push_back(names[j].c_str()), sp.push_back(seqs[j].c_str()), ln.push_back((uint32_t)seqs[j].size());
	return build_from_seqs(io->w, io->k, io->bucket_bits, io->flag, (int)names.size(), np.data(), sp.data(), ln.data(), n_threads);
}

extern "C" {

int mmg_index_open(const char *path, const mmg_idxopt_t *io, int n_threads, mmg_index **out)
{
	*out = 0;
	FILE *fp = fopen(path, "rb");
	if (!fp) { mmg_set_error("cannot open '%s'", path); return MMG_EIO; }
	char magic[4];
	size_t n = fread(magic, 1, 4, fp);
	mmg_index *idx = 0;
	if (n == 4 && memcmp(magic, "MMI\2", 4) == 0) {
		idx = load_mmi(fp);
		fclose(fp);
	} else {
		fclose(fp);
		idx = load_fasta(path, io, n_threads);
	}
	if (!idx) { if (!mmg_last_error()[0]) mmg_set_error("failed to read index or FASTA '%s'", path); return MMG_EIO; }
	*out = idx;
	return MMG_OK;
}

int mmg_index_build(const mmg_idxopt_t *io, int n_seq, const char *const *names, const char *const *seqs, const uint32_t *lens, int n_threads, mmg_index **out)
{
	*out = build_from_seqs(io->w, io->k, io->bucket_bits, io->flag, n_seq, names, seqs, lens, n_threads);
	return *out ? MMG_OK : MMG_EUNSUP;
}

int mmg_index_build_on(const mmg_idxopt_t *io, int n_seq, const char *const *names, const char *const *seqs, const uint32_t *lens, int n_threads, int device, mmg_index **out)
{
	*out = build_from_seqs(io->w, io->k, io->bucket_bits, io->flag, n_seq, names, seqs, lens, n_threads, device);
	return *out ? MMG_OK : MMG_EUNSUP;
}

void mmg_index_destroy(mmg_index *idx) { if (idx) { mmg_index_free_device(idx); delete idx; } }

int mmg_index_info(const mmg_index *idx, int32_t o[5])
{
	o[0] = idx->k, o[1] = idx->w, o[2] = idx->b, o[3] = idx->flag, o[4] = (int32_t)idx->n_seq;
	return MMG_OK;
}

const char *mmg_index_seq_name(const mmg_index *idx, uint32_t i) { return i < idx->n_seq ? idx->names[i].c_str() : 0; }
uint32_t mmg_index_seq_len(const mmg_index *idx, uint32_t i) { return i < idx->n_seq ? idx->lens[i] : 0; }

int mmg_index_name2id(const mmg_index *idx, const char *name)
{
	for (uint32_t i = 0; i < idx->n_seq; ++i)
		if (idx->names[i] == name) return (int)i;
	return -1;
}

int mmg_index_getseq(const mmg_index *idx, uint32_t rid, uint32_t st, uint32_t en, uint8_t *seq)
{
	if (rid >= idx->n_seq || st >= idx->lens[rid] || idx->S.empty()) return -1;
	if (en > idx->lens[rid]) en = idx->lens[rid];
	uint64_t o = idx->offs[rid];
	for (uint64_t i = o + st; i < o + en; ++i)
		seq[i - o - st] = idx->S[i >> 3] >> ((i & 7) << 2) & 0xf;
	return (int)(en - st);
}

uint64_t mmg_index_entries(const mmg_index *idx, uint64_t *minier, uint64_t *pos, uint64_t cap)
{
	uint64_t n = 0;
	if (mmg_index_ensure_host(const_cast<mmg_index*>(idx))) return 0;
	for (size_t s = 0; s < idx->hkeys.size(); ++s) {
		uint64_t key = idx->hkeys[s];
		if (key == MMG_EMPTY_KEY) continue;
		if (key & 1) {
			if (minier && n < cap) minier[n] = key >> 1, pos[n] = idx->hvals[s];
			++n;
		} else {
			uint32_t cnt = (uint32_t)idx->hvals[s];
			for (uint32_t j = 0; j < cnt; ++j) {
				if (minier && n < cap) minier[n] = key >> 1, pos[n] = idx->pos[(idx->hvals[s] >> 32) + j];
				++n;
			}
		}
	}
	return n;
}

int mmg_index_dump(const mmg_index *idx, const char *path)
{ /* index.c: mm_idx_dump -- .mmi v2; hash entries go out in key order per bucket */
	int rc0 = mmg_index_ensure_host(const_cast<mmg_index*>(idx));
	if (rc0) return rc0;
	FILE *fp = fopen(path, "wb");
	if (!fp) { mmg_set_error("cannot write '%s'", path); return MMG_EIO; }
	uint32_t x[5] = { (uint32_t)idx->w, (uint32_t)idx->k, (uint32_t)idx->b, idx->n_seq, (uint32_t)idx->flag };
	fwrite("MMI\2", 1, 4, fp);
	fwrite(x, 4, 5, fp);
	for (uint32_t i = 0; i < idx->n_seq; ++i) {
		uint8_t l = (uint8_t)idx->names[i].size();
		fwrite(&l, 1, 1, fp);
		fwrite(idx->names[i].data(), 1, l, fp);
		fwrite(&idx->lens[i], 4, 1, fp);
	}
	uint64_t bmask = ((uint64_t)1 << idx->b) - 1;
	std::vector<std::pair<uint64_t, uint64_t> > ent; /* (bucket<<? ...) sort by bucket then key */
	std::vector<uint64_t> order;
	for (size_t s = 0; s < idx->hkeys.size(); ++s) if (idx->hkeys[s] != MMG_EMPTY_KEY) order.push_back(s);
	std::sort(order.begin(), order.end(), [&](uint64_t a, uint64_t c) {
		uint64_t ma = idx->hkeys[a] >> 1, mc = idx->hkeys[c] >> 1;
		if ((ma & bmask) != (mc & bmask)) return (ma & bmask) < (mc & bmask);
		return ma < mc;
	});
	size_t o = 0;
	for (uint64_t bkt = 0; bkt <= bmask; ++bkt) {
		size_t o0 = o;
		std::vector<uint64_t> p;
		std::vector<uint64_t> kv;
		while (o < order.size() && ((idx->hkeys[order[o]] >> 1) & bmask) == bkt) {
			uint64_t key = idx->hkeys[order[o]], val = idx->hvals[order[o]];
			uint64_t k2 = ((key >> 1) >> idx->b) << 1 | (key & 1);
			if (!(key & 1)) {
				uint32_t cnt = (uint32_t)val;
				uint64_t off = p.size();
				for (uint32_t j = 0; j < cnt; ++j) p.push_back(idx->pos[(val >> 32) + j]);
				val = off << 32 | cnt;
			}
			kv.push_back(k2), kv.push_back(val);
			++o;
		}
		int32_t n = (int32_t)p.size();
		uint32_t size = (uint32_t)(o - o0);
		fwrite(&n, 4, 1, fp);
		if (n) fwrite(p.data(), 8, n, fp);
		fwrite(&size, 4, 1, fp);
		if (size) fwrite(kv.data(), 8, kv.size(), fp);
	}
	if (!(idx->flag & MMG_I_NO_SEQ)) fwrite(idx->S.data(), 4, idx->S.size(), fp);
	fclose(fp);
	return MMG_OK;
}

} // extern "C"
