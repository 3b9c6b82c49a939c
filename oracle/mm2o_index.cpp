/* mm2o_index.cpp -- ORACLE (test infrastructure only).
 * Restates minimap2 v2.26 index.c: mm_idx_load / mm_idx_dump (.mmi v2 format,
 * SURVEY.md appendix B), mm_idx_gen + worker_post (index construction),
 * mm_idx_get, mm_idx_cal_max_occ, mm_idx_getseq, mm_idx_name2id.
 * Reference call sites: /root/reference/src/lib.rs:398-412 (reader open/read/
 * close), :414 (mm_mapopt_update -> mm_idx_cal_max_occ), :716 (name2id),
 * :747 (getseq).  Pinned by resources/test/test.mmi <-> test.fa.
 *
 * Upstream keeps one khash per bucket; its slot order is implementation-defined
 * and unused on the mapping path, so buckets here are sorted key tables.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include "mm2o.h"
#include "mm2o_sort.h"
#include <thread>
#include <atomic>

#define MM_IDX_MAGIC "MMI\2"

#define mm_seq4_set(s, i, c) ((s)[(i)>>3] |= (uint32_t)(c) << (((i)&7)<<2))
#define mm_seq4_get(s, i)    ((s)[(i)>>3] >> (((i)&7)<<2) & 0xf)

void mm_idx_destroy(mm_idx_t *mi) { delete mi; }

const uint64_t *mm_idx_get(const mm_idx_t *mi, uint64_t minier, int *n)
{
	int mask = (1 << mi->b) - 1;
	const mm_idx_bucket_t *b = &mi->B[minier & mask];
	*n = 0;
	if (b->keys.empty()) return 0;
	uint64_t key = minier >> mi->b << 1;
	// kh_get with idx_eq(a, b) = (a>>1 == b>>1)
	std::vector<uint64_t>::const_iterator it = std::lower_bound(b->keys.begin(), b->keys.end(), key);
	if (it == b->keys.end() || (*it >> 1) != (key >> 1)) return 0;
	size_t k = it - b->keys.begin();
	if (b->keys[k] & 1) { // special casing when there is only one k-mer
		*n = 1;
		return &b->vals[k];
	} else {
		*n = (uint32_t)b->vals[k];
		return &b->p[b->vals[k] >> 32];
	}
}

int32_t mm_idx_cal_max_occ(const mm_idx_t *mi, float f)
{
	size_t n = 0;
	uint32_t thres;
	if (f <= 0.) return INT32_MAX;
	for (int i = 0; i < 1 << mi->b; ++i) n += mi->B[i].keys.size();
	if (n == 0) return INT32_MAX;
	std::vector<uint32_t> a(n);
	n = 0;
	for (int i = 0; i < 1 << mi->b; ++i) {
		const mm_idx_bucket_t *b = &mi->B[i];
		for (size_t k = 0; k < b->keys.size(); ++k)
			a[n++] = b->keys[k] & 1 ? 1 : (uint32_t)b->vals[k];
	}
	thres = ks_ksmall_uint32_t(n, a.data(), (uint32_t)((1. - f) * n)) + 1;
	return thres;
}

int mm_idx_getseq(const mm_idx_t *mi, uint32_t rid, uint32_t st, uint32_t en, uint8_t *seq)
{
	uint64_t i, st1, en1;
	if (rid >= mi->n_seq || st >= mi->seq[rid].len) return -1;
	if (en > mi->seq[rid].len) en = mi->seq[rid].len;
	st1 = mi->seq[rid].offset + st;
	en1 = mi->seq[rid].offset + en;
	for (i = st1; i < en1; ++i)
		seq[i - st1] = mm_seq4_get(mi->S, i);
	return en - st;
}

int mm_idx_getseq_rev(const mm_idx_t *mi, uint32_t rid, uint32_t st, uint32_t en, uint8_t *seq)
{
	uint64_t i, st1, en1;
	const mm_idx_seq_t *s;
	if (rid >= mi->n_seq || st >= mi->seq[rid].len) return -1;
	s = &mi->seq[rid];
	if (en > s->len) en = s->len;
	st1 = s->offset + (s->len - en);
	en1 = s->offset + (s->len - st);
	for (i = st1; i < en1; ++i) {
		uint8_t c = mm_seq4_get(mi->S, i);
		seq[en1 - i - 1] = c < 4 ? 3 - c : c;
	}
	return en - st;
}

int mm_idx_name2id(const mm_idx_t *mi, const char *name)
{
	for (uint32_t i = 0; i < mi->n_seq; ++i)
		if (mi->seq[i].name == name) return (int)i;
	return -1;
}

/* index.c: worker_post -- turn the minimizer list of one bucket into its table */
static void bucket_post(mm_idx_t *mi, mm_idx_bucket_t *b, mm128_v &a)
{
	size_t j, start_a, n;
	if (a.empty()) return;
	radix_sort_128x(a.data(), a.data() + a.size());
	for (j = 1, n = 1, start_a = 0; j <= a.size(); ++j) {
		if (j == a.size() || a[j].x >> 8 != a[j - 1].x >> 8) {
			const mm128_t *p = &a[j - 1];
			uint64_t key = p->x >> 8 >> mi->b << 1;
			if (n == 1) {
				b->keys.push_back(key | 1);
				b->vals.push_back(p->y);
			} else {
				size_t start_p = b->p.size();
				for (size_t k = 0; k < n; ++k) b->p.push_back(a[start_a + k].y);
				radix_sort_64(&b->p[start_p], &b->p[start_p] + n); // sort by position; needed as in-place radix_sort_128x() is not stable
				b->keys.push_back(key);
				b->vals.push_back((uint64_t)start_p << 32 | n);
			}
			start_a = j, n = 1;
		} else ++n;
	}
	// a[] is sorted by x, hence keys are already ascending by key>>1
	a.clear();
}

/* index.c: mm_idx_gen (single part; mappy-rs only reads the first part,
 * src/lib.rs:407) */
mm_idx_t *mm_idx_build(int w, int k, int b, int flag, int n_seq, const char **names, const char **seqs, const uint32_t *lens)
{
	mm_idx_t *mi = new mm_idx_t();
	mi->w = w < 1 ? 1 : w, mi->k = k, mi->b = b, mi->flag = flag, mi->n_seq = n_seq, mi->n_alt = 0;
	mi->B.resize((size_t)1 << b);
	uint64_t sum_len = 0;
	for (int i = 0; i < n_seq; ++i) {
		mm_idx_seq_t s;
		s.name = names[i], s.offset = sum_len, s.len = lens[i], s.is_alt = 0;
		mi->seq.push_back(s);
		sum_len += lens[i];
	}
	if (!(flag & MM_I_NO_SEQ)) {
		mi->S.assign((sum_len + 7) / 8, 0);
		for (int i = 0; i < n_seq; ++i) {
			uint64_t o = mi->seq[i].offset;
			for (uint32_t j = 0; j < lens[i]; ++j) {
				uint64_t c = seq_nt4_table[(uint8_t)seqs[i][j]];
				mm_seq4_set(mi->S, o + j, c);
			}
		}
	}
	std::vector<mm128_v> A((size_t)1 << b);
	int mask = (1 << b) - 1;
	/* The contigs are sketched and the buckets post-processed by a few threads (upstream: kt_pipeline /
	 * kt_for); the index does not depend on the order in which minimizers reach a bucket, because every
	 * position run is sorted in bucket_post. */
	int n_thr = (int)std::thread::hardware_concurrency();
	if (n_thr < 1) n_thr = 1;
	if (n_thr > 32) n_thr = 32;
	if (sum_len < 50000000) n_thr = 1;
	std::vector<mm128_v> per(n_seq);
	{
		std::atomic<int> next(0);
		auto work = [&]() {
			for (;;) {
				int i = next.fetch_add(1);
				if (i >= n_seq) break;
				if (lens[i] > 0) mm_sketch(seqs[i], lens[i], mi->w, mi->k, i, flag & MM_I_HPC, &per[i]);
			}
		};
		std::vector<std::thread> th;
		for (int t = 1; t < n_thr && t < n_seq; ++t) th.emplace_back(work);
		work();
		for (auto &t : th) t.join();
	}
	for (int i = 0; i < n_seq; ++i) {
		const mm128_v &a = per[i];
		for (size_t j = 0; j < a.size(); ++j) // index.c: mm_idx_add
			A[a[j].x >> 8 & mask].push_back(a[j]);
		mm128_v().swap(per[i]);
	}
	{
		std::atomic<size_t> next(0);
		auto work = [&]() {
			for (;;) {
				size_t i = next.fetch_add(64);
				if (i >= A.size()) break;
				for (size_t q = i; q < i + 64 && q < A.size(); ++q) { bucket_post(mi, &mi->B[q], A[q]); mm128_v().swap(A[q]); }
			}
		};
		std::vector<std::thread> th;
		for (int t = 1; t < n_thr; ++t) th.emplace_back(work);
		work();
		for (auto &t : th) t.join();
	}
	return mi;
}

mm_idx_t *mm_idx_from_fasta(const char *fn, int w, int k, int b, int flag)
{
	FILE *fp = fopen(fn, "rb");
	if (!fp) return 0;
	std::string buf;
	char tmp[65536];
	size_t nr;
	while ((nr = fread(tmp, 1, sizeof(tmp), fp)) > 0) buf.append(tmp, nr);
	fclose(fp);
	std::vector<std::string> names, seqs;
	size_t i = 0;
	while (i < buf.size()) {
		size_t e = buf.find('\n', i);
		if (e == std::string::npos) e = buf.size();
		size_t l = e;
		while (l > i && (buf[l - 1] == '\r' || buf[l - 1] == ' ')) --l;
		if (l > i && buf[i] == '>') { // name = header up to the first whitespace (kseq.h)
			size_t p = i + 1;
			while (p < l && buf[p] != ' ' && buf[p] != '\t') ++p;
			names.push_back(buf.substr(i + 1, p - i - 1));
			seqs.push_back(std::string());
		} else if (l > i && !seqs.empty()) seqs.back().append(buf, i, l - i);
		i = e + 1;
	}
	std::vector<const char*> np, sp;
	std::vector<uint32_t> ln;
	for (size_t j = 0; j < names.size(); ++j) np.push_back(names[j].c_str()), sp.push_back(seqs[j].c_str()), ln.push_back((uint32_t)seqs[j].size());
	return mm_idx_build(w, k, b, flag, (int)names.size(), np.data(), sp.data(), ln.data());
}

/* index.c: mm_idx_load -- .mmi v2 (one part) */
mm_idx_t *mm_idx_load(const char *fn)
{
	FILE *fp = fopen(fn, "rb");
	if (!fp) return 0;
	char magic[4];
	uint32_t x[5];
	uint64_t sum_len = 0;
	if (fread(magic, 1, 4, fp) != 4 || strncmp(magic, MM_IDX_MAGIC, 4) != 0) { fclose(fp); return 0; }
	if (fread(x, 4, 5, fp) != 5) { fclose(fp); return 0; }
	mm_idx_t *mi = new mm_idx_t();
	mi->w = x[0], mi->k = x[1], mi->b = x[2], mi->n_seq = x[3], mi->flag = x[4], mi->n_alt = 0;
	mi->B.resize((size_t)1 << mi->b);
	for (uint32_t i = 0; i < mi->n_seq; ++i) {
		uint8_t l;
		mm_idx_seq_t s;
		if (fread(&l, 1, 1, fp) != 1) goto fail;
		if (l) {
			s.name.resize(l);
			if (fread(&s.name[0], 1, l, fp) != l) goto fail;
		}
		if (fread(&s.len, 4, 1, fp) != 1) goto fail;
		s.offset = sum_len, s.is_alt = 0;
		sum_len += s.len;
		mi->seq.push_back(s);
	}
	for (int i = 0; i < 1 << mi->b; ++i) {
		mm_idx_bucket_t *b = &mi->B[i];
		int32_t n;
		uint32_t size;
		if (fread(&n, 4, 1, fp) != 1) goto fail;
		b->p.resize(n);
		if (n && fread(b->p.data(), 8, n, fp) != (size_t)n) goto fail;
		if (fread(&size, 4, 1, fp) != 1) goto fail;
		if (size == 0) continue;
		std::vector<std::pair<uint64_t, uint64_t> > kv(size);
		for (uint32_t j = 0; j < size; ++j) {
			uint64_t y[2];
			if (fread(y, 8, 2, fp) != 2) goto fail;
			kv[j].first = y[0], kv[j].second = y[1];
		}
		std::sort(kv.begin(), kv.end());
		b->keys.resize(size), b->vals.resize(size);
		for (uint32_t j = 0; j < size; ++j) b->keys[j] = kv[j].first, b->vals[j] = kv[j].second;
	}
	if (!(mi->flag & MM_I_NO_SEQ)) {
		mi->S.resize((sum_len + 7) / 8);
		if (!mi->S.empty() && fread(mi->S.data(), 4, mi->S.size(), fp) != mi->S.size()) goto fail;
	}
	fclose(fp);
	return mi;
fail:
	fclose(fp);
	delete mi;
	return 0;
}

/* index.c: mm_idx_dump */
int mm_idx_dump(const char *fn, const mm_idx_t *mi)
{
	FILE *fp = fopen(fn, "wb");
	if (!fp) return -1;
	uint64_t sum_len = 0;
	uint32_t x[5];
	x[0] = mi->w, x[1] = mi->k, x[2] = mi->b, x[3] = mi->n_seq, x[4] = mi->flag;
	fwrite(MM_IDX_MAGIC, 1, 4, fp);
	fwrite(x, 4, 5, fp);
	for (uint32_t i = 0; i < mi->n_seq; ++i) {
		uint8_t l = (uint8_t)mi->seq[i].name.size();
		fwrite(&l, 1, 1, fp);
		fwrite(mi->seq[i].name.data(), 1, l, fp);
		fwrite(&mi->seq[i].len, 4, 1, fp);
		sum_len += mi->seq[i].len;
	}
	for (int i = 0; i < 1 << mi->b; ++i) {
		const mm_idx_bucket_t *b = &mi->B[i];
		int32_t n = (int32_t)b->p.size();
		uint32_t size = (uint32_t)b->keys.size();
		fwrite(&n, 4, 1, fp);
		fwrite(b->p.data(), 8, n, fp);
		fwrite(&size, 4, 1, fp);
		for (uint32_t j = 0; j < size; ++j) {
			uint64_t y[2] = { b->keys[j], b->vals[j] };
			fwrite(y, 8, 2, fp);
		}
	}
	if (!(mi->flag & MM_I_NO_SEQ))
		fwrite(mi->S.data(), 4, (sum_len + 7) / 8, fp);
	fclose(fp);
	return 0;
}
