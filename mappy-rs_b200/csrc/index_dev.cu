/* index_dev.cu -- minimizer index construction on the device.
 *
 * Replaces mm_idx_gen() behind `Aligner("ref.fa")` / mm_idx_reader_read on a FASTA
 * (/root/reference/src/lib.rs:395-410; upstream index.c mm_idx_gen, worker_pipeline,
 * worker_post, minimap2 v2.26): sketch every contig, group the minimizers by key, store one
 * table entry per distinct key and the position runs sorted ascending.
 *
 * Device formulation:
 *   1. contigs are uploaded once; the 4-bit packed sequence S[] is produced from them on the device;
 *   2. contigs are cut into segments that the mapping path's own sketch kernel (sketch.cu, segment mode)
 *      processes as independent "reads": a segment's run starts w+k+8 bases early and ends w+k+8 bases late.
 *      The selection after an event depends on the last w events and on l (saturating at w+k) only, and a
 *      selected minimizer is written at most w events after its own position, so every record whose position
 *      lies inside the segment proper is produced exactly as a whole-contig run produces it; records outside
 *      [seg_lo, seg_hi) belong to a neighbour and are dropped.  (Needs odd k, where no k-mer equals its
 *      reverse complement; even k takes the host builder.)
 *   3. (key, position) pairs are sorted by position, then stably by key (two LSD radix sorts: CUB, a library
 *      sort as cuBLAS is a library GEMM; this is one-off setup, not the mapping hot path);
 *   4. run heads give the distinct keys; a scan gives each multi-occurrence key its slice of pos[]; the
 *      open-addressing table (mmg_internal.h) is filled with atomicCAS inserts;
 *   5. the occurrence histogram mm_idx_cal_max_occ needs is taken on the way.
 * The index stays resident on the device it was built on; mmg_aligner_create() on that device uses it in
 * place, and host copies of the tables are downloaded only if a host-side consumer asks (dump, entries).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include "mmg_internal.h"
#include "dev_common.cuh"
#include "stages.h"
#ifndef MMG_EMU
#include <cub/cub.cuh>
#endif

#define CKI(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { mmg_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); rc = MMG_ECUDA; goto fail; } } while (0)

#define IDX_SEG_MAIN 32768
#define IDX_TILE 4096          /* elements per block in the head/rank kernels (256 threads x 16) */

/* ---- 1. 4-bit packing of the concatenated contigs (index.c: mm_seq4_set) ---- */
__global__ void pack4_kernel(const char *seq, uint64_t n, uint32_t *S)
{
	const uint64_t wi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, p0 = wi * 8;
	if (p0 >= n) return;
	uint32_t v = 0;
	for (int j = 0; j < 8 && p0 + j < n; ++j) {
		const unsigned c = (unsigned char)seq[p0 + j], d = (c & 0xdfu) - 'A';
		const bool ok = d < 21u && ((0x180045u >> d) & 1u);
		unsigned code = (c >> 1) & 3u;
		code ^= code >> 1;
		v |= (ok ? code : 4u) << (4 * j);
	}
	S[wi] = v;
}

/* ---- 2. records of a segment that lie in [lo, hi) -> (key, rid<<32 | pos<<1 | strand) ---- */
struct SegMeta { uint32_t rid, run0, lo, hi; }; /* contig, contig coordinate of the run's first base, kept range */

__global__ void __launch_bounds__(256)
seg_filter_kernel(const uint64_t *mz_x, const uint32_t *mz_y, const uint32_t *n_mz, const SegMeta *meta, uint32_t seg_cap, uint32_t n_seg,
                  const uint64_t *keep_off, uint32_t *n_keep, uint64_t *out_key, uint64_t *out_val, uint64_t out_base, int write)
{
	const int lane = mmg_lane();
	const uint32_t lt = mmg_lanemask_lt();
	const uint32_t r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
	if (r >= n_seg) return;
	const SegMeta m = meta[r];
	const uint64_t *x = mz_x + (uint64_t)r * seg_cap;
	const uint32_t *y = mz_y + (uint64_t)r * seg_cap;
	const int n = (int)n_mz[r];
	uint64_t o = write ? out_base + keep_off[r] : 0;
	int cnt = 0;
	for (int j0 = 0; j0 < n; j0 += 32) {
		const int j = j0 + lane;
		uint32_t yy = j < n ? y[j] : 0, pos = m.run0 + (yy >> 1);
		const bool keep = j < n && pos >= m.lo && pos < m.hi;
		const uint32_t km = __ballot_sync(MMG_FULL, keep);
		if (write && keep) {
			const uint64_t d = o + cnt + __popc(km & lt);
			out_key[d] = x[j] >> 8;
			out_val[d] = (uint64_t)m.rid << 32 | (uint64_t)pos << 1 | (yy & 1u);
		}
		cnt += __popc(km);
	}
	if (!write && lane == 0) n_keep[r] = (uint32_t)cnt;
}

/* ---- 4. heads, ranks, table ---- */
/* per tile: number of run heads and of elements of multi-occurrence runs */
__global__ void __launch_bounds__(256)
head_count_kernel(const uint64_t *key, uint64_t n, uint32_t *blk_heads, uint32_t *blk_multi)
{
	__shared__ uint32_t s_h, s_m;
	if (threadIdx.x == 0) s_h = 0, s_m = 0;
	__syncthreads();
	const uint64_t t0 = (uint64_t)blockIdx.x * IDX_TILE;
	uint32_t h = 0, m = 0;
	for (int j = 0; j < IDX_TILE / 256; ++j) {
		const uint64_t i = t0 + (uint64_t)j * 256 + threadIdx.x;
		if (i < n) {
			const uint64_t k = key[i];
			const bool head = i == 0 || key[i - 1] != k, tail = i + 1 == n || key[i + 1] != k;
			h += head, m += !(head && tail);
		}
	}
	atomicAdd(&s_h, h), atomicAdd(&s_m, m);
	__syncthreads();
	if (threadIdx.x == 0) blk_heads[blockIdx.x] = s_h, blk_multi[blockIdx.x] = s_m;
}

/* block-wide exclusive scan of one value per thread (256 threads) */
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t *s_warp, uint32_t *total)
{
	const int lane = mmg_lane(), wib = threadIdx.x >> 5;
	uint32_t x = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t y = __shfl_up_sync(MMG_FULL, x, o);
		if (lane >= o) x += y;
	}
	if (lane == 31) s_warp[wib] = x;
	__syncthreads();
	uint32_t pre = 0, tot = 0;
	for (int q = 0; q < 8; ++q) { uint32_t t = s_warp[q]; if (q < wib) pre += t; tot += t; }
	__syncthreads();
	*total = tot;
	return pre + x - v;
}

/* head_pos[rank] = index of the run head, head_off[rank] = offset of the run in pos[];
 * pos[] receives the positions of multi-occurrence runs in sorted order */
__global__ void __launch_bounds__(256)
head_rank_kernel(const uint64_t *key, const uint64_t *val, uint64_t n, const uint64_t *blk_heads_off, const uint64_t *blk_multi_off,
                 uint64_t *head_pos, uint64_t *head_off, uint64_t *pos)
{
	__shared__ uint32_t s_warp[8];
	const uint64_t t0 = (uint64_t)blockIdx.x * IDX_TILE;
	uint64_t hbase = blk_heads_off[blockIdx.x], mbase = blk_multi_off[blockIdx.x];
	/* thread t owns the 16 consecutive elements t0 + 16 t .. so that ranks follow the element order */
	const uint64_t i0 = t0 + (uint64_t)threadIdx.x * (IDX_TILE / 256);
	uint32_t hmask = 0, mmask = 0;
	for (int j = 0; j < IDX_TILE / 256; ++j) {
		const uint64_t i = i0 + j;
		if (i < n) {
			const uint64_t k = key[i];
			const bool head = i == 0 || key[i - 1] != k, tail = i + 1 == n || key[i + 1] != k;
			hmask |= (uint32_t)head << j, mmask |= (uint32_t)!(head && tail) << j;
		}
	}
	uint32_t th, tm;
	uint32_t ho = block_excl_scan_256((uint32_t)__popc(hmask), s_warp, &th);
	uint32_t mo = block_excl_scan_256((uint32_t)__popc(mmask), s_warp, &tm);
	for (int j = 0; j < IDX_TILE / 256; ++j) {
		const uint64_t i = i0 + j;
		if (hmask >> j & 1u) { head_pos[hbase + ho] = i, head_off[hbase + ho] = mbase + mo; ++ho; }
		if (mmask >> j & 1u) { pos[mbase + mo] = val[i]; ++mo; }
	}
}

#define OCC_HIST 65536
__global__ void __launch_bounds__(256)
table_fill_kernel(const uint64_t *key, const uint64_t *val, const uint64_t *head_pos, const uint64_t *head_off, uint64_t n_keys,
                  mmg_u128 *tab, uint32_t hbits, unsigned long long *hist, unsigned long long *n_big, uint32_t *big_cnt, uint32_t big_cap)
{
	const uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const bool act = h < n_keys;
	uint64_t cnt = 0;
	if (act) {
		const uint64_t i = head_pos[h], k = key[i];
		cnt = head_pos[h + 1] - i;
		const uint64_t kv = cnt == 1 ? (k << 1 | 1ULL) : (k << 1), vv = cnt == 1 ? val[i] : (head_off[h] << 32 | cnt);
		const uint64_t m = ((uint64_t)1 << hbits) - 1;
		uint64_t s = ((k * 0x9E3779B97F4A7C15ULL) >> (64 - hbits)) & ~(uint64_t)1;
		for (;; s = (s + 1) & m) {
			unsigned long long old = atomicCAS((unsigned long long*)&tab[s].x, (unsigned long long)MMG_EMPTY_KEY, (unsigned long long)kv);
			if (old == (unsigned long long)MMG_EMPTY_KEY) { tab[s].y = vv; break; }
		}
	}
	/* occurrence histogram (mm_idx_cal_max_occ): singletons are the bulk, count them per warp */
	const uint32_t ones = __ballot_sync(MMG_FULL, act && cnt == 1);
	if (mmg_lane() == 0 && ones) atomicAdd(&hist[1], (unsigned long long)__popc(ones));
	if (act && cnt > 1) {
		if (cnt < OCC_HIST) atomicAdd(&hist[cnt], 1ULL);
		else { unsigned long long q = atomicAdd(n_big, 1ULL); if (q < big_cap) big_cnt[q] = (uint32_t)cnt; }
	}
}

__global__ void fill_u64_kernel(uint64_t *p, uint64_t v, uint64_t n)
{
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void fill_tab_kernel(mmg_u128 *t, uint64_t n)
{
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) t[i].x = MMG_EMPTY_KEY, t[i].y = 0;
}

/* ---- 3. sort of (key, val) by val, then stably by key ---- */
static int sort_pairs(uint64_t **key, uint64_t **val, uint64_t **key_alt, uint64_t **val_alt, uint64_t n, int key_bits, int val_bits)
{
	if (n < 2) return MMG_OK;
#ifdef MMG_EMU
	std::vector<std::pair<uint64_t, uint64_t> > a(n);
	for (uint64_t i = 0; i < n; ++i) a[i] = std::make_pair((*key)[i], (*val)[i]);
	std::sort(a.begin(), a.end());
	for (uint64_t i = 0; i < n; ++i) (*key)[i] = a[i].first, (*val)[i] = a[i].second;
	return MMG_OK;
#else
	if (n > 0x7fffffffULL) { mmg_set_error("index too large for the device sort (%llu minimizers)", (unsigned long long)n); return MMG_EUNSUP; }
	void *tmp = 0;
	size_t tmp_bytes = 0, need = 0;
	cub::DoubleBuffer<uint64_t> kb(*key, *key_alt), vb(*val, *val_alt);
	cub::DeviceRadixSort::SortPairs(0, tmp_bytes, vb, kb, (int)n, 0, val_bits);
	cub::DeviceRadixSort::SortPairs(0, need, kb, vb, (int)n, 0, key_bits);
	if (need > tmp_bytes) tmp_bytes = need;
	if (cudaMalloc(&tmp, tmp_bytes + 16) != cudaSuccess) { mmg_set_error("cudaMalloc failed for the sort scratch"); return MMG_ENOMEM; }
	cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, vb, kb, (int)n, 0, val_bits);       /* by position */
	if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kb, vb, (int)n, 0, key_bits); /* stably by key */
	if (e == cudaSuccess) e = cudaDeviceSynchronize();
	cudaFree(tmp);
	if (e != cudaSuccess) { mmg_set_error("device radix sort failed: %s", cudaGetErrorString(e)); return MMG_ECUDA; }
	*key = kb.Current(), *key_alt = kb.Alternate(), *val = vb.Current(), *val_alt = vb.Alternate();
	return MMG_OK;
#endif
}

static int bits_of(uint64_t v) { int b = 0; while (v) ++b, v >>= 1; return b ? b : 1; }

void mmg_index_free_device(mmg_index *idx)
{
	if (idx->dev_device < 0) return;
	int cur = 0;
	cudaGetDevice(&cur);
	cudaSetDevice(idx->dev_device);
	cudaFree(idx->dev_htab), cudaFree(idx->dev_pos), cudaFree(idx->dev_S), cudaFree(idx->dev_seq_off), cudaFree(idx->dev_seq_len);
	idx->dev_htab = 0, idx->dev_pos = 0, idx->dev_S = 0, idx->dev_seq_off = 0, idx->dev_seq_len = 0, idx->dev_device = -1;
	cudaSetDevice(cur);
}

/* host copies of the device-resident tables, for the host-side consumers (dump, entries, lookup) */
int mmg_index_ensure_host(mmg_index *idx)
{
	if (idx->host_tables || idx->dev_device < 0) return MMG_OK;
	int cur = 0;
	cudaGetDevice(&cur);
	cudaSetDevice(idx->dev_device);
	const size_t ns = (size_t)1 << idx->hbits;
	std::vector<mmg_u128> tab(ns);
	bool ok = cudaMemcpy(tab.data(), idx->dev_htab, ns * sizeof(mmg_u128), cudaMemcpyDeviceToHost) == cudaSuccess;
	idx->hkeys.resize(ns), idx->hvals.resize(ns);
	for (size_t i = 0; i < ns; ++i) idx->hkeys[i] = tab[i].x, idx->hvals[i] = tab[i].y;
	idx->pos.resize(idx->n_pos);
	if (ok && idx->n_pos) ok = cudaMemcpy(idx->pos.data(), idx->dev_pos, idx->n_pos * 8, cudaMemcpyDeviceToHost) == cudaSuccess;
	cudaSetDevice(cur);
	if (!ok) { mmg_set_error("cannot copy the index tables to the host: %s", cudaGetErrorString(cudaGetLastError())); return MMG_ECUDA; }
	idx->host_tables = true;
	return MMG_OK;
}

/* mm_idx_gen on the device.  Returns MMG_EUNSUP (no error text) when the configuration needs the host builder. */
int mmg_index_build_device(int w, int k, int b, int flag, int n_seq, const char *const *names, const char *const *seqs, const uint32_t *lens,
                           int device, mmg_index **out)
{
	*out = 0;
	if ((flag & MMG_I_HPC) || !(k & 1) || k > 28 || w < 1 || w > 255 || n_seq <= 0) return MMG_EUNSUP;
	int n_dev = 0;
	if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) return MMG_EUNSUP;
	int rc = MMG_OK;
	cudaSetDevice(device);
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, device);
	const int n_sms = prop.multiProcessorCount;
	mmg_index *idx = new mmg_index();
	idx->k = k, idx->w = w, idx->b = b, idx->flag = flag, idx->n_seq = n_seq;
	idx->offs.assign(n_seq + 1, 0);
	for (int i = 0; i < n_seq; ++i) {
		idx->names.push_back(names[i]);
		idx->lens.push_back(lens[i]);
		idx->offs[i + 1] = idx->offs[i] + lens[i];
	}
	const uint64_t sum_len = idx->offs[n_seq];
	const uint32_t margin = (uint32_t)(w + k + 8), seg_cap = IDX_SEG_MAIN + 2 * margin;
	/* segments */
	std::vector<SegMeta> meta;
	std::vector<uint64_t> seg_beg;
	std::vector<uint32_t> seg_len;
	for (int i = 0; i < n_seq; ++i)
		for (uint32_t lo = 0; lo < lens[i]; lo += IDX_SEG_MAIN) {
			const uint32_t hi = lens[i] - lo > IDX_SEG_MAIN ? lo + IDX_SEG_MAIN : lens[i];
			const uint32_t run0 = lo > margin ? lo - margin : 0, run1 = lens[i] - hi > margin ? hi + margin : lens[i];
			SegMeta m = { (uint32_t)i, run0, lo, hi };
			meta.push_back(m), seg_beg.push_back(idx->offs[i] + run0), seg_len.push_back(run1 - run0);
		}
	const uint64_t n_seg = meta.size();
	const uint32_t pass_segs = 2048;
	char *d_seq = 0;
	uint32_t *d_S = 0, *d_mzy = 0, *d_nmz = 0, *d_nkeep = 0, *d_seglen = 0, *d_work = 0, *d_blkh = 0, *d_blkm = 0, *d_big = 0;
	uint64_t *d_mzx = 0, *d_segbeg = 0, *d_keepoff = 0, *d_key = 0, *d_val = 0, *d_key2 = 0, *d_val2 = 0, *d_blkho = 0, *d_blkmo = 0;
	uint64_t *d_headpos = 0, *d_headoff = 0, *d_pos = 0, *d_soff = 0;
	uint32_t *d_slen = 0;
	unsigned long long *d_hist = 0;
	mmg_u128 *d_tab = 0;
	SegMeta *d_meta = 0;
	uint64_t cap = 0, n_all = 0, n_keys = 0, n_multi = 0, n_tiles = 0;
	std::vector<unsigned long long> hist;
	CKI(cudaMalloc((void**)&d_seq, sum_len + 64));
	for (int i = 0; i < n_seq; ++i)
		if (lens[i]) CKI(cudaMemcpy(d_seq + idx->offs[i], seqs[i], lens[i], cudaMemcpyHostToDevice));
	if (!(flag & MMG_I_NO_SEQ)) {
		const uint64_t nw = (sum_len + 7) / 8;
		CKI(cudaMalloc((void**)&d_S, (nw + 1) * 4));
		if (nw) MMG_LAUNCH(pack4_kernel, (unsigned)((nw + 255) / 256), 256, 0, 0, (const char*)d_seq, sum_len, d_S);
		idx->S.resize(nw);
		if (nw) CKI(cudaMemcpy(idx->S.data(), d_S, nw * 4, cudaMemcpyDeviceToHost)); /* host copy: mm_idx_getseq, cs/MD */
	}
	CKI(cudaMalloc((void**)&d_mzx, (uint64_t)pass_segs * seg_cap * 8));
	CKI(cudaMalloc((void**)&d_mzy, (uint64_t)pass_segs * seg_cap * 4));
	CKI(cudaMalloc((void**)&d_nmz, pass_segs * 4));
	CKI(cudaMalloc((void**)&d_nkeep, pass_segs * 4));
	CKI(cudaMalloc((void**)&d_keepoff, (pass_segs + 1) * 8));
	CKI(cudaMalloc((void**)&d_segbeg, pass_segs * 8));
	CKI(cudaMalloc((void**)&d_seglen, pass_segs * 4));
	CKI(cudaMalloc((void**)&d_meta, pass_segs * sizeof(SegMeta)));
	CKI(cudaMalloc((void**)&d_work, 4));
	cap = (uint64_t)((double)sum_len * 2.2 / (w + 1)) + ((uint64_t)1 << 20);
	CKI(cudaMalloc((void**)&d_key, cap * 8));
	CKI(cudaMalloc((void**)&d_val, cap * 8));
	for (uint64_t s0 = 0; s0 < n_seg; s0 += pass_segs) {
		const uint32_t ns = (uint32_t)(n_seg - s0 < pass_segs ? n_seg - s0 : pass_segs);
		CKI(cudaMemcpy(d_segbeg, &seg_beg[s0], (size_t)ns * 8, cudaMemcpyHostToDevice));
		CKI(cudaMemcpy(d_seglen, &seg_len[s0], (size_t)ns * 4, cudaMemcpyHostToDevice));
		CKI(cudaMemcpy(d_meta, &meta[s0], (size_t)ns * sizeof(SegMeta), cudaMemcpyHostToDevice));
		CKI(cudaMemset(d_work, 0, 4));
		ChunkDev c;
		memset(&c, 0, sizeof(c));
		c.n_reads = ns, c.seq = d_seq, c.seg_beg = d_segbeg, c.seg_len = d_seglen, c.seg_cap = seg_cap;
		c.mz_x = d_mzx, c.mz_y = d_mzy, c.n_mz = d_nmz;
		DevIndex di;
		memset(&di, 0, sizeof(di));
		di.k = k, di.w = w;
		launch_sketch(c, di, n_sms, 0, d_work);
		const int fgrid = (int)((ns + 7) / 8);
		MMG_LAUNCH(seg_filter_kernel, fgrid, 256, 0, 0, (const uint64_t*)d_mzx, (const uint32_t*)d_mzy, (const uint32_t*)d_nmz, (const SegMeta*)d_meta, seg_cap, ns,
		           (const uint64_t*)d_keepoff, d_nkeep, d_key, d_val, (uint64_t)0, 0);
		launch_scan_u32(d_nkeep, d_keepoff, ns, 0);
		uint64_t n_pass = 0;
		CKI(cudaMemcpy(&n_pass, d_keepoff + ns, 8, cudaMemcpyDeviceToHost));
		if (n_all + n_pass > cap) { /* grow */
			uint64_t ncap = cap + cap / 2 > n_all + n_pass ? cap + cap / 2 : n_all + n_pass + (cap >> 2);
			uint64_t *nk = 0, *nv = 0;
			CKI(cudaMalloc((void**)&nk, ncap * 8));
			if (cudaMalloc((void**)&nv, ncap * 8) != cudaSuccess) { cudaFree(nk); mmg_set_error("cudaMalloc failed while growing the minimizer arrays"); rc = MMG_ENOMEM; goto fail; }
			cudaMemcpy(nk, d_key, n_all * 8, cudaMemcpyDeviceToDevice), cudaMemcpy(nv, d_val, n_all * 8, cudaMemcpyDeviceToDevice);
			cudaFree(d_key), cudaFree(d_val);
			d_key = nk, d_val = nv, cap = ncap;
		}
		MMG_LAUNCH(seg_filter_kernel, fgrid, 256, 0, 0, (const uint64_t*)d_mzx, (const uint32_t*)d_mzy, (const uint32_t*)d_nmz, (const SegMeta*)d_meta, seg_cap, ns,
		           (const uint64_t*)d_keepoff, d_nkeep, d_key, d_val, n_all, 1);
		n_all += n_pass;
	}
	CKI(cudaDeviceSynchronize());
	cudaFree(d_mzx), cudaFree(d_mzy), cudaFree(d_seq);
	d_mzx = 0, d_mzy = 0, d_seq = 0;
	/* sort */
	CKI(cudaMalloc((void**)&d_key2, (n_all + 1) * 8));
	CKI(cudaMalloc((void**)&d_val2, (n_all + 1) * 8));
	{
		uint32_t max_len = 0;
		for (int i = 0; i < n_seq; ++i) max_len = lens[i] > max_len ? lens[i] : max_len;
		const int val_bits = 32 + bits_of((uint64_t)(n_seq - 1));
		(void)max_len;
		if ((rc = sort_pairs(&d_key, &d_val, &d_key2, &d_val2, n_all, 2 * k, val_bits))) goto fail;
	}
	cudaFree(d_key2), cudaFree(d_val2);
	d_key2 = 0, d_val2 = 0;
	/* heads and ranks */
	n_tiles = (n_all + IDX_TILE - 1) / IDX_TILE;
	CKI(cudaMalloc((void**)&d_blkh, (n_tiles + 1) * 4));
	CKI(cudaMalloc((void**)&d_blkm, (n_tiles + 1) * 4));
	CKI(cudaMalloc((void**)&d_blkho, (n_tiles + 2) * 8));
	CKI(cudaMalloc((void**)&d_blkmo, (n_tiles + 2) * 8));
	if (n_tiles) {
		MMG_LAUNCH(head_count_kernel, (unsigned)n_tiles, 256, 0, 0, (const uint64_t*)d_key, n_all, d_blkh, d_blkm);
		launch_scan_u32(d_blkh, d_blkho, (uint32_t)n_tiles, 0);
		launch_scan_u32(d_blkm, d_blkmo, (uint32_t)n_tiles, 0);
		CKI(cudaMemcpy(&n_keys, d_blkho + n_tiles, 8, cudaMemcpyDeviceToHost));
		CKI(cudaMemcpy(&n_multi, d_blkmo + n_tiles, 8, cudaMemcpyDeviceToHost));
	}
	CKI(cudaMalloc((void**)&d_headpos, (n_keys + 2) * 8));
	CKI(cudaMalloc((void**)&d_headoff, (n_keys + 2) * 8));
	CKI(cudaMalloc((void**)&d_pos, (n_multi + 1) * 8));
	if (n_tiles) {
		MMG_LAUNCH(head_rank_kernel, (unsigned)n_tiles, 256, 0, 0, (const uint64_t*)d_key, (const uint64_t*)d_val, n_all, (const uint64_t*)d_blkho, (const uint64_t*)d_blkmo,
		           d_headpos, d_headoff, d_pos);
		CKI(cudaMemcpy(d_headpos + n_keys, &n_all, 8, cudaMemcpyHostToDevice));
	}
	/* table */
	{
		uint32_t hb = 4;
		const uint64_t inv_load = getenv("MMG_TABLE_INV_LOAD") ? (uint64_t)atoi(getenv("MMG_TABLE_INV_LOAD")) : 4; /* slots per key (tuning experiment hook) */
		while (((uint64_t)1 << hb) < n_keys * (inv_load < 2 ? 2 : inv_load)) ++hb;
		if (n_multi >> 32) { mmg_set_error("index has %llu multi-occurrence positions: more than 2^32 are not supported", (unsigned long long)n_multi); rc = MMG_ECUDA; goto fail; }
		idx->hbits = hb, idx->n_keys = n_keys, idx->n_pos = n_multi;
		const uint64_t nslots = (uint64_t)1 << hb;
		const uint32_t big_cap = 1u << 16;
		CKI(cudaMalloc((void**)&d_tab, nslots * sizeof(mmg_u128)));
		CKI(cudaMalloc((void**)&d_hist, (OCC_HIST + 2) * 8));
		CKI(cudaMalloc((void**)&d_big, big_cap * 4));
		CKI(cudaMemset(d_hist, 0, (OCC_HIST + 2) * 8));
		MMG_LAUNCH(fill_tab_kernel, n_sms * 8, 256, 0, 0, d_tab, nslots);
		if (n_keys)
			MMG_LAUNCH(table_fill_kernel, (unsigned)((n_keys + 255) / 256), 256, 0, 0, (const uint64_t*)d_key, (const uint64_t*)d_val, (const uint64_t*)d_headpos, (const uint64_t*)d_headoff,
			           n_keys, d_tab, hb, d_hist, d_hist + OCC_HIST, d_big, big_cap);
		CKI(cudaDeviceSynchronize());
		hist.resize(OCC_HIST + 1);
		CKI(cudaMemcpy(hist.data(), d_hist, (OCC_HIST + 1) * 8, cudaMemcpyDeviceToHost));
		const uint64_t nb = hist[OCC_HIST];
		if (nb > big_cap) { mmg_set_error("more than %u minimizers occur over %d times", big_cap, OCC_HIST); rc = MMG_EUNSUP; goto fail; }
		idx->occ_hist.assign(hist.begin(), hist.begin() + OCC_HIST);
		idx->occ_big.resize(nb);
		if (nb) CKI(cudaMemcpy(idx->occ_big.data(), d_big, nb * 4, cudaMemcpyDeviceToHost));
		std::sort(idx->occ_big.begin(), idx->occ_big.end());
	}
	CKI(cudaMalloc((void**)&d_soff, (size_t)(n_seq + 1) * 8));
	CKI(cudaMalloc((void**)&d_slen, (size_t)(n_seq + 1) * 4));
	CKI(cudaMemcpy(d_soff, idx->offs.data(), (size_t)(n_seq + 1) * 8, cudaMemcpyHostToDevice));
	CKI(cudaMemcpy(d_slen, idx->lens.data(), (size_t)n_seq * 4, cudaMemcpyHostToDevice));
	idx->dev_device = device, idx->dev_htab = d_tab, idx->dev_pos = d_pos, idx->dev_S = d_S, idx->dev_seq_off = d_soff, idx->dev_seq_len = d_slen;
	idx->host_tables = false;
	d_tab = 0, d_pos = 0, d_S = 0, d_soff = 0, d_slen = 0;
	*out = idx;
	idx = 0;
fail:
	cudaFree(d_seq), cudaFree(d_S), cudaFree(d_mzx), cudaFree(d_mzy), cudaFree(d_nmz), cudaFree(d_nkeep), cudaFree(d_keepoff), cudaFree(d_segbeg);
	cudaFree(d_seglen), cudaFree(d_meta), cudaFree(d_work), cudaFree(d_key), cudaFree(d_val), cudaFree(d_key2), cudaFree(d_val2);
	cudaFree(d_blkh), cudaFree(d_blkm), cudaFree(d_blkho), cudaFree(d_blkmo), cudaFree(d_headpos), cudaFree(d_headoff), cudaFree(d_pos);
	cudaFree(d_tab), cudaFree(d_hist), cudaFree(d_big), cudaFree(d_soff), cudaFree(d_slen);
	if (idx) delete idx;
	return rc;
}

/* ---- INT32 issue peak (denominator of the integer rooflines) ---- */
__global__ void __launch_bounds__(256)
int32_peak_kernel(uint32_t *out, int iters)
{
	uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
	const uint32_t b = blockIdx.x | 1u;
	for (int i = 0; i < iters; ++i) { /* 8 independent chains, 2 integer instructions each (add, xor) */
		a0 = (a0 + b) ^ a1, a1 = (a1 + b) ^ a2, a2 = (a2 + b) ^ a3, a3 = (a3 + b) ^ a4;
		a4 = (a4 + b) ^ a5, a5 = (a5 + b) ^ a6, a6 = (a6 + b) ^ a7, a7 = (a7 + b) ^ a0;
	}
	out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

extern "C" int mmg_debug_int32_peak(int device, double *gops)
{
	*gops = 0;
	int n_dev = 0;
	if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) { mmg_set_error("no CUDA device"); return MMG_ENODEV; }
	cudaSetDevice(device);
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, device);
	const int grid = prop.multiProcessorCount * 8, iters = 1 << 14;
	uint32_t *d = 0;
	if (cudaMalloc((void**)&d, (size_t)grid * 256 * 4) != cudaSuccess) { mmg_set_error("cudaMalloc failed"); return MMG_ENOMEM; }
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0), cudaEventCreate(&e1);
	double best = 0;
	for (int rep = 0; rep < 4; ++rep) {
		cudaEventRecord(e0, 0);
		MMG_LAUNCH(int32_peak_kernel, grid, 256, 0, 0, d, iters);
		cudaEventRecord(e1, 0);
		cudaEventSynchronize(e1);
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		const double g = ms > 0 ? (double)grid * 256 * iters * 16.0 / (ms * 1e-3) / 1e9 : 0;
		if (rep > 0 && g > best) best = g;
	}
	cudaEventDestroy(e0), cudaEventDestroy(e1), cudaFree(d);
	*gops = best;
	return MMG_OK;
}
