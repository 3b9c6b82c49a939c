/* seed.cu -- query-side minimizer filter, index lookup, seed selection and
 * anchor expansion (north-star (c), first half), one warp per read.
 *
 * Replaces, on the mm_map path (/root/reference/src/lib.rs:482,587; minimap2
 * v2.26): seed.c mm_seed_mz_flt, mm_seed_collect_all (-> index.c mm_idx_get),
 * mm_seed_select, mm_collect_matches and the expansion loop of map.c
 * collect_seed_hits.  Anchors leave this file unsorted, in upstream's order
 * (seed order, then the index's position order) because the downstream sort
 * must reproduce radix_sort_128x's treatment of equal keys.
 *
 * Lookup: one 16-byte load per probe into the flat open-addressing table that
 * replaces upstream's 2^b khash buckets; position runs are gathered with one
 * lane per output anchor (load-balanced search over the seeds' prefix sums), so
 * stores are coalesced even for high-occurrence seeds.
 * Bound: HBM random access (32-byte sectors) - see DESIGN.md.
 */
#include "dev_common.cuh"
#include "stages.h"
#ifdef MMG_EMU
#include <stdio.h>
#include <stdlib.h>
#endif

#define SEED_NCNT 1024          /* query-occurrence counters per warp */
#define MAX_MAX_HIGH_OCC 128    /* seed.c */
#define SEED_UNROLL 4

/* index.c mm_idx_get on the flat table: linear probing that starts at an even slot and reads the slot pair of
 * one 32-byte sector per round trip (load factor <= 0.25: a miss - most minimizers of a noisy read - almost
 * always ends at the first pair).  The first pair is loaded by dev_idx_probe so that a warp can have several
 * independent probes in flight before it looks at any of them. */
struct IdxProbe { mmg_u128 e0, e1; uint64_t s; };

__device__ __forceinline__ IdxProbe dev_idx_probe(const DevIndex &di, uint64_t minier)
{
	IdxProbe p;
	p.s = ((minier * 0x9E3779B97F4A7C15ULL) >> (64 - di.hbits)) & ~(uint64_t)1;
	p.e0 = di.htab[p.s], p.e1 = di.htab[p.s + 1];
	return p;
}

__device__ __forceinline__ bool dev_idx_resolve(const DevIndex &di, uint64_t minier, IdxProbe p, uint32_t *n, uint64_t *val)
{
	const uint64_t m = ((uint64_t)1 << di.hbits) - 1;
	for (;;) {
		mmg_u128 e;
		if (p.e0.x == MMG_INF64) return false;
		if (p.e0.x >> 1 == minier) e = p.e0;
		else {
			if (p.e1.x == MMG_INF64) return false;
			if (p.e1.x >> 1 != minier) {
				p.s = (p.s + 2) & m;
				p.e0 = di.htab[p.s], p.e1 = di.htab[p.s + 1];
				continue;
			}
			e = p.e1;
		}
		if (e.x & 1) *n = 1, *val = e.y;          /* the value is the position word itself */
		else *n = (uint32_t)e.y, *val = e.y >> 32; /* offset into pos[] */
		return true;
	}
}

__device__ __forceinline__ void heap_down(uint64_t *l, int i, int n)
{ /* ksort.h ks_heapdown: max-heap */
	int k = i;
	uint64_t tmp = l[i];
	while ((k = (k << 1) + 1) < n) {
		if (k != n - 1 && l[k] < l[k + 1]) ++k;
		if (l[k] < tmp) break;
		l[i] = l[k]; i = k;
	}
	l[i] = tmp;
}

/* seed.c: mm_seed_select.  A streak is a maximal run [st, en) of seeds that occur more than max_occ times; upstream
 * keeps the (pe - ps) / dist least frequent of each streak (a heap) and drops the rest.  The streaks are found by the
 * warp (ballots over 32 seeds at a time); the heap of a streak runs on lane 0 - streaks are short and rare. */
__device__ void dev_seed_streak(int st, int en, int n, const uint32_t *sn, const uint32_t *sq, uint32_t *meta, int len, int max_max_occ, int dist, uint64_t *b)
{
	int ps = st == 0 ? 0 : (int)(sq[st - 1] >> 1);
	int pe = en == n ? len : (int)(sq[en] >> 1);
	int j, k;
	int max_high_occ = (int)((double)(pe - ps) / dist + .499);
	if (max_high_occ > 0) {
		if (max_high_occ > MAX_MAX_HIGH_OCC) max_high_occ = MAX_MAX_HIGH_OCC;
		for (j = st, k = 0; j < en && k < max_high_occ; ++j, ++k)
			b[k] = (uint64_t)sn[j] << 32 | (uint32_t)j;
		for (int h = (k >> 1) - 1; h >= 0; --h) heap_down(b, h, k);
		for (; j < en; ++j) {
			if (sn[j] < (uint32_t)(b[0] >> 32)) {
				b[0] = (uint64_t)sn[j] << 32 | (uint32_t)j;
				heap_down(b, 0, k);
			}
		}
		for (j = 0; j < k; ++j) meta[(uint32_t)b[j]] |= 2u;
	}
	for (j = st; j < en; ++j) meta[j] ^= 2u;
	for (j = st; j < en; ++j)
		if ((int)sn[j] > max_max_occ) meta[j] |= 2u;
}

__device__ void dev_seed_select(int n, const uint32_t *sn, const uint32_t *sq, uint32_t *meta, int len, int max_occ, int max_max_occ, int dist, uint64_t *b)
{ /* all lanes */
	if (n == 0 || n == 1) return;
	const int lane = mmg_lane();
	int run_st = -1;
	for (int i0 = 0; i0 < n; i0 += 32) {
		const int i = i0 + lane;
		const uint32_t hm = __ballot_sync(MMG_FULL, i < n && (int)sn[i] > max_occ);
		int p = 0; /* bits below p are done */
		while (p < 32) {
			const uint32_t up = ~0u << p;
			if (run_st < 0) {
				const uint32_t mm = hm & up;
				if (!mm) break;
				const int bpos = __ffs((int)mm) - 1;
				run_st = i0 + bpos, p = bpos + 1;
			} else {
				const uint32_t zz = ~hm & up;
				if (!zz) break; /* the streak runs into the next 32 seeds */
				const int bpos = __ffs((int)zz) - 1;
				if (lane == 0) dev_seed_streak(run_st, i0 + bpos, n, sn, sq, meta, len, max_max_occ, dist, b);
				__syncwarp();
				run_st = -1, p = bpos + 1;
			}
		}
	}
	if (run_st >= 0) { /* only when n is a multiple of 32 and the last seed is in a streak */
		if (lane == 0) dev_seed_streak(run_st, n, n, sn, sq, meta, len, max_max_occ, dist, b);
		__syncwarp();
	}
}

__global__ void __launch_bounds__(SEED_WARPS * 32, 3)
seed_kernel(ChunkDev c, DevIndex di, DevOpt o, uint32_t *work)
{
	__shared__ uint32_t s_cnt[SEED_WARPS][SEED_NCNT];
	__shared__ uint64_t s_heap[SEED_WARPS][MAX_MAX_HIGH_OCC];
	const int lane = mmg_lane(), wib = threadIdx.x >> 5;
	const uint32_t lt = mmg_lanemask_lt();
	unsigned long long tot_seed = 0, tot_hit = 0;

	for (;;) {
		uint32_t r = mmg_next_item(work);
		if (r >= c.n_reads) break;
		r = mmg_read_of(c, r);
		const uint64_t base = c.off[r] - c.off0;
		const int qlen = (int)(c.off[r + 1] - c.off[r]);
		int n = (int)c.n_mz[r];
		uint64_t *mx = c.mz_x + base;
		uint32_t *my = c.mz_y + base;

		/* ---- mm_seed_mz_flt: drop minimizers that are too frequent IN THE QUERY ---- */
		if (n > o.mid_occ && o.q_occ_frac > 0.0f && o.mid_occ > 0) {
			uint32_t *cnt = s_cnt[wib];
			for (int j = lane; j < SEED_NCNT; j += 32) cnt[j] = 0;
			__syncwarp();
			for (int i = lane; i < n; i += 32)
				atomicAdd(&cnt[(uint32_t)((mx[i] * 0x9E3779B97F4A7C15ULL) >> 54)], 1u);
			__syncwarp();
			uint32_t mxc = 0;
			for (int j = lane; j < SEED_NCNT; j += 32) mxc = max(mxc, cnt[j]);
			mxc = __reduce_max_sync(MMG_FULL, mxc);
			if ((int)mxc > o.mid_occ) { /* some key MAY exceed the threshold: count those exactly */
				const float thr = (float)n * o.q_occ_frac;
				int n_new = 0;
				/* mark in sd_meta (free scratch at this point), then compact in order */
				uint32_t *mark = c.sd_meta + base;
				for (int i = lane; i < n; i += 32) {
					uint64_t x = mx[i];
					uint32_t d = 0;
					if ((int)cnt[(uint32_t)((x * 0x9E3779B97F4A7C15ULL) >> 54)] > o.mid_occ) {
						int cn = 0;
						for (int j = 0; j < n; ++j) cn += (mx[j] == x);
						d = (cn > o.mid_occ && (float)cn > thr) ? 1u : 0u;
					}
					mark[i] = d;
				}
				__syncwarp();
				for (int i0 = 0; i0 < n; i0 += 32) {
					int i = i0 + lane;
					bool keep = i < n && mark[i] == 0;
					uint64_t x = i < n ? mx[i] : 0;
					uint32_t y = i < n ? my[i] : 0;
					uint32_t km = __ballot_sync(MMG_FULL, keep);
					__syncwarp();
					if (keep) { int d = n_new + __popc(km & lt); mx[d] = x, my[d] = y; }
					n_new += __popc(km);
					__syncwarp();
				}
				n = n_new;
			}
			__syncwarp();
		}

		/* ---- mm_seed_collect_all: index lookup, tandem flag, ordered compaction ---- */
		uint64_t *sv = c.sd_val + base;
		uint32_t *sn = c.sd_n + base, *sq = c.sd_qpos + base, *sm = c.sd_meta + base;
		int n_m0 = 0, n_high = 0;
		for (int i0 = 0; i0 < n; i0 += 32 * SEED_UNROLL) { /* SEED_UNROLL x 32 table probes in flight per warp */
			uint64_t xs[SEED_UNROLL];
			IdxProbe pr[SEED_UNROLL];
#pragma unroll
			for (int u = 0; u < SEED_UNROLL; ++u) {
				const int i = i0 + 32 * u + lane;
				xs[u] = i < n ? mx[i] : 0;
			}
#pragma unroll
			for (int u = 0; u < SEED_UNROLL; ++u) {
				const int i = i0 + 32 * u + lane;
				if (i < n) pr[u] = dev_idx_probe(di, xs[u] >> 8);
			}
#pragma unroll
			for (int u = 0; u < SEED_UNROLL; ++u) {
				const int i = i0 + 32 * u + lane;
				if (i0 + 32 * u >= n) break;
				bool hit = false;
				uint32_t hn = 0, meta = 0;
				uint64_t hv = 0;
				const uint64_t x = xs[u];
				/* neighbours for the tandem flag: from the adjacent lanes, the warp's edges from memory */
				uint64_t xp = __shfl_up_sync(MMG_FULL, x, 1), xn = __shfl_down_sync(MMG_FULL, x, 1);
				if (i < n) {
					if (lane == 0 && i > 0) xp = mx[i - 1];
					if (lane == 31 && i < n - 1) xn = mx[i + 1];
					hit = dev_idx_resolve(di, x >> 8, pr[u], &hn, &hv);
					if (hit) {
						bool tandem = (i > 0 && xp >> 8 == x >> 8) || (i < n - 1 && xn >> 8 == x >> 8);
						meta = (uint32_t)(x & 0xff) << 8 | (tandem ? 1u : 0u);
					}
				}
				uint32_t hm = __ballot_sync(MMG_FULL, hit);
				if (hit) {
					int d = n_m0 + __popc(hm & lt);
					sv[d] = hv, sn[d] = hn, sq[d] = my[i], sm[d] = meta;
				}
				n_m0 += __popc(hm);
				n_high += __popc(__ballot_sync(MMG_FULL, hit && (int)hn > o.mid_occ));
			}
		}
		__syncwarp();

		/* ---- mm_seed_select / plain occurrence cut ---- */
		if (n_high > 0) {
			if (o.occ_dist > 0 && o.max_max_occ > o.mid_occ) {
				dev_seed_select(n_m0, sn, sq, sm, qlen, o.mid_occ, o.max_max_occ, o.occ_dist, s_heap[wib]);
			} else {
				for (int i = lane; i < n_m0; i += 32) if ((int)sn[i] > o.mid_occ) sm[i] |= 2u;
			}
			__syncwarp();
		}

		/* ---- mm_collect_matches tail: rep_len, n_a, kept seeds ---- */
		int n_m = n_m0, rep_len = 0;
		unsigned long long n_a = 0;
		if (n_high > 0) {
			/* 32 seeds per step.  Kept seeds are compacted in place (a seed moves to an index <= its own, and every lane
			 * has read its seed before any lane writes).  rep_len is the length of the union of the dropped seeds'
			 * query intervals [en - span, en); en ascends, so upstream's running (rep_st, rep_en) closes a segment
			 * exactly when a dropped seed starts after the previous dropped seed's end:
			 * rep_len = en(last dropped) + sum over those seeds of (previous en - st). */
			int k = 0, carry_en = 0, acc = 0;
			unsigned na = 0;
			for (int i0 = 0; i0 < n_m0; i0 += 32) {
				const int i = i0 + lane;
				const bool in = i < n_m0;
				uint64_t v = 0;
				uint32_t nn = 0, qq = 0, mm = 0;
				if (in) v = sv[i], nn = sn[i], qq = sq[i], mm = sm[i];
				const bool flg = in && (mm & 2u), kp = in && !(mm & 2u);
				const uint32_t fm = __ballot_sync(MMG_FULL, flg), km = __ballot_sync(MMG_FULL, kp);
				const int en = (int)(qq >> 1) + 1, st = en - (int)(mm >> 8);
				const uint32_t below = fm & lt;
				int pen = __shfl_sync(MMG_FULL, en, below ? 31 - __clz((int)below) : 0);
				if (!below) pen = carry_en;
				if (flg && st > pen) acc += pen - st;
				if (fm) carry_en = __shfl_sync(MMG_FULL, en, 31 - __clz((int)fm));
				__syncwarp();
				if (kp) { const int d = k + __popc(km & lt); sv[d] = v, sn[d] = nn, sq[d] = qq, sm[d] = mm; na += nn; }
				k += __popc(km);
				__syncwarp();
			}
			rep_len = __reduce_add_sync(MMG_FULL, acc) + carry_en;
			n_m = k;
			n_a = __reduce_add_sync(MMG_FULL, na);
		} else {
			unsigned s = 0;
			for (int i = lane; i < n_m0; i += 32) s += sn[i];
			n_a = __reduce_add_sync(MMG_FULL, s);
		}
		if (lane == 0) {
			c.n_mz[r] = (uint32_t)n;
			c.n_seed[r] = (uint32_t)n_m;
			c.n_a[r] = (uint32_t)n_a;
			c.rep_len[r] = rep_len;
		}
		tot_seed += n_m, tot_hit += n_a;
		__syncwarp();
	}
	if (lane == 0 && (tot_seed | tot_hit)) {
		atomicAdd(&c.stats[2], tot_seed);
		atomicAdd(&c.stats[3], tot_hit);
	}
}

/* ---- anchor expansion: map.c collect_seed_hits inner loops ---------------- */
__global__ void __launch_bounds__(SEED_WARPS * 32)
expand_kernel(ChunkDev c, DevIndex di, DevOpt o, uint32_t r0, uint32_t r1, uint32_t *work)
{
	const int lane = mmg_lane();
	for (;;) {
		uint32_t r = r0 + mmg_next_item(work);
		if (r >= r1) break;
		r = mmg_read_of(c, r);
		const uint64_t base = c.off[r] - c.off0;
		const int qlen = (int)(c.off[r + 1] - c.off[r]);
		const int n_m = (int)c.n_seed[r];
		const uint64_t *sv = c.sd_val + base;
		const uint32_t *sn = c.sd_n + base, *sq = c.sd_qpos + base, *sm = c.sd_meta + base;
		uint64_t *ax = c.ax + (c.a_off[r] - c.a_off0), *ay = c.ay + (c.a_off[r] - c.a_off0);
		const bool filt = (c.flags[r] & 4u) != 0; /* anchor_filter_kernel dropped the isolated anchors of this read */
		const uint32_t *bits = filt ? c.keep_bits + (c.af_off[r] >> 5) + r : 0;
		const uint64_t *hits = filt ? c.hit_scratch + c.af_off[r] : 0; /* the index positions gathered by the filter */
		const uint32_t lt = mmg_lanemask_lt();
		uint32_t out = 0, full = 0; /* anchors written / anchors enumerated so far */
		for (int i0 = 0; i0 < n_m; i0 += 32) {
			int i = i0 + lane;
			int cnt = i < n_m ? (int)sn[i] : 0, tot;
			uint64_t val = i < n_m ? sv[i] : 0;
			uint32_t qp = i < n_m ? sq[i] : 0, meta = i < n_m ? sm[i] : 0;
			int ex = mmg_warp_excl_scan(cnt, &tot);
			for (int t0 = 0; t0 < tot; t0 += 32) {
				int t = t0 + lane;
				/* find the seed that owns output t: largest lane s with ex[s] <= t (ex is non-decreasing) */
				int s = 0;
#pragma unroll
				for (int d = 16; d; d >>= 1) {
					int cand = s + d;
					int exc = __shfl_sync(MMG_FULL, ex, cand & 31);
					if (cand < 32 && exc <= t) s = cand;
				}
				/* skip empty seeds that share the same prefix: the owner is the last lane with ex<=t and cnt>0;
				 * the binary search lands on the last lane with ex <= t, which has cnt > 0 whenever t < tot */
				int s_ex = __shfl_sync(MMG_FULL, ex, s);
				int s_cnt = __shfl_sync(MMG_FULL, cnt, s);
				uint64_t s_val = __shfl_sync(MMG_FULL, val, s);
				uint32_t s_qp = __shfl_sync(MMG_FULL, qp, s), s_meta = __shfl_sync(MMG_FULL, meta, s);
				bool keep = t < tot;
				if (keep && filt) { const uint32_t g = full + (uint32_t)t; keep = (bits[g >> 5] >> (g & 31)) & 1u; }
				const uint32_t km = __ballot_sync(MMG_FULL, keep);
				if (keep) {
					int kk = t - s_ex;
					uint64_t rr = filt ? hits[full + (uint32_t)t] & 0x7fffffffffffffffULL : s_cnt == 1 ? s_val : di.pos[s_val + kk];
					uint32_t rpos = (uint32_t)rr >> 1, span = s_meta >> 8;
					uint64_t x, y;
					if ((rr & 1) == (s_qp & 1)) { /* forward strand */
						x = (rr & 0xffffffff00000000ULL) | rpos;
						y = (uint64_t)span << 32 | (s_qp >> 1);
					} else {
						x = 1ULL << 63 | (rr & 0xffffffff00000000ULL) | rpos;
						y = (uint64_t)span << 32 | (uint32_t)(qlen - (int)((s_qp >> 1) + 1 - span) - 1);
					}
					if (s_meta & 1u) y |= 1ULL << 42; /* MM_SEED_TANDEM */
					const uint32_t d = out + (uint32_t)__popc(km & lt);
					ax[d] = x, ay[d] = y;
				}
				out += (uint32_t)__popc(km);
			}
			full += (uint32_t)tot;
		}
		__syncwarp();
	}
}

/* ---- isolated-anchor filter --------------------------------------------------------------------------------
 * On a large reference most index hits of a read are random: anchors with no other anchor of the same strand and
 * contig within max_dist_x on either side.  Such an anchor has an empty predecessor window in mm_lchain_dp, is in
 * no other anchor's window, ends with f = span < min_chain_score and cnt = 1 < min_cnt, so it is in no chain, is
 * no chain end candidate, and leaves every heuristic state of the DP (st, max_ii, n_skip, t[]) as it found it:
 * dropping it before the sort changes nothing downstream (the re-chaining and alignment stages only see chained
 * anchors).  One thing could notice: upstream's unstable radix sort orders EQUAL keys by the whole input
 * permutation.  Equal keys need two minimizers of the read with the same hash, so reads with a repeated
 * minimizer hash keep all their anchors.  The host enables the filter only if min_cnt >= 2 and
 * min_chain_score > k.
 * One CTA per read: (0) repeated-hash test over the read's minimizers (fingerprint set), (1) anchors are counted
 * marked in position bins of width >= max_dist_x (two shared-memory bitmaps over 2^17 hashed bins: "occupied" and
 * "occupied twice"; a collision only keeps more), (2) an anchor is kept if its bin is occupied twice or an
 * adjacent bin is occupied; the keep bits go to c.keep_bits, the kept count
 * replaces c.n_a[r]. */
#define AF_THREADS 256
#define AF_TAB 4096            /* words: fingerprint set of phase 0, then the two bin bitmaps (AF_TAB * 32 bins each) */
#define AF_BIN_BITS 17
#define AF_MAX_SEEDS 3072
#define AF_MAX_ANCHORS 65536
#define AF_MIN_ANCHORS 64
#define AF_ROUNDS 2             /* hash rounds: each re-tests the survivors of the previous one with another hash */
#define AF_GATHER 8             /* independent pos[] reads in flight per thread */

/* hashed slot of the position bin (strand, contig, (pos >> shift) + delta) in a table of 2^bits bins */
__device__ __forceinline__ uint32_t af_bin_slot(uint64_t rr, bool rev, int shift, int delta, uint32_t seed, int bits)
{
	const uint32_t a = ((uint32_t)(rr >> 32) << 1 | (uint32_t)rev) * 0x9E3779B1u + seed;
	const uint32_t b = (((uint32_t)rr >> 1 >> shift) + (uint32_t)delta) * 0x85EBCA77u;
	uint32_t h = (a ^ b) * 0xC2B2AE3Du;
	h ^= h >> 15;
	return (h * 0x27D4EB2Fu) >> (32 - bits);
}

__device__ __forceinline__ bool af_keep(const uint32_t *occ, const uint32_t *two, uint32_t b0, uint32_t bm, uint32_t bp)
{
	return ((two[b0 >> 5] >> (b0 & 31)) | (occ[bm >> 5] >> (bm & 31)) | (occ[bp >> 5] >> (bp & 31))) & 1u;
}

__global__ void __launch_bounds__(AF_THREADS)
anchor_filter_kernel(ChunkDev c, DevIndex di, DevOpt o, uint32_t *work)
{
	MMG_DYN_SMEM(smem_raw);
	uint32_t *s_tab = (uint32_t*)smem_raw, *s_two = s_tab + AF_TAB, *s_pre = s_two + AF_TAB, *s_bits = s_pre + AF_MAX_SEEDS + 1;
	__shared__ uint32_t s_item, s_dup, s_keep, s_warp[AF_THREADS / 32];
	const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
	unsigned long long dropped = 0;
	for (;;) {
		if (tid == 0) s_item = atomicAdd(work, 1u), s_dup = 0, s_keep = 0;
		__syncthreads();
		if (s_item >= c.n_reads) break;
		const uint32_t r = mmg_read_of(c, s_item);
		const uint32_t n_full = c.n_a[r];
		const int n_m = (int)c.n_seed[r], n_mz = (int)c.n_mz[r];
		/* worth it only where the index returns several hits per seed (large references): most of them are random */
		if (n_full < AF_MIN_ANCHORS || n_full < 2u * (uint32_t)n_m || n_full > AF_MAX_ANCHORS || n_m > AF_MAX_SEEDS || n_mz > AF_MAX_SEEDS) { __syncthreads(); continue; }
		const uint64_t base = c.off[r] - c.off0;
		const int qlen = (int)(c.off[r + 1] - c.off[r]);
		/* (0) a minimizer hash that occurs twice in the read?  open-addressing set of 32-bit fingerprints */
		int fbits = 9;
		while ((1 << fbits) < 2 * n_mz) ++fbits;
		const uint32_t fmask = (1u << fbits) - 1;
		for (int j = tid; j <= (int)fmask; j += AF_THREADS) s_tab[j] = 0;
		__syncthreads();
		for (int i = tid; i < n_mz; i += AF_THREADS) {
			const uint64_t h = (c.mz_x[base + i] >> 8) * 0x9E3779B97F4A7C15ULL;
			const uint32_t fp = (uint32_t)(h >> 32) | 1u;
			for (uint32_t sl = (uint32_t)(h >> 20) & fmask;; sl = (sl + 1) & fmask) {
				const uint32_t old = atomicCAS(&s_tab[sl], 0u, fp);
				if (old == 0) break;
				if (old == fp) { s_dup = 1; break; }
			}
		}
		__syncthreads();
		if (s_dup) { __syncthreads(); continue; }
		/* chaining distance of this read (map.c mm_map_frag, as chain.cu) -> bin width 2^shift >= max_dist_x */
		int32_t max_dist_x;
		if (o.max_gap_ref > 0) max_dist_x = o.max_gap_ref;
		else if (o.max_frag_len > 0) { max_dist_x = o.max_frag_len - qlen; if (max_dist_x < o.max_gap) max_dist_x = o.max_gap; }
		else max_dist_x = o.max_gap;
		if (max_dist_x < o.bw) max_dist_x = o.bw;
		int shift = 0;
		while (shift < 30 && (1 << shift) <= max_dist_x) ++shift;
		const uint64_t *sv = c.sd_val + base;
		const uint32_t *sn = c.sd_n + base, *sq = c.sd_qpos + base;
		uint64_t *hits = c.hit_scratch + c.af_off[r];
		const int n_words = (int)((n_full + 31) >> 5);
		{ /* exclusive prefix of the seed occurrence counts: anchor index of a seed's first hit */
			uint32_t carry = 0;
			for (int i0 = 0; i0 < n_m; i0 += AF_THREADS) {
				const int i = i0 + tid;
				const uint32_t v = i < n_m ? sn[i] : 0;
				uint32_t x = v;
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
					uint32_t y = __shfl_up_sync(MMG_FULL, x, d);
					if (lane >= d) x += y;
				}
				if (lane == 31) s_warp[wib] = x;
				__syncthreads();
				uint32_t pre = carry, tot = 0;
				for (int q = 0; q < AF_THREADS / 32; ++q) { if (q < wib) pre += s_warp[q]; tot += s_warp[q]; }
				if (i < n_m) s_pre[i] = pre + x - v;
				carry += tot;
				__syncthreads();
			}
		}
		/* Two rounds over the read's hits; the second re-tests the survivors of the first with another hash, so
		 * that what a collision kept by accident is (almost always) dropped after all.  The index positions are
		 * gathered ONCE (random 8-byte reads of di.pos are the expensive part) into hit_scratch, in anchor order,
		 * with the seed's strand in bit 63; everything after that streams through it. */
		uint32_t n_in = n_full;
		for (int round = 0; round < AF_ROUNDS; ++round) {
			int bits = 12; /* 32 bins per candidate: ~9 % of the isolated anchors survive a round by collision */
			while (bits < AF_BIN_BITS && (1u << bits) < 32u * n_in) ++bits;
			const uint32_t seed = (uint32_t)round * 0x68E31DA4u;
			for (int j = tid; j < (1 << (bits - 5)); j += AF_THREADS) s_tab[j] = 0, s_two[j] = 0;
			__syncthreads();
			if (round == 0) {
				/* one thread per HIT (the seed that owns hit g is found by binary search in the prefix sums), AF_GATHER
				 * independent random reads of pos[] in flight per thread before any of them is used */
				for (uint32_t gb = tid; gb < n_full; gb += AF_THREADS * AF_GATHER) {
					uint64_t rr[AF_GATHER], qb[AF_GATHER];
#pragma unroll
					for (int u = 0; u < AF_GATHER; ++u) {
						const uint32_t g = gb + (uint32_t)u * AF_THREADS;
						rr[u] = 0, qb[u] = 0;
						if (g < n_full) {
							int lo = 0, hi = n_m - 1; /* largest i with s_pre[i] <= g */
							while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_pre[mid] <= g) lo = mid; else hi = mid - 1; }
							const uint32_t cnt = sn[lo];
							const uint64_t val = sv[lo];
							qb[u] = (uint64_t)(sq[lo] & 1u) << 63;
							rr[u] = cnt == 1 ? val : di.pos[val + (g - s_pre[lo])];
						}
					}
#pragma unroll
					for (int u = 0; u < AF_GATHER; ++u) {
						const uint32_t g = gb + (uint32_t)u * AF_THREADS;
						if (g < n_full) {
							hits[g] = rr[u] | qb[u];
							const uint32_t b0 = af_bin_slot(rr[u], (rr[u] & 1) != (qb[u] >> 63), shift, 0, seed, bits);
							if (atomicOr(&s_tab[b0 >> 5], 1u << (b0 & 31)) >> (b0 & 31) & 1u) atomicOr(&s_two[b0 >> 5], 1u << (b0 & 31));
						}
					}
				}
			} else {
				for (uint32_t g = tid; g < n_full; g += AF_THREADS) {
					if (!((s_bits[g >> 5] >> (g & 31)) & 1u)) continue;
					const uint64_t w = hits[g], rr = w & 0x7fffffffffffffffULL;
					const uint32_t b0 = af_bin_slot(rr, (rr & 1) != (w >> 63), shift, 0, seed, bits);
					if (atomicOr(&s_tab[b0 >> 5], 1u << (b0 & 31)) >> (b0 & 31) & 1u) atomicOr(&s_two[b0 >> 5], 1u << (b0 & 31));
				}
			}
			__syncthreads();
			uint32_t kept = 0;
			for (uint32_t g0 = (uint32_t)wib * 32; g0 < n_full; g0 += AF_THREADS) { /* a warp decides 32 consecutive anchors: one word */
				const uint32_t g = g0 + lane;
				bool keep = g < n_full && (round == 0 || ((s_bits[g >> 5] >> (g & 31)) & 1u));
				if (keep) {
					const uint64_t w = hits[g], rr = w & 0x7fffffffffffffffULL;
					const bool rev = (rr & 1) != (w >> 63);
					keep = af_keep(s_tab, s_two, af_bin_slot(rr, rev, shift, 0, seed, bits), af_bin_slot(rr, rev, shift, -1, seed, bits),
					               af_bin_slot(rr, rev, shift, 1, seed, bits));
				}
				const uint32_t km = __ballot_sync(MMG_FULL, keep);
				__syncwarp();
				if (lane == 0) s_bits[g0 >> 5] = km;
				kept += (uint32_t)__popc(km);
			}
			if (lane == 0) s_warp[wib] = kept;
			__syncthreads();
			n_in = 0;
			for (int q = 0; q < AF_THREADS / 32; ++q) n_in += s_warp[q];
			__syncthreads();
		}
		uint32_t *bits_out = c.keep_bits + (c.af_off[r] >> 5) + r;
		for (int j = tid; j < n_words; j += AF_THREADS) bits_out[j] = s_bits[j];
		if (tid == 0) {
			c.n_a[r] = n_in;
			c.flags[r] |= 4u;
			dropped += n_full - n_in;
		}
		__syncthreads();
	}
	if (tid == 0 && dropped) atomicAdd(&c.stats[4], dropped), atomicAdd(&c.stats[10], dropped); /* n_anchor counts what upstream would have sorted */
}

/* ---- exclusive scan of per-read counts (single block; n is a few 10^5) ---- */
__global__ void __launch_bounds__(1024)
scan_u32_kernel(const uint32_t *in, uint64_t *out, uint32_t n)
{
	__shared__ unsigned long long s_warp[32];
	__shared__ unsigned long long s_carry;
	const int lane = mmg_lane(), wib = threadIdx.x >> 5;
	if (threadIdx.x == 0) s_carry = 0;
	__syncthreads();
	for (uint32_t i0 = 0; i0 < n; i0 += 1024) {
		uint32_t i = i0 + threadIdx.x;
		unsigned long long v = i < n ? in[i] : 0, x = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			unsigned long long y = __shfl_up_sync(MMG_FULL, x, o);
			if (lane >= o) x += y;
		}
		if (lane == 31) s_warp[wib] = x;
		__syncthreads();
		if (wib == 0) {
			unsigned long long w = s_warp[lane], ws = w;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				unsigned long long y = __shfl_up_sync(MMG_FULL, ws, o);
				if (lane >= o) ws += y;
			}
			s_warp[lane] = ws - w;
		}
		__syncthreads();
		unsigned long long carry = s_carry;
		if (i < n) out[i] = carry + s_warp[wib] + x - v;
		__syncthreads();
		if (threadIdx.x == 1023) s_carry = carry + s_warp[wib] + x;
		__syncthreads();
	}
	if (threadIdx.x == 0) out[n] = s_carry;
}

int launch_seed(const ChunkDev &c, const DevIndex &di, const DevOpt &o, int n_sms, cudaStream_t st, uint32_t *work)
{
	int grid = n_sms * 5, need = ((int)c.n_reads + SEED_WARPS - 1) / SEED_WARPS;
	if (grid > need) grid = need;
	if (grid < 1) grid = 1;
	MMG_LAUNCH(seed_kernel, grid, SEED_WARPS * 32, 0, st, c, di, o, work);
	return 0;
}

int launch_expand(const ChunkDev &c, const DevIndex &di, const DevOpt &o, uint32_t r0, uint32_t r1, int n_sms, cudaStream_t st, uint32_t *work)
{
	int grid = n_sms * 8, need = ((int)(r1 - r0) + SEED_WARPS - 1) / SEED_WARPS;
	if (grid > need) grid = need;
	if (grid < 1) grid = 1;
	MMG_LAUNCH(expand_kernel, grid, SEED_WARPS * 32, 0, st, c, di, o, r0, r1, work);
	return 0;
}

int launch_anchor_filter(const ChunkDev &c, const DevIndex &di, const DevOpt &o, int n_sms, cudaStream_t st, uint32_t *work)
{
	int grid = n_sms * 4, need = (int)c.n_reads;
	if (grid > need) grid = need;
	if (grid < 1) grid = 1;
	const size_t smem = (size_t)(2 * AF_TAB + AF_MAX_SEEDS + 1 + AF_MAX_ANCHORS / 32) * 4;
	static unsigned char attr_done[64];
	if (mmg_once_per_device(attr_done)) { cudaFuncSetAttribute(anchor_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); }
	MMG_LAUNCH(anchor_filter_kernel, grid, AF_THREADS, smem, st, c, di, o, work);
	return 0;
}

int launch_scan_u32(const uint32_t *in, uint64_t *out, uint32_t n, cudaStream_t st)
{
	MMG_LAUNCH(scan_u32_kernel, 1, 1024, 0, st, in, out, n);
	return 0;
}
