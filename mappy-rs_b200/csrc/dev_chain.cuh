/* dev_chain.cuh -- backtrack + chain compaction, shared by the DP path (chain.cu)
 * and the RMQ re-chain (rmq.cu).  Replaces lchain.c mg_chain_bk_end,
 * mg_chain_backtrack and compact_a (minimap2 v2.26; mm_map path,
 * /root/reference/src/lib.rs:482,587). */
#ifndef MMG_DEV_CHAIN_CUH
#define MMG_DEV_CHAIN_CUH
#include "dev_common.cuh"
#include "dev_sort.cuh"

#ifndef INT32_MIN_
#define INT32_MIN_ (-2147483647 - 1)
#endif

/* lchain.c: mg_chain_bk_end */
__device__ __forceinline__ int dev_chain_bk_end(int32_t max_drop, int32_t zkx, int zky, const int32_t *f, const int32_t *p, int32_t *t)
{
	int i = zky, end_i = -1, max_i = i;
	int32_t max_s = 0;
	if (i < 0 || t[i] != 0) return i;
	do {
		int32_t s;
		t[i] = 2;
		end_i = i = p[i];
		s = i < 0 ? zkx : zkx - f[i];
		if (s > max_s) max_s = s, max_i = i;
		else if (max_s - s > max_drop) break;
	} while (i >= 0 && t[i] == 0);
	for (i = zky; i >= 0 && i != end_i; i = p[i]) t[i] = 0;
	return max_i;
}

/* mg_chain_backtrack + compact_a for one read, called by the whole warp.
 * In: sorted anchors ax/ay[n], f/p[n]; t/v[n] scratch (t is cleared here), zx/zy[2n], cx/cy[n].
 * Out: chained anchors back in ax/ay[0..n_v), chains u[0..n_u) ordered by target position. */
static __device__ void dev_backtrack_compact(int n, uint64_t *ax, uint64_t *ay, const int32_t *f, const int32_t *p, int32_t *t, int32_t *v,
                                             uint64_t *zx, uint64_t *zy, uint64_t *cx, uint64_t *cy, uint64_t *u,
                                             int32_t min_cnt, int32_t min_sc, int32_t max_drop, int *bkt, int *n_u_, int *n_v_)
{
	const int lane = mmg_lane();
	const uint32_t lt = mmg_lanemask_lt();
	int n_z = 0, n_u = 0, n_v = 0;
	for (int i0 = 0; i0 < n; i0 += 32) { /* z[] = (f, index) of every possible chain end, in index order */
		int i = i0 + lane;
		int32_t fv = i < n ? f[i] : INT32_MIN_;
		bool keep = i < n && fv >= min_sc;
		uint32_t km = __ballot_sync(MMG_FULL, keep);
		if (keep) { int d = n_z + __popc(km & lt); zx[d] = (uint64_t)(int64_t)fv, zy[d] = (uint64_t)i; }
		n_z += __popc(km);
		if (i < n) t[i] = 0;
	}
	__syncwarp();
	if (n_z > 0) {
		dev_radix_sort_warp(zx, zy, n_z, bkt, (int*)v); /* v[] is the sort's range stack; it is free again below */
		/* mg_chain_backtrack: chain ends in descending score order.  32 candidates are tested per step; an
		 * end whose anchor is still unused starts a walk.  mg_chain_bk_end's three passes over the chain
		 * (mark, unmark, collect) are one walk here: the path goes to v[] as it is followed, the position of
		 * the best prefix is tracked, and only that prefix is then marked used (by all lanes). */
		for (int kb = n_z - 1; kb >= 0; kb -= 32) {
			const int k = kb - lane;
			const int zi = k >= 0 ? (int)zy[k] : -1;
			const int32_t zkx = k >= 0 ? (int32_t)zx[k] : 0;
			int first = 0; /* lanes below `first` are done */
			for (;;) {
				const bool cand = k >= 0 && lane >= first && t[zi] == 0;
				const uint32_t cm = __ballot_sync(MMG_FULL, cand);
				if (!cm) break;
				const int src = __ffs((int)cm) - 1;
				first = src + 1;
				const int czi = __shfl_sync(MMG_FULL, zi, src);
				const int32_t czkx = __shfl_sync(MMG_FULL, zkx, src);
				int best_len = 0;
				int32_t max_s = 0;
				if (lane == 0) {
					/* the walk is a chain of dependent loads: f, t and p of the next anchor are fetched together,
					 * one round trip per step */
					int m = 0, cur = czi, nxt = p[cur];
					for (;;) {
						v[n_v + m] = cur, ++m;
						int32_t fn = 0, tn = 0, pn = -1;
						if (nxt >= 0) fn = f[nxt], tn = t[nxt], pn = p[nxt];
						const int32_t sc1 = nxt < 0 ? czkx : czkx - fn;
						cur = nxt;
						if (sc1 > max_s) max_s = sc1, best_len = m;
						else if (max_s - sc1 > max_drop) break;
						if (cur < 0 || tn != 0) break;
						nxt = pn;
					}
				}
				best_len = __shfl_sync(MMG_FULL, best_len, 0);
				max_s = __shfl_sync(MMG_FULL, max_s, 0);
				__syncwarp();
				for (int j = lane; j < best_len; j += 32) t[v[n_v + j]] = 1;
				if (max_s >= min_sc && best_len > 0 && best_len >= min_cnt) {
					if (lane == 0) u[n_u] = (uint64_t)max_s << 32 | (uint32_t)best_len;
					++n_u, n_v += best_len;
				}
				__syncwarp();
			}
		}
	}
	__syncwarp();
	if (n_u > 0) {
		int k0 = 0;
		for (int ci = 0; ci < n_u; ++ci) { /* compact_a: chains in forward order into cx/cy */
			int ni = (int)(uint32_t)u[ci];
			for (int j = lane; j < ni; j += 32) {
				int src = v[k0 + (ni - j - 1)];
				cx[k0 + j] = ax[src], cy[k0 + j] = ay[src];
			}
			k0 += ni;
		}
		__syncwarp();
		if (lane == 0) { /* order chains by the target position of their first anchor (radix_sort_128x on w[]) */
			int k = 0;
			for (int ci = 0; ci < n_u; ++ci) {
				zx[ci] = cx[k], zy[ci] = (uint64_t)k << 32 | (uint32_t)ci;
				k += (int)(uint32_t)u[ci];
			}
			dev_radix_sort_128x(zx, zy, n_u, bkt, (int*)t);
			for (int ci = 0; ci < n_u; ++ci) zx[n_u + ci] = u[(uint32_t)zy[ci]];
			for (int ci = 0; ci < n_u; ++ci) u[ci] = zx[n_u + ci];
		}
		__syncwarp();
		k0 = 0;
		for (int ci = 0; ci < n_u; ++ci) {
			int ni = (int)(uint32_t)u[ci], src0 = (int)(zy[ci] >> 32);
			for (int j = lane; j < ni; j += 32) ax[k0 + j] = cx[src0 + j], ay[k0 + j] = cy[src0 + j];
			k0 += ni;
		}
		__syncwarp();
	}
	*n_u_ = n_u, *n_v_ = n_v;
}

#endif
