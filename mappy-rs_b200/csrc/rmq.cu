/* rmq.cu -- long-join re-chaining with range-minimum queries, one warp per read.
 *
 * Replaces, on the mm_map path (/root/reference/src/lib.rs:482,587; minimap2
 * v2.26): the re-chain branch of map.c mm_map_frag (radix_sort_128x of the kept
 * anchors + lchain.c mm_lchain_rmq with bw_long) and the parts of krmq.h it
 * uses.  The branch runs only for reads whose first pass left more than one
 * chain, so it is the rare path; what matters is that it is bit-exact.
 *
 * mm_lchain_rmq keeps the active anchors in an AVL tree keyed by (query pos,
 * index) with a subtree-minimum pointer over pri = -(f + 0.5*pen_gap*(x+y)).
 * The answer to a query with equal priorities depends on the tree's shape, so
 * the tree (insert / erase with rotations / rmq / iterator) is replayed exactly,
 * with 32-bit node indices into a per-read slice of a device arena; one lane
 * walks it while the warp handles the data-parallel parts (sort detection,
 * backtrack copy loops).
 */
#include <stdlib.h>
#include "dev_common.cuh"
#include "dev_sort.cuh"
#include "dev_chain.cuh"
#include "stages.h"

#define RMQ_MAX_DEPTH 64
#define NIL (-1)

struct RNode {
	double pri;
	int32_t y, i;
	int32_t c[2];      /* children */
	int32_t s;         /* node with the minimum pri in this subtree */
	uint32_t size;
	int32_t balance;
	int32_t pad;
};

struct RTree {
	RNode *nd;         /* nd[0] is the scratch "fake" root used by erase */
	int32_t root;
};

__device__ __forceinline__ int rn_cmp(int32_t ay, int32_t ai, const RNode &b)
{
	return ay < b.y ? -1 : ay > b.y ? 1 : (ai > b.i) - (ai < b.i);
}
#define RN_LT2(a, b) (nd[(a)].pri < nd[(b)].pri)
#define RN_SIZE_CHILD(q, d) (nd[(q)].c[(d)] != NIL ? nd[nd[(q)].c[(d)]].size : 0u)

__device__ __forceinline__ void rn_update_min(RNode *nd, int p, int q, int r)
{ /* krmq_update_min(p, q, r) */
	nd[p].s = (q == NIL || RN_LT2(p, nd[q].s)) ? p : nd[q].s;
	nd[p].s = (r == NIL || RN_LT2(nd[p].s, nd[r].s)) ? nd[p].s : nd[r].s;
}

__device__ int rn_rotate1(RNode *nd, int p, int dir)
{
	int opp = 1 - dir;
	int q = nd[p].c[opp], s = nd[p].s;
	unsigned size_p = nd[p].size;
	nd[p].size -= nd[q].size - RN_SIZE_CHILD(q, dir);
	nd[q].size = size_p;
	rn_update_min(nd, p, nd[p].c[dir], nd[q].c[dir]);
	nd[q].s = s;
	nd[p].c[opp] = nd[q].c[dir];
	nd[q].c[dir] = p;
	return q;
}

__device__ int rn_rotate2(RNode *nd, int p, int dir)
{
	int b1, opp = 1 - dir;
	int q = nd[p].c[opp], r = nd[q].c[dir], s = nd[p].s;
	unsigned size_x_dir = RN_SIZE_CHILD(r, dir);
	nd[r].size = nd[p].size;
	nd[p].size -= nd[q].size - size_x_dir;
	nd[q].size -= size_x_dir + 1;
	rn_update_min(nd, p, nd[p].c[dir], nd[r].c[dir]);
	rn_update_min(nd, q, nd[q].c[opp], nd[r].c[opp]);
	nd[r].s = s;
	nd[p].c[opp] = nd[r].c[dir];
	nd[r].c[dir] = p;
	nd[q].c[dir] = nd[r].c[opp];
	nd[r].c[opp] = q;
	b1 = dir == 0 ? +1 : -1;
	if (nd[r].balance == b1) nd[q].balance = 0, nd[p].balance = -b1;
	else if (nd[r].balance == 0) nd[q].balance = nd[p].balance = 0;
	else nd[q].balance = b1, nd[p].balance = 0;
	nd[r].balance = 0;
	return r;
}

__device__ void rn_insert(RTree *t, int x)
{ /* krmq_insert; x.y, x.i, x.pri are set by the caller; keys are unique */
	RNode *nd = t->nd;
	unsigned char stack[RMQ_MAX_DEPTH];
	int path[RMQ_MAX_DEPTH];
	int bp = t->root, bq = NIL, p, q, r = NIL, i, which = 0, top, b1, path_len;
	for (p = bp, q = bq, top = path_len = 0; p != NIL; q = p, p = nd[p].c[which]) {
		int cmp = rn_cmp(nd[x].y, nd[x].i, nd[p]);
		if (cmp == 0) return;
		if (nd[p].balance != 0) bq = q, bp = p, top = 0;
		stack[top++] = which = (cmp > 0);
		path[path_len++] = p;
	}
	nd[x].balance = 0, nd[x].size = 1, nd[x].c[0] = nd[x].c[1] = NIL, nd[x].s = x;
	if (q == NIL) t->root = x;
	else nd[q].c[which] = x;
	if (bp == NIL) return;
	for (i = 0; i < path_len; ++i) ++nd[path[i]].size;
	for (i = path_len - 1; i >= 0; --i) {
		rn_update_min(nd, path[i], nd[path[i]].c[0], nd[path[i]].c[1]);
		if (nd[path[i]].s != x) break;
	}
	for (p = bp, top = 0; p != x; p = nd[p].c[stack[top]], ++top)
		if (stack[top] == 0) --nd[p].balance;
		else ++nd[p].balance;
	if (nd[bp].balance > -2 && nd[bp].balance < 2) return;
	which = (nd[bp].balance < 0);
	b1 = which == 0 ? +1 : -1;
	q = nd[bp].c[1 - which];
	if (nd[q].balance == b1) {
		r = rn_rotate1(nd, bp, which);
		nd[q].balance = nd[bp].balance = 0;
	} else r = rn_rotate2(nd, bp, which);
	if (bq == NIL) t->root = r;
	else nd[bq].c[bp != nd[bq].c[0]] = r;
}

__device__ int rn_find(const RTree *t, int32_t y, int32_t i)
{
	const RNode *nd = t->nd;
	int p = t->root;
	while (p != NIL) {
		int cmp = rn_cmp(y, i, nd[p]);
		if (cmp < 0) p = nd[p].c[0];
		else if (cmp > 0) p = nd[p].c[1];
		else break;
	}
	return p;
}

__device__ int rn_erase(RTree *t, int x)
{ /* krmq_erase of an existing node x; returns the removed node */
	RNode *nd = t->nd;
	int p, path[RMQ_MAX_DEPTH];
	unsigned char dir[RMQ_MAX_DEPTH];
	int i, d = 0, cmp;
	const int fake = 0;
	const int32_t xy = nd[x].y, xi = nd[x].i;
	nd[fake] = nd[t->root];
	nd[fake].c[0] = t->root, nd[fake].c[1] = NIL;
	for (cmp = -1, p = fake; cmp; cmp = rn_cmp(xy, xi, nd[p])) {
		int which = (cmp > 0);
		dir[d] = (unsigned char)which;
		path[d++] = p;
		p = nd[p].c[which];
		if (p == NIL) return NIL;
	}
	for (i = 1; i < d; ++i) --nd[path[i]].size;
	if (nd[p].c[1] == NIL) {
		nd[path[d - 1]].c[dir[d - 1]] = nd[p].c[0];
	} else {
		int q = nd[p].c[1];
		if (nd[q].c[0] == NIL) {
			nd[q].c[0] = nd[p].c[0];
			nd[q].balance = nd[p].balance;
			nd[path[d - 1]].c[dir[d - 1]] = q;
			path[d] = q, dir[d++] = 1;
			nd[q].size = nd[p].size - 1;
		} else {
			int r, e = d++;
			for (;;) {
				dir[d] = 0;
				path[d++] = q;
				r = nd[q].c[0];
				if (nd[r].c[0] == NIL) break;
				q = r;
			}
			nd[r].c[0] = nd[p].c[0];
			nd[q].c[0] = nd[r].c[1];
			nd[r].c[1] = nd[p].c[1];
			nd[r].balance = nd[p].balance;
			nd[path[e - 1]].c[dir[e - 1]] = r;
			path[e] = r, dir[e] = 1;
			for (i = e + 1; i < d; ++i) --nd[path[i]].size;
			nd[r].size = nd[p].size - 1;
		}
	}
	for (i = d - 1; i >= 0; --i)
		rn_update_min(nd, path[i], nd[path[i]].c[0], nd[path[i]].c[1]);
	while (--d > 0) {
		int q = path[d];
		int which, other, b1 = 1, b2 = 2;
		which = dir[d], other = 1 - which;
		if (which) b1 = -b1, b2 = -b2;
		nd[q].balance += b1;
		if (nd[q].balance == b1) break;
		else if (nd[q].balance == b2) {
			int r = nd[q].c[other];
			if (nd[r].balance == -b1) {
				nd[path[d - 1]].c[dir[d - 1]] = rn_rotate2(nd, q, which);
			} else {
				nd[path[d - 1]].c[dir[d - 1]] = rn_rotate1(nd, q, which);
				if (nd[r].balance == 0) {
					nd[r].balance = -b1;
					nd[q].balance = b1;
					break;
				} else nd[r].balance = nd[q].balance = 0;
			}
		}
	}
	t->root = nd[fake].c[0];
	return p;
}

__device__ int rn_rmq(const RTree *t, int32_t lo_y, int32_t lo_i, int32_t up_y, int32_t up_i)
{ /* krmq_rmq: closed interval; returns the node with minimal pri or NIL */
	const RNode *nd = t->nd;
	int p = t->root, path[2][RMQ_MAX_DEPTH], mn;
	int plen[2] = {0, 0}, pcmp[2][RMQ_MAX_DEPTH], i, cmp, lca;
	if (p == NIL) return NIL;
	while (p != NIL) {
		cmp = rn_cmp(lo_y, lo_i, nd[p]);
		path[0][plen[0]] = p, pcmp[0][plen[0]++] = cmp;
		if (cmp < 0) p = nd[p].c[0];
		else if (cmp > 0) p = nd[p].c[1];
		else break;
	}
	p = t->root;
	while (p != NIL) {
		cmp = rn_cmp(up_y, up_i, nd[p]);
		path[1][plen[1]] = p, pcmp[1][plen[1]++] = cmp;
		if (cmp < 0) p = nd[p].c[0];
		else if (cmp > 0) p = nd[p].c[1];
		else break;
	}
	for (i = 0; i < plen[0] && i < plen[1]; ++i)
		if (path[0][i] == path[1][i] && pcmp[0][i] <= 0 && pcmp[1][i] >= 0) break;
	if (i == plen[0] || i == plen[1]) return NIL;
	lca = i, mn = path[0][lca];
	for (i = lca + 1; i < plen[0]; ++i) {
		if (pcmp[0][i] <= 0) {
			int q = path[0][i];
			if (RN_LT2(q, mn)) mn = q;
			if (nd[q].c[1] != NIL && RN_LT2(nd[nd[q].c[1]].s, mn)) mn = nd[nd[q].c[1]].s;
		}
	}
	for (i = lca + 1; i < plen[1]; ++i) {
		if (pcmp[1][i] >= 0) {
			int q = path[1][i];
			if (RN_LT2(q, mn)) mn = q;
			if (nd[q].c[0] != NIL && RN_LT2(nd[nd[q].c[0]].s, mn)) mn = nd[nd[q].c[0]].s;
		}
	}
	return mn;
}

__device__ __forceinline__ float rq_mg_log2(float x)
{
	uint32_t zi = __float_as_uint(x);
	float log_2 = (float)(((zi >> 23) & 255u) - 128u);
	zi &= ~(255u << 23);
	zi += 127u << 23;
	float zf = __uint_as_float(zi);
	float t = __fadd_rn(__fmul_rn(-0.34484843f, zf), 2.02466578f);
	t = __fsub_rn(__fmul_rn(t, zf), 0.67487759f);
	return __fadd_rn(log_2, t);
}

__device__ __forceinline__ int32_t rq_sc_simple(uint64_t aix, uint64_t aiy, uint64_t ajx, uint64_t ajy, float pen_gap, float pen_skip, int32_t *exact, int32_t *width)
{ /* lchain.c comput_sc_simple */
	int32_t dq = (int32_t)aiy - (int32_t)ajy, dr, dd, dg, q_span, sc;
	dr = (int32_t)(aix - ajx);
	*width = dd = dr > dq ? dr - dq : dq - dr;
	dg = dr < dq ? dr : dq;
	q_span = (int32_t)(ajy >> 32 & 0xff);
	sc = q_span < dg ? q_span : dg;
	if (exact) *exact = (dd == 0 && dg <= q_span);
	if (dd || dq > q_span) {
		float lin_pen = __fadd_rn(__fmul_rn(pen_gap, (float)dd), __fmul_rn(pen_skip, (float)dg));
		float log_pen = dd >= 1 ? rq_mg_log2((float)(dd + 1)) : 0.0f;
		sc -= (int)__fadd_rn(lin_pen, __fmul_rn(.5f, log_pen));
	}
	return sc;
}

/* lchain.c mm_lchain_rmq: fills f/p/t; run by one lane */
__device__ void dev_lchain_rmq(int max_dist, int max_dist_inner, int bw, int max_chn_skip, int cap_rmq_size, float pen_gap, float pen_skip,
                               int n, const uint64_t *ax, const uint64_t *ay, int32_t *f, int32_t *p, int32_t *t, RNode *nodes)
{
	RTree outer, inner;
	int st = 0, st_inner = 0, i0 = 0, n_alloc = 1, free_head = NIL; /* node 0 is the fake root */
	outer.nd = inner.nd = nodes;
	outer.root = inner.root = NIL;
	if (max_dist < bw) max_dist = bw;
	if (max_dist_inner < 0) max_dist_inner = 0;
	if (max_dist_inner > max_dist) max_dist_inner = max_dist;
	for (int i = 0; i < n; ++i) t[i] = 0;
#define RN_ALLOC(q) do { if (free_head != NIL) { (q) = free_head; free_head = nodes[free_head].c[0]; } else (q) = n_alloc++; } while (0)
#define RN_FREE(q) do { nodes[(q)].c[0] = free_head; free_head = (q); } while (0)
	for (int i = 0; i < n; ++i) {
		int max_j = -1;
		const uint64_t aix = ax[i], aiy = ay[i];
		int32_t q_span = (int32_t)(aiy >> 32 & 0xff), max_f = q_span;
		if (i0 < i && ax[i0] != aix) { /* add in-range anchors */
			for (int j = i0; j < i; ++j) {
				int q, r;
				RN_ALLOC(q);
				nodes[q].y = (int32_t)ay[j], nodes[q].i = j;
				nodes[q].pri = -((double)f[j] + __dmul_rn(__dmul_rn(0.5, (double)pen_gap), (double)((int32_t)ax[j] + (int32_t)ay[j])));
				rn_insert(&outer, q);
				if (max_dist_inner > 0) {
					RN_ALLOC(r);
					nodes[r].y = nodes[q].y, nodes[r].i = nodes[q].i, nodes[r].pri = nodes[q].pri;
					rn_insert(&inner, r);
				}
			}
			i0 = i;
		}
		while (st < i && ((aix >> 32) != (ax[st] >> 32) || aix > ax[st] + (uint64_t)max_dist || (int)(outer.root != NIL ? nodes[outer.root].size : 0) > cap_rmq_size)) {
			int q = rn_find(&outer, (int32_t)ay[st], st);
			if (q != NIL) { q = rn_erase(&outer, q); RN_FREE(q); }
			++st;
		}
		if (max_dist_inner > 0) {
			while (st_inner < i && ((aix >> 32) != (ax[st_inner] >> 32) || aix > ax[st_inner] + (uint64_t)max_dist_inner || (int)(inner.root != NIL ? nodes[inner.root].size : 0) > cap_rmq_size)) {
				int q = rn_find(&inner, (int32_t)ay[st_inner], st_inner);
				if (q != NIL) { q = rn_erase(&inner, q); RN_FREE(q); }
				++st_inner;
			}
		}
		{
			int q = rn_rmq(&outer, (int32_t)aiy - max_dist, 0x7fffffff, (int32_t)aiy, 0);
			if (q != NIL) {
				int32_t sc, exact, width, n_skip = 0;
				int j = nodes[q].i;
				sc = f[j] + rq_sc_simple(aix, aiy, ax[j], ay[j], pen_gap, pen_skip, &exact, &width);
				if (width <= bw && sc > max_f) max_f = sc, max_j = j;
				if (!exact && inner.root != NIL && (int32_t)aiy > 0) {
					/* krmq_interval(root_inner, (y-1, n)): lower = largest node <= key; then iterate downwards */
					const int32_t ky = (int32_t)aiy - 1, ki = n;
					int stack[RMQ_MAX_DEPTH], top = -1, lo = NIL, pp = inner.root;
					while (pp != NIL) {
						int cmp = rn_cmp(ky, ki, nodes[pp]);
						if (cmp < 0) pp = nodes[pp].c[0];
						else if (cmp > 0) lo = pp, pp = nodes[pp].c[1];
						else { lo = pp; break; }
					}
					if (lo != NIL) {
						/* krmq_itr_find(root_inner, lo): stack = path to lo */
						pp = inner.root;
						while (pp != NIL) {
							stack[++top] = pp;
							int cmp = rn_cmp(nodes[lo].y, nodes[lo].i, nodes[pp]);
							if (cmp < 0) pp = nodes[pp].c[0];
							else if (cmp > 0) pp = nodes[pp].c[1];
							else break;
						}
						while (top >= 0) {
							int qq = stack[top];
							if (nodes[qq].y < (int32_t)aiy - max_dist_inner) break;
							j = nodes[qq].i;
							sc = f[j] + rq_sc_simple(aix, aiy, ax[j], ay[j], pen_gap, pen_skip, 0, &width);
							if (width <= bw) {
								if (sc > max_f) {
									max_f = sc, max_j = j;
									if (n_skip > 0) --n_skip;
								} else if (t[j] == i) {
									if (++n_skip > max_chn_skip) break;
								}
								if (p[j] >= 0) t[p[j]] = i;
							}
							/* krmq_itr_prev */
							{
								int c0 = nodes[stack[top]].c[0];
								if (c0 != NIL) {
									for (pp = c0; pp != NIL; pp = nodes[pp].c[1]) stack[++top] = pp;
								} else {
									int prev;
									do { prev = stack[top--]; } while (top >= 0 && prev == nodes[stack[top]].c[0]);
									if (top < 0) break;
								}
							}
						}
					}
				}
			}
		}
		f[i] = max_f, p[i] = max_j;
	}
#undef RN_ALLOC
#undef RN_FREE
}

/* mm_lchain_rmq with the OUTER tree replaced by what it stands for.  The outer tree only ever holds the anchors
 * [st, i0) - inserted when the target position moves on, erased from the front - so its range-minimum query is a scan
 * of that index range for the smallest priority among the anchors whose query position lies in the interval: 32 anchors
 * per step and a warp reduction instead of ~30 dependent node visits.  The tree matters in one case only: when the
 * minimum is attained twice, krmq's answer depends on the shape of the tree.  Then the function gives up (returns 1)
 * and the caller replays the read with the tree (dev_lchain_rmq).  The inner tree (anchors within max_dist_inner, walked
 * in key order with the n_skip / t[] logic) is kept as it is, on lane 0.  Priorities are compared as order-preserving
 * 64-bit keys of the very doubles upstream computes. */
__device__ int dev_lchain_rmq_warp(int max_dist, int max_dist_inner, int bw, int max_chn_skip, int cap_rmq_size, float pen_gap, float pen_skip,
                                   int n, const uint64_t *ax, const uint64_t *ay, int32_t *f, int32_t *p, int32_t *t, RNode *nodes, unsigned long long *pk, int fake_tie_at)
{
	const int lane = mmg_lane();
	RTree inner;
	int st = 0, st_inner = 0, i0 = 0, n_alloc = 1, free_head = NIL; /* node 0 is the fake root */
	inner.nd = nodes, inner.root = NIL;
	if (max_dist < bw) max_dist = bw;
	if (max_dist_inner < 0) max_dist_inner = 0;
	if (max_dist_inner > max_dist) max_dist_inner = max_dist;
	for (int k = lane; k < n; k += 32) t[k] = 0;
	__syncwarp();
#define RN_ALLOC(q) do { if (free_head != NIL) { (q) = free_head; free_head = nodes[free_head].c[0]; } else (q) = n_alloc++; } while (0)
#define RN_FREE(q) do { nodes[(q)].c[0] = free_head; free_head = (q); } while (0)
#define RN_PRI(j) (-((double)f[(j)] + __dmul_rn(__dmul_rn(0.5, (double)pen_gap), (double)((int32_t)ax[(j)] + (int32_t)ay[(j)]))))
	for (int i = 0; i < n; ++i) {
		const uint64_t aix = ax[i], aiy = ay[i];
		if (i0 < i && ax[i0] != aix) { /* the anchors [i0, i) come into range */
			for (int j = i0 + lane; j < i; j += 32) {
				const unsigned long long b = (unsigned long long)__double_as_longlong(RN_PRI(j));
				pk[j] = b >> 63 ? ~b : b | 0x8000000000000000ull;
			}
			if (lane == 0 && max_dist_inner > 0)
				for (int j = i0; j < i; ++j) {
					int r;
					RN_ALLOC(r);
					nodes[r].y = (int32_t)ay[j], nodes[r].i = j, nodes[r].pri = RN_PRI(j);
					rn_insert(&inner, r);
				}
			i0 = i;
			__syncwarp();
		}
		while (st < i && ((aix >> 32) != (ax[st] >> 32) || aix > ax[st] + (uint64_t)max_dist || i0 - st > cap_rmq_size)) ++st;
		if (max_dist_inner > 0) {
			while (st_inner < i && ((aix >> 32) != (ax[st_inner] >> 32) || aix > ax[st_inner] + (uint64_t)max_dist_inner || i0 - st_inner > cap_rmq_size)) {
				if (lane == 0) {
					int q = rn_find(&inner, (int32_t)ay[st_inner], st_inner);
					if (q != NIL) { q = rn_erase(&inner, q); RN_FREE(q); }
				}
				++st_inner;
			}
		}
		/* krmq_rmq(outer, (y - max_dist, INT_MAX), (y, 0)): closed interval in (query position, index) order */
		const int32_t yi = (int32_t)aiy, ylo = yi - max_dist;
		unsigned long long best = ~0ull;
		int bestj = -1, cnt = 0;
		for (int j = st + lane; j < i0; j += 32) {
			const int32_t yj = (int32_t)ay[j];
			if (yj > ylo && (yj < yi || (yj == yi && j == 0))) {
				const unsigned long long k = pk[j];
				if (k < best) best = k, bestj = j, cnt = 1;
				else if (k == best) ++cnt;
			}
		}
		unsigned long long m = best;
#pragma unroll
		for (int d = 16; d; d >>= 1) { const unsigned long long o = __shfl_xor_sync(MMG_FULL, m, d); m = o < m ? o : m; }
		int qj = -1;
		if (fake_tie_at > 0 && i + 1 == fake_tie_at) return 1;
		if (m != ~0ull) {
			if (__reduce_add_sync(MMG_FULL, best == m ? cnt : 0) > 1) return 1;   /* the minimum is not unique: the tree decides */
			const unsigned w = __ballot_sync(MMG_FULL, best == m);
			qj = __shfl_sync(MMG_FULL, bestj, __ffs((int)w) - 1);
		}
		if (lane == 0) {
			int max_j = -1;
			int32_t q_span = (int32_t)(aiy >> 32 & 0xff), max_f = q_span;
			if (qj >= 0) {
				int32_t sc, exact, width, n_skip = 0;
				int j = qj;
				sc = f[j] + rq_sc_simple(aix, aiy, ax[j], ay[j], pen_gap, pen_skip, &exact, &width);
				if (width <= bw && sc > max_f) max_f = sc, max_j = j;
				if (!exact && inner.root != NIL && (int32_t)aiy > 0) {
					/* krmq_interval(root_inner, (y-1, n)): lower = largest node <= key; then iterate downwards */
					const int32_t ky = (int32_t)aiy - 1, ki = n;
					int stack[RMQ_MAX_DEPTH], top = -1, lo = NIL, pp = inner.root;
					while (pp != NIL) {
						int cmp = rn_cmp(ky, ki, nodes[pp]);
						if (cmp < 0) pp = nodes[pp].c[0];
						else if (cmp > 0) lo = pp, pp = nodes[pp].c[1];
						else { lo = pp; break; }
					}
					if (lo != NIL) {
						pp = inner.root;
						while (pp != NIL) {
							stack[++top] = pp;
							int cmp = rn_cmp(nodes[lo].y, nodes[lo].i, nodes[pp]);
							if (cmp < 0) pp = nodes[pp].c[0];
							else if (cmp > 0) pp = nodes[pp].c[1];
							else break;
						}
						while (top >= 0) {
							int qq = stack[top];
							if (nodes[qq].y < (int32_t)aiy - max_dist_inner) break;
							j = nodes[qq].i;
							sc = f[j] + rq_sc_simple(aix, aiy, ax[j], ay[j], pen_gap, pen_skip, 0, &width);
							if (width <= bw) {
								if (sc > max_f) {
									max_f = sc, max_j = j;
									if (n_skip > 0) --n_skip;
								} else if (t[j] == i) {
									if (++n_skip > max_chn_skip) break;
								}
								if (p[j] >= 0) t[p[j]] = i;
							}
							{ /* krmq_itr_prev */
								int c0 = nodes[stack[top]].c[0];
								if (c0 != NIL) {
									for (pp = c0; pp != NIL; pp = nodes[pp].c[1]) stack[++top] = pp;
								} else {
									int prev;
									do { prev = stack[top--]; } while (top >= 0 && prev == nodes[stack[top]].c[0]);
									if (top < 0) break;
								}
							}
						}
					}
				}
			}
			f[i] = max_f, p[i] = max_j;
		}
		__syncwarp();
	}
#undef RN_ALLOC
#undef RN_FREE
#undef RN_PRI
	return 0;
}

__global__ void __launch_bounds__(CHAIN_WARPS * 32)
rechain_kernel(ChunkDev c, DevOpt o, uint32_t r0, uint32_t r1, RNode *nodes, int serial, uint32_t *work)
{
	__shared__ int s_bkt[CHAIN_WARPS][512];
	const int lane = mmg_lane(), wib = threadIdx.x >> 5;
	unsigned long long tot_rechain = 0;
	for (;;) {
		uint32_t r = r0 + mmg_next_item(work);
		if (r >= r1) break;
		r = mmg_read_of(c, r);
		int n_u = (int)c.n_u[r];
		if (!(o.bw_long > o.bw && !(o.flag & MMG_F_NO_LJOIN) && n_u > 1)) continue;
		const uint64_t ab = c.a_off[r] - c.a_off0;
		const int qlen = (int)(c.off[r + 1] - c.off[r]);
		uint64_t *ax = c.bx + ab, *ay = c.by + ab;
		const int32_t st = (int32_t)ay[0], en = (int32_t)ay[(uint32_t)c.u[ab] - 1];
		if (!(qlen - (en - st) > o.rmq_rescue_size || (float)(en - st) > __fmul_rn((float)qlen, o.rmq_rescue_ratio))) continue;
		int n = (int)c.n_v[r], n_v = 0;
		int32_t *f = c.f + ab, *p = c.p + ab, *t = c.t + ab, *v = c.v + ab;
		RNode *nd = nodes + 2 * ab + 2 * (uint64_t)r;   /* disjoint per read: 2 nodes per anchor + 2 (the arena holds 2 * (anchors + reads) nodes) */
		if (lane == 0) {
			dev_radix_sort_128x(ax, ay, n, s_bkt[wib], (int*)v);
			c.flags[r] |= 2u;
		}
		__syncwarp();
		/* the warp form needs n + 1 nodes for its one tree; the keys go into the upper half of the read's slice */
		if ((serial & 1) || dev_lchain_rmq_warp(o.max_gap, o.rmq_inner_dist, o.bw_long, o.max_chain_skip, o.rmq_size_cap, o.chn_pen_gap, o.chn_pen_skip,
		                                        n, ax, ay, f, p, t, nd, (unsigned long long*)(nd + n + 2), serial >> 1)) {
			__syncwarp();
			if (lane == 0)
				dev_lchain_rmq(o.max_gap, o.rmq_inner_dist, o.bw_long, o.max_chain_skip, o.rmq_size_cap, o.chn_pen_gap, o.chn_pen_skip, n, ax, ay, f, p, t, nd);
		}
		__syncwarp();
		dev_backtrack_compact(n, ax, ay, f, p, t, v, c.zx + 2 * ab, c.zy + 2 * ab, c.cx + ab, c.cy + ab, c.u + ab,
		                      o.min_cnt, o.min_chain_score, o.bw_long, s_bkt[wib], &n_u, &n_v);
		if (lane == 0) c.n_u[r] = (uint32_t)n_u, c.n_v[r] = (uint32_t)n_v;
		tot_rechain += 1;
		__syncwarp();
	}
	if (lane == 0 && tot_rechain) atomicAdd(&c.stats[9], tot_rechain);
}

/* MM_F_RMQ (the asm5 / asm10 / asm20 presets): map.c mm_map_frag chains with mm_lchain_rmq instead of mm_lchain_dp.
 * Same routine as the re-chain step, on the sorted anchors of the read, band bw; backtrack_kernel follows as usual. */
__global__ void __launch_bounds__(CHAIN_WARPS * 32)
chain_rmq_kernel(ChunkDev c, DevOpt o, uint32_t r0, uint32_t r1, RNode *nodes, int serial, uint32_t *work)
{
	const int lane = mmg_lane();
	for (;;) {
		uint32_t r = r0 + mmg_next_item(work);
		if (r >= r1) break;
		r = mmg_read_of(c, r);
		const int n = (int)c.n_a[r];
		const uint64_t ab = c.a_off[r] - c.a_off0;
		RNode *nd = nodes + 2 * ab + 2 * (uint64_t)r;
		if (n > 0 && ((serial & 1) || dev_lchain_rmq_warp(o.max_gap, o.rmq_inner_dist, o.bw, o.max_chain_skip, o.rmq_size_cap, o.chn_pen_gap, o.chn_pen_skip,
		                                                  n, c.bx + ab, c.by + ab, c.f + ab, c.p + ab, c.t + ab, nd, (unsigned long long*)(nd + n + 2), serial >> 1))) {
			__syncwarp();
			if (lane == 0)
				dev_lchain_rmq(o.max_gap, o.rmq_inner_dist, o.bw, o.max_chain_skip, o.rmq_size_cap, o.chn_pen_gap, o.chn_pen_skip,
				               n, c.bx + ab, c.by + ab, c.f + ab, c.p + ab, c.t + ab, nd);
		}
		__syncwarp();
	}
}

/* MMG_RMQ_SERIAL=1: every read through the tree replay (the form the warp version is checked against);
 * MMG_RMQ_SERIAL=2k (k > 0): the warp version pretends a tie at its k-th anchor, so that tests walk the
 * "abandon the read half way and replay it" path, which real ties take too rarely to rely on.  Read per launch. */
static int rmq_serial(void)
{
	const char *e = getenv("MMG_RMQ_SERIAL");
	return e ? atoi(e) : 0;
}

int launch_chain_rmq(const ChunkDev &c, const DevOpt &o, uint32_t r0, uint32_t r1, void *nodes, int n_sms, cudaStream_t st, uint32_t *work)
{
	int grid = n_sms * 8, need = ((int)(r1 - r0) + CHAIN_WARPS - 1) / CHAIN_WARPS;
	if (grid > need) grid = need;
	if (grid < 1) grid = 1;
	MMG_LAUNCH(chain_rmq_kernel, grid, CHAIN_WARPS * 32, 0, st, c, o, r0, r1, (RNode*)nodes, rmq_serial(), work);
	return 0;
}

int launch_rechain(const ChunkDev &c, const DevOpt &o, uint32_t r0, uint32_t r1, void *nodes, int n_sms, cudaStream_t st, uint32_t *work)
{
	int grid = n_sms * 8, need = ((int)(r1 - r0) + CHAIN_WARPS - 1) / CHAIN_WARPS;
	if (grid > need) grid = need;
	if (grid < 1) grid = 1;
	MMG_LAUNCH(rechain_kernel, grid, CHAIN_WARPS * 32, 0, st, c, o, r0, r1, (RNode*)nodes, rmq_serial(), work);
	return 0;
}
