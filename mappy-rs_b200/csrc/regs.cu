/* regs.cu -- chains to regions, primary/secondary selection, divergence
 * estimate and mapping quality on the device (north-star (e)), one warp per read.
 *
 * Replaces, on the mm_map path (/root/reference/src/lib.rs:482,587; minimap2
 * v2.26): hit.c mm_gen_regs / mm_reg_set_coor / mm_cal_fuzzy_len, map.c
 * chain_post (mm_set_parent, mm_select_sub, mm_sync_regs), esterr.c mm_est_err,
 * hit.c mm_filter_strand_retained and mm_set_mapq.  A read has a handful of
 * regions, so the logic is serial on lane 0 (the fuzzy-length sums run on the
 * whole warp); thousands of reads are in flight on other warps.
 *
 * Floating point: mm_set_mapq uses float arithmetic and glibc logf().  Device
 * code uses explicit round-to-nearest multiplies/adds (no FMA contraction) and
 * dev_logf(), a restatement of glibc 2.39's logf algorithm (table + degree-3
 * polynomial evaluated in double) that the CPU test-suite checks against the
 * host libm over every positive normal float.
 */
#include <string.h>
#include "dev_common.cuh"
#include "dev_sort.cuh"
#include "dev_regs.cuh"
#include "stages.h"

__global__ void __launch_bounds__(CHAIN_WARPS * 32)
regs_kernel(ChunkDev c, DevIndex di, DevOpt o, uint32_t r0, uint32_t r1, uint64_t regs_cap, uint32_t *work)
{
	__shared__ int s_bkt[CHAIN_WARPS][512];
	const int lane = mmg_lane(), wib = threadIdx.x >> 5;
	unsigned long long tot_kept = 0, tot_regs = 0;
	for (;;) {
		uint32_t r = r0 + mmg_next_item(work);
		if (r >= r1) break;
		r = mmg_read_of(c, r);
		const int n_u = (int)c.n_u[r];
		const uint64_t ab = c.a_off[r] - c.a_off0, rb = c.r_off[r];
		const int qlen = (int)(c.off[r + 1] - c.off[r]);
		int n_regs = 0;
		if (n_u > 0 && rb + (uint64_t)n_u <= regs_cap) {
			DevReg *regs = c.regs + rb;
			const uint64_t *ax = c.bx + ab, *ay = c.by + ab;
			uint64_t *zx = c.zx + 2 * ab, *zy = c.zy + 2 * ab;
			uint32_t hash = dev_wang_hash((uint32_t)qlen) + dev_wang_hash((uint32_t)o.seed); /* qname == NULL */
			hash = dev_wang_hash(hash);
			dev_gen_regs(hash, qlen, n_u, c.u + ab, ax, ay, regs, zx, zy, s_bkt[wib], (int*)(c.t + ab));
			n_regs = n_u;
			if (lane == 0) {
				if (!(o.flag & MMG_F_ALL_CHAINS)) { /* map.c chain_post */
					dev_set_parent(o.mask_level, o.mask_len, n_regs, regs, o.a * 2 + o.b, o.alt_drop, zx, (int*)zy);
					dev_select_sub(o.pri_ratio, di.k * 2, o.best_n, 1, (int)(o.max_gap * 0.8), &n_regs, regs, (int*)zy);
				}
			}
			n_regs = __shfl_sync(MMG_FULL, n_regs, 0);
			__syncwarp();
			{
				const uint64_t base = c.off[r] - c.off0;
				dev_est_err(di, qlen, n_regs, regs, ax, ay, (int)c.n_seed[r], c.sd_qpos + base, c.sd_meta + base);
			}
			if (lane == 0) {
				n_regs = dev_filter_strand_retained(n_regs, regs);
				if (!(o.flag & MMG_F_CIGAR))
					dev_set_mapq(n_regs, regs, o.min_chain_score, o.a, c.rep_len[r]);
			}
			n_regs = __shfl_sync(MMG_FULL, n_regs, 0);
		} else if (n_u > 0) {
			if (lane == 0) atomicOr(&c.flags[r], 0x80000000u), atomicOr(c.err, 0x80000000u); /* region arena overflow: reported by the host */
		}
		if (lane == 0) c.n_regs[r] = (uint32_t)n_regs;
		tot_kept += c.n_v[r], tot_regs += n_regs;
		__syncwarp();
	}
	if (lane == 0 && (tot_kept | tot_regs)) {
		atomicAdd(&c.stats[6], tot_kept);
		if (!(o.flag & MMG_F_CIGAR)) atomicAdd(&c.stats[8], tot_regs); /* with CIGAR the final count is taken after alignment (ext_final_kernel) */
	}
}

/* dense hit records for the device->host copy */
__global__ void __launch_bounds__(256)
pack_hits_kernel(ChunkDev c, uint32_t r0, uint32_t r1, mmg_hit_t *hits)
{
	for (uint32_t r = r0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < r1; r += gridDim.x * (blockDim.x >> 5)) {
		const int n = (int)c.n_regs[r];
		const DevReg *regs = c.regs + c.r_off[r];
		mmg_hit_t *h = hits + c.h_off[r];
		for (int i = mmg_lane(); i < n; i += 32) {
			const DevReg g = regs[i];
			mmg_hit_t o;
			o.rid = g.rid, o.rs = g.rs, o.re = g.re, o.qs = g.qs, o.qe = g.qe;
			o.mlen = g.mlen, o.blen = g.blen;
			o.score = g.score, o.score0 = g.score0, o.cnt = g.cnt, o.subsc = g.subsc, o.n_sub = g.n_sub;
			o.parent = g.parent, o.id = g.id;
			o.dp_score = g.dp_score, o.dp_max = g.dp_max, o.dp_max2 = g.dp_max2;
			o.n_ambi = g.n_ambi;
			o.nm = REG_HASP(g) ? g.blen - g.mlen + g.n_ambi : 0;
			o.hash = g.hash, o.div = g.div;
			o.rev = (uint8_t)REG_REV(g), o.mapq = (uint8_t)REG_MAPQ(g), o.is_primary = (uint8_t)(g.parent == g.id);
			o.flags = (uint8_t)((REG_SAMPRI(g) ? 1 : 0) | (REG_INV(g) ? 2 : 0) | (REG_SRET(g) ? 4 : 0) | ((g.bits >> 8 & 1u) ? 8 : 0) | ((g.bits >> 9 & 1u) ? 16 : 0) | (REG_HASP(g) ? 32 : 0));
			o.n_cigar = g.n_cigar, o.cigar_off = g.cigar_off;
			h[i] = o;
			((uint32_t*)&h[i])[23] = 0; /* the padding word between n_cigar and cigar_off: a batch mapped twice returns identical bytes */
		}
	}
}

/* test hook: dev_logf over an array (the mapq path's only transcendental) */
__global__ void logf_kernel(const float *x, float *y, uint64_t n)
{
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) y[i] = dev_logf(x[i]);
}

extern "C" int mmg_debug_logf(const float *x, float *y, uint64_t n)
{
	float *dx = 0, *dy = 0;
	int n_dev = 0;
	if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) return MMG_ENODEV;
	if (cudaMalloc((void**)&dx, n * 4) != cudaSuccess || cudaMalloc((void**)&dy, n * 4) != cudaSuccess) return MMG_ENOMEM;
	cudaMemcpy(dx, x, n * 4, cudaMemcpyHostToDevice);
	MMG_LAUNCH(logf_kernel, 64, 256, 0, (cudaStream_t)0, (const float*)dx, dy, n);
	cudaError_t e = cudaMemcpy(y, dy, n * 4, cudaMemcpyDeviceToHost);
	cudaFree(dx); cudaFree(dy);
	return e == cudaSuccess ? MMG_OK : MMG_ECUDA;
}

int launch_regs(const ChunkDev &c, const DevIndex &di, const DevOpt &o, uint32_t r0, uint32_t r1, uint64_t regs_cap, int n_sms, cudaStream_t st, uint32_t *work)
{
	int grid = n_sms * 12, need = ((int)(r1 - r0) + CHAIN_WARPS - 1) / CHAIN_WARPS;
	if (grid > need) grid = need;
	if (grid < 1) grid = 1;
	MMG_LAUNCH(regs_kernel, grid, CHAIN_WARPS * 32, 0, st, c, di, o, r0, r1, regs_cap, work);
	return 0;
}

int launch_pack_hits(const ChunkDev &c, uint32_t r0, uint32_t r1, mmg_hit_t *hits, int n_sms, cudaStream_t st)
{
	int grid = n_sms * 4, need = ((int)(r1 - r0) + 7) / 8;
	if (grid > need) grid = need;
	if (grid < 1) grid = 1;
	MMG_LAUNCH(pack_hits_kernel, grid, 256, 0, st, c, r0, r1, hits);
	return 0;
}
