/* dev_common.cuh -- device-side types and helpers shared by the stage kernels.
 * Compiled by nvcc for sm_100a (product) and, for the CPU test-suite only, by
 * g++ against tests/emu/mmg_emu.h (MMG_EMU). */
#ifndef MMG_DEV_COMMON_CUH
#define MMG_DEV_COMMON_CUH

#ifdef MMG_EMU
#include "mmg_emu.h"
#define MMG_LAUNCH(kernel, grid, block, smem, stream, ...) \
	(emu_kernel_name = #kernel, emu_launch(dim3(grid), dim3(block), (smem), [=]() { kernel(__VA_ARGS__); }))
#define MMG_DYN_SMEM(name) unsigned char *name = emu_dyn_smem
#else
#include <cuda_runtime.h>
#define MMG_LAUNCH(kernel, grid, block, smem, stream, ...) \
	kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define MMG_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

#include <stdint.h>
#include "../../include/mmg.h"

/* true the first time it is called on the current device for this flag array: kernel attributes
 * (cudaFuncSetAttribute) are per device, and a multi-device aligner launches every kernel on several */
static inline bool mmg_once_per_device(unsigned char *done /* [64] */)
{
	int dev = 0;
	cudaGetDevice(&dev);
	if (dev < 0 || dev >= 64 || done[dev]) return false;
	done[dev] = 1;
	return true;
}

#define MMG_FULL 0xffffffffu
#define MMG_INF64 0xffffffffffffffffULL

struct __align__(16) mmg_u128 { uint64_t x, y; };

/* GPU-resident index (north-star (a)); layout described in mmg_internal.h */
struct DevIndex {
	int32_t k, w, b, flag;
	uint32_t n_seq, hbits;
	const mmg_u128 *htab;     /* {key, val} slots */
	const uint64_t *pos;
	const uint32_t *S;        /* 4-bit packed reference */
	const uint64_t *seq_off;  /* n_seq + 1 */
	const uint32_t *seq_len;
};

/* mapping options the kernels need (from mmg_mapopt_t, after mm_mapopt_update) */
struct DevOpt {
	int64_t flag;
	int32_t seed;
	int32_t bw, bw_long, max_gap, max_gap_ref, max_frag_len;
	int32_t max_chain_skip, max_chain_iter, min_cnt, min_chain_score;
	float chn_pen_gap, chn_pen_skip;      /* chain_gap_scale*0.01*k, computed on the host in double like map.c */
	int32_t rmq_size_cap, rmq_inner_dist, rmq_rescue_size;
	float rmq_rescue_ratio;
	float mask_level; int32_t mask_len; float pri_ratio; int32_t best_n;
	float alt_drop;
	int32_t a, b, q, e, q2, e2, sc_ambi, zdrop, zdrop_inv, end_bonus, min_dp_max, min_ksw_len;
	int32_t anchor_ext_len, anchor_ext_shift;
	float max_clip_ratio;
	float q_occ_frac;
	int32_t mid_occ, max_max_occ, occ_dist, max_qlen;
	int64_t max_sw_mat;
};

/* One region record on the device: mm_reg1_t + the mm_extra_t scalars */
struct DevReg {
	int32_t id, cnt, rid, score;
	int32_t qs, qe, rs, re;
	int32_t parent, subsc;
	int32_t as;
	int32_t mlen, blen;
	int32_t n_sub, score0;
	uint32_t hash;
	float div;
	uint32_t bits;     /* mapq:8 | split:2<<8 | rev<<10 | inv<<11 | sam_pri<<12 | strand_retained<<13 | split_inv<<14 | has_p<<15 */
	int32_t dp_score, dp_max, dp_max2, n_ambi;
	uint32_t n_cigar; uint32_t pad;
	uint64_t cigar_off;
};
#define REG_MAPQ(r)   ((r).bits & 0xffu)
#define REG_REV(r)    (((r).bits >> 10) & 1u)
#define REG_INV(r)    (((r).bits >> 11) & 1u)
#define REG_SAMPRI(r) (((r).bits >> 12) & 1u)
#define REG_SRET(r)   (((r).bits >> 13) & 1u)
#define REG_HASP(r)   (((r).bits >> 15) & 1u)

/* Device buffers of one chunk of reads (SoA, sized once per aligner and reused).
 * Per-read slices of the minimizer/seed arrays start at the read's base offset
 * (a read of L bases yields at most L minimizers); anchor-sized arrays are
 * sliced by the exclusive scan a_off[]. */
struct ChunkDev {
	uint32_t n_reads;
	const char *seq;          /* concatenated ASCII bases */
	const uint64_t *off;      /* n_reads + 1: absolute offsets into seq */
	uint64_t off0;            /* off[0]: per-read slices of the base-sized arrays start at off[r] - off0 */
	const uint32_t *order;    /* work order of the reads of the chunk, longest first (0 = index order): every per-read kernel claims
	                           * item i and works on read order[i], so that a launch does not end with its longest reads */
	const uint64_t *seg_beg;  /* segment mode of the sketch kernel (index construction), else 0 */
	const uint32_t *seg_len;
	uint32_t seg_cap;
	/* sketch */
	uint64_t *mz_x; uint32_t *mz_y; uint32_t *n_mz;
	/* seeds */
	uint64_t *sd_val; uint32_t *sd_n; uint32_t *sd_qpos; uint32_t *sd_meta; /* meta = span<<8 | flt<<1 | tandem */
	uint32_t *n_seed; uint32_t *n_a; int32_t *rep_len;
	uint64_t *a_off;          /* n_reads + 1 (exclusive scan of n_a) */
	uint64_t a_off0;          /* anchor-sized arrays are sliced at a_off[r] - a_off0 (sub-range of a chunk) */
	uint64_t *af_off;         /* n_reads + 1: scan of the anchor counts BEFORE the isolated-anchor filter */
	uint64_t *hit_scratch;    /* index position words of the reads the filter looked at, at af_off[r] (bit 63 = strand of the seed) */
	uint32_t *keep_bits;      /* one bit per unfiltered anchor of a filtered read (flags bit 2), at word (af_off[r] >> 5) + r */
	/* anchors */
	uint64_t *ax, *ay, *bx, *by;
	int32_t *f, *p, *t, *v;
	uint64_t *zx, *zy;        /* sort scratch (2 x anchors) */
	/* chains */
	uint64_t *cx, *cy, *u;
	uint32_t *n_u, *n_v;
	uint64_t *r_off;          /* n_reads + 1 (exclusive scan of n_u) */
	DevReg *regs; uint32_t *n_regs;
	uint64_t *h_off;          /* n_reads + 1 (exclusive scan of n_regs) */
	/* counters */
	unsigned long long *stats; /* MMG_N_STATS */
	uint32_t *work;            /* dynamic work counters, one per kernel launch */
	uint32_t *err;             /* one word: OR of every arena-overflow flag raised while the chunk ran (read back by the host) */
	uint32_t *big_list;        /* reads deferred by a kernel's small-tile pass to its large-tile pass */
	uint32_t *tie_list;        /* reads whose anchors have equal keys (sort stage) */
	uint32_t *flags;           /* per read: bit0 = anchor ties (exact re-sort done), bit1 = re-chained, bit2 = isolated anchors dropped */
};

__device__ __forceinline__ int mmg_lane() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t mmg_lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }

/* warp-cooperative claim of the next work item (persistent warps) */
__device__ __forceinline__ uint32_t mmg_next_item(uint32_t *counter)
{
	uint32_t r = 0;
	if (mmg_lane() == 0) r = atomicAdd(counter, 1u);
	return __shfl_sync(MMG_FULL, r, 0);
}

__device__ __forceinline__ uint32_t mmg_read_of(const ChunkDev &c, uint32_t item) { return c.order ? c.order[item] : item; }

__device__ __forceinline__ int mmg_warp_excl_scan(int v, int *total)
{
	int lane = mmg_lane(), x = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		int y = __shfl_up_sync(MMG_FULL, x, o);
		if (lane >= o) x += y;
	}
	*total = __shfl_sync(MMG_FULL, x, 31);
	return x - v;
}

#endif
