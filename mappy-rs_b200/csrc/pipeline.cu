/* pipeline.cu -- the device pipeline behind mmg_map_batch().
 *
 * Replaces, for the mappy-rs host, the per-read `mm_map` calls of `Aligner.map`
 * and of the `map_batch` worker threads (/root/reference/src/lib.rs:482-488,
 * 541-636): a batch of reads is cut into chunks sized for the device arenas,
 * and every chunk runs sketch -> seed -> expand -> sort -> chain -> backtrack
 * -> (re-chain) -> regions/mapq as one kernel per stage with one warp (or CTA)
 * per read.  The index is uploaded once at mmg_aligner_create() and stays
 * resident (north-star (a)).  There is no CPU mapping path.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>
#include <mutex>
#include <thread>
#include <string>
#include "mmg_internal.h"
#include "dev_common.cuh"
#include "stages.h"
#include "extend.h"

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { mmg_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); return MMG_ECUDA; } } while (0)

/* MMG_DEBUG_SYNC=1: synchronise after every stage and name the stage whose kernels faulted (fault isolation only) */
static bool debug_sync() { static int v = -1; if (v < 0) v = getenv("MMG_DEBUG_SYNC") ? 1 : 0; return v == 1; }
static void debug_check(const char *where, int device)
{
	if (!debug_sync()) return;
	int cur = 0;
	cudaGetDevice(&cur);
	cudaSetDevice(device);
	cudaError_t e = cudaDeviceSynchronize();
	fprintf(stderr, "[mmg debug] %s: device %d: %s\n", where, device, cudaGetErrorString(e));
	cudaSetDevice(cur);
}

static const char *g_stage_names[MMG_N_STAGES] = { "h2d", "sketch", "seed", "scan", "expand", "sort", "chain_dp", "backtrack", "rechain", "regs", "extend", "d2h" };

/* Pinned host blocks that receive the hit records of streamed batches directly (no staging copy, no page faults on a
 * fresh allocation); a block goes back to the pool when its batch is destroyed.  Reference-counted so that a batch
 * may outlive the aligner. */
struct HostPool {
	std::mutex mu;
	std::vector<std::pair<void*, uint64_t> > blocks; /* (pointer, bytes) */
	bool dead;
	int refs;
	HostPool() : dead(false), refs(1) {}
};

static void *pool_acquire(HostPool *hp, uint64_t bytes, uint64_t *got)
{
	{
		std::lock_guard<std::mutex> g(hp->mu);
		int best = -1;
		for (size_t i = 0; i < hp->blocks.size(); ++i)
			if (hp->blocks[i].second >= bytes && (best < 0 || hp->blocks[i].second < hp->blocks[best].second)) best = (int)i;
		++hp->refs;
		if (best >= 0) {
			void *p = hp->blocks[best].first;
			*got = hp->blocks[best].second;
			hp->blocks.erase(hp->blocks.begin() + best);
			return p;
		}
	}
	void *p = 0;
	bytes += bytes >> 2;
	if (cudaMallocHost(&p, bytes) != cudaSuccess) { std::lock_guard<std::mutex> g(hp->mu); --hp->refs; return 0; }
	*got = bytes;
	return p;
}

static void pool_release(HostPool *hp, void *p, uint64_t bytes)
{
	bool del = false;
	{
		std::lock_guard<std::mutex> g(hp->mu);
		if (hp->dead) cudaFreeHost(p);
		else hp->blocks.push_back(std::make_pair(p, bytes));
		del = --hp->refs == 0;
	}
	if (del) delete hp;
}

struct mmg_aligner {
	/* A multi-device aligner (mmg_aligner_create_multi) is a group: one full single-device aligner per GPU in `subs`,
	 * the index replicated on each; mmg_map_batch on the group shards the reads by bases and gathers in read order. */
	std::vector<mmg_aligner*> subs;
	HostPool *pool;
	const mmg_index *idx;
	mmg_mapopt_t mo;
	int device, n_sms;
	cudaStream_t stream;
	DevIndex di;
	DevOpt dopt;
	std::vector<void*> dev_allocs;     /* index + arenas */
	std::vector<uint64_t> dev_alloc_bytes;
	uint64_t dev_bytes = 0;
	/* arena capacities */
	uint64_t cap_bases, cap_anchors, cap_regs, cap_keep_words;
	int dual_min;                      /* fewest reads of a chunk for which the two-stream split is used */
	int dual_stream;                   /* 1 = the two halves of a chunk run expand..re-chain on two streams (tuning knob "dual_stream") */
	cudaStream_t st2;
	cudaEvent_t ev_fork, ev_join;
	int ramp_shift;                    /* streamed mode: the first chunk is 1/2^ramp_shift of the arena (tuning knob "ramp_shift") */
	int anchor_filter;                 /* 1 = drop isolated anchors before the sort (seed.cu anchor_filter_kernel) */
	uint32_t cap_reads;
	bool arenas_ready, caps_auto;
	ChunkDev cd;                       /* arena pointers */
	unsigned char *rmq_nodes;          /* AVL node arena of the re-chain stage */
	uint32_t *d_order;                 /* longest-first work order of the chunk's reads */
	ExtBufs xb;                        /* extension stage arenas (allocated when MM_F_CIGAR is set) */
	uint64_t *cg_read_off;
	uint64_t cap_tb, cap_cg, cap_jobs, big_per_warp;
	int profile;
	double stage_ms[MMG_N_STAGES];
	uint64_t stage_launches[MMG_N_STAGES];
	cudaEvent_t ev0, ev1, ev_run0, ev_run1;
	double last_run_ms;
	/* per-stage timing without host synchronisation: event pairs are recorded on the compute stream
	 * and read back when the run has finished */
	std::vector<cudaEvent_t> ev_pool;
	std::vector<int> ev_stage;
	size_t ev_used;
	/* streamed mode (mmg_map_batch): chunk k+1 is copied in on s_in while chunk k is computed on
	 * `stream`, and the results of finished sub-ranges are copied out on s_out */
	cudaStream_t s_in, s_out;
	bool stream_ready;
	char *in_bases[2]; uint64_t *in_off[2]; uint64_t *h_in_off[2];
	cudaEvent_t ev_in[2], ev_free[2];
	struct ResSlot {
		mmg_hit_t *d_hits, *h_hits; uint64_t hits_cap;
		uint32_t *d_cigar, *h_cigar; uint64_t cigar_cap;
		uint32_t *d_nregs, *h_nregs; uint64_t nregs_cap;
		cudaEvent_t ev_packed, ev_out;
		bool pending; uint64_t n_hits, n_cigar, hit_base, cigar_base; uint32_t read0, n_reads;
	} rs[2];
	unsigned long long *d_stats_pool;
	unsigned long long *h_ctl;         /* pinned control words the compute stream copies counters into (16) */
	uint64_t n_sub;                    /* sub-ranges issued by the current streamed call */
};

struct mmg_batch {
	uint32_t n_reads;
	uint64_t n_bases;
	const char *h_bases;               /* caller's buffers (valid until upload returns) */
	const uint64_t *h_off;
	char *d_bases; uint64_t *d_off;    /* device-resident inputs */
	mmg_hit_t *d_hits; uint64_t hits_cap, n_hits_dev;
	uint32_t *d_nregs;                 /* per read */
	unsigned long long *d_stats;
	uint32_t *d_cigar; uint64_t cigar_cap, n_cigar_dev;
	std::vector<uint64_t> off;         /* host copy of offsets */
	std::vector<uint64_t> hit_off;
	std::vector<mmg_hit_t> hits;
	std::vector<uint32_t> cigar;
	uint64_t stats[MMG_N_STATS];
	bool uploaded, ran, fetched;
	bool streamed;                     /* inputs/results go through the aligner's streaming slots */
	HostPool *pool; mmg_hit_t *ph; uint64_t ph_bytes; /* streamed: the hit records live in a pinned pool block */
	uint32_t *pc; uint64_t pc_bytes;                  /* streamed, CIGAR mode: so do the CIGAR operations */
	/* debug: arenas of the LAST chunk stay valid until the next run */
	uint32_t dbg_r0, dbg_r1;
};

template<typename T> static int dev_alloc(mmg_aligner *al, T **p, uint64_t n)
{
	void *q = 0;
	if (n == 0) n = 1;
	if (const char *lim = getenv("MMG_ALLOC_LIMIT")) { /* test hook: a device with this many bytes left for this aligner */
		if (al->dev_bytes + n * sizeof(T) > strtoull(lim, 0, 10)) { mmg_set_error("cudaMalloc of %llu bytes refused (MMG_ALLOC_LIMIT)", (unsigned long long)(n * sizeof(T))); return MMG_ENOMEM; }
	}
	cudaError_t e = cudaMalloc(&q, n * sizeof(T));
	if (e != cudaSuccess) { mmg_set_error("cudaMalloc of %llu bytes failed: %s", (unsigned long long)(n * sizeof(T)), cudaGetErrorString(e)); return MMG_ENOMEM; }
	al->dev_allocs.push_back(q), al->dev_alloc_bytes.push_back(n * sizeof(T)), al->dev_bytes += n * sizeof(T);
	*p = (T*)q;
	if (getenv("MMG_POISON")) cudaMemset(q, 0xCD, n * sizeof(T)), cudaDeviceSynchronize(); /* test hook: nothing may depend on what cudaMalloc returns */
	return MMG_OK;
}

/* `src`: a device copy of the same index on GPU `src_dev` (another aligner's, or the device the index was built on):
 * replicated over NVLink with peer copies instead of a second trip through the host */
static int upload_index(mmg_aligner *al, const DevIndex *src = 0, int src_dev = -1)
{
	const mmg_index *idx = al->idx;
	DevIndex &di = al->di;
	di.k = idx->k, di.w = idx->w, di.b = idx->b, di.flag = idx->flag, di.n_seq = idx->n_seq, di.hbits = idx->hbits;
	if (idx->dev_device == al->device && idx->dev_htab) { /* built on this device (index_dev.cu): used in place */
		di.htab = (const mmg_u128*)idx->dev_htab, di.pos = idx->dev_pos, di.S = idx->dev_S, di.seq_off = idx->dev_seq_off, di.seq_len = idx->dev_seq_len;
		return MMG_OK;
	}
	DevIndex from;
	if (!src && idx->dev_htab && idx->dev_device >= 0) {
		from.htab = (const mmg_u128*)idx->dev_htab, from.pos = idx->dev_pos, from.S = idx->dev_S, from.seq_off = idx->dev_seq_off, from.seq_len = idx->dev_seq_len;
		src = &from, src_dev = idx->dev_device;
	}
	if (src && src_dev >= 0 && src_dev != al->device) {
		const size_t nslots = (size_t)1 << idx->hbits, n_pos = idx->dev_htab ? (size_t)idx->n_pos : idx->pos.size(), n_S = idx->S.size();
		mmg_u128 *d_tab; uint64_t *d_pos, *d_soff; uint32_t *d_S, *d_slen;
		int rc;
		if ((rc = dev_alloc(al, &d_tab, nslots)) || (rc = dev_alloc(al, &d_pos, n_pos)) || (rc = dev_alloc(al, &d_S, n_S)) ||
		    (rc = dev_alloc(al, &d_soff, idx->offs.size())) || (rc = dev_alloc(al, &d_slen, idx->lens.size()))) return rc;
		/* peer access is switched on for the copies only (NVLink DMA instead of staging through the host) and off again:
		 * nothing on the mapping path touches another device's memory */
		int can = 0, enabled = 0;
		if (cudaDeviceCanAccessPeer(&can, al->device, src_dev) == cudaSuccess && can) {
			enabled = cudaDeviceEnablePeerAccess(src_dev, 0) == cudaSuccess;
			cudaGetLastError();
		}
		CK(cudaMemcpyPeer(d_tab, al->device, src->htab, src_dev, nslots * sizeof(mmg_u128)));
		if (n_pos) CK(cudaMemcpyPeer(d_pos, al->device, src->pos, src_dev, n_pos * 8));
		if (n_S) CK(cudaMemcpyPeer(d_S, al->device, src->S, src_dev, n_S * 4));
		CK(cudaDeviceSynchronize());   /* peer copies are asynchronous to the host, and the aligner's streams do not wait for the null stream */
		if (enabled) { cudaDeviceDisablePeerAccess(src_dev); cudaGetLastError(); }
		CK(cudaMemcpy(d_soff, idx->offs.data(), idx->offs.size() * 8, cudaMemcpyHostToDevice));
		if (!idx->lens.empty()) CK(cudaMemcpy(d_slen, idx->lens.data(), idx->lens.size() * 4, cudaMemcpyHostToDevice));
		di.htab = d_tab, di.pos = d_pos, di.S = d_S, di.seq_off = d_soff, di.seq_len = d_slen;
		return MMG_OK;
	}
	{ int rc0 = mmg_index_ensure_host(const_cast<mmg_index*>(idx)); if (rc0) return rc0; }
	size_t nslots = idx->hkeys.size();
	std::vector<mmg_u128> tab(nslots);
	for (size_t i = 0; i < nslots; ++i) tab[i].x = idx->hkeys[i], tab[i].y = idx->hvals[i];
	mmg_u128 *d_tab; uint64_t *d_pos, *d_soff; uint32_t *d_S, *d_slen;
	int rc;
	if ((rc = dev_alloc(al, &d_tab, nslots))) return rc;
	if ((rc = dev_alloc(al, &d_pos, idx->pos.size()))) return rc;
	if ((rc = dev_alloc(al, &d_S, idx->S.size()))) return rc;
	if ((rc = dev_alloc(al, &d_soff, idx->offs.size()))) return rc;
	if ((rc = dev_alloc(al, &d_slen, idx->lens.size()))) return rc;
	CK(cudaMemcpy(d_tab, tab.data(), nslots * sizeof(mmg_u128), cudaMemcpyHostToDevice));
	if (!idx->pos.empty()) CK(cudaMemcpy(d_pos, idx->pos.data(), idx->pos.size() * 8, cudaMemcpyHostToDevice));
	if (!idx->S.empty()) CK(cudaMemcpy(d_S, idx->S.data(), idx->S.size() * 4, cudaMemcpyHostToDevice));
	CK(cudaMemcpy(d_soff, idx->offs.data(), idx->offs.size() * 8, cudaMemcpyHostToDevice));
	if (!idx->lens.empty()) CK(cudaMemcpy(d_slen, idx->lens.data(), idx->lens.size() * 4, cudaMemcpyHostToDevice));
	di.htab = d_tab, di.pos = d_pos, di.S = d_S, di.seq_off = d_soff, di.seq_len = d_slen;
	return MMG_OK;
}

static void fill_devopt(mmg_aligner *al)
{
	const mmg_mapopt_t &m = al->mo;
	DevOpt &o = al->dopt;
	memset(&o, 0, sizeof(o));
	o.flag = m.flag, o.seed = m.seed;
	o.bw = m.bw, o.bw_long = m.bw_long, o.max_gap = m.max_gap, o.max_gap_ref = m.max_gap_ref, o.max_frag_len = m.max_frag_len;
	o.max_chain_skip = m.max_chain_skip, o.max_chain_iter = m.max_chain_iter, o.min_cnt = m.min_cnt, o.min_chain_score = m.min_chain_score;
	/* map.c mm_map_frag: float = float * double * int, evaluated in double */
	o.chn_pen_gap = (float)(m.chain_gap_scale * 0.01 * al->idx->k);
	o.chn_pen_skip = (float)(m.chain_skip_scale * 0.01 * al->idx->k);
	o.rmq_size_cap = m.rmq_size_cap, o.rmq_inner_dist = m.rmq_inner_dist, o.rmq_rescue_size = m.rmq_rescue_size, o.rmq_rescue_ratio = m.rmq_rescue_ratio;
	o.mask_level = m.mask_level, o.mask_len = m.mask_len, o.pri_ratio = m.pri_ratio, o.best_n = m.best_n, o.alt_drop = m.alt_drop;
	o.a = m.a, o.b = m.b, o.q = m.q, o.e = m.e, o.q2 = m.q2, o.e2 = m.e2, o.sc_ambi = m.sc_ambi;
	o.zdrop = m.zdrop, o.zdrop_inv = m.zdrop_inv, o.end_bonus = m.end_bonus, o.min_dp_max = m.min_dp_max, o.min_ksw_len = m.min_ksw_len;
	o.anchor_ext_len = m.anchor_ext_len, o.anchor_ext_shift = m.anchor_ext_shift, o.max_clip_ratio = m.max_clip_ratio;
	o.q_occ_frac = m.q_occ_frac, o.mid_occ = m.mid_occ, o.max_max_occ = m.max_max_occ, o.occ_dist = m.occ_dist, o.max_qlen = m.max_qlen;
	o.max_sw_mat = m.max_sw_mat;
}

static int alloc_arenas_once(mmg_aligner *al)
{
	ChunkDev &c = al->cd;
	const uint64_t B = al->cap_bases, A = al->cap_anchors, R = al->cap_reads, G = al->cap_regs;
	int rc = 0;
#define AL(ptr, n) if ((rc = dev_alloc(al, &(ptr), (n)))) return rc
	AL(c.mz_x, B); AL(c.mz_y, B); AL(c.n_mz, R);
	AL(c.sd_val, B); AL(c.sd_n, B); AL(c.sd_qpos, B); AL(c.sd_meta, B);
	AL(c.n_seed, R); AL(c.n_a, R); AL(c.rep_len, R); AL(c.a_off, R + 1);
	AL(c.ax, A); AL(c.ay, A); AL(c.bx, A); AL(c.by, A);
	AL(c.f, A); AL(c.p, A); AL(c.t, A); AL(c.v, A);
	AL(c.zx, 2 * A); AL(c.zy, 2 * A);
	AL(c.cx, A); AL(c.cy, A); AL(c.u, A);
	AL(c.n_u, R); AL(c.n_v, R); AL(c.r_off, R + 1);
	AL(c.regs, G); AL(c.n_regs, R); AL(c.h_off, R + 1);
	AL(c.work, 64); AL(c.err, 4); AL(c.flags, R); AL(c.big_list, R); AL(c.tie_list, R);
	AL(al->d_order, R); AL(c.af_off, R + 1); AL(c.keep_bits, al->cap_keep_words); AL(c.hit_scratch, al->cap_keep_words * 32);
	AL(al->rmq_nodes, (2 * A + 2 * R + 2) * RMQ_NODE_BYTES);
	if (al->mo.flag & MMG_F_CIGAR) {
		ExtBufs &x = al->xb;
		x.cap_jobs = al->cap_jobs, x.cap_tb = al->cap_tb, x.cap_cg = al->cap_cg, x.big_per_warp = al->big_per_warp;
		AL(x.jobs, x.cap_jobs); AL(x.n_jobs, 4); AL(x.xregs, G); AL(x.regs_tmp, G); AL(x.n_sq, R);
		AL(x.tb, x.cap_tb + 64); AL(x.jcigar, x.cap_cg); AL(x.rcigar, x.cap_cg);
		AL(x.tb_base, 2); AL(x.cg_base, 2); AL(x.n_pending, 4); AL(x.reg_cap, R); AL(x.ovf, x.cap_jobs + 16); AL(x.ovf_n, 4);
		AL(x.big, (uint64_t)al->n_sms * 32 * x.big_per_warp); /* one slice per resident warp of the prep (8x4 per SM) and DP (4x4 per SM) grids */
		AL(al->cg_read_off, R + 1);
		x.xr_off = c.r_off;
	}
#undef AL
	al->arenas_ready = true;
	return MMG_OK;
}

/* The default arena sizes assume the aligner has the device to itself (a B200: 121 GB with CIGAR).  When the memory is
 * not there - another aligner on the same GPU, a smaller device - the allocation is rolled back and repeated one size
 * down: chunks of 192, 96, 48 Mbases, then with a smaller traceback arena.  Sizes set through mmg_aligner_set are the
 * caller's: they are tried once. */
static int alloc_arenas(mmg_aligner *al)
{
	if (al->arenas_ready) return MMG_OK;
	for (;;) {
		const size_t mark = al->dev_allocs.size();
		int rc = alloc_arenas_once(al);
#ifndef MMG_EMU
		if (rc == MMG_OK && al->caps_auto) { /* what is still to come: two input slots of a chunk each, the result slots, per-batch buffers */
			size_t free_b = 0, total_b = 0;
			const uint64_t reserve = 2 * al->cap_bases + ((uint64_t)6 << 30);
			if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b < reserve) {
				mmg_set_error("only %llu MB of device memory would be left after the arenas", (unsigned long long)(free_b >> 20));
				al->arenas_ready = false, rc = MMG_ENOMEM;
			}
		}
#endif
		if (rc == MMG_OK) return MMG_OK;
		while (al->dev_allocs.size() > mark) { cudaFree(al->dev_allocs.back()); al->dev_bytes -= al->dev_alloc_bytes.back(); al->dev_allocs.pop_back(), al->dev_alloc_bytes.pop_back(); }
		cudaGetLastError();
		if (rc != MMG_ENOMEM || !al->caps_auto) return rc;
		if (al->cap_bases > ((uint64_t)48 << 20)) {
			al->cap_bases >>= 1, al->cap_anchors >>= 1;
			if (al->cap_reads > (1u << 17)) al->cap_reads >>= 1;
			if (al->cap_regs > ((uint64_t)4 << 20)) al->cap_regs >>= 1;
			if (al->cap_keep_words > ((uint64_t)1 << 22)) al->cap_keep_words >>= 1;
			al->cap_cg = (uint64_t)3 * al->cap_bases, al->cap_jobs = al->cap_bases / 48;
		} else if (al->cap_tb > ((uint64_t)4 << 30)) al->cap_tb >>= 1;
		else return rc;
	}
}

/* counting sort of the chunk's reads by length / 64, longest first (one CTA; ties in any order) */
__global__ void __launch_bounds__(1024)
read_order_kernel(const uint64_t *off, uint32_t n, uint32_t *order)
{
	__shared__ uint32_t s_bin[4096];
	const uint32_t tid = threadIdx.x;
	for (uint32_t k = tid; k < 4096; k += 1024) s_bin[k] = 0;
	__syncthreads();
	for (uint32_t i = tid; i < n; i += 1024) {
		const uint64_t l = (off[i + 1] - off[i]) >> 6;
		atomicAdd(&s_bin[4095 - (l > 4095 ? 4095u : (uint32_t)l)], 1u);
	}
	__syncthreads();
	if (tid < 32) { /* exclusive prefix over the 4096 bins: 128 consecutive bins per lane */
		uint32_t sum = 0;
		for (uint32_t k = 0; k < 128; ++k) sum += s_bin[tid * 128 + k];
		uint32_t x = sum;
		for (int d = 1; d < 32; d <<= 1) {
			uint32_t y = __shfl_up_sync(MMG_FULL, x, d);
			if ((int)tid >= d) x += y;
		}
		uint32_t run = x - sum;
		for (uint32_t k = 0; k < 128; ++k) { const uint32_t v = s_bin[tid * 128 + k]; s_bin[tid * 128 + k] = run; run += v; }
	}
	__syncthreads();
	for (uint32_t i = tid; i < n; i += 1024) {
		const uint64_t l = (off[i + 1] - off[i]) >> 6;
		const uint32_t p = atomicAdd(&s_bin[4095 - (l > 4095 ? 4095u : (uint32_t)l)], 1u); /* rank, longest first */
		order[(p & 1u) ? (n + 1) / 2 + p / 2 : p / 2] = i;   /* even ranks, then odd ranks: two balanced halves, each longest first */
	}
}

__global__ void reg_cap_kernel(const uint32_t *n_u, uint32_t *cap, uint32_t n)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) cap[i] = n_u[i] ? 2 * n_u[i] + 4 : 0;
}

static int aligner_create_on(const mmg_index *idx, const mmg_mapopt_t *mo, int device, const DevIndex *src, int src_dev, mmg_aligner **out)
{
	*out = 0;
	int n_dev = 0;
	if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) {
		mmg_set_error("no CUDA device: libmmg has no CPU mapping path");
		return MMG_ENODEV;
	}
	if (device < 0 || device >= n_dev) { mmg_set_error("device %d out of range (%d devices)", device, n_dev); return MMG_EINVAL; }
	if ((mo->flag & MMG_F_CIGAR) && (idx->flag & MMG_I_NO_SEQ)) { mmg_set_error("index has no sequence but CIGAR was requested"); return MMG_ENOSEQ; }
	/* MM_F_NO_DIAG (0x1) and MM_F_NO_DUAL (0x2) compare the query NAME with the target names; mappy-rs passes qname = NULL
	 * (crate minimap2 Aligner::map), so they have no effect on this path and are accepted (preset ava-ont sets them) */
	const int64_t unsup = 0x80LL | 0x100LL | 0x200LL | 0x1000LL | 0x2000LL | MMG_F_FOR_ONLY | MMG_F_REV_ONLY | 0x400000LL | 0x20000000LL | 0x100000000LL;
	if (mo->flag & unsup) { mmg_set_error("mapping flag 0x%llx selects a code path outside the supported long-read path", (unsigned long long)(mo->flag & unsup)); return MMG_EUNSUP; }
	if (idx->flag & MMG_I_HPC) { mmg_set_error("homopolymer-compressed indexes are not supported"); return MMG_EUNSUP; }
	if (mo->flag & MMG_F_CIGAR) {
		/* the limits inside which ksw_extd2_sse / ksw_extz2_sse compute what they are meant to: their 8-bit lanes wrap beyond
		 * (q+e)+(q2+e2) <= 127 (minimap2's mm_check_opt states this one; mappy-rs does not call it, src/lib.rs:339-385),
		 * and both kernels return without a result when a substitution costs more than 2*(q+e).  Outside them upstream's
		 * output is an artefact of the wrap-around, so such scorings are refused instead of reproduced. */
		if (mo->transition != 0 && mo->transition != mo->b) { mmg_set_error("a separate transition score is not supported (mappy-rs cannot set one)"); return MMG_EUNSUP; }
		const int qe = mo->q + mo->e, qe2 = mo->q2 + mo->e2, worst = std::max(std::max(mo->b, mo->sc_ambi), mo->transition);
		if (mo->a < 1 || mo->b < 0 || mo->q < 0 || mo->e < 1 || mo->q2 < 0 || mo->e2 < 1 || mo->sc_ambi < 0 || mo->transition < 0) {
			mmg_set_error("scoring: a, e, e2 must be positive and b, q, q2, sc_ambi, transition non-negative"); return MMG_EINVAL; }
		if (qe + qe2 > 127 || mo->a + std::max(qe, qe2) > 127) { mmg_set_error("scoring system violates (O1+E1)+(O2+E2) <= 127 and A+max(O+E) <= 127: minimap2's 8-bit DP lanes wrap beyond it"); return MMG_EINVAL; }
		if (worst > 2 * qe) { mmg_set_error("scoring: a mismatch / ambiguous-base penalty above 2*(O1+E1) makes minimap2's DP kernels return no alignment"); return MMG_EINVAL; }
	}
	if (idx->offs.back() >= ((uint64_t)1 << 35)) { mmg_set_error("references of 2^35 bases or more are not supported"); return MMG_EUNSUP; }
	CK(cudaSetDevice(device));
#define CKA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { mmg_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); mmg_aligner_destroy(al); return MMG_ECUDA; } } while (0)
	mmg_aligner *al = new mmg_aligner();
	al->idx = idx, al->mo = *mo, al->device = device;
	cudaDeviceProp prop;
	CKA(cudaGetDeviceProperties(&prop, device));
	al->n_sms = prop.multiProcessorCount;
	CKA(cudaStreamCreateWithFlags(&al->stream, cudaStreamNonBlocking));
	CKA(cudaEventCreate(&al->ev0));
	CKA(cudaEventCreate(&al->ev1));
	CKA(cudaEventCreate(&al->ev_run0));
	CKA(cudaEventCreate(&al->ev_run1));
	al->last_run_ms = 0;
	al->pool = new HostPool();
	al->h_ctl = 0;
	CKA(cudaMallocHost((void**)&al->h_ctl, 16 * sizeof(unsigned long long)));
	memset(al->h_ctl, 0, 16 * sizeof(unsigned long long));
	al->ev_used = 0, al->s_in = 0, al->s_out = 0, al->stream_ready = false, al->d_stats_pool = 0, al->n_sub = 0;
	memset(al->in_bases, 0, sizeof(al->in_bases)), memset(al->in_off, 0, sizeof(al->in_off)), memset(al->h_in_off, 0, sizeof(al->h_in_off));
	memset(al->ev_in, 0, sizeof(al->ev_in)), memset(al->ev_free, 0, sizeof(al->ev_free)), memset(al->rs, 0, sizeof(al->rs));
	al->arenas_ready = false, al->caps_auto = true;
	al->profile = 0;
	al->cap_bases = (uint64_t)96 << 20, al->cap_reads = 1u << 17, al->cap_anchors = (uint64_t)48 << 20, al->cap_regs = (uint64_t)4 << 20;
	al->cap_keep_words = (uint64_t)1 << 22; /* unfiltered anchors per chunk the isolated-anchor filter can look at: 2^27, 2^29 on a 180 GB device */
	if (prop.totalGlobalMem >= ((size_t)120 << 30)) {
		/* B200 (180 GB): chunks of 384 Mbases / 256 M anchors (~60 GB of arenas; with CIGAR + 32 GB of traceback slices
		 * and 5 GB of CIGAR arena).  Every stage kernel ends with a tail of a few long-running reads (long reads,
		 * equal-key replays, the last batches of the DP kernels); fewer, larger chunks pay that tail fewer times per
		 * batch: configs[2] with CIGAR 225 k -> 246 k reads/s against 96-Mbase chunks. */
		al->cap_bases = (uint64_t)384 << 20, al->cap_reads = 1u << 19, al->cap_anchors = (uint64_t)256 << 20, al->cap_regs = (uint64_t)16 << 20;
		al->cap_keep_words = (uint64_t)1 << 24;
	}
	al->anchor_filter = 1;
	al->dual_stream = 0, al->dual_min = 4096; /* measured: 118.2 vs 119.7 ms per step on configs[2] - kept as an option, off by default */
	CKA(cudaStreamCreateWithFlags(&al->st2, cudaStreamNonBlocking));
	CKA(cudaEventCreateWithFlags(&al->ev_fork, cudaEventDisableTiming));
	CKA(cudaEventCreateWithFlags(&al->ev_join, cudaEventDisableTiming));
	al->ramp_shift = 1;
	al->cap_tb = (uint64_t)32 << 30, al->cap_cg = (uint64_t)3 * al->cap_bases, al->cap_jobs = al->cap_bases / 48, al->big_per_warp = (uint64_t)1 << 20;
	memset(&al->xb, 0, sizeof(al->xb));
	al->cg_read_off = 0;
	memset(al->stage_ms, 0, sizeof(al->stage_ms));
	memset(al->stage_launches, 0, sizeof(al->stage_launches));
	memset(&al->cd, 0, sizeof(al->cd));
	int rc = upload_index(al, src, src_dev);
	if (rc) { mmg_aligner_destroy(al); return rc; }
	fill_devopt(al);
	*out = al;
	return MMG_OK;
}

extern "C" {

int mmg_aligner_create(const mmg_index *idx, const mmg_mapopt_t *mo, int device, mmg_aligner **out)
{
	return aligner_create_on(idx, mo, device, 0, -1, out);
}

int mmg_aligner_create_multi(const mmg_index *idx, const mmg_mapopt_t *mo, const int *devices, int n_dev, mmg_aligner **out)
{
	*out = 0;
	if (n_dev < 1 || !devices) { mmg_set_error("mmg_aligner_create_multi needs at least one device"); return MMG_EINVAL; }
	for (int a = 0; a < n_dev; ++a)
		for (int b = 0; b < a; ++b)
			if (devices[a] == devices[b]) { mmg_set_error("device %d is listed twice", devices[a]); return MMG_EINVAL; }
	mmg_aligner *g = new mmg_aligner();
	g->idx = idx, g->mo = *mo, g->device = devices[0];
	/* the first member uploads the index (or uses it in place where it was built); the others copy it from a peer */
	int first = 0;
	for (int a = 0; a < n_dev; ++a) if (devices[a] == idx->dev_device) first = a;
	g->subs.assign(n_dev, (mmg_aligner*)0);
	debug_check("create_multi: before", devices[first]);
	int rc = aligner_create_on(idx, mo, devices[first], 0, -1, &g->subs[first]);
	debug_check("create_multi: first member created", devices[first]);
	for (int a = 0; a < n_dev && !rc; ++a)
		if (a != first) {
			rc = aligner_create_on(idx, mo, devices[a], &g->subs[first]->di, devices[first], &g->subs[a]);
			debug_check("create_multi: member created (its device)", devices[a]);
			debug_check("create_multi: member created (source device)", devices[first]);
		}
	if (rc) { mmg_aligner_destroy(g); return rc; }
	*out = g;
	return MMG_OK;
}

void mmg_aligner_destroy(mmg_aligner *al)
{
	if (!al) return;
	mmg_stream_shutdown(al);   /* stops the streaming worker (mmg_submit / mmg_next), if any */
	if (al->subs.empty()) debug_check("aligner_destroy: entry", al->device);
	if (!al->subs.empty()) {
		for (size_t a = 0; a < al->subs.size(); ++a) mmg_aligner_destroy(al->subs[a]);
		delete al;
		return;
	}
	cudaSetDevice(al->device);
	for (size_t i = 0; i < al->dev_allocs.size(); ++i) cudaFree(al->dev_allocs[i]);
	if (al->stream) cudaStreamDestroy(al->stream);
	if (al->st2) cudaStreamDestroy(al->st2);
	if (al->ev_fork) cudaEventDestroy(al->ev_fork);
	if (al->ev_join) cudaEventDestroy(al->ev_join);
	if (al->s_in) cudaStreamDestroy(al->s_in);
	if (al->s_out) cudaStreamDestroy(al->s_out);
	for (int k = 0; k < 2; ++k) {
		if (al->h_in_off[k]) cudaFreeHost(al->h_in_off[k]);
		if (al->ev_in[k]) cudaEventDestroy(al->ev_in[k]);
		if (al->ev_free[k]) cudaEventDestroy(al->ev_free[k]);
		mmg_aligner::ResSlot &r = al->rs[k];
		if (r.ev_packed) cudaEventDestroy(r.ev_packed);
		if (r.ev_out) cudaEventDestroy(r.ev_out);
		if (r.d_hits) cudaFree(r.d_hits);
		if (r.d_cigar) cudaFree(r.d_cigar);
		if (r.d_nregs) cudaFree(r.d_nregs);
		if (r.h_hits) cudaFreeHost(r.h_hits);
		if (r.h_cigar) cudaFreeHost(r.h_cigar);
		if (r.h_nregs) cudaFreeHost(r.h_nregs);
	}
	for (size_t i = 0; i < al->ev_pool.size(); ++i) cudaEventDestroy(al->ev_pool[i]);
	if (al->pool) {
		bool del = false;
		{
			std::lock_guard<std::mutex> g(al->pool->mu);
			for (size_t i = 0; i < al->pool->blocks.size(); ++i) cudaFreeHost(al->pool->blocks[i].first);
			al->pool->blocks.clear();
			al->pool->dead = true;
			del = --al->pool->refs == 0;
		}
		if (del) delete al->pool;
	}
	if (al->h_ctl) cudaFreeHost(al->h_ctl);
	if (al->ev0) cudaEventDestroy(al->ev0);
	if (al->ev1) cudaEventDestroy(al->ev1);
	if (al->ev_run0) cudaEventDestroy(al->ev_run0);
	if (al->ev_run1) cudaEventDestroy(al->ev_run1);
	delete al;
}

int mmg_host_alloc(size_t bytes, void **out)
{
	*out = 0;
	cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
	if (e != cudaSuccess) { mmg_set_error("cudaMallocHost of %llu bytes failed: %s", (unsigned long long)bytes, cudaGetErrorString(e)); return MMG_ENOMEM; }
	return MMG_OK;
}
void mmg_host_free(void *p) { if (p) cudaFreeHost(p); }

int mmg_aligner_set(mmg_aligner *al, const char *key, int64_t v)
{
	if (!al->subs.empty()) {
		for (size_t a = 0; a < al->subs.size(); ++a) { int rc = mmg_aligner_set(al->subs[a], key, v); if (rc) return rc; }
		return MMG_OK;
	}
	if (strcmp(key, "profile") == 0) { al->profile = (int)v; return MMG_OK; }
	if (strcmp(key, "sort_small_max") == 0) { mmg_sort_set_small_max((int)v); return MMG_OK; }
	if (strcmp(key, "anchor_filter") == 0) { al->anchor_filter = v != 0; return MMG_OK; }
	if (strcmp(key, "dual_stream") == 0) { al->dual_stream = v != 0; return MMG_OK; }
	if (strcmp(key, "dual_min") == 0) { al->dual_min = v < 2 ? 2 : (int)v; return MMG_OK; }
	if (strcmp(key, "ramp_shift") == 0) { al->ramp_shift = v < 0 ? 0 : v > 6 ? 6 : (int)v; return MMG_OK; }
	if (al->arenas_ready) { mmg_set_error("arena sizes are fixed after the first batch"); return MMG_EINVAL; }
	al->caps_auto = false;   /* the caller sizes the arenas: no stepping down on its behalf */
	if (strcmp(key, "chunk_bases") == 0) al->cap_bases = (uint64_t)v;
	else if (strcmp(key, "chunk_reads") == 0) al->cap_reads = (uint32_t)v;
	else if (strcmp(key, "anchor_cap") == 0) al->cap_anchors = (uint64_t)v;
	else if (strcmp(key, "regs_cap") == 0) al->cap_regs = (uint64_t)v;
	else if (strcmp(key, "keep_words") == 0) al->cap_keep_words = (uint64_t)v;
	else if (strcmp(key, "tb_cap") == 0) al->cap_tb = (uint64_t)v;
	else if (strcmp(key, "cigar_cap") == 0) al->cap_cg = (uint64_t)v;
	else if (strcmp(key, "jobs_cap") == 0) al->cap_jobs = (uint64_t)v;
	else if (strcmp(key, "big_per_warp") == 0) al->big_per_warp = (uint64_t)v;
	else { mmg_set_error("unknown key '%s'", key); return MMG_EINVAL; }
	return MMG_OK;
}

int mmg_batch_upload(mmg_aligner *al, const char *bases, const uint64_t *offsets, uint32_t n_reads, mmg_batch **out)
{
	*out = 0;
	if (!al->subs.empty()) { mmg_set_error("upload / run / fetch time one device: use mmg_map_batch on a multi-device aligner"); return MMG_EUNSUP; }
	CK(cudaSetDevice(al->device));
	mmg_batch *b = new mmg_batch();
	b->n_reads = n_reads, b->h_bases = bases, b->h_off = offsets;
	b->n_bases = n_reads ? offsets[n_reads] - offsets[0] : 0;
	b->off.assign(offsets, offsets + n_reads + 1);
	b->d_bases = 0, b->d_off = 0, b->d_hits = 0, b->d_nregs = 0, b->d_stats = 0, b->d_cigar = 0, b->cigar_cap = 0, b->n_cigar_dev = 0;
	b->uploaded = b->ran = b->fetched = false;
	b->streamed = false;
	b->pool = 0, b->ph = 0, b->ph_bytes = 0, b->pc = 0, b->pc_bytes = 0;
	memset(b->stats, 0, sizeof(b->stats));
	for (uint32_t i = 0; i < n_reads; ++i) {
		uint64_t l = offsets[i + 1] - offsets[i];
		if (l > 0x7fffffffULL || l > al->cap_bases) { mmg_set_error("read %u is longer than the chunk capacity (%llu bases)", i, (unsigned long long)al->cap_bases); delete b; return MMG_EINVAL; }
	}
	if (cudaMalloc((void**)&b->d_bases, b->n_bases + 64) != cudaSuccess || cudaMalloc((void**)&b->d_off, (size_t)(n_reads + 1) * 8) != cudaSuccess ||
	    cudaMalloc((void**)&b->d_nregs, (size_t)(n_reads + 1) * 4) != cudaSuccess || cudaMalloc((void**)&b->d_stats, MMG_N_STATS * 8) != cudaSuccess) {
		mmg_set_error("cudaMalloc failed for the batch inputs");
		mmg_batch_destroy(b);
		return MMG_ENOMEM;
	}
	b->hits_cap = (uint64_t)n_reads * (uint64_t)(al->mo.best_n + 4 > 6 ? al->mo.best_n + 4 : 6) + 1024;   /* primaries + best_n secondaries; the streamed path grows on demand */
	if (cudaMalloc((void**)&b->d_hits, b->hits_cap * sizeof(mmg_hit_t)) != cudaSuccess) { mmg_set_error("cudaMalloc failed for the result pool"); mmg_batch_destroy(b); return MMG_ENOMEM; }
	if (al->mo.flag & MMG_F_CIGAR) {
		b->cigar_cap = b->n_bases / 2 + ((uint64_t)1 << 16) + (uint64_t)n_reads * 8;
		if (cudaMalloc((void**)&b->d_cigar, b->cigar_cap * 4) != cudaSuccess) { mmg_set_error("cudaMalloc failed for the CIGAR pool"); mmg_batch_destroy(b); return MMG_ENOMEM; }
	}
	if (al->profile) cudaEventRecord(al->ev0, al->stream);
	/* offsets are rebased so that the device buffer starts at 0 */
	std::vector<uint64_t> rel(n_reads + 1);
	for (uint32_t i = 0; i <= n_reads; ++i) rel[i] = offsets[i] - offsets[0];
	b->off = rel;
	CK(cudaMemcpyAsync(b->d_bases, bases + offsets[0], b->n_bases, cudaMemcpyHostToDevice, al->stream));
	CK(cudaMemcpyAsync(b->d_off, rel.data(), (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, al->stream));
	CK(cudaMemsetAsync(b->d_stats, 0, MMG_N_STATS * 8, al->stream));
	if (al->profile) cudaEventRecord(al->ev1, al->stream);
	CK(cudaStreamSynchronize(al->stream));
	if (al->profile) { float ms = 0; cudaEventElapsedTime(&ms, al->ev0, al->ev1); al->stage_ms[ST_H2D] += ms; }
	b->uploaded = true;
	*out = b;
	return MMG_OK;
}

static cudaEvent_t stage_event(mmg_aligner *al, int stage)
{
	if (al->ev_used == al->ev_pool.size()) {
		cudaEvent_t e = 0;
		cudaEventCreate(&e);
		al->ev_pool.push_back(e), al->ev_stage.push_back(stage);
	}
	al->ev_stage[al->ev_used] = stage;
	return al->ev_pool[al->ev_used++];
}
/* called after the compute stream has been synchronised */
static void stage_collect(mmg_aligner *al)
{
	for (size_t i = 0; i + 1 < al->ev_used; i += 2) {
		float ms = 0;
		if (cudaEventElapsedTime(&ms, al->ev_pool[i], al->ev_pool[i + 1]) == cudaSuccess) al->stage_ms[al->ev_stage[i + 1]] += ms;
	}
	al->ev_used = 0;
}
#define STAGE_BEGIN() do { if (al->profile) cudaEventRecord(stage_event(al, -1), st); } while (0)
#define STAGE_END(id) do { al->stage_launches[id] += 1; if (al->profile) cudaEventRecord(stage_event(al, id), st); \
	if (debug_sync()) { cudaError_t e_ = cudaStreamSynchronize(st); if (e_ == cudaSuccess) e_ = cudaGetLastError(); \
		if (e_ != cudaSuccess) { mmg_set_error("stage %s on device %d failed: %s (%s:%d)", g_stage_names[id], al->device, cudaGetErrorString(e_), __FILE__, __LINE__); return MMG_ECUDA; } } } while (0)

/* Base-level alignment of the regions of reads [s0, s1): rounds of prep -> DP jobs -> stitch until no
 * region is left pending (a z-drop split creates a region that is aligned in the next round).  Within a round the
 * kernels read the job count from device memory, so the host synchronises once per round (to learn whether another
 * round is needed and whether an arena overflowed), not between the kernels. */
static int run_extension(mmg_aligner *al, mmg_batch *b, ChunkDev &c, uint32_t s0, uint32_t s1, uint32_t *work, int *wi_, uint64_t *n_cg_sub)
{
	cudaStream_t st = al->stream;
	ExtBufs &xb = al->xb;
	int wi = *wi_;
	uint32_t n_jobs_prev = 0;
	unsigned long long cg_end = 0;
	CK(cudaMemsetAsync(xb.n_jobs, 0, 4, st));
	int round = 0;
	for (; round < 64; ++round) {
		if (wi + EXT_DP_COUNTERS + 3 > 64) { CK(cudaMemsetAsync(c.work, 0, 64 * 4, st)); wi = 0; }
		STAGE_BEGIN();
		launch_ext_prep(c, al->di, al->dopt, xb, s0, s1, round, al->n_sms, st, work + wi++);
		if (debug_sync()) { cudaError_t e_ = cudaStreamSynchronize(st); if (e_ != cudaSuccess) { mmg_set_error("ext_prep_kernel (round %d) on device %d failed: %s", round, al->device, cudaGetErrorString(e_)); return MMG_ECUDA; } }
		al->h_ctl[0] = al->h_ctl[1] = cg_end;   /* pinned: the CIGAR slices of this round's jobs start where the last round ended; [1] becomes the new end */
		CK(cudaMemcpyAsync(xb.cg_base, al->h_ctl, 16, cudaMemcpyHostToDevice, st));
		launch_ext_job_scan(xb, n_jobs_prev, al->n_sms, st);
		if (debug_sync()) { cudaError_t e_ = cudaStreamSynchronize(st); if (e_ != cudaSuccess) { mmg_set_error("ext_job_scan_kernel (round %d) on device %d failed: %s", round, al->device, cudaGetErrorString(e_)); return MMG_ECUDA; } }
		CK(cudaMemsetAsync(xb.ovf_n, 0, 16, st));
		if (launch_ext_dp(c, al->di, al->dopt, xb, n_jobs_prev, al->n_sms, st, work + wi)) { mmg_set_error("a DP kernel faulted (round %d, device %d): see stderr", round, al->device); return MMG_ECUDA; }
		wi += EXT_DP_COUNTERS;
		CK(cudaMemsetAsync(xb.n_pending, 0, 4, st));
		launch_ext_stitch(c, al->di, al->dopt, xb, s0, s1, round, al->n_sms, st, work + wi++);
		CK(cudaMemcpyAsync(al->h_ctl + 1, xb.cg_base + 1, 8, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(al->h_ctl + 2, xb.n_jobs, 4, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(al->h_ctl + 3, xb.n_pending, 4, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(al->h_ctl + 4, c.err, 4, cudaMemcpyDeviceToHost, st));
		STAGE_END(ST_EXTEND);
		CK(cudaStreamSynchronize(st));
		cg_end = al->h_ctl[1];
		const uint32_t n_jobs = (uint32_t)al->h_ctl[2], n_pending = (uint32_t)al->h_ctl[3], err = (uint32_t)al->h_ctl[4];
		if (n_jobs > xb.cap_jobs || (err & 0x40000000u)) { mmg_set_error("extension job arena overflow (%u jobs > jobs_cap)", n_jobs); return MMG_ENOMEM; }
		if (cg_end > xb.cap_cg) { mmg_set_error("extension CIGAR arena overflow (%llu > cigar_cap)", cg_end); return MMG_ENOMEM; }
		if (err & 0x10000000u) { mmg_set_error("one alignment needs more traceback memory than tb_cap / 4 (%llu bytes)", (unsigned long long)(xb.cap_tb / 4)); return MMG_ENOMEM; }
		if (err & 0x20000000u) { mmg_set_error("one alignment is longer than the per-warp DP slice (big_per_warp)"); return MMG_ENOMEM; }
		n_jobs_prev = n_jobs;
		if (n_pending == 0) break;
	}
	if (round == 64) { mmg_set_error("internal: regions still pending after 64 alignment rounds"); return MMG_ECUDA; }
	STAGE_BEGIN();
	launch_ext_final(c, al->di, al->dopt, xb, s0, s1, al->n_sms, st, work + wi++);
	launch_scan_u32(xb.n_sq + s0, al->cg_read_off + s0, s1 - s0, st);
	STAGE_END(ST_EXTEND);
	CK(cudaMemcpyAsync(al->h_ctl + 5, al->cg_read_off + s1, 8, cudaMemcpyDeviceToHost, st)); /* read by run_chunk after its final synchronisation */
	(void)n_cg_sub;
	*wi_ = wi;
	return MMG_OK;
}

} // extern "C"

/* ---- streamed mode plumbing ------------------------------------------------------------------ */

static int stream_setup(mmg_aligner *al)
{
	if (al->stream_ready) return MMG_OK;
	CK(cudaStreamCreateWithFlags(&al->s_in, cudaStreamNonBlocking));
	CK(cudaStreamCreateWithFlags(&al->s_out, cudaStreamNonBlocking));
	for (int k = 0; k < 2; ++k) {
		int rc;
		if ((rc = dev_alloc(al, &al->in_bases[k], al->cap_bases + 64))) return rc;
		if ((rc = dev_alloc(al, &al->in_off[k], (uint64_t)al->cap_reads + 1))) return rc;
		CK(cudaMallocHost((void**)&al->h_in_off[k], ((size_t)al->cap_reads + 1) * 8));
		CK(cudaEventCreateWithFlags(&al->ev_in[k], cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&al->ev_free[k], cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&al->rs[k].ev_packed, cudaEventDisableTiming));
		CK(cudaEventCreateWithFlags(&al->rs[k].ev_out, cudaEventDisableTiming));
	}
	if (cudaMalloc((void**)&al->d_stats_pool, MMG_N_STATS * 8) != cudaSuccess) { mmg_set_error("cudaMalloc failed"); return MMG_ENOMEM; }
	al->dev_allocs.push_back(al->d_stats_pool), al->dev_alloc_bytes.push_back(0);
	al->stream_ready = true;
	return MMG_OK;
}

/* h = 0: device staging only (hits and CIGARs are copied out straight into the batch's pinned pool block) */
template<typename T> static int slot_grow(T **d, T **h, uint64_t *cap, uint64_t need)
{
	if (need <= *cap) return MMG_OK;
	uint64_t n = *cap ? *cap : 1024;
	while (n < need) n *= 2;
	if (*d) cudaFree(*d);
	if (h && *h) cudaFreeHost(*h);
	*d = 0, *cap = 0;
	if (h) *h = 0;
	if (cudaMalloc((void**)d, n * sizeof(T)) != cudaSuccess || (h && cudaMallocHost((void**)h, n * sizeof(T)) != cudaSuccess)) {
		mmg_set_error("cannot allocate %llu bytes for a result slot", (unsigned long long)(n * sizeof(T)));
		return MMG_ENOMEM;
	}
	if (getenv("MMG_POISON")) cudaMemset(*d, 0xCD, n * sizeof(T)), cudaDeviceSynchronize();
	*cap = n;
	return MMG_OK;
}

/* move the finished results of a slot from pinned staging into the batch */
static int slot_drain(mmg_aligner *al, mmg_batch *b, int k)
{
	mmg_aligner::ResSlot &r = al->rs[k];
	if (!r.pending) return MMG_OK;
	CK(cudaEventSynchronize(r.ev_out));
	uint64_t acc = r.hit_base;
	for (uint32_t i = 0; i < r.n_reads; ++i) b->hit_off[r.read0 + i] = acc, acc += r.h_nregs[i];
	if (acc != r.hit_base + r.n_hits) { mmg_set_error("internal: hit count mismatch in a result slot (%llu vs %llu)", (unsigned long long)acc, (unsigned long long)(r.hit_base + r.n_hits)); return MMG_ECUDA; }
	r.pending = false;
	return MMG_OK;
}

/* One chunk of reads [r0, r1) of batch b through every stage.  c.seq / c.off / c.off0 are set by the caller
 * (whole-batch device buffers in resident mode, an input slot in streamed mode). */
static int run_chunk(mmg_aligner *al, mmg_batch *b, ChunkDev &c, uint32_t r0, std::vector<uint64_t> &h_aoff)
{
	cudaStream_t st = al->stream;
	uint32_t *work = c.work;
	int wi = 0;
	CK(cudaMemsetAsync(c.work, 0, 64 * 4, st));
	CK(cudaMemsetAsync(c.err, 0, 4, st));
	CK(cudaMemsetAsync(c.flags, 0, (size_t)c.n_reads * 4, st));
	/* work order: reads by descending length, so that a kernel's persistent warps take the long reads first and the
	 * launch does not end with a few of them still running (computed on the device: a host-side table would have to
	 * queue behind the next chunk's input on the copy engine) */
	if (c.n_reads) MMG_LAUNCH(read_order_kernel, 1, 1024, 0, st, c.off, c.n_reads, al->d_order);
	c.order = al->d_order;
	STAGE_BEGIN(); launch_sketch(c, al->di, al->n_sms, st, work + wi++); STAGE_END(ST_SKETCH);
	STAGE_BEGIN(); launch_seed(c, al->di, al->dopt, al->n_sms, st, work + wi++); STAGE_END(ST_SEED);
	STAGE_BEGIN(); launch_scan_u32(c.n_a, c.a_off, c.n_reads, st); STAGE_END(ST_SCAN);
	uint64_t a_total = 0;
	CK(cudaMemcpyAsync(&a_total, c.a_off + c.n_reads, 8, cudaMemcpyDeviceToHost, st));
	CK(cudaStreamSynchronize(st));
	/* isolated-anchor filter (exact under these conditions, see seed.cu): worth its two passes over the hits only
	 * when reads carry many anchors, i.e. on large references */
	if (al->anchor_filter && !(al->mo.flag & MMG_F_RMQ) && al->mo.min_cnt >= 2 && al->mo.min_chain_score > al->idx->k && c.n_reads > 0 &&
	    a_total > (uint64_t)c.n_reads * 64 && (a_total >> 5) + c.n_reads + 1 <= al->cap_keep_words) {
		STAGE_BEGIN();
		CK(cudaMemcpyAsync(c.af_off, c.a_off, (size_t)(c.n_reads + 1) * 8, cudaMemcpyDeviceToDevice, st));
		launch_anchor_filter(c, al->di, al->dopt, al->n_sms, st, work + wi++);
		launch_scan_u32(c.n_a, c.a_off, c.n_reads, st);
		STAGE_END(ST_EXPAND);
		CK(cudaMemcpyAsync(&a_total, c.a_off + c.n_reads, 8, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
	}
	const bool split = a_total > al->cap_anchors;
	if (split) c.order = 0; /* the order is a permutation of the whole chunk, not of a sub-range */
	if (split) { /* the per-read offsets are only needed to cut sub-ranges */
		h_aoff.resize(c.n_reads + 1);
		CK(cudaMemcpyAsync(h_aoff.data(), c.a_off, (size_t)(c.n_reads + 1) * 8, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
	}
	/* sub-ranges whose anchors fit the anchor-sized arenas */
	for (uint32_t s0 = 0; s0 < c.n_reads;) {
		uint32_t s1 = s0;
		if (!split) s1 = c.n_reads, c.a_off0 = 0;
		else {
			while (s1 < c.n_reads && h_aoff[s1 + 1] - h_aoff[s0] <= al->cap_anchors) ++s1;
			if (s1 == s0) { mmg_set_error("read %u has %llu anchors, more than anchor_cap", r0 + s0, (unsigned long long)(h_aoff[s0 + 1] - h_aoff[s0])); return MMG_ENOMEM; }
			c.a_off0 = h_aoff[s0];
		}
		if (wi + 24 > 64) { CK(cudaMemsetAsync(c.work, 0, 64 * 4, st)); wi = 0; }
		if (al->dual_stream && !al->profile && !split && c.order && s1 - s0 >= (uint32_t)al->dual_min) {
			/* Every kernel of expand -> sort -> chain -> backtrack -> re-chain ends with a tail in which a few long reads
			 * (or equal-key replays) keep a handful of SMs busy.  The reads are cut in two interleaved halves (the work
			 * order puts the even ranks first, the odd ranks second) that go through these stages on two streams: the
			 * tail of one half's kernel is filled by the other half's next kernel.  Same arenas (the halves own disjoint
			 * slices), separate work counters and deferred-read lists.  Per-stage events would overlap, so stage timing
			 * ("profile") keeps the single-stream order. */
			const uint32_t mid = s0 + (s1 - s0 + 1) / 2;
			ChunkDev c2 = c;
			c2.big_list = c.big_list + al->cap_reads / 2, c2.tie_list = c.tie_list + al->cap_reads / 2;
			CK(cudaEventRecord(al->ev_fork, st));
			CK(cudaStreamWaitEvent(al->st2, al->ev_fork, 0));
			for (int h = 0; h < 2; ++h) {
				const ChunkDev &ch = h ? c2 : c;
				cudaStream_t sh = h ? al->st2 : st;
				const uint32_t a = h ? mid : s0, bnd = h ? s1 : mid;
				launch_expand(ch, al->di, al->dopt, a, bnd, al->n_sms, sh, work + wi++);
				launch_sort(ch, al->di, a, bnd, al->n_sms, sh, work + wi); wi += 5;
				if (al->mo.flag & MMG_F_RMQ) launch_chain_rmq(ch, al->dopt, a, bnd, al->rmq_nodes, al->n_sms, sh, work + wi++);
				else launch_chain(ch, al->dopt, a, bnd, al->n_sms, sh, work + wi++);
				launch_backtrack(ch, al->dopt, a, bnd, al->n_sms, sh, work + wi++);
				launch_rechain(ch, al->dopt, a, bnd, al->rmq_nodes, al->n_sms, sh, work + wi++);
			}
			CK(cudaEventRecord(al->ev_join, al->st2));
			CK(cudaStreamWaitEvent(st, al->ev_join, 0));
			al->stage_launches[ST_EXPAND] += 2, al->stage_launches[ST_SORT] += 2, al->stage_launches[ST_CHAIN] += 2, al->stage_launches[ST_BACKTRACK] += 2, al->stage_launches[ST_RECHAIN] += 2;
		} else {
		STAGE_BEGIN(); launch_expand(c, al->di, al->dopt, s0, s1, al->n_sms, st, work + wi++); STAGE_END(ST_EXPAND);
		STAGE_BEGIN(); launch_sort(c, al->di, s0, s1, al->n_sms, st, work + wi); wi += 5; STAGE_END(ST_SORT);
		STAGE_BEGIN();
		if (al->mo.flag & MMG_F_RMQ) launch_chain_rmq(c, al->dopt, s0, s1, al->rmq_nodes, al->n_sms, st, work + wi++);   /* asm presets: mm_lchain_rmq */
		else launch_chain(c, al->dopt, s0, s1, al->n_sms, st, work + wi++);
		STAGE_END(ST_CHAIN);
		STAGE_BEGIN(); launch_backtrack(c, al->dopt, s0, s1, al->n_sms, st, work + wi++); STAGE_END(ST_BACKTRACK);
		STAGE_BEGIN(); launch_rechain(c, al->dopt, s0, s1, al->rmq_nodes, al->n_sms, st, work + wi++); STAGE_END(ST_RECHAIN);
		}
		const bool with_cigar = (al->mo.flag & MMG_F_CIGAR) != 0;
		STAGE_BEGIN();
		if (with_cigar) { /* region slices leave room for the pieces z-drop splits insert */
			MMG_LAUNCH(reg_cap_kernel, (int)((s1 - s0 + 255) / 256), 256, 0, st, (const uint32_t*)(c.n_u + s0), al->xb.reg_cap + s0, s1 - s0);
			launch_scan_u32(al->xb.reg_cap + s0, c.r_off + s0, s1 - s0, st);
		} else launch_scan_u32(c.n_u + s0, c.r_off + s0, s1 - s0, st);
		STAGE_END(ST_SCAN);
		STAGE_BEGIN(); launch_regs(c, al->di, al->dopt, s0, s1, al->cap_regs, al->n_sms, st, work + wi++); STAGE_END(ST_REGS);
		uint64_t n_cg_sub = 0;
		if (with_cigar) {
			int rc2 = run_extension(al, b, c, s0, s1, work, &wi, &n_cg_sub);
			if (rc2) return rc2;
		}
		STAGE_BEGIN();
		launch_scan_u32(c.n_regs + s0, c.h_off + s0, s1 - s0, st);
		STAGE_END(ST_SCAN);
		CK(cudaMemcpyAsync(al->h_ctl + 6, c.h_off + s1, 8, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(al->h_ctl + 7, c.r_off + s1, 8, cudaMemcpyDeviceToHost, st));
		CK(cudaMemcpyAsync(al->h_ctl + 8, c.err, 4, cudaMemcpyDeviceToHost, st));
		CK(cudaStreamSynchronize(st));
		const uint64_t n_hits_sub = al->h_ctl[6], n_regs_sub = al->h_ctl[7];
		if (with_cigar) n_cg_sub = al->h_ctl[5];
		if ((uint32_t)al->h_ctl[8] & 0x80000000u) { mmg_set_error("region arena overflow (regs_cap)"); return MMG_ENOMEM; }
		if ((uint32_t)al->h_ctl[8]) { mmg_set_error("device arena overflow (flags 0x%x)", (unsigned)al->h_ctl[8]); return MMG_ENOMEM; }
		if (n_regs_sub > al->cap_regs) { mmg_set_error("region arena overflow (%llu > regs_cap)", (unsigned long long)n_regs_sub); return MMG_ENOMEM; }
		if (!b->streamed) {
			if (b->n_hits_dev + n_hits_sub > b->hits_cap) { mmg_set_error("result pool overflow (%llu hits)", (unsigned long long)(b->n_hits_dev + n_hits_sub)); return MMG_ENOMEM; }
			STAGE_BEGIN();
			launch_pack_hits(c, s0, s1, b->d_hits + b->n_hits_dev, al->n_sms, st);
			if (with_cigar) {
				if (b->n_cigar_dev + n_cg_sub > b->cigar_cap) { mmg_set_error("CIGAR pool overflow (%llu ops)", (unsigned long long)(b->n_cigar_dev + n_cg_sub)); return MMG_ENOMEM; }
				launch_pack_cigar(c, al->xb, s0, s1, b->d_hits + b->n_hits_dev, b->d_cigar + b->n_cigar_dev, b->n_cigar_dev, al->cg_read_off, al->n_sms, st);
			}
			CK(cudaMemcpyAsync(b->d_nregs + r0 + s0, c.n_regs + s0, (size_t)(s1 - s0) * 4, cudaMemcpyDeviceToDevice, st));
			STAGE_END(ST_REGS);
		} else {
			/* results of this sub-range leave through a slot: pack on the compute stream, copy out on s_out */
			const int k = (int)(al->n_sub++ & 1);
			mmg_aligner::ResSlot &r = al->rs[k];
			int rc;
			if ((rc = slot_drain(al, b, k))) return rc;
			if ((rc = slot_grow(&r.d_hits, (mmg_hit_t**)0, &r.hits_cap, n_hits_sub + 1))) return rc;
			if ((b->n_hits_dev + n_hits_sub) * sizeof(mmg_hit_t) > b->ph_bytes) { /* rare: more hits than reserved */
				uint64_t nb = 0;
				mmg_hit_t *np = (mmg_hit_t*)pool_acquire(b->pool, 2 * (b->n_hits_dev + n_hits_sub) * sizeof(mmg_hit_t), &nb);
				if (!np) { mmg_set_error("cannot allocate pinned memory for the results"); return MMG_ENOMEM; }
				CK(cudaStreamSynchronize(al->s_out));
				memcpy(np, b->ph, b->n_hits_dev * sizeof(mmg_hit_t));
				pool_release(b->pool, b->ph, b->ph_bytes);
				b->ph = np, b->ph_bytes = nb;
			}
			if (with_cigar && (b->n_cigar_dev + n_cg_sub) * 4 > b->pc_bytes) { /* rare: more operations than reserved */
				uint64_t nb = 0;
				uint32_t *np = (uint32_t*)pool_acquire(b->pool, 2 * (b->n_cigar_dev + n_cg_sub) * 4, &nb);
				if (!np) { mmg_set_error("cannot allocate pinned memory for the results"); return MMG_ENOMEM; }
				CK(cudaStreamSynchronize(al->s_out));
				if (b->pc) { memcpy(np, b->pc, b->n_cigar_dev * 4); pool_release(b->pool, b->pc, b->pc_bytes); }
				b->pc = np, b->pc_bytes = nb;
			}
			if ((rc = slot_grow(&r.d_nregs, &r.h_nregs, &r.nregs_cap, (uint64_t)(s1 - s0) + 1))) return rc;
			if (with_cigar && (rc = slot_grow(&r.d_cigar, (uint32_t**)0, &r.cigar_cap, n_cg_sub + 1))) return rc;
			STAGE_BEGIN();
			launch_pack_hits(c, s0, s1, r.d_hits, al->n_sms, st);
			if (with_cigar) launch_pack_cigar(c, al->xb, s0, s1, r.d_hits, r.d_cigar, b->n_cigar_dev, al->cg_read_off, al->n_sms, st);
			CK(cudaMemcpyAsync(r.d_nregs, c.n_regs + s0, (size_t)(s1 - s0) * 4, cudaMemcpyDeviceToDevice, st));
			STAGE_END(ST_REGS);
			CK(cudaEventRecord(r.ev_packed, st));
			CK(cudaStreamWaitEvent(al->s_out, r.ev_packed, 0));
			if (n_hits_sub) CK(cudaMemcpyAsync(b->ph + b->n_hits_dev, r.d_hits, n_hits_sub * sizeof(mmg_hit_t), cudaMemcpyDeviceToHost, al->s_out));
			if (n_cg_sub) CK(cudaMemcpyAsync(b->pc + b->n_cigar_dev, r.d_cigar, n_cg_sub * 4, cudaMemcpyDeviceToHost, al->s_out));
			CK(cudaMemcpyAsync(r.h_nregs, r.d_nregs, (size_t)(s1 - s0) * 4, cudaMemcpyDeviceToHost, al->s_out));
			CK(cudaEventRecord(r.ev_out, al->s_out));
			/* the next sub-range that packs into this slot must not start before the copy-out is done */
			CK(cudaStreamWaitEvent(st, r.ev_out, 0));
			r.pending = true, r.n_hits = n_hits_sub, r.n_cigar = n_cg_sub, r.hit_base = b->n_hits_dev, r.cigar_base = b->n_cigar_dev;
			r.read0 = r0 + s0, r.n_reads = s1 - s0;
		}
		b->n_cigar_dev += n_cg_sub;
		b->n_hits_dev += n_hits_sub;
		b->dbg_r0 = r0 + s0, b->dbg_r1 = r0 + s1;
		s0 = s1;
	}
	return MMG_OK;
}

/* ramp: streamed mode starts with a smaller chunk (1/2 of the arena by default) so that the first kernels start after
 * a short copy-in and the copy of every later chunk hides behind the compute of its predecessor */
static void cut_chunks(const mmg_aligner *al, const mmg_batch *b, std::vector<uint32_t> &cuts, bool ramp)
{
	const uint32_t n = b->n_reads;
	if (al->mo.flag & MMG_F_CIGAR) ramp = false; /* extension dominates and scales with the bases: equal chunks, the copy-in is negligible */
	/* pass 0 fills every chunk to the arena (that fixes the number of chunks); pass 1 cuts the same number of chunks
	 * evenly, so that the last one is not a sliver that pays a full set of kernel tails for little work */
	size_t n_greedy = 0;
	for (int pass = 0; pass < 2; ++pass) {
		cuts.clear();
		cuts.push_back(0);
		int shift = ramp ? al->ramp_shift : 0;
		size_t k = 0;
		for (uint32_t r0 = 0; r0 < n; ++k) {
			uint32_t r1 = r0;
			uint64_t cap = al->cap_bases >> shift;
			if (pass == 1 && n_greedy > 1 && k < n_greedy) { /* even share of what is left, the ramped first chunk counting for its fraction */
				const double w = shift ? 1.0 / (double)(1 << shift) : 1.0, left_w = w + (double)(n_greedy - 1 - k);
				const uint64_t share = (uint64_t)((double)(b->off[n] - b->off[r0]) * w / left_w) + 1;
				if (share < cap) cap = share;
			}
			while (r1 < n && r1 - r0 < al->cap_reads && b->off[r1 + 1] - b->off[r0] <= cap) ++r1;
			if (r1 == r0) { /* a read longer than the ramped / even size (reads longer than the arena were rejected earlier) */
				while (r1 < n && r1 - r0 < al->cap_reads && b->off[r1 + 1] - b->off[r0] <= al->cap_bases) ++r1;
				if (r1 > r0 + 1 && pass == 1) r1 = r0 + 1;
			}
			cuts.push_back(r1);
			r0 = r1;
			shift = 0; /* one small first chunk, then full chunks */
		}
		if (pass == 0) { n_greedy = cuts.size() - 1; if (n_greedy <= 1) break; }
	}
}

extern "C" {

int mmg_batch_run(mmg_aligner *al, mmg_batch *b)
{
	if (!al->subs.empty()) { mmg_set_error("upload / run / fetch time one device: use mmg_map_batch on a multi-device aligner"); return MMG_EUNSUP; }
	if (!b->uploaded) { mmg_set_error("batch not uploaded"); return MMG_EINVAL; }
	CK(cudaSetDevice(al->device));
	int rc = alloc_arenas(al);
	if (rc) return rc;
	cudaStream_t st = al->stream;
	memset(al->stage_ms, 0, sizeof(al->stage_ms));
	memset(al->stage_launches, 0, sizeof(al->stage_launches));
	al->ev_used = 0;
	b->n_hits_dev = 0, b->n_cigar_dev = 0;
	CK(cudaMemsetAsync(b->d_stats, 0, MMG_N_STATS * 8, st));
	CK(cudaEventRecord(al->ev_run0, st));
	std::vector<uint64_t> h_aoff;
	std::vector<uint32_t> cuts;
	cut_chunks(al, b, cuts, false);
	for (size_t k = 0; k + 1 < cuts.size(); ++k) {
		const uint32_t r0 = cuts[k], r1 = cuts[k + 1];
		ChunkDev c = al->cd;
		c.n_reads = r1 - r0;
		c.seq = b->d_bases, c.off = b->d_off + r0, c.off0 = b->off[r0];
		c.stats = b->d_stats;
		if ((rc = run_chunk(al, b, c, r0, h_aoff))) return rc;
	}
	CK(cudaEventRecord(al->ev_run1, st));
	CK(cudaStreamSynchronize(st));
	CK(cudaGetLastError());
	{ float ms = 0; cudaEventElapsedTime(&ms, al->ev_run0, al->ev_run1); al->last_run_ms = ms; }
	stage_collect(al);
	b->ran = true;
	return MMG_OK;
}

int mmg_batch_fetch(mmg_aligner *al, mmg_batch *b)
{
	if (!b->ran) { mmg_set_error("batch not run"); return MMG_EINVAL; }
	if (b->streamed) { b->fetched = true; return MMG_OK; }
	if (!al->subs.empty()) { mmg_set_error("upload / run / fetch time one device: use mmg_map_batch on a multi-device aligner"); return MMG_EUNSUP; }
	CK(cudaSetDevice(al->device));
	cudaStream_t st = al->stream;
	std::vector<uint32_t> nregs(b->n_reads + 1);
	b->hits.resize(b->n_hits_dev);
	CK(cudaEventRecord(al->ev0, st));
	if (b->n_reads) CK(cudaMemcpyAsync(nregs.data(), b->d_nregs, (size_t)b->n_reads * 4, cudaMemcpyDeviceToHost, st));
	if (b->n_hits_dev) CK(cudaMemcpyAsync(b->hits.data(), b->d_hits, b->n_hits_dev * sizeof(mmg_hit_t), cudaMemcpyDeviceToHost, st));
	b->cigar.resize(b->n_cigar_dev);
	if (b->n_cigar_dev) CK(cudaMemcpyAsync(b->cigar.data(), b->d_cigar, b->n_cigar_dev * 4, cudaMemcpyDeviceToHost, st));
	CK(cudaMemcpyAsync(b->stats, b->d_stats, MMG_N_STATS * 8, cudaMemcpyDeviceToHost, st));
	CK(cudaEventRecord(al->ev1, st));
	CK(cudaStreamSynchronize(st));
	{ float ms = 0; cudaEventElapsedTime(&ms, al->ev0, al->ev1); al->stage_ms[ST_D2H] += ms; al->stage_launches[ST_D2H] += 1; }
	b->hit_off.resize(b->n_reads + 1);
	uint64_t acc = 0;
	for (uint32_t i = 0; i < b->n_reads; ++i) b->hit_off[i] = acc, acc += nregs[i];
	b->hit_off[b->n_reads] = acc;
	if (acc != b->n_hits_dev) { mmg_set_error("internal: hit count mismatch (%llu vs %llu)", (unsigned long long)acc, (unsigned long long)b->n_hits_dev); return MMG_ECUDA; }
	b->fetched = true;
	return MMG_OK;
}

/* Streamed mapping of a host batch: the H2D copy of chunk k+1 (s_in) and the D2H copy of finished
 * results (s_out) overlap the kernels of chunk k; no per-batch device allocation. */
static int map_batch_streamed(mmg_aligner *al, mmg_batch *b)
{
	int rc;
	if ((rc = alloc_arenas(al)) || (rc = stream_setup(al))) return rc;
	cudaStream_t st = al->stream;
	memset(al->stage_ms, 0, sizeof(al->stage_ms));
	memset(al->stage_launches, 0, sizeof(al->stage_launches));
	al->ev_used = 0, al->n_sub = 0;
	al->rs[0].pending = al->rs[1].pending = false;
	b->n_hits_dev = 0, b->n_cigar_dev = 0;
	b->hit_off.assign((size_t)b->n_reads + 1, 0);
	b->pool = al->pool;
	b->ph = (mmg_hit_t*)pool_acquire(b->pool, ((uint64_t)b->n_reads + (b->n_reads >> 3) + 1024) * sizeof(mmg_hit_t), &b->ph_bytes);
	if (!b->ph) { b->pool = 0; mmg_set_error("cannot allocate pinned memory for the results"); return MMG_ENOMEM; }
	if (al->mo.flag & MMG_F_CIGAR) { /* ~0.09 operations per base on 8 %-error reads; grown on demand */
		b->pc = (uint32_t*)pool_acquire(b->pool, (b->n_bases / 8 + 65536) * 4, &b->pc_bytes);
		if (!b->pc) { mmg_set_error("cannot allocate pinned memory for the results"); return MMG_ENOMEM; }
	}
	std::vector<uint32_t> cuts;
	cut_chunks(al, b, cuts, true);
	const size_t n_chunks = cuts.size() - 1;
	CK(cudaMemsetAsync(al->d_stats_pool, 0, MMG_N_STATS * 8, st));
	CK(cudaEventRecord(al->ev_run0, st));
	auto copy_in = [&](size_t k) -> int {
		const int slot = (int)(k & 1);
		const uint32_t r0 = cuts[k], r1 = cuts[k + 1];
		if (k >= 2) CK(cudaEventSynchronize(al->ev_in[slot])); /* the host staging of the offsets is free again */
		uint64_t *ho = al->h_in_off[slot];
		const uint64_t base = b->off[r0];
		for (uint32_t i = r0; i <= r1; ++i) ho[i - r0] = b->off[i] - base;
		if (k >= 2) CK(cudaStreamWaitEvent(al->s_in, al->ev_free[slot], 0)); /* chunk k-2 has finished reading the slot */
		CK(cudaMemcpyAsync(al->in_bases[slot], b->h_bases + b->h_off[0] + base, b->off[r1] - base, cudaMemcpyHostToDevice, al->s_in));
		CK(cudaMemcpyAsync(al->in_off[slot], ho, (size_t)(r1 - r0 + 1) * 8, cudaMemcpyHostToDevice, al->s_in));
		CK(cudaEventRecord(al->ev_in[slot], al->s_in));
		return MMG_OK;
	};
	std::vector<uint64_t> h_aoff;
	if (n_chunks && (rc = copy_in(0))) return rc;
	for (size_t k = 0; k < n_chunks; ++k) {
		const int slot = (int)(k & 1);
		if (k + 1 < n_chunks && (rc = copy_in(k + 1))) return rc;
		CK(cudaStreamWaitEvent(st, al->ev_in[slot], 0));
		ChunkDev c = al->cd;
		c.n_reads = cuts[k + 1] - cuts[k];
		c.seq = al->in_bases[slot], c.off = al->in_off[slot], c.off0 = 0;
		c.stats = al->d_stats_pool;
		if ((rc = run_chunk(al, b, c, cuts[k], h_aoff))) return rc;
		CK(cudaEventRecord(al->ev_free[slot], st));
	}
	CK(cudaMemcpyAsync(b->stats, al->d_stats_pool, MMG_N_STATS * 8, cudaMemcpyDeviceToHost, st));
	CK(cudaEventRecord(al->ev_run1, st));
	CK(cudaStreamSynchronize(st));
	/* the slots were filled in issue order: drain the older one first so the vectors grow monotonically */
	const int older = (int)(al->n_sub & 1);
	if ((rc = slot_drain(al, b, older)) || (rc = slot_drain(al, b, older ^ 1))) return rc;
	CK(cudaStreamSynchronize(al->s_out));
	CK(cudaGetLastError());
	{ float ms = 0; cudaEventElapsedTime(&ms, al->ev_run0, al->ev_run1); al->last_run_ms = ms; }
	stage_collect(al);
	b->hit_off[b->n_reads] = b->n_hits_dev;
	b->ran = b->fetched = true;
	return MMG_OK;
}

/* Multi-device mmg_map_batch: the reads are cut into contiguous shards of (nearly) equal BASES, one per device; every
 * device maps its shard with its own streams on its own host thread (no collective: the path has no exchange step),
 * and the results are gathered into one batch in read order - what N workers sharing one index do in the reference
 * (/root/reference/src/lib.rs:545-553). */
static int map_batch_group(mmg_aligner *g, const char *bases, const uint64_t *offsets, uint32_t n_reads, mmg_batch **out)
{
	const size_t nd = g->subs.size();
	std::vector<uint32_t> cut(nd + 1, 0);
	const uint64_t total = n_reads ? offsets[n_reads] - offsets[0] : 0;
	for (size_t d = 1; d < nd; ++d) { /* first read whose start offset reaches the d-th share of the bases */
		const uint64_t target = offsets[0] + total / nd * d + total % nd * d / nd;
		uint32_t i = (uint32_t)(std::lower_bound(offsets, offsets + n_reads + 1, target) - offsets);
		cut[d] = i < cut[d - 1] ? cut[d - 1] : i > n_reads ? n_reads : i;
	}
	cut[nd] = n_reads;
	std::vector<mmg_batch*> sb(nd, (mmg_batch*)0);
	std::vector<int> rcs(nd, 0);
	std::vector<std::string> errs(nd);
	std::vector<std::thread> th;
	for (size_t d = 0; d < nd; ++d)
		th.emplace_back([&, d]() {
			rcs[d] = mmg_map_batch(g->subs[d], bases, offsets + cut[d], cut[d + 1] - cut[d], &sb[d]);
			if (rcs[d]) errs[d] = mmg_last_error();
		});
	for (size_t d = 0; d < nd; ++d) th[d].join();
	for (size_t d = 0; d < nd; ++d)
		if (rcs[d]) {
			mmg_set_error("device %d: %s", g->subs[d]->device, errs[d].c_str());
			for (size_t k = 0; k < nd; ++k) mmg_batch_destroy(sb[k]);
			return rcs[d];
		}
	mmg_batch *b = new mmg_batch();
	b->n_reads = n_reads, b->n_bases = total, b->h_bases = bases, b->h_off = offsets;
	b->d_bases = 0, b->d_off = 0, b->d_hits = 0, b->d_nregs = 0, b->d_stats = 0, b->d_cigar = 0, b->cigar_cap = 0, b->hits_cap = 0;
	b->uploaded = false, b->ran = b->fetched = true, b->streamed = true;
	b->pool = g->subs[0]->pool, b->ph = 0, b->ph_bytes = 0, b->pc = 0, b->pc_bytes = 0, b->dbg_r0 = b->dbg_r1 = 0;
	memset(b->stats, 0, sizeof(b->stats));
	std::vector<uint64_t> hbase(nd + 1, 0), cbase(nd + 1, 0);
	for (size_t d = 0; d < nd; ++d) hbase[d + 1] = hbase[d] + sb[d]->n_hits_dev, cbase[d + 1] = cbase[d] + sb[d]->n_cigar_dev;
	b->n_hits_dev = hbase[nd], b->n_cigar_dev = cbase[nd];
	b->ph = (mmg_hit_t*)pool_acquire(b->pool, (b->n_hits_dev + 1) * sizeof(mmg_hit_t), &b->ph_bytes);
	if (b->n_cigar_dev) b->pc = (uint32_t*)pool_acquire(b->pool, b->n_cigar_dev * 4, &b->pc_bytes);
	if (!b->ph || (b->n_cigar_dev && !b->pc)) {
		mmg_set_error("cannot allocate host memory for the gathered results");
		for (size_t k = 0; k < nd; ++k) mmg_batch_destroy(sb[k]);
		mmg_batch_destroy(b);
		return MMG_ENOMEM;
	}
	b->hit_off.resize((size_t)n_reads + 1);
	th.clear();
	for (size_t d = 0; d < nd; ++d)
		th.emplace_back([&, d]() { /* gather: every shard moves its own records; CIGAR offsets are rebased */
			const mmg_batch *s = sb[d];
			for (uint32_t i = 0; i < s->n_reads; ++i) b->hit_off[cut[d] + i] = hbase[d] + s->hit_off[i];
			mmg_hit_t *dst = b->ph + hbase[d];
			if (s->n_hits_dev) memcpy(dst, s->ph, s->n_hits_dev * sizeof(mmg_hit_t));
			if (cbase[d]) for (uint64_t i = 0; i < s->n_hits_dev; ++i) dst[i].cigar_off += cbase[d];
			if (s->n_cigar_dev) memcpy(b->pc + cbase[d], s->pc, s->n_cigar_dev * 4);
		});
	for (size_t d = 0; d < nd; ++d) th[d].join();
	b->hit_off[n_reads] = b->n_hits_dev;
	g->last_run_ms = 0;
	memset(g->stage_ms, 0, sizeof(g->stage_ms)), memset(g->stage_launches, 0, sizeof(g->stage_launches));
	for (size_t d = 0; d < nd; ++d) {
		for (int k = 0; k < MMG_N_STATS; ++k) b->stats[k] += sb[d]->stats[k];
		if (g->subs[d]->last_run_ms > g->last_run_ms) g->last_run_ms = g->subs[d]->last_run_ms;   /* devices run side by side */
		for (int k = 0; k < MMG_N_STAGES; ++k) {
			if (g->subs[d]->stage_ms[k] > g->stage_ms[k]) g->stage_ms[k] = g->subs[d]->stage_ms[k];
			g->stage_launches[k] += g->subs[d]->stage_launches[k];
		}
		mmg_batch_destroy(sb[d]);
	}
	*out = b;
	return MMG_OK;
}

int mmg_map_batch(mmg_aligner *al, const char *bases, const uint64_t *offsets, uint32_t n_reads, mmg_batch **out)
{
	*out = 0;
	if (!al->subs.empty()) return map_batch_group(al, bases, offsets, n_reads, out);
	CK(cudaSetDevice(al->device));
	for (uint32_t i = 0; i < n_reads; ++i) {
		uint64_t l = offsets[i + 1] - offsets[i];
		if (l > 0x7fffffffULL || l > al->cap_bases) { mmg_set_error("read %u is longer than the chunk capacity (%llu bases)", i, (unsigned long long)al->cap_bases); return MMG_EINVAL; }
	}
	mmg_batch *b = new mmg_batch();
	b->n_reads = n_reads, b->h_bases = bases, b->h_off = offsets;
	b->n_bases = n_reads ? offsets[n_reads] - offsets[0] : 0;
	b->off.resize((size_t)n_reads + 1);
	for (uint32_t i = 0; i <= n_reads; ++i) b->off[i] = offsets[i] - offsets[0];
	b->d_bases = 0, b->d_off = 0, b->d_hits = 0, b->d_nregs = 0, b->d_stats = 0, b->d_cigar = 0, b->cigar_cap = 0, b->n_cigar_dev = 0;
	b->hits_cap = 0, b->n_hits_dev = 0;
	b->uploaded = b->ran = b->fetched = false;
	b->streamed = true;
	b->pool = 0, b->ph = 0, b->ph_bytes = 0, b->pc = 0, b->pc_bytes = 0;
	b->dbg_r0 = b->dbg_r1 = 0;
	memset(b->stats, 0, sizeof(b->stats));
	int rc = map_batch_streamed(al, b);
	if (rc) {
		cudaStreamSynchronize(al->stream), cudaStreamSynchronize(al->s_in), cudaStreamSynchronize(al->s_out);
		mmg_batch_destroy(b);
		return rc;
	}
	*out = b;
	return MMG_OK;
}

void mmg_batch_destroy(mmg_batch *b)
{
	if (!b) return;
	if (b->ph && b->pool) pool_release(b->pool, b->ph, b->ph_bytes);
	if (b->pc && b->pool) pool_release(b->pool, b->pc, b->pc_bytes);
	if (b->d_bases) cudaFree(b->d_bases);
	if (b->d_off) cudaFree(b->d_off);
	if (b->d_hits) cudaFree(b->d_hits);
	if (b->d_nregs) cudaFree(b->d_nregs);
	if (b->d_stats) cudaFree(b->d_stats);
	if (b->d_cigar) cudaFree(b->d_cigar);
	delete b;
}

uint32_t mmg_batch_n_reads(const mmg_batch *b) { return b->n_reads; }
uint64_t mmg_batch_n_hits(const mmg_batch *b) { return b->streamed ? b->n_hits_dev : b->hits.size(); }
const uint64_t *mmg_batch_hit_off(const mmg_batch *b) { return b->hit_off.data(); }
const mmg_hit_t *mmg_batch_hits(const mmg_batch *b) { return b->streamed ? b->ph : b->hits.data(); }
uint64_t mmg_batch_n_cigar(const mmg_batch *b) { return b->streamed ? b->n_cigar_dev : b->cigar.size(); }
const uint32_t *mmg_batch_cigar(const mmg_batch *b) { return b->streamed ? b->pc : b->cigar.data(); }
int mmg_batch_stats(const mmg_batch *b, uint64_t out[MMG_N_STATS]) { memcpy(out, b->stats, sizeof(b->stats)); return MMG_OK; }

int mmg_stage_times(const mmg_aligner *al, double ms[MMG_N_STAGES], uint64_t launches[MMG_N_STAGES])
{
	memcpy(ms, al->stage_ms, sizeof(al->stage_ms));
	memcpy(launches, al->stage_launches, sizeof(al->stage_launches));
	return MMG_OK;
}
double mmg_last_run_ms(const mmg_aligner *al) { return al->last_run_ms; }
const char *mmg_stage_name(int s) { return s >= 0 && s < MMG_N_STAGES ? g_stage_names[s] : 0; }

int64_t mmg_debug_dump(mmg_aligner *al, mmg_batch *b, int which, uint64_t *x, uint64_t *y, uint64_t cap, uint64_t *off)
{
	/* valid for the reads of the last processed sub-range only (tests use single-chunk batches) */
	const ChunkDev &c = al->cd;
	uint32_t n = b->dbg_r1 - b->dbg_r0;
	if (b->dbg_r0 != 0 || n != b->n_reads) { mmg_set_error("debug dump needs a batch that fits one chunk"); return MMG_EINVAL; }
	std::vector<uint32_t> cnt(n);
	std::vector<uint64_t> aoff(n + 1);
	if (cudaMemcpy(aoff.data(), c.a_off, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost) != cudaSuccess) return MMG_ECUDA;
	const uint32_t *dcnt = which == 0 ? c.n_mz : which == 1 ? c.n_a : which == 2 ? c.n_v : c.n_u;
	if (cudaMemcpy(cnt.data(), dcnt, (size_t)n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return MMG_ECUDA;
	uint64_t tot = 0;
	for (uint32_t i = 0; i < n; ++i) {
		off[i] = tot;
		uint64_t src = which == 0 ? b->off[i] - b->off[0] : aoff[i] - aoff[0];
		if (tot + cnt[i] <= cap && cnt[i]) {
			if (which == 0) {
				std::vector<uint32_t> y32(cnt[i]);
				cudaMemcpy(x + tot, c.mz_x + src, (size_t)cnt[i] * 8, cudaMemcpyDeviceToHost);
				cudaMemcpy(y32.data(), c.mz_y + src, (size_t)cnt[i] * 4, cudaMemcpyDeviceToHost);
				for (uint32_t j = 0; j < cnt[i]; ++j) y[tot + j] = y32[j];
			} else if (which == 1 || which == 2) {
				cudaMemcpy(x + tot, c.bx + src, (size_t)cnt[i] * 8, cudaMemcpyDeviceToHost);
				cudaMemcpy(y + tot, c.by + src, (size_t)cnt[i] * 8, cudaMemcpyDeviceToHost);
			} else {
				cudaMemcpy(x + tot, c.u + src, (size_t)cnt[i] * 8, cudaMemcpyDeviceToHost);
			}
		}
		tot += cnt[i];
	}
	off[n] = tot;
	return (int64_t)tot;
}

const char *mmg_version(void)
{
#ifdef MMG_EMU
	return "mmg 0.1 (SIMT-emulated test build)";
#else
	return "mmg 0.1 (sm_100a)";
#endif
}
int mmg_sizeof_hit(void) { return (int)sizeof(mmg_hit_t); }

} // extern "C"
