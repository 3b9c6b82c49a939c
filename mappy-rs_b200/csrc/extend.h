/* extend.h -- job / region records and arenas of the extension stage (extend.cu). */
#ifndef MMG_EXTEND_H
#define MMG_EXTEND_H
#include "dev_common.cuh"

#define EXT_DP_WARPS 4
#define EXT_SMEM_PER_WARP 12288   /* bytes of shared memory per DP warp (18 B per target column + the query): jobs up to ~640 x 640 bases, longer ones use the global slice */
#define EXT_DPK_BYTES 96         /* per-warp slot of the pass constants (struct DpK) behind the DP slices */
#define EXT_LEFT 0
#define EXT_FILL 1
#define EXT_RIGHT 2
#define EXT_INV 3
#define EXT_PENDING 1
#define EXT_DONE 2
#define EXT_INV_PENDING 3

/* One banded DP (align.c mm_align_pair call) */
struct ExtJob {
	uint32_t read, reg;
	int32_t rid;
	int32_t qs, qe, rs, re;     /* query interval in qseq0[rev] coordinates, target interval */
	int32_t w, zdrop, end_bonus, flag;
	uint8_t kind, rev, zdropped, reach_end, zdrop_code, pad[3];
	uint64_t tb_size, tb_off;   /* traceback bytes the job needs (a warp keeps them in its own slice of the arena; tb_off unused) */
	uint32_t cg_size, n_cigar;
	uint64_t cg_off;            /* cigar slice (u32 units) */
	int32_t max, max_q, max_t, mqe, mqe_t, score;  /* ksw_extz_t */
};

/* Per-region state of mm_align1 that has to survive between the kernels */
struct ExtReg {
	int32_t state;
	int32_t as1, cnt1;          /* anchors kept by mm_fix_bad_ends */
	int32_t rs, qs, re, qe;     /* first / last kept anchor ends */
	int32_t rs0, qs0, re0, qe0; /* extension windows */
	uint32_t job0; int32_t n_jobs;
	int32_t pad[3];
};

struct ExtBufs {
	ExtJob *jobs; uint64_t cap_jobs;
	uint32_t *n_jobs;           /* device counter */
	ExtReg *xregs; uint64_t *xr_off; /* region slices: the same offsets as ChunkDev::regs */
	DevReg *regs_tmp;
	uint32_t *n_sq;             /* per read: anchors after mm_squeeze_a */
	uint8_t *tb; uint64_t cap_tb;   /* traceback arena: one slice per resident warp (a traceback lives only while its job runs) */
	uint32_t *jcigar, *rcigar; uint64_t cap_cg;
	unsigned long long *tb_base, *cg_base; /* [0] = base of the current round, [1] = end after the scan */
	unsigned char *big; uint64_t big_per_warp; /* global DP arrays for jobs that do not fit shared memory */
	uint32_t *n_pending;        /* regions created by splits, to be aligned in the next round */
	uint32_t *ovf, *ovf_n;      /* jobs whose traceback did not fit the slice of a DP pass, one list segment per pass; ovf_n[4] */
	uint32_t *reg_cap;          /* per read: capacity of its region slice */
};

int launch_ext_prep(const ChunkDev &c, const DevIndex &di, const DevOpt &o, const ExtBufs &xb, uint32_t r0, uint32_t r1, int round, int n_sms, cudaStream_t st, uint32_t *work);
/* j1 is read from the device (*xb.n_jobs) by these two: no host round trip between prep, scan, DP and stitch */
int launch_ext_job_scan(const ExtBufs &xb, uint32_t j0, int n_sms, cudaStream_t st);
int launch_ext_dp(const ChunkDev &c, const DevIndex &di, const DevOpt &o, const ExtBufs &xb, uint32_t j0, int n_sms, cudaStream_t st, uint32_t *work);
#define EXT_DP_COUNTERS 6   /* claim counters launch_ext_dp uses */
int launch_ext_stitch(const ChunkDev &c, const DevIndex &di, const DevOpt &o, const ExtBufs &xb, uint32_t r0, uint32_t r1, int round, int n_sms, cudaStream_t st, uint32_t *work);
int launch_ext_final(const ChunkDev &c, const DevIndex &di, const DevOpt &o, const ExtBufs &xb, uint32_t r0, uint32_t r1, int n_sms, cudaStream_t st, uint32_t *work);
int launch_pack_cigar(const ChunkDev &c, const ExtBufs &xb, uint32_t r0, uint32_t r1, mmg_hit_t *hits, uint32_t *cigar_out, uint64_t cigar_base, uint64_t *cg_read_off, int n_sms, cudaStream_t st);

#endif
