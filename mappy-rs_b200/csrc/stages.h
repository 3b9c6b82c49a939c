/* stages.h -- host launchers of the stage kernels (one .cu per stage). */
#ifndef MMG_STAGES_H
#define MMG_STAGES_H
#include "dev_common.cuh"

#define SKETCH_WARPS 8
#define SEED_WARPS 8
#define CHAIN_WARPS 4
#define SORT_THREADS 256
#define SORT_SMEM_ELEMS 8192
#define SORT_SMALL_ELEMS 2048

enum { ST_H2D = 0, ST_SKETCH, ST_SEED, ST_SCAN, ST_EXPAND, ST_SORT, ST_CHAIN, ST_BACKTRACK, ST_RECHAIN, ST_REGS, ST_EXTEND, ST_D2H };

int launch_sketch(const ChunkDev &c, const DevIndex &di, int n_sms, cudaStream_t st, uint32_t *work);
int launch_seed(const ChunkDev &c, const DevIndex &di, const DevOpt &o, int n_sms, cudaStream_t st, uint32_t *work);
int launch_anchor_filter(const ChunkDev &c, const DevIndex &di, const DevOpt &o, int n_sms, cudaStream_t st, uint32_t *work);
int launch_scan_u32(const uint32_t *in, uint64_t *out, uint32_t n, cudaStream_t st);  /* out[n] = total */
int launch_expand(const ChunkDev &c, const DevIndex &di, const DevOpt &o, uint32_t r0, uint32_t r1, int n_sms, cudaStream_t st, uint32_t *work);
int launch_sort(const ChunkDev &c, const DevIndex &di, uint32_t r0, uint32_t r1, int n_sms, cudaStream_t st, uint32_t *work);
void mmg_sort_set_small_max(int v); /* reads with more anchors take the radix pass (tuning knob "sort_small_max") */
int launch_chain(const ChunkDev &c, const DevOpt &o, uint32_t r0, uint32_t r1, int n_sms, cudaStream_t st, uint32_t *work);
int launch_chain_rmq(const ChunkDev &c, const DevOpt &o, uint32_t r0, uint32_t r1, void *nodes, int n_sms, cudaStream_t st, uint32_t *work);
int launch_backtrack(const ChunkDev &c, const DevOpt &o, uint32_t r0, uint32_t r1, int n_sms, cudaStream_t st, uint32_t *work);
#define RMQ_NODE_BYTES 40 /* sizeof(RNode) in rmq.cu; the arena holds 2 * (anchors + reads) nodes */
int launch_rechain(const ChunkDev &c, const DevOpt &o, uint32_t r0, uint32_t r1, void *nodes, int n_sms, cudaStream_t st, uint32_t *work);
int launch_regs(const ChunkDev &c, const DevIndex &di, const DevOpt &o, uint32_t r0, uint32_t r1, uint64_t regs_cap, int n_sms, cudaStream_t st, uint32_t *work);
int launch_pack_hits(const ChunkDev &c, uint32_t r0, uint32_t r1, mmg_hit_t *hits, int n_sms, cudaStream_t st);

#endif
