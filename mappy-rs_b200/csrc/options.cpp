/* options.cpp -- presets and derived thresholds.
 * Replaces mm_set_opt(NULL,..) / mm_set_opt(preset,..) (/root/reference/src/lib.rs:333,336)
 * and mm_mapopt_update (:414); constants per minimap2 v2.26 options.c
 * (SURVEY.md appendix A.1). */
#include <string.h>
#include <limits.h>
#include "mmg_internal.h"

static void idxopt_defaults(mmg_idxopt_t *io)
{
	memset(io, 0, sizeof(*io));
	io->k = 15, io->w = 10, io->flag = 0, io->bucket_bits = 14;
	io->mini_batch_size = 50000000;
	io->batch_size = 8000000000ULL;
}

static void mapopt_defaults(mmg_mapopt_t *mo)
{
	memset(mo, 0, sizeof(*mo));
	mo->seed = 11;
	mo->mid_occ_frac = 2e-4f, mo->min_mid_occ = 10, mo->max_mid_occ = 1000000;
	mo->sdust_thres = 0, mo->q_occ_frac = 0.01f;
	mo->min_cnt = 3, mo->min_chain_score = 40;
	mo->bw = 500, mo->bw_long = 20000;
	mo->max_gap = 5000, mo->max_gap_ref = -1;
	mo->max_chain_skip = 25, mo->max_chain_iter = 5000;
	mo->rmq_inner_dist = 1000, mo->rmq_size_cap = 100000, mo->rmq_rescue_size = 1000, mo->rmq_rescue_ratio = 0.1f;
	mo->chain_gap_scale = 0.8f, mo->chain_skip_scale = 0.0f;
	mo->max_max_occ = 4095, mo->occ_dist = 500;
	mo->mask_level = 0.5f, mo->mask_len = INT_MAX, mo->pri_ratio = 0.8f, mo->best_n = 5;
	mo->alt_drop = 0.15f;
	mo->a = 2, mo->b = 4, mo->q = 4, mo->e = 2, mo->q2 = 24, mo->e2 = 1;
	mo->transition = 0, mo->sc_ambi = 1;
	mo->zdrop = 400, mo->zdrop_inv = 200, mo->end_bonus = -1;
	mo->min_dp_max = mo->min_chain_score * mo->a;
	mo->min_ksw_len = 200, mo->anchor_ext_len = 20, mo->anchor_ext_shift = 6;
	mo->max_clip_ratio = 1.0f;
	mo->mini_batch_size = 500000000, mo->max_sw_mat = 100000000, mo->cap_kalloc = 1000000000;
	mo->rank_min_len = 500, mo->rank_frac = 0.9f;
	mo->pe_ori = 0, mo->pe_bonus = 33;
}

extern "C" int mmg_set_opt(const char *preset, mmg_idxopt_t *io, mmg_mapopt_t *mo)
{
	if (preset == 0) {
		idxopt_defaults(io);
		mapopt_defaults(mo);
		return MMG_OK;
	}
	struct P { const char *name; int kind; };
	if (!strcmp(preset, "map-ont") || !strcmp(preset, "lr")) return MMG_OK; /* same as the defaults */
	if (!strcmp(preset, "map-hifi") || !strcmp(preset, "map-ccs")) {
		io->flag = 0, io->k = 19, io->w = 19;
		mo->max_gap = 10000;
		mo->a = 1, mo->b = 4, mo->q = 6, mo->q2 = 26, mo->e = 2, mo->e2 = 1;
		mo->occ_dist = 500;
		mo->min_mid_occ = 50, mo->max_mid_occ = 500;
		mo->min_dp_max = 200;
		return MMG_OK;
	}
	if (!strcmp(preset, "ava-ont")) { /* options.c; NO_DIAG | NO_DUAL need a query name, which this path never has */
		io->flag = 0, io->k = 15, io->w = 5;
		mo->flag |= 0x001 | 0x002 | MMG_F_ALL_CHAINS | MMG_F_NO_LJOIN;
		mo->min_chain_score = 100, mo->pri_ratio = 0.0f, mo->max_chain_skip = 25;
		mo->bw = mo->bw_long = 2000;
		mo->occ_dist = 0;
		return MMG_OK;
	}
	if (!strncmp(preset, "asm", 3)) { /* assembly-to-reference: chaining by range-minimum query (MM_F_RMQ) */
		int a, b, q, q2, e, w = 19;
		if (!strcmp(preset, "asm5")) a = 1, b = 19, q = 39, q2 = 81, e = 3;
		else if (!strcmp(preset, "asm10")) a = 1, b = 9, q = 16, q2 = 41, e = 2;
		else if (!strcmp(preset, "asm20")) a = 1, b = 4, q = 6, q2 = 26, e = 2, w = 10;
		else { mmg_set_error("unknown preset '%s'", preset); return MMG_EINVAL; }
		io->flag = 0, io->k = 19, io->w = (short)w;
		mo->bw = 1000, mo->bw_long = 100000;
		mo->max_gap = 10000;
		mo->flag |= MMG_F_RMQ;
		mo->min_mid_occ = 50, mo->max_mid_occ = 500;
		mo->min_dp_max = 200;
		mo->best_n = 50;
		mo->a = a, mo->b = b, mo->q = q, mo->q2 = q2, mo->e = e, mo->e2 = 1, mo->zdrop = mo->zdrop_inv = 200;
		return MMG_OK;
	}
	mmg_set_error("preset '%s' is outside the supported path (map-ont / lr, map-hifi / map-ccs, ava-ont, asm5 / asm10 / asm20)", preset);
	return MMG_EINVAL;
}

extern "C" int mmg_mapopt_update(mmg_mapopt_t *mo, const mmg_index *idx)
{
	if (mo->mid_occ <= 0) {
		mo->mid_occ = mmg_index_cal_max_occ(idx, mo->mid_occ_frac);
		if (mo->mid_occ < mo->min_mid_occ) mo->mid_occ = mo->min_mid_occ;
		if (mo->max_mid_occ > mo->min_mid_occ && mo->mid_occ > mo->max_mid_occ) mo->mid_occ = mo->max_mid_occ;
	}
	if (mo->bw_long < mo->bw) mo->bw_long = mo->bw;
	return MMG_OK;
}
