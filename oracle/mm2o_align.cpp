/* placeholder; replaced below */
#include <stdio.h>
#include <stdlib.h>
#include "mm2o.h"
mm_reg1_t *mm_align_skeleton(const mm_mapopt_t *opt, const mm_idx_t *mi, int qlen, const char *qstr, int *n_regs_, mm_reg1_t *regs, mm128_t *a, mm2o_stats_t *st)
{ fprintf(stderr, "mm_align_skeleton: not built yet\n"); abort(); }
std::string mm_gen_cs(const mm_idx_t *mi, const mm_reg1_t *r, const char *seq, int no_iden) { return std::string(); }
std::string mm_gen_MD(const mm_idx_t *mi, const mm_reg1_t *r, const char *seq) { return std::string(); }
