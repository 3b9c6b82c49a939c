/* simgen.c -- synthetic reference / read generator.
 * Test + benchmark infrastructure; not part of the product path.
 * PRNG: splitmix64 (self-contained so C and any re-implementation agree). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t sm64(uint64_t *s)
{
	uint64_t z = (*s += 0x9e3779b97f4a7c15ULL);
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
	return z ^ (z >> 31);
}
static inline double u01(uint64_t *s) { return (sm64(s) >> 11) * (1.0 / 9007199254740992.0); }

void sim_reference(uint64_t seed, uint64_t len, char *out)
{
	uint64_t s = seed * 0x2545F4914F6CDD1DULL + 1, i = 0;
	while (i < len) {
		uint64_t r = sm64(&s);
		for (int j = 0; j < 32 && i < len; ++j, r >>= 2) out[i++] = "ACGT"[r & 3];
	}
}

/* plant n_rep copies of random segments (len in [lmin,lmax]) with divergence dv: repeat-stress variant */
void sim_plant_repeats(uint64_t seed, char *ref, uint64_t len, int n_rep, int lmin, int lmax, double dv)
{
	uint64_t s = seed * 0x9E3779B97F4A7C15ULL + 7;
	for (int r = 0; r < n_rep; ++r) {
		uint64_t l = lmin + sm64(&s) % (uint64_t)(lmax - lmin + 1);
		if (l * 2 >= len) continue;
		uint64_t src = sm64(&s) % (len - l), dst = sm64(&s) % (len - l);
		for (uint64_t i = 0; i < l; ++i) {
			char c = ref[src + i];
			if (u01(&s) < dv) c = "ACGT"[sm64(&s) & 3];
			ref[dst + i] = c;
		}
	}
}

/* Repeat FAMILIES (SURVEY.md section 8d stress variant): n_fam random consensus sequences of lmin..lmax bases, each
 * copied `copies` times to random places with per-base divergence dv - high-copy k-mers that exercise
 * mm_seed_select / mid_occ, equal anchor keys and the re-chaining path. */
void sim_plant_families(uint64_t seed, char *ref, uint64_t len, int n_fam, int copies, int lmin, int lmax, double dv)
{
	uint64_t s = seed * 0x9E3779B97F4A7C15ULL + 11;
	char *cons = (char*)malloc((size_t)lmax + 1);
	for (int f = 0; f < n_fam; ++f) {
		uint64_t l = lmin + sm64(&s) % (uint64_t)(lmax - lmin + 1);
		if (l * 2 >= len) continue;
		for (uint64_t i = 0; i < l; ++i) cons[i] = "ACGT"[sm64(&s) & 3];
		for (int c = 0; c < copies; ++c) {
			uint64_t dst = sm64(&s) % (len - l);
			int rev = (int)(sm64(&s) & 1);
			for (uint64_t i = 0; i < l; ++i) {
				char b = rev ? cons[l - 1 - i] : cons[i];
				if (rev) b = b == 'A' ? 'T' : b == 'C' ? 'G' : b == 'G' ? 'C' : 'A';
				if (u01(&s) < dv) b = "ACGT"[sm64(&s) & 3];
				ref[dst + i] = b;
			}
		}
	}
	free(cons);
}

static inline char comp(char c) { switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return c; } }

/* Simulate reads from a concatenated reference with contig table (coff[n_ctg+1]).
 * out must hold n_reads * (len_max * 1.25 + 64) bytes; offsets has n_reads+1 entries.
 * truth: per read {ctg, start, end, strand}. Length distribution: uniform [len_min,len_max]
 * (len_sd <= 0) or normal(len_mean,len_sd) clipped to [len_min,len_max]. */
uint64_t sim_reads(uint64_t seed, const char *ref, int n_ctg, const uint64_t *coff, uint32_t n_reads,
                   int len_min, int len_max, double len_mean, double len_sd,
                   double p_sub, double p_ins, double p_del, char *out, uint64_t *offsets, int64_t *truth)
{
	uint64_t s = seed * 0xD1342543DE82EF95ULL + 3, o = 0, total = coff[n_ctg];
	for (uint32_t r = 0; r < n_reads; ++r) {
		int L;
		if (len_sd > 0) {
			double u1 = u01(&s), u2 = u01(&s);
			if (u1 < 1e-300) u1 = 1e-300;
			double z = __builtin_sqrt(-2.0 * __builtin_log(u1)) * __builtin_cos(6.283185307179586 * u2);
			L = (int)(len_mean + len_sd * z);
			if (L < len_min) L = len_min;
			if (L > len_max) L = len_max;
		} else L = len_min + (int)(sm64(&s) % (uint64_t)(len_max - len_min + 1));
		/* contig proportional to length */
		uint64_t g = sm64(&s) % total;
		int c = 0;
		while (c + 1 < n_ctg && coff[c + 1] <= g) ++c;
		uint64_t clen = coff[c + 1] - coff[c];
		if ((uint64_t)L > clen) L = (int)clen;
		uint64_t st = clen == (uint64_t)L ? 0 : sm64(&s) % (clen - L + 1);
		int strand = (int)(sm64(&s) & 1);
		const char *p = ref + coff[c] + st;
		uint64_t o0 = o;
		{
			const uint32_t t_del = (uint32_t)(p_del * 1048576.0), t_sub = (uint32_t)(p_sub * 1048576.0), t_ins = (uint32_t)(p_ins * 1048576.0);
			for (int i = 0; i < L; ++i) {
				uint64_t rr = sm64(&s); /* one draw per base: 3 x 20-bit thresholds + 4 spare bits */
				if ((uint32_t)(rr & 0xfffff) < t_del) continue;
				char b = p[i];
				if ((uint32_t)(rr >> 20 & 0xfffff) < t_sub) { int k = (int)((rr >> 60) % 3); const char *alt = b == 'A' ? "CGT" : b == 'C' ? "AGT" : b == 'G' ? "ACT" : "ACG"; b = alt[k]; }
				out[o++] = b;
				if ((uint32_t)(rr >> 40 & 0xfffff) < t_ins) out[o++] = "ACGT"[sm64(&s) & 3];
			}
		}
		if (strand) {
			uint64_t i = o0, j = o ? o - 1 : 0;
			while (i < j) { char a = comp(out[i]), b = comp(out[j]); out[i++] = b; out[j--] = a; }
			if (i == j && o > o0) out[i] = comp(out[i]);
		}
		offsets[r] = o0;
		if (truth) truth[4 * r] = c, truth[4 * r + 1] = (int64_t)st, truth[4 * r + 2] = (int64_t)st + L, truth[4 * r + 3] = strand;
	}
	offsets[n_reads] = o;
	return o;
}
