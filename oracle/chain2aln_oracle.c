/*
 * oracle/chain2aln_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of BWA-MEM 0.7.8's mem_chain2aln (bwa-0.7.8/bwamem.c:730-878), cal_max_gap
 * (bwamem.c:544-551) and bns_get_seq (bntseq.c:355-376), strictly sequential like the reference:
 * one seed at a time, the DP (ksw_oracle_extend2, oracle/ksw_oracle.c) is called only for the seeds
 * the containment test lets through, the right extension starts from the left score.
 * Pinned against the reference's own mem_chain2aln compiled from its sources (oracle/_ref/libbwa_ref.so,
 * oracle/ref_chain_shim.c) and against tests/golden/chain2aln_golden.npz generated from it.
 * Used only by tests/ as the checker of the product's batched driver (bwamem_ext.c).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int ksw_oracle_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                       int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                       int *qle, int *tle, int *gtle, int *gscore, int *max_off, int64_t *cells, int32_t *rows);

typedef struct { int64_t rbeg; int32_t qbeg, len; } o_seed_t;
typedef struct {
	int64_t rb, re; int qb, qe; int score, truesc, sub, csub, sub_n, w, seedcov, secondary; uint64_t hash;
} o_reg_t;
typedef struct { int a, b, o_del, e_del, o_ins, e_ins, pen_clip5, pen_clip3, w, zdrop; int8_t mat[25]; } o_opt_t;

static int gap_budget(const o_opt_t *o, int qlen)                 /* cal_max_gap */
{
	int d = (int)((double)(qlen * o->a - o->o_del) / o->e_del + 1.);
	int i = (int)((double)(qlen * o->a - o->o_ins) / o->e_ins + 1.);
	int l = d > i ? d : i;
	l = l > 1 ? l : 1;
	return l < o->w << 1 ? l : o->w << 1;
}

static uint8_t *ref_window(int64_t l_pac, const uint8_t *pac, int64_t beg, int64_t end, int64_t *len)   /* bns_get_seq */
{
	uint8_t *seq = 0;
	*len = 0;
	if (end < beg) { int64_t t = beg; beg = end; end = t; }
	if (end > 2 * l_pac) end = 2 * l_pac;
	if (beg < 0) beg = 0;
	if (beg >= l_pac || end <= l_pac) {
		int64_t k, l = 0;
		*len = end - beg;
		seq = malloc((size_t)(end - beg) + 1);
		if (beg >= l_pac) {
			for (k = 2 * l_pac - 1 - beg; k > 2 * l_pac - 1 - end; --k)
				seq[l++] = (uint8_t)(3 - ((pac[k >> 2] >> ((3 - (k & 3)) * 2)) & 3));
		} else {
			for (k = beg; k < end; ++k) seq[l++] = (uint8_t)((pac[k >> 2] >> ((3 - (k & 3)) * 2)) & 3);
		}
	}
	return seq;
}

static int by_key(const void *a, const void *b)
{
	uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
	return (x > y) - (x < y);
}

/* appends to regs[*n_regs...]; returns the number of ksw calls made (for statistics) */
static int chain2aln(const o_opt_t *o, int64_t l_pac, const uint8_t *pac, int l_query, const uint8_t *query,
                     int n_seeds, const o_seed_t *seeds, o_reg_t *regs, int64_t *n_regs, int64_t first_reg)
{
	int64_t rmax0 = 2 * l_pac, rmax1 = 0, rlen;
	uint8_t *rseq;
	uint64_t *srt;
	int i, k, calls = 0;
	if (n_seeds == 0) return 0;
	for (i = 0; i < n_seeds; ++i) {
		const o_seed_t *t = &seeds[i];
		int64_t b = t->rbeg - (t->qbeg + gap_budget(o, t->qbeg));
		int64_t e = t->rbeg + t->len + ((l_query - t->qbeg - t->len) + gap_budget(o, l_query - t->qbeg - t->len));
		if (b < rmax0) rmax0 = b;
		if (e > rmax1) rmax1 = e;
	}
	if (rmax0 < 0) rmax0 = 0;
	if (rmax1 > 2 * l_pac) rmax1 = 2 * l_pac;
	if (rmax0 < l_pac && l_pac < rmax1) {
		if (seeds[0].rbeg < l_pac) rmax1 = l_pac; else rmax0 = l_pac;
	}
	rseq = ref_window(l_pac, pac, rmax0, rmax1, &rlen);
	srt = malloc(8 * (size_t)n_seeds);
	for (i = 0; i < n_seeds; ++i) srt[i] = (uint64_t)seeds[i].len << 32 | (uint32_t)i;
	qsort(srt, (size_t)n_seeds, 8, by_key);
	for (k = n_seeds - 1; k >= 0; --k) {
		const o_seed_t *s = &seeds[(uint32_t)srt[k]];
		o_reg_t *a;
		int64_t r;
		int aw0 = o->w, aw1 = o->w, max_off = 0;
		for (r = first_reg; r < *n_regs; ++r) {
			const o_reg_t *p = &regs[r];
			int64_t rd; int qd, w, g;
			if (s->rbeg < p->rb || s->rbeg + s->len > p->re || s->qbeg < p->qb || s->qbeg + s->len > p->qe) continue;
			qd = s->qbeg - p->qb; rd = s->rbeg - p->rb;
			g = gap_budget(o, qd < rd ? qd : rd);
			w = g < o->w ? g : o->w;
			if (qd - rd < w && rd - qd < w) break;
			qd = p->qe - (s->qbeg + s->len); rd = p->re - (s->rbeg + s->len);
			g = gap_budget(o, qd < rd ? qd : rd);
			w = g < o->w ? g : o->w;
			if (qd - rd < w && rd - qd < w) break;
		}
		if (r < *n_regs) {
			for (i = k + 1; i < n_seeds; ++i) {
				const o_seed_t *t;
				if (srt[i] == 0) continue;
				t = &seeds[(uint32_t)srt[i]];
				if (t->len < s->len * .95) continue;
				if (s->qbeg <= t->qbeg && s->qbeg + s->len - t->qbeg >= s->len >> 2 && t->qbeg - s->qbeg != t->rbeg - s->rbeg) break;
				if (t->qbeg <= s->qbeg && t->qbeg + t->len - s->qbeg >= s->len >> 2 && s->qbeg - t->qbeg != s->rbeg - t->rbeg) break;
			}
			if (i == n_seeds) { srt[k] = 0; continue; }
		}
		a = &regs[(*n_regs)++];
		memset(a, 0, sizeof(*a));
		a->w = o->w; a->score = a->truesc = -1;
		if (s->qbeg) {
			int qle, tle, gtle, gscore, t;
			int64_t tmp = s->rbeg - rmax0;
			uint8_t *qs = malloc((size_t)s->qbeg), *rs = malloc((size_t)tmp + 1);
			for (i = 0; i < s->qbeg; ++i) qs[i] = query[s->qbeg - 1 - i];
			for (i = 0; i < tmp; ++i) rs[i] = rseq[tmp - 1 - i];
			for (t = 0; t < 2; ++t) {
				int prev = a->score;
				aw0 = o->w << t;
				a->score = ksw_oracle_extend2(s->qbeg, qs, (int)tmp, rs, 5, o->mat, o->o_del, o->e_del, o->o_ins, o->e_ins, aw0,
				                              o->pen_clip5, o->zdrop, s->len * o->a, &qle, &tle, &gtle, &gscore, &max_off, 0, 0);
				++calls;
				if (a->score == prev || max_off < (aw0 >> 1) + (aw0 >> 2)) break;
			}
			if (gscore <= 0 || gscore <= a->score - o->pen_clip5) { a->qb = s->qbeg - qle; a->rb = s->rbeg - tle; a->truesc = a->score; }
			else { a->qb = 0; a->rb = s->rbeg - gtle; a->truesc = gscore; }
			free(qs); free(rs);
		} else { a->score = a->truesc = s->len * o->a; a->qb = 0; a->rb = s->rbeg; }
		if (s->qbeg + s->len != l_query) {
			int qle, tle, gtle, gscore, t, sc0 = a->score;
			int qe = s->qbeg + s->len;
			int64_t re = s->rbeg + s->len - rmax0;
			for (t = 0; t < 2; ++t) {
				int prev = a->score;
				aw1 = o->w << t;
				a->score = ksw_oracle_extend2(l_query - qe, query + qe, (int)(rmax1 - rmax0 - re), rseq + re, 5, o->mat, o->o_del, o->e_del,
				                              o->o_ins, o->e_ins, aw1, o->pen_clip3, o->zdrop, sc0, &qle, &tle, &gtle, &gscore, &max_off, 0, 0);
				++calls;
				if (a->score == prev || max_off < (aw1 >> 1) + (aw1 >> 2)) break;
			}
			if (gscore <= 0 || gscore <= a->score - o->pen_clip3) { a->qe = qe + qle; a->re = rmax0 + re + tle; a->truesc += a->score - sc0; }
			else { a->qe = l_query; a->re = rmax0 + re + gtle; a->truesc += gscore - sc0; }
		} else { a->qe = l_query; a->re = s->rbeg + s->len; }
		for (i = 0, a->seedcov = 0; i < n_seeds; ++i) {
			const o_seed_t *t = &seeds[i];
			if (t->qbeg >= a->qb && t->qbeg + t->len <= a->qe && t->rbeg >= a->rb && t->rbeg + t->len <= a->re) a->seedcov += t->len;
		}
		a->w = aw0 > aw1 ? aw0 : aw1;
	}
	free(srt); free(rseq);
	return calls;
}

/* flat form, same argument meaning as b200_chain2aln_flat (include/bwamem_b200.h); n_calls = DP calls made */
int oracle_chain2aln_flat(const o_opt_t *opt, int64_t l_pac, const uint8_t *pac, int n_reads, const int64_t *read_off,
                          const int32_t *read_len, const uint8_t *qpool, int n_chains, const int32_t *chain_read,
                          const int64_t *chain_seed0, const int32_t *chain_nseeds, const o_seed_t *seeds, int64_t out_cap,
                          o_reg_t *out, int32_t *out_read, int64_t *n_out, int64_t *n_calls)
{
	int r, c = 0;
	int64_t n = 0, calls = 0;
	for (r = 0; r < n_reads; ++r) {
		const int64_t first = n;
		int64_t k;
		for (; c < n_chains && chain_read[c] == r; ++c) {
			if (n + chain_nseeds[c] > out_cap) return -1;
			calls += chain2aln(opt, l_pac, pac, read_len[r], qpool + read_off[r], chain_nseeds[c], seeds + chain_seed0[c], out, &n, first);
		}
		for (k = first; k < n; ++k) out_read[k] = r;
	}
	*n_out = n;
	if (n_calls) *n_calls = calls;
	return 0;
}
