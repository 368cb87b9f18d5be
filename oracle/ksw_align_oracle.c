/*
 * oracle/ksw_align_oracle.c — TEST INFRASTRUCTURE (never linked into the product library).
 *
 * CPU restatement of the reference's local-alignment routine ksw_align2 (bwa-0.7.8/ksw.c:329-354) and of the two
 * striped SSE2 kernels it drives, ksw_u8 (ksw.c:110-236) and ksw_i16 (ksw.c:238-320), with the query profile of
 * ksw_qinit (ksw.c:62-108).  mem_matesw calls it for mate rescue (bwamem_pair.c:150).
 *
 * The reference computes on 128-bit vectors with the query STRIPED over the lanes (lane l of vector j holds column
 * j + l*slen, ksw.c:87-89).  Its results are not those of a textbook Smith-Waterman in three observable ways, so this
 * restatement keeps the vectors and only replaces each SSE2 instruction by a loop over the lanes:
 *   - E(i+1,j) is computed from the H of the main loop, before the lazy-F loop corrects H across lane boundaries
 *     (ksw.c:166-170, 181: "we disallow adjacent insertion and then deletion");
 *   - the lazy-F loop stops as soon as no lane can improve (ksw.c:190), which matters when o_ins == 0;
 *   - the padding columns (>= qlen, profile score 0) take part in the row maximum that feeds the second-best list
 *     (ksw.c:196-205), and the byte kernel saturates at 255.
 * Parity: pinned against the compiled reference (oracle/_ref/libksw_ref.so, ksw_ref_align_batch) by
 * tests/test_align.py on fuzzed jobs in both score widths and every KSW_X* flag combination.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define XBYTE  0x10000   /* ksw.h:6-9 */
#define XSTOP  0x20000
#define XSUBO  0x40000
#define XSTART 0x80000

typedef struct { int32_t score, te, qe, score2, te2, tb, qb, pad; } aln_res_t;

static inline int imax2(int a, int b) { return a > b ? a : b; }
static inline int imin2(int a, int b) { return a < b ? a : b; }

/* one call of ksw_u8 (size 1) or ksw_i16 (size 2) including the profile of ksw_qinit */
/* closed: 0 = the lazy-F loop as the reference runs it; 1 = its closed form, the claim the GPU kernel rests on when o_ins >= 1
 * (H' = max(H, carry), carry into lane l = max over l' < l of f_end(l') - e_ins * slen * (l - l' - 1), decaying by e_ins per
 * column): tests/test_align.py checks on the CPU that both give the same seven outputs */
static aln_res_t striped_pass(int size, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                              int o_del, int e_del, int o_ins, int e_ins, int xtra, int closed)
{
	const int p = size == 1 ? 16 : 8;                       /* values per vector, ksw.c:67 */
	const int slen = (qlen + p - 1) / p, nlen = slen * p;   /* ksw.c:68 */
	const int vmask = size == 1 ? 0xff : 0xffff;
	uint8_t shift8 = 127, mdiff = 0;
	int a, i, j, l, qmax, shift;
	int *prof, *base, *H0, *H1, *E, *Hmax, *h, *f, *mx, *swap;
	int *sb_score = 0, *sb_row = 0;                         /* the second-best list: one (row maximum, row) entry per run of rows */
	int n_b = 0, te = -1, gmax = 0;
	const int minsc = (xtra & XSUBO) ? (xtra & 0xffff) : 0x10000;   /* ksw.c:127-128 */
	const int endsc = (xtra & XSTOP) ? (xtra & 0xffff) : 0x10000;
	/* _mm_set1_epi8 / _mm_set1_epi16 truncate the gap costs (ksw.c:131-134, 251-254) */
	const int oe_del_v = (o_del + e_del) & vmask, e_del_v = e_del & vmask, oe_ins_v = (o_ins + e_ins) & vmask, e_ins_v = e_ins & vmask;
	aln_res_t r = { 0, -1, -1, -1, -1, -1, -1, 0 };          /* g_defr, ksw.c:43 */

	for (a = 0; a < m * m; ++a) {                           /* ksw.c:77-80 */
		if (mat[a] < (int8_t)shift8) shift8 = (uint8_t)mat[a];
		if (mat[a] > (int8_t)mdiff) mdiff = (uint8_t)mat[a];
	}
	qmax = mdiff;
	shift8 = (uint8_t)(256 - shift8);
	shift = shift8;
	prof = (int *)malloc(sizeof(int) * (size_t)(m * nlen + 1));
	for (a = 0; a < m; ++a)                                  /* ksw.c:87-106 */
		for (j = 0; j < slen; ++j)
			for (l = 0; l < p; ++l) {
				const int k = j + l * slen;
				const int sc = k >= qlen ? 0 : mat[a * m + query[k]];
				prof[(a * slen + j) * p + l] = size == 1 ? (uint8_t)(int8_t)(sc + shift) : sc;
			}
	base = (int *)calloc((size_t)nlen * 4 + 3 * p + 4, sizeof(int));
	H0 = base; H1 = H0 + nlen; E = H1 + nlen; Hmax = E + nlen; h = Hmax + nlen; f = h + p; mx = f + p;

	for (i = 0; i < tlen; ++i) {
		const int *S = prof + (size_t)target[i] * nlen;
		int imax = 0, k, stop = 0;
		for (l = 0; l < p; ++l) { h[l] = l == 0 ? 0 : H0[(slen - 1) * p + l - 1]; f[l] = 0; mx[l] = 0; }   /* ksw.c:147-148 */
		for (j = 0; j < slen; ++j)
			for (l = 0; l < p; ++l) {
				int hv, e = E[j * p + l], t;
				if (size == 1) {                             /* ksw.c:156-174 */
					hv = imin2(255, h[l] + S[j * p + l]);
					hv = imax2(0, hv - shift);
				} else {                                     /* ksw.c:268 (signed saturating add) */
					hv = h[l] + S[j * p + l];
					hv = hv > 32767 ? 32767 : (hv < -32768 ? -32768 : hv);
				}
				hv = imax2(hv, e);
				hv = imax2(hv, f[l]);
				mx[l] = imax2(mx[l], hv);
				H1[j * p + l] = hv;
				e = imax2(0, e - e_del_v);
				t = imax2(0, hv - oe_del_v);
				E[j * p + l] = imax2(e, t);
				f[l] = imax2(0, f[l] - e_ins_v);
				t = imax2(0, hv - oe_ins_v);
				f[l] = imax2(f[l], t);
				h[l] = H0[j * p + l];
			}
		if (closed) {
			int g = 0;                                       /* carry into lane l */
			stop = 1;
			for (l = 0; l < p; ++l) {
				if (l > 0) g = imax2(imax2(0, g - e_ins_v * slen), f[l - 1]);
				for (j = 0; j < slen; ++j) H1[j * p + l] = imax2(H1[j * p + l], g - e_ins_v * j);
			}
		}
		for (k = 0; k < 16 && !stop; ++k) {                  /* the lazy-F loop, ksw.c:182-192 / 282-291 (16 rounds in both kernels) */
			for (l = p - 1; l > 0; --l) f[l] = f[l - 1];
			f[0] = 0;
			for (j = 0; j < slen; ++j) {
				int any = 0;
				for (l = 0; l < p; ++l) {
					int hv = imax2(H1[j * p + l], f[l]);
					H1[j * p + l] = hv;
					hv = imax2(0, hv - oe_ins_v);
					f[l] = imax2(0, f[l] - e_ins_v);
					if (f[l] > hv) any = 1;
				}
				if (!any) { stop = 1; break; }
			}
		}
		for (l = 0; l < p; ++l) imax = imax2(imax, mx[l]);
		if (imax >= minsc) {
			/* ksw.c:196-205: a row whose maximum reaches the threshold opens a new entry unless it directly follows the row
			 * the last entry names; in that case the entry is overwritten only by a larger maximum (and then names this row) */
			const int follows = n_b > 0 && sb_row[n_b - 1] + 1 == i;
			if (!follows) {
				if (!sb_score) { sb_score = (int *)malloc(sizeof(int) * (size_t)(tlen + 1)); sb_row = (int *)malloc(sizeof(int) * (size_t)(tlen + 1)); }
				sb_score[n_b] = imax; sb_row[n_b] = i; ++n_b;
			} else if (sb_score[n_b - 1] < imax) { sb_score[n_b - 1] = imax; sb_row[n_b - 1] = i; }
		}
		if (imax > gmax) {                                   /* ksw.c:206-211 */
			gmax = imax; te = i;
			memcpy(Hmax, H1, sizeof(int) * (size_t)nlen);
			if ((size == 1 && gmax + shift >= 255) || gmax >= endsc) break;
		}
		swap = H1; H1 = H0; H0 = swap;
	}
	r.score = size == 1 ? (gmax + shift < 255 ? gmax : 255) : gmax;   /* ksw.c:214, 309 */
	r.te = te;
	if (size == 2 || r.score != 255) {                       /* ksw.c:216-232, 310-326 */
		int max = -1;
		for (i = 0; i < nlen; ++i) {                         /* memory order: i = j*p + l holds column j + l*slen */
			const int col = i / p + i % p * slen;
			if (Hmax[i] > max) max = Hmax[i], r.qe = col;
			else if (Hmax[i] == max && col < r.qe) r.qe = col;
		}
		if (n_b > 0) {
			/* the best entry at least ceil(score / max) rows away from the end of the best hit; the first one wins ties */
			const int d = (r.score + qmax - 1) / qmax;
			for (i = 0; i < n_b; ++i)
				if ((sb_row[i] < te - d || sb_row[i] > te + d) && sb_score[i] > r.score2) r.score2 = sb_score[i], r.te2 = sb_row[i];
		}
	}
	free(sb_score); free(sb_row); free(prof); free(base);
	return r;
}

static void revseq(int l, uint8_t *s) { int i; for (i = 0; i < l >> 1; ++i) { const uint8_t t = s[i]; s[i] = s[l - 1 - i]; s[l - 1 - i] = t; } }

/* ksw_align2 with qry == NULL (ksw.c:329-354): out = {score, te, qe, score2, te2, tb, qb}.  Domain: qlen >= 1, the
 * matrix has a positive entry, and no overflow of the byte kernel (score 255), where the reference goes on with qe = -1. */
static int align2_impl(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                       int o_del, int e_del, int o_ins, int e_ins, int xtra, int32_t *out, int closed)
{
	const int size = (xtra & XBYTE) ? 1 : 2;
	aln_res_t r = striped_pass(size, qlen, query, tlen, target, m, mat, o_del, e_del, o_ins, e_ins, xtra, closed), rr;
	if (!((xtra & XSTART) == 0 || ((xtra & XSUBO) && r.score < (xtra & 0xffff))) && r.qe >= 0) {
		uint8_t *q = (uint8_t *)malloc((size_t)qlen + 1), *t = (uint8_t *)malloc((size_t)tlen + 1);
		memcpy(q, query, (size_t)qlen); memcpy(t, target, (size_t)tlen);
		revseq(r.qe + 1, q); revseq(r.te + 1, t);            /* ksw.c:343 */
		/* NB the second pass still runs over all tlen rows (ksw.c:345), the reversed prefix first */
		rr = striped_pass(size, r.qe + 1, q, tlen, t, m, mat, o_del, e_del, o_ins, e_ins, XSTOP | r.score, closed);
		if (r.score == rr.score) r.tb = r.te - rr.te, r.qb = r.qe - rr.qe;
		free(q); free(t);
	}
	out[0] = r.score; out[1] = r.te; out[2] = r.qe; out[3] = r.score2; out[4] = r.te2; out[5] = r.tb; out[6] = r.qb;
	return r.score;
}

int ksw_oracle_align2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                      int o_del, int e_del, int o_ins, int e_ins, int xtra, int32_t *out)
{
	return align2_impl(qlen, query, tlen, target, m, mat, o_del, e_del, o_ins, e_ins, xtra, out, 0);
}

/* the same with the lazy-F loop replaced by its closed form (valid for o_ins >= 1; see striped_pass) */
int ksw_oracle_align2_closed_form(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                                  int o_del, int e_del, int o_ins, int e_ins, int xtra, int32_t *out)
{
	return align2_impl(qlen, query, tlen, target, m, mat, o_del, e_del, o_ins, e_ins, xtra, out, 1);
}

/* ---- batch driver (same job / result records as include/ksw_b200.h: ksw_b200_ajob_t / ksw_b200_ares_t) ---- */
typedef struct { int8_t mat[25]; int32_t m, o_del, e_del, o_ins, e_ins, zdrop, end_bonus; } ocfg_t;
typedef struct { uint64_t q_off, t_off; int32_t qlen, tlen, xtra, pad; } ajob_t;
typedef struct { const ocfg_t *cfg; const ajob_t *jobs; const uint8_t *qpool, *tpool; aln_res_t *res; int64_t n, begin, stride; int closed; } aarg_t;

static void *aworker(void *p_)
{
	aarg_t *a = (aarg_t *)p_;
	const ocfg_t *c = a->cfg;
	int64_t k;
	for (k = a->begin; k < a->n; k += a->stride) {
		const ajob_t *j = &a->jobs[k];
		int32_t o[7];
		align2_impl(j->qlen, a->qpool + j->q_off, j->tlen, a->tpool + j->t_off, c->m, c->mat, c->o_del, c->e_del, c->o_ins,
		            c->e_ins, j->xtra, o, a->closed);
		memcpy(&a->res[k], o, sizeof(o));
		a->res[k].pad = 0;
	}
	return 0;
}

static int align_batch_impl(const ocfg_t *cfg, int64_t n, const ajob_t *jobs, const uint8_t *qpool, const uint8_t *tpool,
                            aln_res_t *res, int n_threads, int closed)
{
	int t;
	pthread_t *tid; aarg_t *args;
	if (n_threads < 1) n_threads = 1;
	tid = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
	args = (aarg_t *)malloc(sizeof(aarg_t) * n_threads);
	for (t = 0; t < n_threads; ++t) {
		aarg_t a = { cfg, jobs, qpool, tpool, res, n, t, n_threads, closed };
		args[t] = a;
		pthread_create(&tid[t], 0, aworker, &args[t]);
	}
	for (t = 0; t < n_threads; ++t) pthread_join(tid[t], 0);
	free(tid); free(args);
	return 0;
}

int ksw_oracle_align_batch(const ocfg_t *cfg, int64_t n, const ajob_t *jobs, const uint8_t *qpool, const uint8_t *tpool,
                           aln_res_t *res, int n_threads)
{
	return align_batch_impl(cfg, n, jobs, qpool, tpool, res, n_threads, 0);
}

int ksw_oracle_align_batch_closed_form(const ocfg_t *cfg, int64_t n, const ajob_t *jobs, const uint8_t *qpool, const uint8_t *tpool,
                                       aln_res_t *res, int n_threads)
{
	return align_batch_impl(cfg, n, jobs, qpool, tpool, res, n_threads, 1);
}
