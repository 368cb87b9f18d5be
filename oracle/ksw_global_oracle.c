/*
 * oracle/ksw_global_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the banded global alignment with backtrace of BWA 0.7.8
 * (reference: bwa-0.7.8/ksw.c:501-584 `ksw_global2`, wrapper `ksw_global` ksw.c:586-589, CIGAR run-length
 * accumulation `push_cigar` ksw.c:486-499).  Checker for tests/ only.
 *
 * Parity pinning: the reference ships no golden vectors; this restatement is pinned by differential
 * testing against the reference's own ksw.c compiled unmodified (oracle/_ref/libksw_ref.so, driver ksw_ref_global_batch
 * in ref_shim.c) and by the fixture tests/golden/ksw_global_golden.npz generated from that compiled reference.
 *
 * Written differently from the reference on purpose: separate H and E rows instead of a struct array, no query
 * profile, three direction planes instead of one packed byte, and the CIGAR is built by counting runs first.  The
 * direction planes are addressed exactly like the reference's z (row i, column j at i*n_col + j - beg_i), because a
 * path is only defined by what it reads there.  Domain: the band must hold the end cell (|tlen - qlen| <= w), as
 * bwa_gen_cigar2 guarantees (bwa.c:124-126); outside it the reference reads cells it never wrote.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define G_NEG (-0x40000000)   /* MINUS_INF, ksw.c:36 */

typedef struct { uint64_t q_off, t_off; int32_t qlen, tlen, w, reserved; } oracle_gjob_t;
typedef struct { int32_t score, n_cigar; int64_t cigar_off; } oracle_gres_t;
typedef struct { int8_t mat[25]; int32_t m, o_del, e_del, o_ins, e_ins, zdrop, end_bonus; } oracle_gcfg_t;

/* cigar: room for qlen + tlen + 2 operations (len<<4|op; 0=M 1=I 2=D); returns the score, *n_cigar operations */
int ksw_oracle_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                       int o_del, int e_del, int o_ins, int e_ins, int w, int *n_cigar, uint32_t *cigar)
{
	const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
	const size_t zn = (size_t)n_col * (size_t)(tlen > 0 ? tlen : 0) + (size_t)qlen + 2;
	int32_t *H = (int32_t *)malloc(sizeof(int32_t) * (size_t)(qlen + 1));
	int32_t *E = (int32_t *)malloc(sizeof(int32_t) * (size_t)(qlen + 1));
	/* from which state each state of a cell is entered: dh 0 = diagonal, 1 = E, 2 = F; de 1 = E extended; df 1 = F extended */
	uint8_t *dh = (uint8_t *)calloc(zn, 1), *de = (uint8_t *)calloc(zn, 1), *df = (uint8_t *)calloc(zn, 1);
	int i, j, score;
	H[0] = 0; E[0] = G_NEG;                                                  /* ksw.c:520 */
	for (j = 1; j <= qlen; ++j) {
		H[j] = j <= w ? -(o_ins + e_ins * j) : G_NEG;                        /* ksw.c:521-523 */
		E[j] = G_NEG;
	}
	for (i = 0; i < tlen; ++i) {
		const int8_t *row = mat + (int)target[i] * m;
		const int beg = i > w ? i - w : 0, end = i + w + 1 < qlen ? i + w + 1 : qlen;
		int32_t left = beg == 0 ? -(o_del + e_del * (i + 1)) : G_NEG;        /* H(i, beg-1), ksw.c:533 */
		int32_t F = G_NEG;
		uint8_t *ph = dh + (size_t)i * n_col, *pe = de + (size_t)i * n_col, *pf = df + (size_t)i * n_col;
		for (j = beg; j < end; ++j) {
			const int32_t diag = H[j] + row[query[j]];                       /* M(i,j) */
			int32_t best = diag;
			uint8_t from = 0;
			if (E[j] > best) { best = E[j]; from = 1; }                       /* ties keep the diagonal (ksw.c:546) */
			if (F > best) { best = F; from = 2; }                             /* ties keep what was chosen (ksw.c:548) */
			H[j] = left; left = best;
			ph[j - beg] = from;
			/* E(i+1,j) and F(i,j+1): extending beats opening only when strictly better (ksw.c:552,556) */
			pe[j - beg] = (uint8_t)(E[j] - e_del > diag - (o_del + e_del));
			E[j] = pe[j - beg] ? E[j] - e_del : diag - (o_del + e_del);
			pf[j - beg] = (uint8_t)(F - e_ins > diag - (o_ins + e_ins));
			F = pf[j - beg] ? F - e_ins : diag - (o_ins + e_ins);
		}
		H[end] = left; E[end] = G_NEG;                                        /* ksw.c:558 */
	}
	score = H[qlen];
	if (n_cigar && cigar) {
		/* walk back from the last cell (ksw.c:565-573): state 0 = H, 1 = E (a deletion column), 2 = F (an insertion) */
		int pass, n_runs = 0;
		for (pass = 0; pass < 2; ++pass) {
			int state = 0, r = 0, cur = -1, len = 0, op, step;
			i = tlen - 1;
			j = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
			while (i >= 0 && j >= 0) {
				const size_t c = (size_t)i * n_col + (size_t)(j - (i > w ? i - w : 0));
				if (state == 0) state = dh[c];
				else if (state == 1) state = de[c] ? 1 : 0;
				else state = df[c] ? 2 : 0;
				if (state == 0) { op = 0; --i; --j; }
				else if (state == 1) { op = 2; --i; }
				else { op = 1; --j; }
				if (op != cur) {
					if (cur >= 0) { if (pass) cigar[n_runs - 1 - r] = (uint32_t)len << 4 | (uint32_t)cur; ++r; }
					cur = op; len = 0;
				}
				++len;
			}
			for (step = 0; step < 2; ++step) {                                /* ksw.c:574-575: leading D, then leading I */
				const int rest = step == 0 ? i + 1 : j + 1;
				op = step == 0 ? 2 : 1;
				if (rest <= 0) continue;
				if (op != cur) {
					if (cur >= 0) { if (pass) cigar[n_runs - 1 - r] = (uint32_t)len << 4 | (uint32_t)cur; ++r; }
					cur = op; len = 0;
				}
				len += rest;
			}
			if (cur >= 0) { if (pass) cigar[n_runs - 1 - r] = (uint32_t)len << 4 | (uint32_t)cur; ++r; }
			n_runs = r;
		}
		*n_cigar = n_runs;
	}
	free(H); free(E); free(dh); free(de); free(df);
	return score;
}

typedef struct {
	const oracle_gcfg_t *cfg; const oracle_gjob_t *jobs; const uint8_t *qpool, *tpool;
	oracle_gres_t *res; uint32_t *cigar; int64_t n, begin, stride;
} garg_t;

static void *gworker(void *p)
{
	garg_t *a = (garg_t *)p;
	const oracle_gcfg_t *c = a->cfg;
	int64_t k;
	for (k = a->begin; k < a->n; k += a->stride) {
		const oracle_gjob_t *j = &a->jobs[k];
		int nc = 0;
		a->res[k].score = ksw_oracle_global2(j->qlen, a->qpool + j->q_off, j->tlen, a->tpool + j->t_off, c->m, c->mat,
		                                     c->o_del, c->e_del, c->o_ins, c->e_ins, j->w, &nc, a->cigar + a->res[k].cigar_off);
		a->res[k].n_cigar = nc;
	}
	return 0;
}

/* res[k].cigar_off must be preset by the caller to a slot with room for qlen + tlen + 2 operations */
int ksw_oracle_global_batch(const oracle_gcfg_t *cfg, int64_t n, const oracle_gjob_t *jobs, const uint8_t *qpool,
                            const uint8_t *tpool, oracle_gres_t *res, uint32_t *cigar, int n_threads)
{
	int t;
	pthread_t *tid; garg_t *args;
	if (n_threads < 1) n_threads = 1;
	tid = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
	args = (garg_t *)malloc(sizeof(garg_t) * n_threads);
	for (t = 0; t < n_threads; ++t) {
		garg_t a = { cfg, jobs, qpool, tpool, res, cigar, n, t, n_threads };
		args[t] = a;
		pthread_create(&tid[t], 0, gworker, &args[t]);
	}
	for (t = 0; t < n_threads; ++t) pthread_join(tid[t], 0);
	free(tid); free(args);
	return 0;
}
