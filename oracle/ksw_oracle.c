/*
 * oracle/ksw_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the banded affine-gap extension of BWA-MEM 0.7.8
 * (reference: bwa-0.7.8/ksw.c:379-476 `ksw_extend2`, wrapper `ksw_extend`
 * ksw.c:478-481).  It is used only as the checker by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
 * legs.  Nothing under bwa_mem_quickassist_b200/ may link, import or call it.
 *
 * Parity pinning: the reference ships no golden vectors for this path
 * (SURVEY.md §4).  This restatement is pinned by differential testing against
 * the reference's own ksw.c compiled unmodified into oracle/_ref/libksw_ref.so
 * (oracle/Makefile) and by the golden fixture tests/golden/ksw_extend_golden.npz
 * that was generated from that compiled reference (tests/golden/make_golden.py).
 *
 * The restatement is deliberately written differently from the reference
 * (two int arrays instead of an {h,e} struct array, no query profile, explicit
 * helper functions) so that an accidental shared bug is unlikely, but every
 * observable quirk of 0.7.8 is kept (each is tagged with the reference line).
 *
 * In addition to the six reference outputs it counts the DP cells the
 * reference loop visits (sum over executed rows of end-beg, ksw.c:418-421) and
 * the executed rows; these are the units of the GCUPS metric.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

typedef struct {
	int32_t score, qle, tle, gtle, gscore, max_off;
} oracle_res_t;

typedef struct {
	uint64_t q_off, t_off;      /* byte offsets into the query / target pools   */
	int32_t qlen, tlen, h0, w;  /* w is the caller's band (before the clamp)    */
} oracle_job_t;

typedef struct {
	int8_t mat[25];
	int32_t m;
	int32_t o_del, e_del, o_ins, e_ins;
	int32_t zdrop, end_bonus;
} oracle_cfg_t;

static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* ksw.c:398-406 — the band is narrowed to the longest gap the query could pay for.
 * The double division and the truncating cast are part of the contract. */
int ksw_oracle_clamp_w(int qlen, int m, const int8_t *mat, int o_del, int e_del,
                       int o_ins, int e_ins, int w, int end_bonus)
{
	int i, best = 0, lim;
	for (i = 0; i < m * m; ++i) best = imax(best, mat[i]);
	lim = (int)((double)(qlen * best + end_bonus - o_ins) / e_ins + 1.);
	w = imin(w, imax(lim, 1));
	lim = (int)((double)(qlen * best + end_bonus - o_del) / e_del + 1.);
	w = imin(w, imax(lim, 1));
	return w;
}

/* One extension.  H[j] plays eh[j].h (= H(i-1,j-1) when row i starts), E[j] plays
 * eh[j].e (= E(i,j)).  `cells`/`rows` may be NULL. */
int ksw_oracle_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                       int m, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
                       int w, int end_bonus, int zdrop, int h0,
                       int *qle, int *tle, int *gtle, int *gscore, int *max_off,
                       int64_t *cells, int32_t *rows)
{
	const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
	int32_t *H, *E;
	int i, j, lo, hi;
	int best, best_i, best_j, end_i, end_sc, off;
	int64_t ncell = 0;
	int32_t nrow = 0;

	if (h0 < 0) h0 = 0;                                           /* ksw.c:384 */
	H = (int32_t *)calloc((size_t)qlen + 1, sizeof(int32_t));
	E = (int32_t *)calloc((size_t)qlen + 1, sizeof(int32_t));
	/* row -1: a pure insertion ramp falling from h0 (ksw.c:394-396) */
	H[0] = h0;
	if (qlen >= 1) H[1] = h0 > oe_ins ? h0 - oe_ins : 0;
	for (j = 2; j <= qlen && H[j - 1] > e_ins; ++j) H[j] = H[j - 1] - e_ins;

	w = ksw_oracle_clamp_w(qlen, m, mat, o_del, e_del, o_ins, e_ins, w, end_bonus);

	best = h0; best_i = best_j = -1; end_i = -1; end_sc = -1; off = 0; /* ksw.c:408-410 */
	lo = 0; hi = qlen;
	for (i = 0; i < tlen; ++i) {
		const int8_t *srow = mat + (int)target[i] * m;
		int left = h0 - (o_del + e_del * (i + 1));  /* H(i,-1), used at column lo even when lo>0 (ksw.c:415-416,429) */
		int f = 0, rmax = 0, rarg = -1;
		if (left < 0) left = 0;
		lo = imax(lo, i - w);                                      /* ksw.c:418 */
		hi = imin(imin(hi, i + w + 1), qlen);                      /* ksw.c:419-420 */
		++nrow;
		if (hi > lo) ncell += hi - lo;
		for (j = lo; j < hi; ++j) {
			int diag = H[j], e = E[j], h, t;
			H[j] = left;                                           /* ksw.c:429 */
			h = diag + srow[query[j]];                             /* no "diag ? … : 0" guard in 0.7.8 (ksw.c:430) */
			h = imax(imax(h, e), f);
			left = h;
			if (!(rmax > h)) rarg = j;                             /* ties go to the last column (ksw.c:434) */
			rmax = imax(rmax, h);
			t = imax(h - oe_del, 0);
			E[j] = imax(e - e_del, t);                             /* E(i+1,j) */
			t = imax(h - oe_ins, 0);
			f = imax(f - e_ins, t);                                /* F(i,j+1) */
		}
		H[hi] = left; E[hi] = 0;                                   /* ksw.c:446 */
		/* ksw.c:447 tests the loop variable after the loop: when the row was empty it
		 * still equals lo, so the test is max(lo,hi)==qlen, not hi==qlen. */
		if ((hi > lo ? hi : lo) == qlen) {
			if (!(end_sc > left)) end_i = i;                       /* ties go to the last row (ksw.c:448) */
			end_sc = imax(end_sc, left);
		}
		if (rmax == 0) break;                                      /* ksw.c:451 */
		if (rmax > best) {                                         /* strict (ksw.c:452) */
			best = rmax; best_i = i; best_j = rarg;
			off = imax(off, abs(rarg - i));
		} else if (zdrop > 0) {                                    /* ksw.c:455-461 */
			int di = i - best_i, dj = rarg - best_j;
			if (di > dj) {
				if (best - rmax - (di - dj) * e_del > zdrop) break;
			} else {
				if (best - rmax - (dj - di) * e_ins > zdrop) break;
			}
		}
		/* ksw.c:463-466 — shrink to the non-zero run around the row maximum; hi may grow by one */
		for (j = rarg; j >= lo && H[j]; --j) ;
		lo = j + 1;
		for (j = rarg + 2; j <= hi && H[j]; ++j) ;
		hi = j;
	}
	free(H); free(E);
	if (qle) *qle = best_j + 1;
	if (tle) *tle = best_i + 1;
	if (gtle) *gtle = end_i + 1;
	if (gscore) *gscore = end_sc;
	if (max_off) *max_off = off;
	if (cells) *cells = ncell;
	if (rows) *rows = nrow;
	return best;
}

int ksw_oracle_extend(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                      int m, const int8_t *mat, int gapo, int gape, int w, int end_bonus,
                      int zdrop, int h0, int *qle, int *tle, int *gtle, int *gscore, int *max_off)
{
	return ksw_oracle_extend2(qlen, query, tlen, target, m, mat, gapo, gape, gapo, gape, w,
	                          end_bonus, zdrop, h0, qle, tle, gtle, gscore, max_off, 0, 0);
}

/* ---- batch driver (pthread over jobs) used for fixtures and the CPU baseline ---- */
typedef struct {
	const oracle_cfg_t *cfg;
	const oracle_job_t *jobs;
	const uint8_t *qpool, *tpool;
	oracle_res_t *res;
	int64_t *cells;
	int64_t n, begin, stride;
} batch_arg_t;

static void *batch_worker(void *p)
{
	batch_arg_t *a = (batch_arg_t *)p;
	const oracle_cfg_t *c = a->cfg;
	int64_t k;
	for (k = a->begin; k < a->n; k += a->stride) {
		const oracle_job_t *jb = &a->jobs[k];
		oracle_res_t *r = &a->res[k];
		int64_t nc = 0;
		r->score = ksw_oracle_extend2(jb->qlen, a->qpool + jb->q_off, jb->tlen, a->tpool + jb->t_off,
		                              c->m, c->mat, c->o_del, c->e_del, c->o_ins, c->e_ins, jb->w,
		                              c->end_bonus, c->zdrop, jb->h0, &r->qle, &r->tle, &r->gtle,
		                              &r->gscore, &r->max_off, &nc, 0);
		if (a->cells) a->cells[k] = nc;
	}
	return 0;
}

/* returns 0; `cells` (per job, may be NULL) receives the visited-cell counts */
int ksw_oracle_extend_batch(const oracle_cfg_t *cfg, int64_t n, const oracle_job_t *jobs,
                            const uint8_t *qpool, const uint8_t *tpool, oracle_res_t *res,
                            int64_t *cells, int n_threads)
{
	int t;
	pthread_t *tid;
	batch_arg_t *args;
	if (n_threads < 1) n_threads = 1;
	tid = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
	args = (batch_arg_t *)malloc(sizeof(batch_arg_t) * n_threads);
	for (t = 0; t < n_threads; ++t) {
		batch_arg_t a = { cfg, jobs, qpool, tpool, res, cells, n, t, n_threads };
		args[t] = a;
		pthread_create(&tid[t], 0, batch_worker, &args[t]);
	}
	for (t = 0; t < n_threads; ++t) pthread_join(tid[t], 0);
	free(tid); free(args);
	return 0;
}
