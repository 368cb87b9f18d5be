/*
 * oracle/ref_shim.c — TEST INFRASTRUCTURE.  Batch driver around the REFERENCE's own
 * ksw_extend2 (bwa-0.7.8/ksw.c:379, compiled unmodified from /root/reference by
 * oracle/Makefile into oracle/_ref/libksw_ref.so).  Contains no algorithm: it only
 * loops over jobs (optionally on several pthreads) and calls the reference symbol.
 */
#include <stdint.h>
#include <stdlib.h>
#include <pthread.h>

/* prototype as declared in the reference header bwa-0.7.8/ksw.h:108 */
int ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m,
                const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int w,
                int end_bonus, int zdrop, int h0, int *qle, int *tle, int *gtle,
                int *gscore, int *max_off);

typedef struct { int32_t score, qle, tle, gtle, gscore, max_off; } ref_res_t;
typedef struct { uint64_t q_off, t_off; int32_t qlen, tlen, h0, w; } ref_job_t;
typedef struct { int8_t mat[25]; int32_t m, o_del, e_del, o_ins, e_ins, zdrop, end_bonus; } ref_cfg_t;

typedef struct {
	const ref_cfg_t *cfg; const ref_job_t *jobs; const uint8_t *qpool, *tpool;
	ref_res_t *res; int64_t n, begin, stride;
} arg_t;

static void *worker(void *p)
{
	arg_t *a = (arg_t *)p;
	const ref_cfg_t *c = a->cfg;
	int64_t k;
	for (k = a->begin; k < a->n; k += a->stride) {
		const ref_job_t *j = &a->jobs[k];
		ref_res_t *r = &a->res[k];
		r->score = ksw_extend2(j->qlen, a->qpool + j->q_off, j->tlen, a->tpool + j->t_off, c->m, c->mat,
		                       c->o_del, c->e_del, c->o_ins, c->e_ins, j->w, c->end_bonus, c->zdrop, j->h0,
		                       &r->qle, &r->tle, &r->gtle, &r->gscore, &r->max_off);
	}
	return 0;
}

int ksw_ref_extend_batch(const ref_cfg_t *cfg, int64_t n, const ref_job_t *jobs, const uint8_t *qpool,
                         const uint8_t *tpool, ref_res_t *res, int n_threads)
{
	int t;
	pthread_t *tid; arg_t *args;
	if (n_threads < 1) n_threads = 1;
	tid = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
	args = (arg_t *)malloc(sizeof(arg_t) * n_threads);
	for (t = 0; t < n_threads; ++t) {
		arg_t a = { cfg, jobs, qpool, tpool, res, n, t, n_threads };
		args[t] = a;
		pthread_create(&tid[t], 0, worker, &args[t]);
	}
	for (t = 0; t < n_threads; ++t) pthread_join(tid[t], 0);
	free(tid); free(args);
	return 0;
}

/* ---- the same for the reference's ksw_global2 (bwa-0.7.8/ksw.c:501, prototype ksw.h:84) ---- */
int ksw_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                int o_del, int e_del, int o_ins, int e_ins, int w, int *n_cigar, uint32_t **cigar);

typedef struct { uint64_t q_off, t_off; int32_t qlen, tlen, w, reserved; } ref_gjob_t;
typedef struct { int32_t score, n_cigar; int64_t cigar_off; } ref_gres_t;
typedef struct {
	const ref_cfg_t *cfg; const ref_gjob_t *jobs; const uint8_t *qpool, *tpool;
	ref_gres_t *res; uint32_t *cigar; int64_t n, begin, stride;
} garg_t;

static void *gworker(void *p)
{
	garg_t *a = (garg_t *)p;
	const ref_cfg_t *c = a->cfg;
	int64_t k;
	for (k = a->begin; k < a->n; k += a->stride) {
		const ref_gjob_t *j = &a->jobs[k];
		int nc = 0, x;
		uint32_t *cig = 0;
		a->res[k].score = ksw_global2(j->qlen, a->qpool + j->q_off, j->tlen, a->tpool + j->t_off, c->m, c->mat,
		                              c->o_del, c->e_del, c->o_ins, c->e_ins, j->w, &nc, &cig);
		a->res[k].n_cigar = nc;
		for (x = 0; x < nc; ++x) a->cigar[a->res[k].cigar_off + x] = cig[x];   /* copy out, then give the array back */
		free(cig);
	}
	return 0;
}

/* res[k].cigar_off must be preset by the caller to a slot with room for qlen + tlen + 2 operations */
int ksw_ref_global_batch(const ref_cfg_t *cfg, int64_t n, const ref_gjob_t *jobs, const uint8_t *qpool,
                         const uint8_t *tpool, ref_gres_t *res, uint32_t *cigar, int n_threads)
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

/* ---- the same for the reference's ksw_align2 (bwa-0.7.8/ksw.c:329, prototype and kswr_t ksw.h:30-36, 62) ---- */
#include <string.h>
typedef struct { int score; int te, qe; int score2, te2; int tb, qb; } ref_kswr_t;   /* layout of kswr_t, ksw.h:30-36 */
struct _kswq_t;
ref_kswr_t ksw_align2(int qlen, uint8_t *query, int tlen, uint8_t *target, int m, const int8_t *mat, int o_del, int e_del,
                      int o_ins, int e_ins, int xtra, struct _kswq_t **qry);

typedef struct { uint64_t q_off, t_off; int32_t qlen, tlen, xtra, pad; } ref_ajob_t;
typedef struct { int32_t score, te, qe, score2, te2, tb, qb, pad; } ref_ares_t;
typedef struct {
	const ref_cfg_t *cfg; const ref_ajob_t *jobs; const uint8_t *qpool, *tpool;
	ref_ares_t *res; int64_t n, begin, stride;
} aarg_t;

static void *aworker(void *p)
{
	aarg_t *a = (aarg_t *)p;
	const ref_cfg_t *c = a->cfg;
	int64_t k;
	for (k = a->begin; k < a->n; k += a->stride) {
		const ref_ajob_t *j = &a->jobs[k];
		/* the reference reverses both sequences in place and back (ksw.c:343-346): give it private copies */
		uint8_t *q = (uint8_t *)malloc((size_t)j->qlen + 16), *t = (uint8_t *)malloc((size_t)j->tlen + 16);
		ref_kswr_t r;
		memcpy(q, a->qpool + j->q_off, (size_t)j->qlen); memcpy(t, a->tpool + j->t_off, (size_t)j->tlen);
		r = ksw_align2(j->qlen, q, j->tlen, t, c->m, c->mat, c->o_del, c->e_del, c->o_ins, c->e_ins, j->xtra, 0);
		a->res[k].score = r.score; a->res[k].te = r.te; a->res[k].qe = r.qe; a->res[k].score2 = r.score2; a->res[k].te2 = r.te2;
		a->res[k].tb = r.tb; a->res[k].qb = r.qb; a->res[k].pad = 0;
		free(q); free(t);
	}
	return 0;
}

int ksw_ref_align_batch(const ref_cfg_t *cfg, int64_t n, const ref_ajob_t *jobs, const uint8_t *qpool,
                        const uint8_t *tpool, ref_ares_t *res, int n_threads)
{
	int t;
	pthread_t *tid; aarg_t *args;
	if (n_threads < 1) n_threads = 1;
	tid = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
	args = (aarg_t *)malloc(sizeof(aarg_t) * n_threads);
	for (t = 0; t < n_threads; ++t) {
		aarg_t a = { cfg, jobs, qpool, tpool, res, n, t, n_threads };
		args[t] = a;
		pthread_create(&tid[t], 0, aworker, &args[t]);
	}
	for (t = 0; t < n_threads; ++t) pthread_join(tid[t], 0);
	free(tid); free(args);
	return 0;
}
