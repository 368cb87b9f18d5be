/* tests/emu/ext_stub.c — TEST INFRASTRUCTURE.  Lets the host-side extension driver (bwamem_ext.c: plan / passes /
 * replay and the rounds scheduler) be tested on a machine without a GPU: this stub provides the three ksw_b200 entry
 * points the driver calls, answering every batch with the CPU oracle.  It is linked only into tests/emu/libext_emu.so;
 * the product library never contains it. */
#include <stdint.h>
#include "../../include/ksw_b200.h"

int ksw_oracle_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                       int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                       int *qle, int *tle, int *gtle, int *gscore, int *max_off, int64_t *cells, int32_t *rows);
int ksw_oracle_clamp_w(int qlen, int m, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus);

int64_t ext_stub_jobs = 0, ext_stub_calls = 0;

int ksw_b200_extend_batch(ksw_b200_ctx_t *ctx, const ksw_b200_cfg_t *c, int64_t n, const ksw_b200_job_t *jobs,
                          const uint8_t *qpool, const uint8_t *tpool, ksw_b200_res_t *res)
{
	int64_t k;
	(void)ctx;
	++ext_stub_calls; ext_stub_jobs += n;
	for (k = 0; k < n; ++k) {
		const ksw_b200_job_t *j = &jobs[k];
		ksw_b200_res_t *r = &res[k];
		r->score = ksw_oracle_extend2(j->qlen, qpool + j->q_off, j->tlen, tpool + j->t_off, c->m, c->mat, c->o_del, c->e_del,
		                              c->o_ins, c->e_ins, j->w, c->end_bonus, c->zdrop, j->h0, &r->qle, &r->tle, &r->gtle,
		                              &r->gscore, &r->max_off, 0, 0);
	}
	return 0;
}

int ksw_b200_clamp_w(int qlen, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus)
{
	return ksw_oracle_clamp_w(qlen, 5, mat, o_del, e_del, o_ins, e_ins, w, end_bonus);
}

/* device-reference mode: the stub keeps the .pac pointer and slices it the way bns_get_seq does (bntseq.c:355-376) */
#include <stdlib.h>
static const uint8_t *stub_pac;
static int64_t stub_l_pac;
int ksw_b200_ref_set(ksw_b200_ctx_t *ctx, const uint8_t *pac, int64_t l_pac) { (void)ctx; stub_pac = pac; stub_l_pac = l_pac; return 0; }

int ksw_b200_extend_batch_ref(ksw_b200_ctx_t *ctx, const ksw_b200_cfg_t *c, int64_t n, const ksw_b200_rjob_t *jobs,
                              const uint8_t *qpool, size_t qpool_bytes, ksw_b200_res_t *res)
{
	int64_t k, i;
	(void)ctx; (void)qpool_bytes;
	++ext_stub_calls; ext_stub_jobs += n;
	for (k = 0; k < n; ++k) {
		const ksw_b200_rjob_t *j = &jobs[k];
		ksw_b200_res_t *r = &res[k];
		uint8_t *q = malloc(j->qlen > 0 ? j->qlen : 1), *t = malloc(j->tlen > 0 ? j->tlen : 1);
		for (i = 0; i < j->qlen; ++i) q[i] = qpool[(int64_t)j->q_off + j->q_step * i];
		for (i = 0; i < j->tlen; ++i) {
			const int64_t x = j->t_pos + j->t_step * i;
			const int64_t p = x >= stub_l_pac ? (stub_l_pac << 1) - 1 - x : x;
			const int b = stub_pac[p >> 2] >> ((~p & 3) << 1) & 3;
			t[i] = (uint8_t)(x >= stub_l_pac ? 3 - b : b);
		}
		r->score = ksw_oracle_extend2(j->qlen, q, j->tlen, t, c->m, c->mat, c->o_del, c->e_del, c->o_ins, c->e_ins, j->w,
		                              c->end_bonus, c->zdrop, j->h0, &r->qle, &r->tle, &r->gtle, &r->gscore, &r->max_off, 0, 0);
		free(q); free(t);
	}
	return 0;
}

int ksw_b200_queue_extend_ref(ksw_b200_queue_t *q, const ksw_b200_cfg_t *c, int64_t n, const ksw_b200_rjob_t *jobs,
                              const uint8_t *qpool, size_t qpool_bytes, ksw_b200_res_t *res)
{
	(void)q;
	return ksw_b200_extend_batch_ref(0, c, n, jobs, qpool, qpool_bytes, res);
}
