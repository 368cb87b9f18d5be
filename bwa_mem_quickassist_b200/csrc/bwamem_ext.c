/*
 * bwamem_ext.c — host C: the batched restructuring of BWA-MEM 0.7.8's mem_chain2aln
 * (bwa-0.7.8/bwamem.c:730-878) on top of the batched extension entry ksw_b200_extend_batch.
 * Interface and rationale: include/bwamem_b200.h.
 *
 * plan   -> every seed of every registered chain gets a left job (reversed query prefix against the
 *           reversed reference prefix, bwamem.c:813-817) and a right job (query suffix against the
 *           reference suffix, bwamem.c:844,854).  The read is stored once forward and once reversed,
 *           the chain's reference window once forward and once reversed, so a job is two offsets.
 * run    -> pass L, retries of L with a doubled band, pass R with h0 := left score, retries of R.
 * replay -> the reference's per-seed loop with the DP calls replaced by table look-ups.
 *
 * No DP is computed on the host in this file.
 */
#include <assert.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/bwamem_b200.h"

#define B200_MAX_BAND_TRY 2          /* MAX_BAND_TRY, bwamem.c:493 */

/* ---- growable arrays ------------------------------------------------------------------ */
#define VEC(type) struct { type *a; size_t n, m; }
#define vec_reserve(v, need)                                                        \
	do {                                                                            \
		if ((v).m < (size_t)(need)) {                                               \
			size_t m__ = (v).m ? (v).m : 64;                                        \
			while (m__ < (size_t)(need)) m__ += m__ >> 1;                           \
			(v).a = realloc((v).a, m__ * sizeof(*(v).a));                           \
			if (!(v).a) abort();                                                    \
			(v).m = m__;                                                            \
		}                                                                           \
	} while (0)
#define vec_push(v, x) do { vec_reserve(v, (v).n + 1); (v).a[(v).n++] = (x); } while (0)

typedef struct {
	int32_t l_query;
	uint64_t q_fwd, q_rev;             /* offsets of the read and of its reversal in qpool */
} read_rec_t;

typedef struct {
	int32_t read, n;                   /* owning read, number of seeds */
	size_t seed0;                      /* first seed / first ext record of this chain */
	int64_t rmax0, rmax1;              /* reference window (bwamem.c:740-755) */
	uint64_t r_fwd, r_rev;             /* offsets of the window and of its reversal in tpool */
} chain_rec_t;

typedef struct {
	int32_t has_left, has_right;
	int32_t aw_left, aw_right;         /* band actually used (aw[0], aw[1], bwamem.c:819,847) */
	ksw_b200_res_t left, right;
} ext_rec_t;

/* rounds mode (SURVEY.md 7.3-3 strategy B): the timeline of a read = its chains and the regions that other code
 * (mem_chain2aln_short in the reference flow) inserted between them, in order */
typedef struct {
	int32_t chain;                     /* >= 0: a registered chain; -1: a fixed region */
	b200_alnreg_t reg;
} item_rec_t;

typedef struct {
	size_t item, item_end;             /* current / one-past-last timeline item of the read */
	int32_t k;                         /* position in the sorted seed order of the current chain; -1: chain not started */
	int32_t phase;                     /* 0: decide on seed k, 1: left result pending, 2: right result pending */
	int32_t attempt;                   /* band try (0 or 1) */
	int32_t aw0, aw1, sc0;
	size_t cur;                        /* index of the region under construction in av */
	uint64_t *order;                   /* sorted seeds of the current chain */
	ksw_b200_res_t res;                /* the result that just arrived */
	b200_alnreg_v av;
} read_state_t;

struct b200_ext_plan {
	b200_ext_opt_t opt;
	int64_t l_pac;
	const uint8_t *pac;
	VEC(read_rec_t) reads;
	VEC(chain_rec_t) chains;
	VEC(b200_seed_t) seeds;
	VEC(ext_rec_t) ext;
	VEC(uint8_t) qpool, tpool;
	/* scratch of run() */
	VEC(ksw_b200_job_t) jobs;
	VEC(ksw_b200_rjob_t) rjobs;         /* device-reference mode: what is actually submitted */
	VEC(ksw_b200_res_t) res;
	VEC(uint32_t) owner;               /* job -> ext record (speculative mode) / read (rounds mode) */
	VEC(item_rec_t) items;             /* rounds mode */
	VEC(size_t) read_item0;            /* first timeline item of every read (+ sentinel) */
	VEC(read_state_t) rstate;
	int64_t st_seeds, st_left, st_right, st_retry, st_rounds;
	/* device-reference mode: the chain windows are not materialised here; a job's t_off holds the
	 * doubled-space coordinate of its first target base (bit 63: the run goes downwards) and the GPU slices the resident
	 * .pac (ksw_b200_extend_batch_ref) */
	int ref_mode, ref_ready;
	ksw_b200_queue_t *queue;           /* device-reference mode: submit through the GPU's shared queue instead of a private context */
};

#define REF_DOWN (1ull << 63)

/* ---- small pieces of the reference logic ---------------------------------------------- */

/* cal_max_gap, bwamem.c:544-551 (double division and truncation are part of the contract) */
static int max_gap_for(const b200_ext_opt_t *o, int qlen)
{
	int l_del = (int)((double)(qlen * o->a - o->o_del) / o->e_del + 1.);
	int l_ins = (int)((double)(qlen * o->a - o->o_ins) / o->e_ins + 1.);
	int l = l_del > l_ins ? l_del : l_ins;
	if (l < 1) l = 1;
	return l < o->w << 1 ? l : o->w << 1;
}

static inline int pac_base(const uint8_t *pac, int64_t k)        /* _get_pac, bntseq.c:192 */
{
	return pac[k >> 2] >> ((~k & 3) << 1) & 3;
}

int64_t b200_get_ref_slice(int64_t l_pac, const uint8_t *pac, int64_t beg, int64_t end, uint8_t *out)
{
	int64_t k, n = 0;
	if (end < beg) { int64_t t = beg; beg = end; end = t; }
	if (end > l_pac << 1) end = l_pac << 1;
	if (beg < 0) beg = 0;
	if (!(beg >= l_pac || end <= l_pac)) return 0;               /* bridges the strand boundary: nothing */
	if (beg >= l_pac) {                                          /* reverse strand: complement, read backwards */
		const int64_t hi = (l_pac << 1) - 1 - beg, lo = (l_pac << 1) - 1 - end;
		for (k = hi; k > lo; --k) out[n++] = (uint8_t)(3 - pac_base(pac, k));
	} else {
		for (k = beg; k < end; ++k) out[n++] = (uint8_t)pac_base(pac, k);
	}
	return n;
}

static int cmp_u64(const void *a, const void *b)
{
	const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
	return x < y ? -1 : x > y;
}

/* ---- plan -------------------------------------------------------------------------------- */

b200_ext_plan_t *b200_ext_plan_create(const b200_ext_opt_t *opt, int64_t l_pac, const uint8_t *pac)
{
	b200_ext_plan_t *p = calloc(1, sizeof(*p));
	if (!p) return 0;
	p->opt = *opt; p->l_pac = l_pac; p->pac = pac;
	{
		const char *e = getenv("KSW_B200_REF");                  /* default of the mode; b200_ext_plan_set_device_ref overrides */
		p->ref_mode = e && e[0] == '1';
	}
	return p;
}

void b200_ext_plan_set_queue(b200_ext_plan_t *p, ksw_b200_queue_t *q) { p->queue = q; }

void b200_ext_plan_set_device_ref(b200_ext_plan_t *p, int on)
{
	if (p->chains.n == 0) { p->ref_mode = on != 0; p->ref_ready = 0; }       /* only between batches */
}

void b200_ext_plan_destroy(b200_ext_plan_t *p)
{
	if (!p) return;
	free(p->reads.a); free(p->chains.a); free(p->seeds.a); free(p->ext.a);
	free(p->qpool.a); free(p->tpool.a); free(p->jobs.a); free(p->rjobs.a); free(p->res.a); free(p->owner.a);
	free(p->items.a); free(p->read_item0.a);
	if (p->rstate.a) {
		size_t r;
		for (r = 0; r < p->rstate.n; ++r) { free(p->rstate.a[r].order); free(p->rstate.a[r].av.a); }
		free(p->rstate.a);
	}
	free(p);
}

void b200_ext_plan_reset(b200_ext_plan_t *p)
{
	p->reads.n = p->chains.n = p->seeds.n = p->ext.n = 0;
	p->qpool.n = p->tpool.n = 0;
	p->items.n = p->read_item0.n = 0;
	{
		size_t r;
		for (r = 0; r < p->rstate.n; ++r) { free(p->rstate.a[r].order); free(p->rstate.a[r].av.a); }
		p->rstate.n = 0;
	}
	p->st_seeds = p->st_left = p->st_right = p->st_retry = p->st_rounds = 0;
}

int b200_ext_plan_add_read(b200_ext_plan_t *p, int l_query, const uint8_t *query)
{
	read_rec_t r;
	int i;
	r.l_query = l_query;
	vec_reserve(p->qpool, p->qpool.n + 2 * (size_t)l_query);
	r.q_fwd = p->qpool.n;
	memcpy(p->qpool.a + p->qpool.n, query, (size_t)l_query);
	p->qpool.n += (size_t)l_query;
	r.q_rev = p->qpool.n;
	for (i = 0; i < l_query; ++i) p->qpool.a[p->qpool.n + i] = query[l_query - 1 - i];
	p->qpool.n += (size_t)l_query;
	vec_push(p->reads, r);
	vec_push(p->read_item0, p->items.n);
	return (int)p->reads.n - 1;
}

int b200_ext_plan_add_chain(b200_ext_plan_t *p, int read, const b200_chain_t *c)
{
	const b200_ext_opt_t *o = &p->opt;
	const int l_query = p->reads.a[read].l_query;
	chain_rec_t ch;
	int64_t lo, hi, rlen, k;
	int i;
	if (c->n == 0) return -1;
	/* the widest reference span any seed of the chain could reach (bwamem.c:740-750) */
	lo = p->l_pac << 1; hi = 0;
	for (i = 0; i < c->n; ++i) {
		const b200_seed_t *t = &c->seeds[i];
		const int64_t b = t->rbeg - (t->qbeg + max_gap_for(o, t->qbeg));
		const int tail = l_query - t->qbeg - t->len;
		const int64_t e = t->rbeg + t->len + (tail + max_gap_for(o, tail));
		if (b < lo) lo = b;
		if (e > hi) hi = e;
	}
	if (lo < 0) lo = 0;
	if (hi > p->l_pac << 1) hi = p->l_pac << 1;
	if (lo < p->l_pac && p->l_pac < hi) {                        /* keep the strand of the seeds (bwamem.c:752-755) */
		if (c->seeds[0].rbeg < p->l_pac) hi = p->l_pac;
		else lo = p->l_pac;
	}
	ch.read = read; ch.n = c->n; ch.seed0 = p->seeds.n; ch.rmax0 = lo; ch.rmax1 = hi;
	if (p->ref_mode) {
		/* the GPU slices the resident .pac: nothing to materialise (the window never bridges the strands, see above) */
		ch.r_fwd = ch.r_rev = 0;
		(void)rlen; (void)k;
	} else {
		/* the window, forward and reversed (bns_get_seq, bwamem.c:757) */
		vec_reserve(p->tpool, p->tpool.n + 2 * (size_t)(hi - lo));
		ch.r_fwd = p->tpool.n;
		rlen = b200_get_ref_slice(p->l_pac, p->pac, lo, hi, p->tpool.a + p->tpool.n);
		assert(rlen == hi - lo);
		p->tpool.n += (size_t)rlen;
		ch.r_rev = p->tpool.n;
		for (k = 0; k < rlen; ++k) p->tpool.a[ch.r_rev + k] = p->tpool.a[ch.r_fwd + (rlen - 1 - k)];
		p->tpool.n += (size_t)rlen;
	}
	/* seeds and their (still empty) extension records */
	vec_reserve(p->seeds, p->seeds.n + (size_t)c->n);
	vec_reserve(p->ext, p->ext.n + (size_t)c->n);
	for (i = 0; i < c->n; ++i) {
		const b200_seed_t *s = &c->seeds[i];
		ext_rec_t x;
		memset(&x, 0, sizeof(x));
		x.has_left = s->qbeg != 0;                               /* bwamem.c:810 */
		x.has_right = s->qbeg + s->len != l_query;               /* bwamem.c:841 */
		x.aw_left = x.aw_right = o->w;
		p->seeds.a[p->seeds.n++] = *s;
		p->ext.a[p->ext.n++] = x;
	}
	vec_push(p->chains, ch);
	p->st_seeds += c->n;
	{
		item_rec_t it;
		memset(&it, 0, sizeof(it));
		it.chain = (int32_t)p->chains.n - 1;
		assert(read == (int)p->reads.n - 1);                    /* the timeline of a read is built before the next read */
		vec_push(p->items, it);
	}
	return (int)p->chains.n - 1;
}

/* ---- run --------------------------------------------------------------------------------- */

static void cfg_from_opt(const b200_ext_opt_t *o, int end_bonus, ksw_b200_cfg_t *cfg)
{
	memcpy(cfg->mat, o->mat, 25);
	cfg->m = 5;
	cfg->o_del = o->o_del; cfg->e_del = o->e_del; cfg->o_ins = o->o_ins; cfg->e_ins = o->e_ins;
	cfg->zdrop = o->zdrop; cfg->end_bonus = end_bonus;
}

/* the left job of seed s of chain ch: qs[i] = query[qbeg-1-i], rs[i] = rseq[tmp-1-i] (bwamem.c:813-817) */
static void left_job(const b200_ext_plan_t *p, const chain_rec_t *ch, const b200_seed_t *s, int w, ksw_b200_job_t *j)
{
	const read_rec_t *rd = &p->reads.a[ch->read];
	const int64_t rlen = ch->rmax1 - ch->rmax0, tmp = s->rbeg - ch->rmax0;
	j->q_off = rd->q_rev + (uint64_t)(rd->l_query - s->qbeg);
	j->qlen = s->qbeg;
	/* rs[i] = rseq[tmp-1-i] = reference base (s->rbeg - 1 - i) of the doubled space */
	j->t_off = p->ref_mode ? ((uint64_t)(s->rbeg - 1) | REF_DOWN) : ch->r_rev + (uint64_t)(rlen - tmp);
	j->tlen = (int32_t)tmp;
	j->h0 = s->len * p->opt.a;
	j->w = w;
}

/* the right job: query + qe against rseq + re (bwamem.c:843-844,854) */
static void right_job(const b200_ext_plan_t *p, const chain_rec_t *ch, const b200_seed_t *s, int h0, int w, ksw_b200_job_t *j)
{
	const read_rec_t *rd = &p->reads.a[ch->read];
	const int qe = s->qbeg + s->len;
	const int64_t re = s->rbeg + s->len - ch->rmax0;
	assert(re >= 0);
	j->q_off = rd->q_fwd + (uint64_t)qe;
	j->qlen = rd->l_query - qe;
	j->t_off = p->ref_mode ? (uint64_t)(s->rbeg + s->len) : ch->r_fwd + (uint64_t)re;
	j->tlen = (int32_t)(ch->rmax1 - ch->rmax0 - re);
	j->h0 = h0;
	j->w = w;
}

/* Trace facility (the batched counterpart of the reference's `-v 4` extension trace, bwamem.c:821-828): with
 * KSW_B200_DUMP=<prefix> every pass appends its job batch to <prefix>.<pid>.bin as
 * [magic "KSWJ"][cfg][n][qpool bytes][tpool bytes][jobs][qpool][tpool] for offline replay (scripts/bench_jobs.py). */
#include <stdio.h>
#include <unistd.h>
#include <pthread.h>
static pthread_mutex_t dump_mu = PTHREAD_MUTEX_INITIALIZER;     /* the worker threads of one process share the file */
static void dump_jobs(const b200_ext_plan_t *p, const ksw_b200_cfg_t *cfg)
{
	const char *pre = getenv("KSW_B200_DUMP");
	char path[4096];
	FILE *f;
	uint64_t hdr[3];
	if (!pre || !*pre) return;
	snprintf(path, sizeof(path), "%s.%d.bin", pre, (int)getpid());
	pthread_mutex_lock(&dump_mu);
	f = fopen(path, "ab");
	if (!f) { pthread_mutex_unlock(&dump_mu); return; }
	hdr[0] = p->jobs.n; hdr[1] = p->qpool.n; hdr[2] = p->tpool.n;
	fwrite("KSWJ", 1, 4, f);
	fwrite(cfg, sizeof(*cfg), 1, f);
	fwrite(hdr, sizeof(hdr), 1, f);
	fwrite(p->jobs.a, sizeof(ksw_b200_job_t), p->jobs.n, f);
	fwrite(p->qpool.a, 1, p->qpool.n, f);
	fwrite(p->tpool.a, 1, p->tpool.n, f);
	fclose(f);
	pthread_mutex_unlock(&dump_mu);
}

/* submits p->jobs[0..n) and waits for p->res[0..n) */
static int submit_jobs(b200_ext_plan_t *p, ksw_b200_ctx_t *ctx, const ksw_b200_cfg_t *cfg, size_t n)
{
	size_t k;
	vec_reserve(p->res, n);
	if (!p->ref_mode) {
		dump_jobs(p, cfg);
		return ksw_b200_extend_batch(ctx, cfg, (int64_t)n, p->jobs.a, p->qpool.a, p->tpool.a, p->res.a);
	}
	if (!p->ref_ready && !p->queue) {
		const int rc = ksw_b200_ref_set(ctx, p->pac, p->l_pac);  /* shared per device; a no-op after the first context */
		if (rc) return rc;
		p->ref_ready = 1;
	}
	vec_reserve(p->rjobs, n);
	for (k = 0; k < n; ++k) {
		const ksw_b200_job_t *j = &p->jobs.a[k];
		ksw_b200_rjob_t *r = &p->rjobs.a[k];
		memset(r, 0, sizeof(*r));
		r->q_off = j->q_off; r->q_step = 1;                       /* the read is stored forward and reversed */
		r->t_pos = (int64_t)(j->t_off & ~REF_DOWN); r->t_step = (j->t_off & REF_DOWN) ? -1 : 1;
		r->qlen = j->qlen; r->tlen = j->tlen; r->h0 = j->h0; r->w = j->w;
	}
	if (p->queue)                                                /* the queue's owner has put the reference on its device */
		return ksw_b200_queue_extend_ref(p->queue, cfg, (int64_t)n, p->rjobs.a, p->qpool.a, p->qpool.n, p->res.a);
	return ksw_b200_extend_batch_ref(ctx, cfg, (int64_t)n, p->rjobs.a, p->qpool.a, p->qpool.n, p->res.a);
}

static int run_jobs(b200_ext_plan_t *p, ksw_b200_ctx_t *ctx, const ksw_b200_cfg_t *cfg)
{
	if (p->jobs.n == 0) return 0;
	return submit_jobs(p, ctx, cfg, p->jobs.n);
}

int b200_ext_plan_run(b200_ext_plan_t *p, ksw_b200_ctx_t *ctx)
{
	const b200_ext_opt_t *o = &p->opt;
	ksw_b200_cfg_t cfg;
	size_t c, k;
	int side, rc;
	for (side = 0; side < 2; ++side) {                            /* 0: left pass, 1: right pass */
		int attempt;
		cfg_from_opt(o, side == 0 ? o->pen_clip5 : o->pen_clip3, &cfg);
		for (attempt = 0; attempt < B200_MAX_BAND_TRY; ++attempt) {
			const int w = o->w << attempt;
			p->jobs.n = p->owner.n = 0;
			for (c = 0; c < p->chains.n; ++c) {
				const chain_rec_t *ch = &p->chains.a[c];
				int i;
				for (i = 0; i < ch->n; ++i) {
					const b200_seed_t *s = &p->seeds.a[ch->seed0 + i];
					ext_rec_t *x = &p->ext.a[ch->seed0 + i];
					ksw_b200_job_t j;
					if (side == 0) {
						if (!x->has_left) continue;
						/* retry only if the best cell sat in the outer quarter of the band (bwamem.c:828);
						 * "score == prev" cannot hold on the first try because prev is -1 */
						if (attempt > 0 && x->left.max_off < (x->aw_left >> 1) + (x->aw_left >> 2)) continue;
						left_job(p, ch, s, w, &j);
					} else {
						const int sc0 = x->has_left ? x->left.score : s->len * o->a;   /* bwamem.c:839,842 */
						if (!x->has_right) continue;
						if (attempt > 0 && (x->right.score == sc0 ||                      /* bwamem.c:856 */
						                    x->right.max_off < (x->aw_right >> 1) + (x->aw_right >> 2))) continue;
						right_job(p, ch, s, sc0, w, &j);
					}
					vec_push(p->jobs, j);
					vec_push(p->owner, (uint32_t)(ch->seed0 + i));
				}
			}
			rc = run_jobs(p, ctx, &cfg);
			if (rc) return rc;
			for (k = 0; k < p->jobs.n; ++k) {
				ext_rec_t *x = &p->ext.a[p->owner.a[k]];
				if (side == 0) { x->left = p->res.a[k]; x->aw_left = w; }
				else { x->right = p->res.a[k]; x->aw_right = w; }
			}
			if (attempt == 0) { if (side == 0) p->st_left += (int64_t)p->jobs.n; else p->st_right += (int64_t)p->jobs.n; }
			else p->st_retry += (int64_t)p->jobs.n;
		}
	}
	return 0;
}

/* ---- replay -------------------------------------------------------------------------------- */

static b200_alnreg_t *av_push(b200_alnreg_v *av)                /* kv_pushp, kvec.h:83 */
{
	if (av->n == av->m) {
		av->m = av->m ? av->m << 1 : 2;
		av->a = realloc(av->a, sizeof(b200_alnreg_t) * av->m);
		if (!av->a) abort();
	}
	return &av->a[av->n++];
}

/* bwamem.c:769-799: 1 if seed order[k] needs no extension (it lies inside, and near the diagonal of, a region found
 * earlier, and no longer overlapping seed of the chain sits on another diagonal) */
static int seed_is_covered(const b200_ext_opt_t *o, const b200_alnreg_v *av, const b200_seed_t *seeds, int n,
                           const uint64_t *order, int k)
{
	const b200_seed_t *s = &seeds[(uint32_t)order[k]];
	size_t i;
	int j;
	for (i = 0; i < av->n; ++i) {
		const b200_alnreg_t *r = &av->a[i];
		int64_t rd;
		int qd, w, g;
		if (s->rbeg < r->rb || s->rbeg + s->len > r->re || s->qbeg < r->qb || s->qbeg + s->len > r->qe) continue;
		qd = s->qbeg - r->qb; rd = s->rbeg - r->rb;
		g = max_gap_for(o, qd < rd ? qd : (int)rd);
		w = g < o->w ? g : o->w;
		if (qd - rd < w && rd - qd < w) break;
		qd = r->qe - (s->qbeg + s->len); rd = r->re - (s->rbeg + s->len);
		g = max_gap_for(o, qd < rd ? qd : (int)rd);
		w = g < o->w ? g : o->w;
		if (qd - rd < w && rd - qd < w) break;
	}
	if (i == av->n) return 0;
	for (j = k + 1; j < n; ++j) {
		const b200_seed_t *t;
		if (order[j] == 0) continue;
		t = &seeds[(uint32_t)order[j]];
		if (t->len < s->len * .95) continue;
		if (s->qbeg <= t->qbeg && s->qbeg + s->len - t->qbeg >= s->len >> 2 && t->qbeg - s->qbeg != t->rbeg - s->rbeg) return 0;
		if (t->qbeg <= s->qbeg && t->qbeg + t->len - s->qbeg >= s->len >> 2 && s->qbeg - t->qbeg != s->rbeg - t->rbeg) return 0;
	}
	return 1;
}

static void region_begin(const b200_ext_opt_t *o, b200_alnreg_t *a)       /* bwamem.c:805-808 */
{
	memset(a, 0, sizeof(*a));
	a->w = o->w;
	a->score = a->truesc = -1;
}

static void region_left(const b200_ext_opt_t *o, b200_alnreg_t *a, const b200_seed_t *s, const ksw_b200_res_t *L)
{
	if (L) {                                                     /* bwamem.c:830-837 */
		a->score = L->score;
		if (L->gscore <= 0 || L->gscore <= a->score - o->pen_clip5) {   /* local end */
			a->qb = s->qbeg - L->qle; a->rb = s->rbeg - L->tle;
			a->truesc = a->score;
		} else {                                                 /* reaches the query start */
			a->qb = 0; a->rb = s->rbeg - L->gtle;
			a->truesc = L->gscore;
		}
	} else {                                                     /* bwamem.c:839 */
		a->score = a->truesc = s->len * o->a; a->qb = 0; a->rb = s->rbeg;
	}
}

static void region_right(const b200_ext_opt_t *o, b200_alnreg_t *a, const b200_seed_t *s, const ksw_b200_res_t *R,
                         int l_query, int64_t rmax0)
{
	if (R) {                                                     /* bwamem.c:858-865 */
		const int sc0 = a->score, qe = s->qbeg + s->len;
		const int64_t re = s->rbeg + s->len - rmax0;
		a->score = R->score;
		if (R->gscore <= 0 || R->gscore <= a->score - o->pen_clip3) {
			a->qe = qe + R->qle; a->re = rmax0 + re + R->tle;
			a->truesc += a->score - sc0;
		} else {
			a->qe = l_query; a->re = rmax0 + re + R->gtle;
			a->truesc += R->gscore - sc0;
		}
	} else {                                                     /* bwamem.c:867 */
		a->qe = l_query; a->re = s->rbeg + s->len;
	}
}

static void region_end(b200_alnreg_t *a, const b200_seed_t *seeds, int n, int aw0, int aw1)   /* bwamem.c:870-875 */
{
	int j;
	a->seedcov = 0;
	for (j = 0; j < n; ++j) {
		const b200_seed_t *t = &seeds[j];
		if (t->qbeg >= a->qb && t->qbeg + t->len <= a->qe && t->rbeg >= a->rb && t->rbeg + t->len <= a->re)
			a->seedcov += t->len;
	}
	a->w = aw0 > aw1 ? aw0 : aw1;
}

static uint64_t *sorted_seed_order(const b200_seed_t *seeds, int n)       /* bwamem.c:760-763 */
{
	uint64_t *order = malloc((size_t)(n > 0 ? n : 1) * 8);
	int k;
	for (k = 0; k < n; ++k) order[k] = (uint64_t)seeds[k].len << 32 | (uint32_t)k;
	qsort(order, (size_t)n, 8, cmp_u64);
	return order;
}

void b200_ext_replay_chain(const b200_ext_plan_t *p, int chain, b200_alnreg_v *av)
{
	const b200_ext_opt_t *o = &p->opt;
	const chain_rec_t *ch;
	const b200_seed_t *seeds;
	const ext_rec_t *ext;
	uint64_t *order;
	int l_query, k, n;
	if (chain < 0) return;
	ch = &p->chains.a[chain];
	seeds = p->seeds.a + ch->seed0;
	ext = p->ext.a + ch->seed0;
	n = ch->n;
	l_query = p->reads.a[ch->read].l_query;
	order = sorted_seed_order(seeds, n);                          /* longest seed first; equal lengths in index order */
	for (k = n - 1; k >= 0; --k) {
		const int si = (int)(uint32_t)order[k];
		const b200_seed_t *s = &seeds[si];
		const ext_rec_t *x = &ext[si];
		b200_alnreg_t *a;
		if (seed_is_covered(o, av, seeds, n, order, k)) { order[k] = 0; continue; }   /* skipped: marked like srt[k] = 0 */
		a = av_push(av);
		region_begin(o, a);
		region_left(o, a, s, x->has_left ? &x->left : 0);
		region_right(o, a, s, x->has_right ? &x->right : 0, l_query, ch->rmax0);
		region_end(a, seeds, n, x->has_left ? x->aw_left : o->w, x->has_right ? x->aw_right : o->w);
	}
	free(order);
}

/* ---- rounds mode: exact, minimal-work scheduling (SURVEY.md 7.3-3 strategy B) ------------------------------------ *
 * Every read walks its own timeline exactly like the reference's sequential code and stops whenever it needs a DP
 * result; one round = one batched GPU call with the pending job of every read that is still active (left and right
 * jobs mixed; the band clamp of each job is computed here with its own end bonus).  Only the seeds the reference would
 * extend are extended.                                                                                                  */

int b200_ext_plan_add_region(b200_ext_plan_t *p, int read, const b200_alnreg_t *reg)
{
	item_rec_t it;
	assert(read == (int)p->reads.n - 1);
	it.chain = -1; it.reg = *reg;
	vec_push(p->items, it);
	return 0;
}

/* advances read r until it needs a job (returns 1 and fills *j) or its timeline is finished (returns 0) */
static int read_advance(b200_ext_plan_t *p, read_state_t *st, ksw_b200_job_t *j)
{
	const b200_ext_opt_t *o = &p->opt;
	for (;;) {
		const item_rec_t *it;
		const chain_rec_t *ch;
		const b200_seed_t *seeds, *s;
		b200_alnreg_t *a;
		int n;
		if (st->item >= st->item_end) return 0;
		it = &p->items.a[st->item];
		if (it->chain < 0) { *av_push(&st->av) = it->reg; ++st->item; continue; }
		ch = &p->chains.a[it->chain];
		seeds = p->seeds.a + ch->seed0; n = ch->n;
		if (!st->order) { st->order = sorted_seed_order(seeds, n); st->k = n - 1; }        /* entering the chain */
		if (st->k < 0) { free(st->order); st->order = 0; ++st->item; continue; }           /* all its seeds are done */
		s = &seeds[(uint32_t)st->order[st->k]];
		if (st->phase == 0) {
			if (seed_is_covered(o, &st->av, seeds, n, st->order, st->k)) {
				st->order[st->k] = 0;                                  /* skipped: marked like srt[k] = 0 */
				--st->k;
				continue;
			}
			st->cur = st->av.n;
			a = av_push(&st->av);
			region_begin(o, a);
			st->aw0 = st->aw1 = o->w; st->attempt = 0;
			if (s->qbeg) {                                            /* left extension needed */
				left_job(p, ch, s, o->w, j);
				j->w = ksw_b200_clamp_w(j->qlen, o->mat, o->o_del, o->e_del, o->o_ins, o->e_ins, o->w, o->pen_clip5);
				st->phase = 1;
				return 1;
			}
			region_left(o, a, s, 0);
			st->phase = 3;                                            /* fall through to the right side */
		}
		a = &st->av.a[st->cur];
		if (st->phase == 1) {                                         /* a left result arrived */
			const int w = o->w << st->attempt;
			st->aw0 = w;
			if (st->attempt + 1 < B200_MAX_BAND_TRY && st->res.max_off >= (w >> 1) + (w >> 2)) {   /* bwamem.c:828 */
				++st->attempt;
				left_job(p, ch, s, 0, j);
				j->w = ksw_b200_clamp_w(j->qlen, o->mat, o->o_del, o->e_del, o->o_ins, o->e_ins, o->w << st->attempt, o->pen_clip5);
				++p->st_retry;
				return 1;
			}
			region_left(o, a, s, &st->res);
			st->phase = 3;
		}
		if (st->phase == 3) {                                         /* start the right side */
			st->attempt = 0;
			if (s->qbeg + s->len != p->reads.a[ch->read].l_query) {
				st->sc0 = a->score;
				right_job(p, ch, s, st->sc0, 0, j);
				j->w = ksw_b200_clamp_w(j->qlen, o->mat, o->o_del, o->e_del, o->o_ins, o->e_ins, o->w, o->pen_clip3);
				st->phase = 2;
				return 1;
			}
			region_right(o, a, s, 0, p->reads.a[ch->read].l_query, ch->rmax0);
			st->phase = 4;
		}
		if (st->phase == 2) {                                         /* a right result arrived */
			const int w = o->w << st->attempt;
			st->aw1 = w;
			if (st->attempt + 1 < B200_MAX_BAND_TRY &&
			    !(st->res.score == st->sc0 || st->res.max_off < (w >> 1) + (w >> 2))) {               /* bwamem.c:856 */
				++st->attempt;
				right_job(p, ch, s, st->sc0, 0, j);
				j->w = ksw_b200_clamp_w(j->qlen, o->mat, o->o_del, o->e_del, o->o_ins, o->e_ins, o->w << st->attempt, o->pen_clip3);
				++p->st_retry;
				return 1;
			}
			region_right(o, a, s, &st->res, p->reads.a[ch->read].l_query, ch->rmax0);
			st->phase = 4;
		}
		/* phase 4: the region is complete */
		region_end(a, seeds, n, st->aw0, st->aw1);
		st->phase = 0;
		--st->k;
	}
}

int b200_ext_plan_run_rounds(b200_ext_plan_t *p, ksw_b200_ctx_t *ctx)
{
	const b200_ext_opt_t *o = &p->opt;
	ksw_b200_cfg_t cfg;
	size_t r, n_reads = p->reads.n, k;
	int rc;
	/* the packer clamps again with cfg.end_bonus; with the larger of the two bonuses that second clamp is a no-op */
	cfg_from_opt(o, o->pen_clip5 > o->pen_clip3 ? o->pen_clip5 : o->pen_clip3, &cfg);
	vec_push(p->read_item0, p->items.n);                           /* sentinel */
	vec_reserve(p->rstate, n_reads + 1);
	p->rstate.n = n_reads;
	for (r = 0; r < n_reads; ++r) {
		read_state_t *st = &p->rstate.a[r];
		memset(st, 0, sizeof(*st));
		st->item = p->read_item0.a[r]; st->item_end = p->read_item0.a[r + 1];
		st->order = 0;
	}
	p->read_item0.n -= 1;
	/* round 0 asks every read; later rounds only the reads whose job was in the previous round */
	p->jobs.n = p->owner.n = 0;
	for (r = 0; r < n_reads; ++r) {
		ksw_b200_job_t j;
		if (read_advance(p, &p->rstate.a[r], &j)) { vec_push(p->jobs, j); vec_push(p->owner, (uint32_t)r); }
	}
	while (p->jobs.n) {
		size_t n_jobs = p->jobs.n, w = 0;
		rc = submit_jobs(p, ctx, &cfg, n_jobs);
		if (rc) return rc;
		++p->st_rounds;
		for (k = 0; k < n_jobs; ++k) {
			read_state_t *st = &p->rstate.a[p->owner.a[k]];
			ksw_b200_job_t j;
			if (st->phase == 1) ++p->st_left; else ++p->st_right;
			st->res = p->res.a[k];
			if (read_advance(p, st, &j)) { p->jobs.a[w] = j; p->owner.a[w] = p->owner.a[k]; ++w; }
		}
		p->jobs.n = p->owner.n = w;
	}
	return 0;
}

/* hands the regions of a read (rounds mode) to the caller, who frees them with free() */
void b200_ext_plan_take_regions(b200_ext_plan_t *p, int read, b200_alnreg_v *out)
{
	read_state_t *st = &p->rstate.a[read];
	*out = st->av;
	st->av.n = st->av.m = 0; st->av.a = 0;
}

int64_t b200_ext_plan_rounds(const b200_ext_plan_t *p) { return p->st_rounds; }

void b200_ext_plan_stats(const b200_ext_plan_t *p, int64_t *n_seeds, int64_t *n_left, int64_t *n_right, int64_t *n_retry)
{
	if (n_seeds) *n_seeds = p->st_seeds;
	if (n_left) *n_left = p->st_left;
	if (n_right) *n_right = p->st_right;
	if (n_retry) *n_retry = p->st_retry;
}

int b200_chain2aln_batch(ksw_b200_ctx_t *ctx, const b200_ext_opt_t *opt, int64_t l_pac, const uint8_t *pac,
                         int n_reads, const b200_read_t *reads, b200_alnreg_v *av)
{
	b200_ext_plan_t *p = b200_ext_plan_create(opt, l_pac, pac);
	VEC(int) handles = {0, 0, 0};
	int r, c, rc;
	size_t h = 0;
	if (!p) return 1;
	for (r = 0; r < n_reads; ++r) {
		const int rd = b200_ext_plan_add_read(p, reads[r].l_query, reads[r].query);
		for (c = 0; c < reads[r].n_chains; ++c) vec_push(handles, b200_ext_plan_add_chain(p, rd, &reads[r].chains[c]));
	}
	rc = b200_ext_plan_run(p, ctx);
	if (rc == 0)
		for (r = 0; r < n_reads; ++r)
			for (c = 0; c < reads[r].n_chains; ++c) b200_ext_replay_chain(p, handles.a[h++], &av[r]);
	free(handles.a);
	b200_ext_plan_destroy(p);
	return rc;
}

/* Flat-array form of b200_chain2aln_batch for bindings without struct-of-pointer support (ctypes, cgo
 * slices): reads = (read_off, read_len) into qpool; chains = (chain_read, chain_seed0, chain_nseeds)
 * into seeds[], listed read by read.  Regions are written read by read into out[0..*n_out) with
 * out_read[k] = the read of region k.  Returns 0, a ksw_b200 error code, or -1 if out_cap is too small. */
int b200_chain2aln_flat(ksw_b200_ctx_t *ctx, const b200_ext_opt_t *opt, int64_t l_pac, const uint8_t *pac,
                        int n_reads, const int64_t *read_off, const int32_t *read_len, const uint8_t *qpool,
                        int n_chains, const int32_t *chain_read, const int64_t *chain_seed0, const int32_t *chain_nseeds,
                        const b200_seed_t *seeds, int64_t out_cap, b200_alnreg_t *out, int32_t *out_read, int64_t *n_out)
{
	b200_ext_plan_t *p = b200_ext_plan_create(opt, l_pac, pac);
	int *handle = malloc(sizeof(int) * (size_t)(n_chains > 0 ? n_chains : 1));
	int r, c = 0, rc;
	int64_t n = 0;
	if (!p || !handle) return 1;
	for (r = 0; r < n_reads; ++r) {
		const int rd = b200_ext_plan_add_read(p, read_len[r], qpool + read_off[r]);
		for (; c < n_chains && chain_read[c] == r; ++c) {
			b200_chain_t ch;
			ch.n = ch.m = chain_nseeds[c]; ch.pos = 0;
			ch.seeds = (b200_seed_t *)(seeds + chain_seed0[c]);
			handle[c] = b200_ext_plan_add_chain(p, rd, &ch);
		}
	}
	rc = b200_ext_plan_run(p, ctx);
	c = 0;
	for (r = 0; rc == 0 && r < n_reads; ++r) {
		b200_alnreg_v av = {0, 0, 0};
		size_t k;
		for (; c < n_chains && chain_read[c] == r; ++c) b200_ext_replay_chain(p, handle[c], &av);
		for (k = 0; k < av.n; ++k) {
			if (n >= out_cap) { rc = -1; break; }
			out[n] = av.a[k]; out_read[n] = r; ++n;
		}
		free(av.a);
	}
	*n_out = n;
	free(handle);
	b200_ext_plan_destroy(p);
	return rc;
}

/* b200_chain2aln_flat with the rounds scheduler; n_jobs_out (may be NULL) = DP jobs actually run */
int b200_chain2aln_flat_rounds(ksw_b200_ctx_t *ctx, const b200_ext_opt_t *opt, int64_t l_pac, const uint8_t *pac,
                               int n_reads, const int64_t *read_off, const int32_t *read_len, const uint8_t *qpool,
                               int n_chains, const int32_t *chain_read, const int64_t *chain_seed0, const int32_t *chain_nseeds,
                               const b200_seed_t *seeds, int64_t out_cap, b200_alnreg_t *out, int32_t *out_read, int64_t *n_out,
                               int64_t *n_jobs_out)
{
	b200_ext_plan_t *p = b200_ext_plan_create(opt, l_pac, pac);
	int r, c = 0, rc;
	int64_t n = 0;
	if (!p) return 1;
	for (r = 0; r < n_reads; ++r) {
		const int rd = b200_ext_plan_add_read(p, read_len[r], qpool + read_off[r]);
		for (; c < n_chains && chain_read[c] == r; ++c) {
			b200_chain_t ch;
			ch.n = ch.m = chain_nseeds[c]; ch.pos = 0;
			ch.seeds = (b200_seed_t *)(seeds + chain_seed0[c]);
			b200_ext_plan_add_chain(p, rd, &ch);
		}
	}
	rc = b200_ext_plan_run_rounds(p, ctx);
	for (r = 0; rc == 0 && r < n_reads; ++r) {
		b200_alnreg_v av;
		size_t k;
		b200_ext_plan_take_regions(p, r, &av);
		for (k = 0; k < av.n; ++k) {
			if (n >= out_cap) { rc = -1; break; }
			out[n] = av.a[k]; out_read[n] = r; ++n;
		}
		free(av.a);
	}
	*n_out = n;
	if (n_jobs_out) *n_jobs_out = p->st_left + p->st_right;     /* retries are counted on the side they belong to */
	b200_ext_plan_destroy(p);
	return rc;
}
