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
	VEC(ksw_b200_res_t) res;
	VEC(uint32_t) owner;               /* job -> ext record */
	int64_t st_seeds, st_left, st_right, st_retry;
};

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
	return p;
}

void b200_ext_plan_destroy(b200_ext_plan_t *p)
{
	if (!p) return;
	free(p->reads.a); free(p->chains.a); free(p->seeds.a); free(p->ext.a);
	free(p->qpool.a); free(p->tpool.a); free(p->jobs.a); free(p->res.a); free(p->owner.a);
	free(p);
}

void b200_ext_plan_reset(b200_ext_plan_t *p)
{
	p->reads.n = p->chains.n = p->seeds.n = p->ext.n = 0;
	p->qpool.n = p->tpool.n = 0;
	p->st_seeds = p->st_left = p->st_right = p->st_retry = 0;
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
	/* the window, forward and reversed (bns_get_seq, bwamem.c:757) */
	vec_reserve(p->tpool, p->tpool.n + 2 * (size_t)(hi - lo));
	ch.r_fwd = p->tpool.n;
	rlen = b200_get_ref_slice(p->l_pac, p->pac, lo, hi, p->tpool.a + p->tpool.n);
	assert(rlen == hi - lo);
	p->tpool.n += (size_t)rlen;
	ch.r_rev = p->tpool.n;
	for (k = 0; k < rlen; ++k) p->tpool.a[ch.r_rev + k] = p->tpool.a[ch.r_fwd + (rlen - 1 - k)];
	p->tpool.n += (size_t)rlen;
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
	j->t_off = ch->r_rev + (uint64_t)(rlen - tmp);
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
	j->t_off = ch->r_fwd + (uint64_t)re;
	j->tlen = (int32_t)(ch->rmax1 - ch->rmax0 - re);
	j->h0 = h0;
	j->w = w;
}

/* Trace facility (the batched counterpart of the reference's `-v 4` extension trace, bwamem.c:821-828): with
 * KSW_B200_DUMP=<prefix> every pass appends its job batch to <prefix>.<pid>.bin as
 * [magic "KSWJ"][cfg][n][qpool bytes][tpool bytes][jobs][qpool][tpool] for offline replay (scripts/bench_jobs.py). */
#include <stdio.h>
#include <unistd.h>
static void dump_jobs(const b200_ext_plan_t *p, const ksw_b200_cfg_t *cfg)
{
	const char *pre = getenv("KSW_B200_DUMP");
	char path[4096];
	FILE *f;
	uint64_t hdr[3];
	if (!pre || !*pre) return;
	snprintf(path, sizeof(path), "%s.%d.bin", pre, (int)getpid());
	f = fopen(path, "ab");
	if (!f) return;
	hdr[0] = p->jobs.n; hdr[1] = p->qpool.n; hdr[2] = p->tpool.n;
	fwrite("KSWJ", 1, 4, f);
	fwrite(cfg, sizeof(*cfg), 1, f);
	fwrite(hdr, sizeof(hdr), 1, f);
	fwrite(p->jobs.a, sizeof(ksw_b200_job_t), p->jobs.n, f);
	fwrite(p->qpool.a, 1, p->qpool.n, f);
	fwrite(p->tpool.a, 1, p->tpool.n, f);
	fclose(f);
}

static int run_jobs(b200_ext_plan_t *p, ksw_b200_ctx_t *ctx, const ksw_b200_cfg_t *cfg)
{
	if (p->jobs.n == 0) return 0;
	dump_jobs(p, cfg);
	vec_reserve(p->res, p->jobs.n);
	return ksw_b200_extend_batch(ctx, cfg, (int64_t)p->jobs.n, p->jobs.a, p->qpool.a, p->tpool.a, p->res.a);
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

void b200_ext_replay_chain(const b200_ext_plan_t *p, int chain, b200_alnreg_v *av)
{
	const b200_ext_opt_t *o = &p->opt;
	const chain_rec_t *ch;
	const b200_seed_t *seeds;
	const ext_rec_t *ext;
	uint64_t *order;
	int l_query, k, n;
	size_t i;
	if (chain < 0) return;
	ch = &p->chains.a[chain];
	seeds = p->seeds.a + ch->seed0;
	ext = p->ext.a + ch->seed0;
	n = ch->n;
	l_query = p->reads.a[ch->read].l_query;
	/* longest seed first; equal lengths in index order (bwamem.c:760-763) */
	order = malloc((size_t)n * 8);
	for (k = 0; k < n; ++k) order[k] = (uint64_t)seeds[k].len << 32 | (uint32_t)k;
	qsort(order, (size_t)n, 8, cmp_u64);

	for (k = n - 1; k >= 0; --k) {
		const int si = (int)(uint32_t)order[k];
		const b200_seed_t *s = &seeds[si];
		const ext_rec_t *x = &ext[si];
		b200_alnreg_t *a;
		int aw0 = o->w, aw1 = o->w, j;
		/* is the seed inside, and near the diagonal of, a region found earlier? (bwamem.c:769-784) */
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
		if (i < av->n) {
			/* contained: extend anyway only if a longer, overlapping seed lies on another diagonal (bwamem.c:785-799) */
			for (j = k + 1; j < n; ++j) {
				const b200_seed_t *t;
				if (order[j] == 0) continue;
				t = &seeds[(uint32_t)order[j]];
				if (t->len < s->len * .95) continue;
				if (s->qbeg <= t->qbeg && s->qbeg + s->len - t->qbeg >= s->len >> 2 && t->qbeg - s->qbeg != t->rbeg - s->rbeg) break;
				if (t->qbeg <= s->qbeg && t->qbeg + t->len - s->qbeg >= s->len >> 2 && s->qbeg - t->qbeg != s->rbeg - t->rbeg) break;
			}
			if (j == n) { order[k] = 0; continue; }              /* skipped: marked like srt[k] = 0 */
		}

		a = av_push(av);
		memset(a, 0, sizeof(*a));
		a->w = o->w;
		a->score = a->truesc = -1;
		if (x->has_left) {                                       /* bwamem.c:810-838 */
			const ksw_b200_res_t *L = &x->left;
			aw0 = x->aw_left;
			a->score = L->score;
			if (L->gscore <= 0 || L->gscore <= a->score - o->pen_clip5) {   /* local end */
				a->qb = s->qbeg - L->qle; a->rb = s->rbeg - L->tle;
				a->truesc = a->score;
			} else {                                             /* reaches the query start */
				a->qb = 0; a->rb = s->rbeg - L->gtle;
				a->truesc = L->gscore;
			}
		} else {
			a->score = a->truesc = s->len * o->a; a->qb = 0; a->rb = s->rbeg;
		}
		if (x->has_right) {                                      /* bwamem.c:841-867 */
			const ksw_b200_res_t *R = &x->right;
			const int sc0 = a->score, qe = s->qbeg + s->len;
			const int64_t re = s->rbeg + s->len - ch->rmax0;
			aw1 = x->aw_right;
			a->score = R->score;
			if (R->gscore <= 0 || R->gscore <= a->score - o->pen_clip3) {
				a->qe = qe + R->qle; a->re = ch->rmax0 + re + R->tle;
				a->truesc += a->score - sc0;
			} else {
				a->qe = l_query; a->re = ch->rmax0 + re + R->gtle;
				a->truesc += R->gscore - sc0;
			}
		} else {
			a->qe = l_query; a->re = s->rbeg + s->len;
		}
		/* seeds fully inside the new region (bwamem.c:870-874) */
		a->seedcov = 0;
		for (j = 0; j < n; ++j) {
			const b200_seed_t *t = &seeds[j];
			if (t->qbeg >= a->qb && t->qbeg + t->len <= a->qe && t->rbeg >= a->rb && t->rbeg + t->len <= a->re)
				a->seedcov += t->len;
		}
		a->w = aw0 > aw1 ? aw0 : aw1;
	}
	free(order);
}

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
