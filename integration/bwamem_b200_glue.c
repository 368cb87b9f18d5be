/*
 * integration/bwamem_b200_glue.c — the reference-side binding: `bwa mem` with its seed-extension
 * pass running on the B200 library.
 *
 * How it is built (integration/Makefile, target _bin/bwa_b200): this one file is compiled INSTEAD of the
 * reference's bwamem.c.  It textually includes the reference's bwamem.c where it lies (nothing is
 * copied) with the fork's mem_process_seqs renamed, and then defines mem_process_seqs again with pass 1
 * restructured as the fork's authors sketched it (bwamem.c:579 `mem_chain2aln_batched`, commented out):
 *
 *   worker1_b200(start, batch_size)            <- kt_for_batch, kthread_batch.c:44 (unchanged)
 *     for every read of the batch: encode, mem_chain, mem_chain_flt       (host, unchanged: bwamem.c:1093-1097)
 *       for every chain: mem_chain2aln_short (host, unchanged); if it falls through, register the chain
 *     b200_ext_plan_run: pass L, pass R (+ band retries) on the GPU        (replaces the ksw_extend2 calls)
 *       [or, with KSW_B200_SCHED=rounds, b200_ext_plan_run_rounds: only the seeds the reference extends, in rounds]
 *     for every read, chains in order: push the short-path region or replay mem_chain2aln from the cached DP
 *     mem_sort_and_dedup, mem_test_and_remove_exact                        (host, unchanged: bwamem.c:1111-1117)
 *
 *     CIGAR look-ahead: for every region of the batch, the global alignments that mem_reg2aln will
 *       ask for in pass 2 (bwamem.c:1187-1200 via bwa_gen_cigar2, bwa.c:118-132; up to three bands per region) are
 *       computed now in batches on the GPU (ksw_b200_global_batch) and stored in a table keyed by the EXACT inputs
 *       of ksw_global2 (query bytes, reference-window bytes, band).
 *   pass 2 (worker2) is the reference's own code; only the ksw_global2 call inside the reference's bwa_gen_cigar2 is
 *   redirected (bwa.c is compiled with -Dksw_global2=b200_global2_hook, nothing is edited) to b200_global2_hook below:
 *   table hit -> the stored score + CIGAR; miss (mate-rescued regions, bwa_fix_xref2) -> one GPU call for that job.
 *   Because the key is the complete input of the DP, a wrong prediction can only cost a miss, never a wrong CIGAR.
 *
 * Everything else (seeding, chaining, pairing, mate rescue, MD/NM, SAM text) is the reference's own code,
 * so the SAM must be byte-identical to stock `bwa mem` apart from @PG.  One extension context per worker
 * thread; worker t uses GPU (t mod #GPUs), so read batches shard over the GPUs of the box with no exchange.
 * `-b` stays the batch knob; the default of 1 read per batch (bwamem.c:68) would mean one GPU round trip per
 * read, so values below KSW_B200_MIN_BATCH (default 4096, env) are raised to it.
 */
#define mem_process_seqs mem_process_seqs_cpu_fork
#include "bwamem.c"                       /* the reference source, found through -I$(REF) */
#undef mem_process_seqs

#include <pthread.h>
#include "bwamem_b200.h"

typedef struct {
	ksw_b200_ctx_t *ctx;               /* private context (KSW_B200_QUEUE=0), else NULL */
	ksw_b200_queue_t *queue;           /* the shared queue of this worker's GPU */
	b200_ext_plan_t *plan;
	const uint8_t *pac;
} b200_thread_t;

/* kt_for_batch creates fresh pthreads for every chunk (kthread_batch.c:54), so the per-worker state is kept
 * in a table indexed by the worker id, not in thread-local storage */
#define B200_MAX_WORKERS 1024
static b200_thread_t b200_workers[B200_MAX_WORKERS];
static int b200_n_gpus = -1;
/* GPUs the read batches shard over: all visible ones, or the first KSW_B200_GPUS of them */
static int b200_gpu_count(void)
{
	int n = ksw_b200_device_count();
	const char *e = getenv("KSW_B200_GPUS");
	if (e && atoi(e) > 0 && atoi(e) < n) n = atoi(e);
	return n;
}
/* wall-clock seconds summed over worker threads: seeding+chaining, planning, GPU passes, replay+dedup */
static double b200_t_seed, b200_t_plan, b200_t_gpu, b200_t_replay, b200_t_init;
static void b200_add_time(double *acc, double dt) { __sync_synchronize(); { static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER; pthread_mutex_lock(&mu); *acc += dt; pthread_mutex_unlock(&mu); } }

/* One submission queue per GPU shared by all workers (cross-thread batch coalescing): whatever the
 * workers submit while the GPU is busy runs as one merged batch, and GPU memory is one context per GPU whatever -t is.
 * KSW_B200_QUEUE=0 (or KSW_B200_REF=0, the queue needs the device-resident reference): one private context per worker. */
#define B200_MAX_GPUS 64
static ksw_b200_queue_t *b200_queue[B200_MAX_GPUS];
static const uint8_t *b200_queue_pac[B200_MAX_GPUS];
static pthread_mutex_t b200_queue_mu[B200_MAX_GPUS];        /* one per GPU: the queues of a multi-GPU box come up side by side */
static pthread_once_t b200_queue_once = PTHREAD_ONCE_INIT;
static void b200_queue_mu_init(void) { int g; for (g = 0; g < B200_MAX_GPUS; ++g) pthread_mutex_init(&b200_queue_mu[g], 0); }
static int b200_queue_on(void)
{
	static int v = -1;
	if (v < 0) {
		const char *e = getenv("KSW_B200_QUEUE"), *r = getenv("KSW_B200_REF");
		v = !(e && e[0] == '0') && !(r && r[0] == '0');
	}
	return v;
}
static ksw_b200_queue_t *b200_queue_for(int gpu, const bntseq_t *bns, const uint8_t *pac)
{
	ksw_b200_queue_t *q;
	pthread_once(&b200_queue_once, b200_queue_mu_init);
	pthread_mutex_lock(&b200_queue_mu[gpu]);
	if (!b200_queue[gpu] && ksw_b200_queue_create(gpu, &b200_queue[gpu]) != 0)
		err_fatal(__func__, "no usable CUDA device: the B200 extension path has no CPU fallback");
	if (pac && b200_queue_pac[gpu] != pac) {
		if (ksw_b200_queue_ref_set(b200_queue[gpu], pac, bns->l_pac) != 0) err_fatal(__func__, "cannot put the reference on GPU %d", gpu);
		b200_queue_pac[gpu] = pac;
	}
	q = b200_queue[gpu];
	pthread_mutex_unlock(&b200_queue_mu[gpu]);
	return q;
}

static b200_thread_t *b200_thread_state(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, int tid)
{
	b200_thread_t *t;
	if (tid < 0 || tid >= B200_MAX_WORKERS) err_fatal(__func__, "more than %d worker threads", B200_MAX_WORKERS);
	t = &b200_workers[tid];
	if (b200_n_gpus < 0) b200_n_gpus = b200_gpu_count();
	if (b200_queue_on()) {
		if (b200_n_gpus < 1) err_fatal(__func__, "no usable CUDA device: the B200 extension path has no CPU fallback");
		t->queue = b200_queue_for(tid % (b200_n_gpus < B200_MAX_GPUS ? b200_n_gpus : B200_MAX_GPUS), bns, pac);
	} else if (!t->ctx) {
		if (b200_n_gpus < 0) b200_n_gpus = b200_gpu_count();
		if (b200_n_gpus < 1 || ksw_b200_ctx_create(tid % b200_n_gpus, &t->ctx) != 0)
			err_fatal(__func__, "no usable CUDA device: the B200 extension path has no CPU fallback");
		ksw_b200_ctx_set_pack_threads(t->ctx, 1);          /* the bwa worker threads are the parallelism */
	}
	if (!t->plan || t->pac != pac) {
		b200_ext_opt_t eo;
		if (t->plan) b200_ext_plan_destroy(t->plan);
		eo.a = opt->a; eo.b = opt->b; eo.o_del = opt->o_del; eo.e_del = opt->e_del; eo.o_ins = opt->o_ins; eo.e_ins = opt->e_ins;
		eo.pen_clip5 = opt->pen_clip5; eo.pen_clip3 = opt->pen_clip3; eo.w = opt->w; eo.zdrop = opt->zdrop;
		memcpy(eo.mat, opt->mat, 25);
		t->plan = b200_ext_plan_create(&eo, bns->l_pac, pac);
		t->pac = pac;
		/* the .pac stays on the GPU (one copy per device, shared by the workers' contexts) and the extension jobs name their
		 * targets by coordinate: no bns_get_seq / window copies in pass 1.  KSW_B200_REF=0: the plan
		 * materialises the windows on the host as before (A/B switch) */
		{
			const char *e = getenv("KSW_B200_REF");
			b200_ext_plan_set_device_ref(t->plan, !(e && e[0] == '0'));
		}
		if (t->queue) b200_ext_plan_set_queue(t->plan, t->queue);
	}
	return t;
}

/* CUDA context creation costs seconds on a large GPU, and so does the first launch of every kernel; start both while
 * `bwa mem` is still loading the index and parsing the first chunk of reads.  (glibc passes argc/argv to constructors.) */
static void *b200_warmup_device(void *arg)
{
	const int d = (int)(intptr_t)arg;
	ksw_b200_ctx_t *c = 0;
	if (ksw_b200_ctx_create(d, &c) != 0) return 0;
	{
		/* one job per kernel (keyed, unkeyed, generic; binning; global alignment; local alignment): with lazy module loading the
		 * first launch of every kernel costs tens of milliseconds, and 16 workers would queue up behind it in their first batch */
		static const int qlens[3] = {20, 200, 600};
		uint8_t seq[600];
		ksw_b200_cfg_t cfg;
		ksw_b200_job_t jobs[3];
		ksw_b200_res_t res[3];
		ksw_b200_gjob_t gj;
		ksw_b200_gres_t gr;
		ksw_b200_ajob_t aj[2];
		ksw_b200_ares_t ar[2];
		const uint32_t *pool;
		int64_t total;
		int i;
		for (i = 0; i < 600; ++i) seq[i] = (uint8_t)((i * 7 + i / 5) & 3);
		memset(&cfg, 0, sizeof(cfg));
		bwa_fill_scmat(1, 4, cfg.mat);
		cfg.m = 5; cfg.o_del = cfg.o_ins = 6; cfg.e_del = cfg.e_ins = 1; cfg.zdrop = 100; cfg.end_bonus = 5;
		for (i = 0; i < 3; ++i) { jobs[i].q_off = 0; jobs[i].t_off = 0; jobs[i].qlen = qlens[i]; jobs[i].tlen = qlens[i]; jobs[i].h0 = 19; jobs[i].w = 100; }
		ksw_b200_extend_batch(c, &cfg, 3, jobs, seq, seq, res);
		gj.q_off = gj.t_off = 0; gj.qlen = gj.tlen = 100; gj.w = 10; gj.reserved = 0;
		ksw_b200_global_batch(c, &cfg, 1, &gj, seq, seq, &gr, &pool, &total);
		for (i = 0; i < 2; ++i) {                                   /* byte kernel and 16-bit kernel */
			aj[i].q_off = 0; aj[i].t_off = 100; aj[i].qlen = 100; aj[i].tlen = 300; aj[i].reserved = 0;
			aj[i].xtra = KSW_XSUBO | KSW_XSTART | (i == 0 ? KSW_XBYTE : 0) | 19;
		}
		ksw_b200_align_batch(c, &cfg, 2, aj, seq, seq, ar);
	}
	ksw_b200_ctx_destroy(c);
	return 0;
}
static void *b200_warmup_thread(void *arg)
{
	/* only the GPUs the run will use (KSW_B200_GPUS), all of them at the same time */
	int d, n = b200_gpu_count();
	pthread_t th[B200_MAX_GPUS];
	(void)arg;
	if (n > B200_MAX_GPUS) n = B200_MAX_GPUS;
	for (d = 0; d < n; ++d) if (pthread_create(&th[d], 0, b200_warmup_device, (void *)(intptr_t)d) != 0) th[d] = 0;
	for (d = 0; d < n; ++d) if (th[d]) pthread_join(th[d], 0);
	return 0;
}
__attribute__((constructor)) static void b200_warmup(int argc, char **argv)
{
	pthread_t th;
	if (argc >= 2 && strcmp(argv[1], "mem") == 0 && !getenv("KSW_B200_NO_WARMUP"))
		if (pthread_create(&th, 0, b200_warmup_thread, 0) == 0) pthread_detach(th);
}

/* ------------------------------------------------------------------------------------------------------------------
 * CIGAR look-ahead table.  Entries are written in pass 1 (several worker threads: lock-free push onto bucket lists,
 * storage from per-worker arenas) and only read in pass 2; the table is emptied at the start of every chunk of reads. */
typedef struct b200_cig_entry {
	struct b200_cig_entry *next;
	uint64_t hash;
	int32_t qlen, tlen, w, score, n_cigar;
	/* followed by: uint32_t cigar[n_cigar]; uint8_t query[qlen]; uint8_t target[tlen] */
} b200_cig_entry_t;

#define B200_CIG_BUCKETS (1u << 21)
#define B200_ARENA_BLOCK ((size_t)4 << 20)
typedef struct { char **blk; int n_blk, m_blk, cur; size_t used; } b200_arena_t;
static b200_cig_entry_t **b200_cig_table;
static b200_arena_t b200_cig_arena[B200_MAX_WORKERS];
static long long b200_cig_hits, b200_cig_misses, b200_cig_jobs;
static double b200_t_cigar;

/* KSW_B200_CIGAR: unset/1 = look-ahead in pass 1 + table (default); 0 = off, pass 2 entirely the reference's (its own
 * ksw_global2); 2 = no look-ahead but the redirected call stays: every look-up misses and goes to the GPU one job at a
 * time (slow; it exists to test the miss path) */
static int b200_cigar_mode(void)
{
	static int v = -1;
	if (v < 0) { const char *e = getenv("KSW_B200_CIGAR"); v = !e ? 1 : (e[0] == '0' ? 0 : (e[0] == '2' ? 2 : 1)); }
	return v;
}
static int b200_cigar_on(void) { return b200_cigar_mode() != 0; }

static void *b200_arena_alloc(b200_arena_t *a, size_t n)
{
	n = (n + 15) & ~(size_t)15;
	if (n > B200_ARENA_BLOCK) return 0;                    /* too large to keep: the caller skips it (computed on demand later) */
	if (a->n_blk == 0 || a->used + n > B200_ARENA_BLOCK) {
		if (a->n_blk && a->cur + 1 < a->n_blk) ++a->cur;
		else {
			if (a->n_blk == a->m_blk) { a->m_blk = a->m_blk ? a->m_blk << 1 : 16; a->blk = (char **)realloc(a->blk, sizeof(char *) * a->m_blk); }
			a->blk[a->n_blk] = (char *)malloc(B200_ARENA_BLOCK);
			a->cur = a->n_blk++;
		}
		a->used = 0;
	}
	a->used += n;
	return a->blk[a->cur] + a->used - n;
}

static void b200_cig_reset(void)                         /* single-threaded: called between chunks */
{
	int t;
	if (!b200_cig_table) b200_cig_table = (b200_cig_entry_t **)calloc(B200_CIG_BUCKETS, sizeof(void *));
	else memset(b200_cig_table, 0, sizeof(void *) * B200_CIG_BUCKETS);
	for (t = 0; t < B200_MAX_WORKERS; ++t) { b200_cig_arena[t].cur = 0; b200_cig_arena[t].used = 0; }   /* blocks are kept and reused */
}

static uint64_t b200_cig_hash(int qlen, const uint8_t *q, int tlen, const uint8_t *t, int w)
{
	uint64_t h = 1469598103934665603ull ^ ((uint64_t)(uint32_t)qlen << 32 | (uint32_t)tlen) ^ ((uint64_t)(uint32_t)w * 0x9E3779B97F4A7C15ull);
	int i;
	for (i = 0; i < qlen; ++i) h = (h ^ q[i]) * 1099511628211ull;
	for (i = 0; i < tlen; ++i) h = (h ^ t[i]) * 1099511628211ull;
	return h;
}

static void b200_cig_insert(int tid, int qlen, const uint8_t *q, int tlen, const uint8_t *t, int w, int score, int n_cigar, const uint32_t *cigar)
{
	b200_cig_entry_t *e = (b200_cig_entry_t *)b200_arena_alloc(&b200_cig_arena[tid], sizeof(b200_cig_entry_t) + 4 * (size_t)n_cigar + qlen + tlen);
	uint32_t *c;
	uint8_t *p;
	b200_cig_entry_t **slot;
	if (!e) return;
	c = (uint32_t *)(e + 1);
	p = (uint8_t *)(c + n_cigar);
	e->hash = b200_cig_hash(qlen, q, tlen, t, w);
	e->qlen = qlen; e->tlen = tlen; e->w = w; e->score = score; e->n_cigar = n_cigar;
	memcpy(c, cigar, 4 * (size_t)n_cigar); memcpy(p, q, qlen); memcpy(p + qlen, t, tlen);
	slot = &b200_cig_table[e->hash & (B200_CIG_BUCKETS - 1)];
	do { e->next = *(b200_cig_entry_t * volatile *)slot; } while (!__sync_bool_compare_and_swap(slot, e->next, e));
}

/* What the reference's bwa_gen_cigar2 calls instead of ksw_global2 (bwa.c:132) in this build.  Same contract: returns
 * the score; the CIGAR array is malloc'd for the caller (who grows it with kputw to append MD, bwa.c:136). */
int b200_global2_hook(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                      int o_del, int e_del, int o_ins, int e_ins, int w, int *n_cigar_, uint32_t **cigar_)
{
	/* only for misses, which are rare: one context for the whole process behind a mutex (kt_for starts fresh threads for
	 * every chunk, so thread-local contexts would pile up) */
	static ksw_b200_ctx_t *ctx;
	static pthread_mutex_t miss_mu = PTHREAD_MUTEX_INITIALIZER;
	int score;
	ksw_b200_cfg_t cfg;
	ksw_b200_gjob_t job;
	ksw_b200_gres_t r;
	const uint32_t *pool = 0;
	int64_t total = 0;
	int k;
	if (!b200_cigar_on()) return ksw_global2(qlen, query, tlen, target, m, mat, o_del, e_del, o_ins, e_ins, w, n_cigar_, cigar_);
	if (b200_cig_table) {
		const uint64_t h = b200_cig_hash(qlen, query, tlen, target, w);
		const b200_cig_entry_t *e;
		for (e = b200_cig_table[h & (B200_CIG_BUCKETS - 1)]; e; e = e->next) {
			const uint32_t *c = (const uint32_t *)(e + 1);
			const uint8_t *p = (const uint8_t *)(c + e->n_cigar);
			if (e->hash != h || e->qlen != qlen || e->tlen != tlen || e->w != w) continue;
			if (memcmp(p, query, qlen) != 0 || memcmp(p + qlen, target, tlen) != 0) continue;
			__sync_fetch_and_add(&b200_cig_hits, 1);
			if (n_cigar_) *n_cigar_ = 0;
			if (n_cigar_ && cigar_) {
				uint32_t *out = (uint32_t *)malloc(4 * (size_t)(e->n_cigar > 0 ? e->n_cigar : 1));
				memcpy(out, c, 4 * (size_t)e->n_cigar);
				*n_cigar_ = e->n_cigar; *cigar_ = out;
			}
			return e->score;
		}
	}
	/* miss: one job on the GPU (there is no CPU fallback in this mode) */
	__sync_fetch_and_add(&b200_cig_misses, 1);
	if (b200_queue_on()) {
		/* through GPU 0's shared queue: concurrent misses of the pass-2 threads are merged into one batch there */
		uint32_t *own = 0;
		if (b200_n_gpus < 0) b200_n_gpus = b200_gpu_count();
		memcpy(cfg.mat, mat, 25);
		cfg.m = m; cfg.o_del = o_del; cfg.e_del = e_del; cfg.o_ins = o_ins; cfg.e_ins = e_ins; cfg.zdrop = 0; cfg.end_bonus = 0;
		job.q_off = 0; job.t_off = 0; job.qlen = qlen; job.tlen = tlen; job.w = w; job.reserved = 0;
		if (b200_n_gpus < 1 || ksw_b200_queue_global(b200_queue_for(0, 0, 0), &cfg, 1, &job, query, target, &r, &own, &total) != 0)
			err_fatal(__func__, "GPU global alignment failed");
		if (n_cigar_) *n_cigar_ = 0;
		if (n_cigar_ && cigar_) {
			if (!own) own = (uint32_t *)malloc(4);
			*n_cigar_ = r.n_cigar; *cigar_ = own;
		} else free(own);
		return r.score;
	}
	pthread_mutex_lock(&miss_mu);
	if (!ctx) {
		if (b200_n_gpus < 0) b200_n_gpus = b200_gpu_count();
		if (b200_n_gpus < 1 || ksw_b200_ctx_create(0, &ctx) != 0) err_fatal(__func__, "no usable CUDA device");
	}
	memcpy(cfg.mat, mat, 25);
	cfg.m = m; cfg.o_del = o_del; cfg.e_del = e_del; cfg.o_ins = o_ins; cfg.e_ins = e_ins; cfg.zdrop = 0; cfg.end_bonus = 0;
	job.q_off = 0; job.t_off = 0; job.qlen = qlen; job.tlen = tlen; job.w = w; job.reserved = 0;
	if (ksw_b200_global_batch(ctx, &cfg, 1, &job, query, target, &r, &pool, &total) != 0)
		err_fatal(__func__, "GPU global alignment failed: %s", ksw_b200_strerror(ctx));
	if (n_cigar_) *n_cigar_ = 0;
	if (n_cigar_ && cigar_) {
		uint32_t *out = (uint32_t *)malloc(4 * (size_t)(r.n_cigar > 0 ? r.n_cigar : 1));
		for (k = 0; k < r.n_cigar; ++k) out[k] = pool[r.cigar_off + k];
		*n_cigar_ = r.n_cigar; *cigar_ = out;
	}
	score = r.score;
	pthread_mutex_unlock(&miss_mu);
	return score;
}

/* The jobs pass 2 will ask for, per region, predicted exactly the way mem_reg2aln + bwa_gen_cigar2 derive them. */
typedef struct { int64_t rb, re; int qb, qe, w2, truesc, last_sc, read; } b200_cig_meta_t;
typedef struct {
	ksw_b200_gjob_t *jobs; b200_cig_meta_t *meta; int64_t n, m;
	uint8_t *q, *t; size_t nq, mq, nt, mt;
} b200_cig_batch_t;

static void b200_cig_push(b200_cig_batch_t *B, const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, const uint8_t *read,
                          const b200_cig_meta_t *mt)
{
	/* bwa_gen_cigar2's prelude (bwa.c:95-126): window, strand reversal, band */
	const int l_query = mt->qe - mt->qb;
	int64_t rlen;
	uint8_t *rseq;
	int i, w, max_gap, max_ins, max_del, min_w;
	if (l_query <= 0 || mt->rb >= mt->re || (mt->rb < bns->l_pac && mt->re > bns->l_pac)) return;
	rseq = bns_get_seq(bns->l_pac, pac, mt->rb, mt->re, &rlen);
	if (mt->re - mt->rb != rlen) { free(rseq); return; }
	if (l_query == mt->re - mt->rb && mt->w2 == 0) { free(rseq); return; }          /* no DP in the reference either (bwa.c:110) */
	max_ins = (int)((double)(((l_query + 1) >> 1) * opt->mat[0] - opt->o_ins) / opt->e_ins + 1.);
	max_del = (int)((double)(((l_query + 1) >> 1) * opt->mat[0] - opt->o_del) / opt->e_del + 1.);
	max_gap = max_ins > max_del ? max_ins : max_del;
	max_gap = max_gap > 1 ? max_gap : 1;
	w = (max_gap + abs((int)rlen - l_query) + 1) >> 1;
	w = w < mt->w2 ? w : mt->w2;
	min_w = abs((int)rlen - l_query) + 3;
	w = w > min_w ? w : min_w;
	if (B->n == B->m) {
		B->m = B->m ? B->m << 1 : 4096;
		B->jobs = (ksw_b200_gjob_t *)realloc(B->jobs, sizeof(ksw_b200_gjob_t) * B->m);
		B->meta = (b200_cig_meta_t *)realloc(B->meta, sizeof(b200_cig_meta_t) * B->m);
	}
	if (B->nq + l_query > B->mq) { B->mq = (B->mq ? B->mq << 1 : 1 << 20) + l_query; B->q = (uint8_t *)realloc(B->q, B->mq); }
	if (B->nt + rlen > B->mt) { B->mt = (B->mt ? B->mt << 1 : 1 << 20) + rlen; B->t = (uint8_t *)realloc(B->t, B->mt); }
	if (mt->rb >= bns->l_pac) {                      /* reverse both (bwa.c:103-108): indels end up left-aligned */
		for (i = 0; i < l_query; ++i) B->q[B->nq + i] = read[mt->qb + l_query - 1 - i];
		for (i = 0; i < rlen; ++i) B->t[B->nt + i] = rseq[rlen - 1 - i];
	} else {
		memcpy(B->q + B->nq, read + mt->qb, l_query);
		memcpy(B->t + B->nt, rseq, rlen);
	}
	B->jobs[B->n].q_off = B->nq; B->jobs[B->n].t_off = B->nt;
	B->jobs[B->n].qlen = l_query; B->jobs[B->n].tlen = (int)rlen; B->jobs[B->n].w = w; B->jobs[B->n].reserved = 0;
	B->meta[B->n] = *mt;
	B->nq += l_query; B->nt += rlen; ++B->n;
	free(rseq);
}

/* runs the collected jobs in up to three rounds of growing bands (the do-while of bwamem.c:1194-1200) and fills the table */
static void b200_cig_rounds(b200_thread_t *t, int tid, worker_t *w, b200_cig_batch_t *first)
{
	const mem_opt_t *opt = w->opt;
	b200_cig_batch_t cur = *first, nxt;
	ksw_b200_cfg_t cfg;
	ksw_b200_gres_t *res = 0;
	int64_t m_res = 0;
	int round;
	memset(&nxt, 0, sizeof(nxt));
	memcpy(cfg.mat, opt->mat, 25);
	cfg.m = 5; cfg.o_del = opt->o_del; cfg.e_del = opt->e_del; cfg.o_ins = opt->o_ins; cfg.e_ins = opt->e_ins; cfg.zdrop = 0; cfg.end_bonus = 0;
	for (round = 0; round < 3 && cur.n > 0; ++round) {
		const uint32_t *pool = 0;
		int64_t total = 0, j;
		if (cur.n > m_res) { m_res = cur.n; res = (ksw_b200_gres_t *)realloc(res, sizeof(ksw_b200_gres_t) * m_res); }
		uint32_t *own_pool = 0;
		if (t->queue) {
			if (ksw_b200_queue_global(t->queue, &cfg, cur.n, cur.jobs, cur.q, cur.t, res, &own_pool, &total) != 0)
				err_fatal(__func__, "GPU global alignment failed: %s", ksw_b200_queue_strerror(t->queue));
			pool = own_pool;
		} else if (ksw_b200_global_batch(t->ctx, &cfg, cur.n, cur.jobs, cur.q, cur.t, res, &pool, &total) != 0)
			err_fatal(__func__, "GPU global alignment failed: %s", ksw_b200_strerror(t->ctx));
		__sync_fetch_and_add(&b200_cig_jobs, cur.n);
		nxt.n = 0; nxt.nq = 0; nxt.nt = 0;
		for (j = 0; j < cur.n; ++j) {
			const ksw_b200_gjob_t *g = &cur.jobs[j];
			b200_cig_meta_t mt = cur.meta[j];
			b200_cig_insert(tid, g->qlen, cur.q + g->q_off, g->tlen, cur.t + g->t_off, g->w, res[j].score, res[j].n_cigar, pool + res[j].cigar_off);
			if (res[j].score == mt.last_sc) continue;                             /* bwamem.c:1198 */
			mt.last_sc = res[j].score; mt.w2 <<= 1;
			if (round + 1 < 3 && res[j].score < mt.truesc - opt->a)
				b200_cig_push(&nxt, opt, w->bns, w->pac, (const uint8_t *)w->seqs[mt.read].seq, &mt);
		}
		{ b200_cig_batch_t x = cur; cur = nxt; nxt = x; }
		free(own_pool);
	}
	free(cur.jobs); free(cur.meta); free(cur.q); free(cur.t);
	free(nxt.jobs); free(nxt.meta); free(nxt.q); free(nxt.t);
	free(res);
	memset(first, 0, sizeof(*first));
}

/* one region as mem_reg2aln will see it (bwamem.c:1183-1191): the first band it tries */
static void b200_cig_push_region(b200_cig_batch_t *B, worker_t *w, int read, int qb, int qe, int64_t rb, int64_t re, int truesc, int reg_w)
{
	const mem_opt_t *opt = w->opt;
	b200_cig_meta_t mt;
	int tmp, w2;
	mt.qb = qb; mt.qe = qe; mt.rb = rb; mt.re = re;
	if (bwa_fix_xref2(opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, opt->w, w->bns, w->pac, (uint8_t *)w->seqs[read].seq,
	                  &mt.qb, &mt.qe, &mt.rb, &mt.re) < 0) return;
	tmp = infer_bw(mt.qe - mt.qb, mt.re - mt.rb, truesc, opt->a, opt->o_del, opt->e_del);
	w2 = infer_bw(mt.qe - mt.qb, mt.re - mt.rb, truesc, opt->a, opt->o_ins, opt->e_ins);
	w2 = w2 > tmp ? w2 : tmp;
	if (w2 > opt->w) w2 = w2 < reg_w ? w2 : reg_w;
	mt.w2 = w2; mt.truesc = truesc; mt.last_sc = -(1 << 30); mt.read = read;
	b200_cig_push(B, opt, w->bns, w->pac, (const uint8_t *)w->seqs[read].seq, &mt);
}

static void b200_cig_lookahead(b200_thread_t *t, int tid, worker_t *w, int start, int batch_size)
{
	const mem_opt_t *opt = w->opt;
	b200_cig_batch_t cur;
	int b;
	size_t k;
	memset(&cur, 0, sizeof(cur));
	for (b = 0; b < batch_size; ++b) {
		const mem_alnreg_v *regs = &w->regs[start + b];
		for (k = 0; k < regs->n; ++k) {
			const mem_alnreg_t *ar = &regs->a[k];
			/* single-end output never prints a region below -T (bwamem.c:1015); the paired path can (mem_sam_pe takes the pair
			 * mem_pair chose, bwamem_pair.c:290-300) */
			if (ar->rb < 0 || ar->re < 0 || (ar->score < opt->T && !(opt->flag & MEM_F_PE))) continue;
			b200_cig_push_region(&cur, w, start + b, ar->qb, ar->qe, ar->rb, ar->re, ar->truesc, ar->w);
		}
	}
	b200_cig_rounds(t, tid, w, &cur);
}

/* ------------------------------------------------------------------------------------------------------------------
 * Mate-rescue look-ahead.  mem_matesw (bwamem_pair.c:109-176) runs one local alignment (ksw_align2)
 * per anchor region and admissible orientation whose pair is not yet consistent.  Its inputs — the mate (or its reverse
 * complement) and the reference window [rb, re) derived from the anchor and the insert-size statistics — are all known
 * once pass 1 and mem_pestat are done, so between mem_pestat and pass 2 the workers compute, for every pair, the jobs
 * mem_sam_pe can ask for (a superset: orientations that a rescue found earlier in the same pair makes it skip are still
 * computed), run them in batches on the GPU (ksw_b200_align_batch) and store the results in a table keyed by the EXACT
 * inputs of ksw_align2 (flags, query bytes, window bytes).  bwamem_pair.c is compiled, unedited, with
 * -Dksw_align2=b200_align2_hook: hit -> the stored kswr_t; miss -> one GPU call for that job.  KSW_B200_RESCUE=0: off
 * (the hook calls the reference's own ksw_align2). */
typedef struct b200_aln_entry {
	struct b200_aln_entry *next;
	uint64_t hash;
	int32_t qlen, tlen, xtra;
	kswr_t r;
	/* followed by: uint8_t query[qlen]; uint8_t target[tlen] */
} b200_aln_entry_t;
static b200_aln_entry_t **b200_aln_table;
static b200_arena_t b200_aln_arena[B200_MAX_WORKERS];
static long long b200_aln_hits, b200_aln_misses, b200_aln_jobs;
static double b200_t_rescue;
static ksw_b200_ctx_t *b200_aln_ctx[B200_MAX_GPUS];
static pthread_mutex_t b200_aln_mu[B200_MAX_GPUS];
static pthread_once_t b200_aln_once = PTHREAD_ONCE_INIT;
static void b200_aln_init(void) { int g; for (g = 0; g < B200_MAX_GPUS; ++g) pthread_mutex_init(&b200_aln_mu[g], 0); }

static int b200_rescue_on(void)
{
	static int v = -1;
	if (v < 0) { const char *e = getenv("KSW_B200_RESCUE"); v = !(e && e[0] == '0'); }
	return v;
}

static void b200_aln_reset(void)                         /* single-threaded: called between chunks */
{
	int t;
	if (!b200_aln_table) b200_aln_table = (b200_aln_entry_t **)calloc(B200_CIG_BUCKETS, sizeof(void *));
	else memset(b200_aln_table, 0, sizeof(void *) * B200_CIG_BUCKETS);
	for (t = 0; t < B200_MAX_WORKERS; ++t) { b200_aln_arena[t].cur = 0; b200_aln_arena[t].used = 0; }
}

static uint64_t b200_aln_hash(int qlen, const uint8_t *q, int tlen, const uint8_t *t, int xtra)
{
	uint64_t h = 1469598103934665603ull ^ ((uint64_t)(uint32_t)qlen << 32 | (uint32_t)tlen) ^ ((uint64_t)(uint32_t)xtra * 0x9E3779B97F4A7C15ull);
	int i;
	for (i = 0; i < qlen; ++i) h = (h ^ q[i]) * 1099511628211ull;
	for (i = 0; i + 8 <= tlen; i += 8) { uint64_t x; memcpy(&x, t + i, 8); h = (h ^ x) * 1099511628211ull; }
	for (; i < tlen; ++i) h = (h ^ t[i]) * 1099511628211ull;
	return h;
}

static const b200_aln_entry_t *b200_aln_find(uint64_t h, int qlen, const uint8_t *q, int tlen, const uint8_t *t, int xtra)
{
	const b200_aln_entry_t *e;
	if (!b200_aln_table) return 0;
	for (e = b200_aln_table[h & (B200_CIG_BUCKETS - 1)]; e; e = e->next) {
		const uint8_t *p = (const uint8_t *)(e + 1);
		if (e->hash != h || e->qlen != qlen || e->tlen != tlen || e->xtra != xtra) continue;
		if (memcmp(p, q, qlen) != 0 || memcmp(p + qlen, t, tlen) != 0) continue;
		return e;
	}
	return 0;
}

static void b200_aln_insert(int tid, int qlen, const uint8_t *q, int tlen, const uint8_t *t, int xtra, const ksw_b200_ares_t *r)
{
	b200_aln_entry_t *e = (b200_aln_entry_t *)b200_arena_alloc(&b200_aln_arena[tid], sizeof(b200_aln_entry_t) + (size_t)qlen + tlen);
	uint8_t *p;
	b200_aln_entry_t **slot;
	if (!e) return;
	p = (uint8_t *)(e + 1);
	e->hash = b200_aln_hash(qlen, q, tlen, t, xtra);
	e->qlen = qlen; e->tlen = tlen; e->xtra = xtra;
	e->r.score = r->score; e->r.te = r->te; e->r.qe = r->qe; e->r.score2 = r->score2; e->r.te2 = r->te2; e->r.tb = r->tb; e->r.qb = r->qb;
	memcpy(p, q, qlen); memcpy(p + qlen, t, tlen);
	slot = &b200_aln_table[e->hash & (B200_CIG_BUCKETS - 1)];
	do { e->next = *(b200_aln_entry_t * volatile *)slot; } while (!__sync_bool_compare_and_swap(slot, e->next, e));
}

/* one batch on GPU `gpu` through that GPU's alignment context (shared by the workers: the GPU runs one batch at a time anyway) */
static void b200_aln_run(int gpu, const mem_opt_t *opt, int64_t n, const ksw_b200_ajob_t *jobs, const uint8_t *q_, const uint8_t *t, ksw_b200_ares_t *res)
{
	ksw_b200_cfg_t cfg;
	pthread_once(&b200_aln_once, b200_aln_init);
	memcpy(cfg.mat, opt->mat, 25);
	cfg.m = 5; cfg.o_del = opt->o_del; cfg.e_del = opt->e_del; cfg.o_ins = opt->o_ins; cfg.e_ins = opt->e_ins; cfg.zdrop = 0; cfg.end_bonus = 0;
	if (b200_queue_on()) {                                   /* through the GPU's shared queue: concurrent batches are merged there */
		ksw_b200_queue_t *q = b200_queue_for(gpu, 0, 0);
		if (ksw_b200_queue_align(q, &cfg, n, jobs, q_, t, res) != 0) err_fatal(__func__, "GPU local alignment failed: %s", ksw_b200_queue_strerror(q));
		return;
	}
	pthread_mutex_lock(&b200_aln_mu[gpu]);
	if (!b200_aln_ctx[gpu]) {
		if (ksw_b200_ctx_create(gpu, &b200_aln_ctx[gpu]) != 0) err_fatal(__func__, "no usable CUDA device: the B200 path has no CPU fallback");
		ksw_b200_ctx_set_pack_threads(b200_aln_ctx[gpu], 2);
	}
	if (ksw_b200_align_batch(b200_aln_ctx[gpu], &cfg, n, jobs, q_, t, res) != 0)
		err_fatal(__func__, "GPU local alignment failed: %s", ksw_b200_strerror(b200_aln_ctx[gpu]));
	pthread_mutex_unlock(&b200_aln_mu[gpu]);
}

static const mem_opt_t *b200_aln_opt;                  /* the options of the running mem_process_seqs (for the miss path) */

kswr_t b200_align2_hook(int qlen, uint8_t *query, int tlen, uint8_t *target, int m, const int8_t *mat, int o_del, int e_del,
                        int o_ins, int e_ins, int xtra, kswq_t **qry)
{
	kswr_t r;
	ksw_b200_ajob_t job;
	ksw_b200_ares_t ar;
	const mem_opt_t *opt = b200_aln_opt;
	if (!b200_rescue_on() || qry || m != 5 || qlen < 1 || qlen > 4096 || !opt || memcmp(mat, opt->mat, 25) != 0 || o_del != opt->o_del ||
	    e_del != opt->e_del || o_ins != opt->o_ins || e_ins != opt->e_ins)
		return ksw_align2(qlen, query, tlen, target, m, mat, o_del, e_del, o_ins, e_ins, xtra, qry);   /* not mem_matesw's call */
	{
		const b200_aln_entry_t *e = b200_aln_find(b200_aln_hash(qlen, query, tlen, target, xtra), qlen, query, tlen, target, xtra);
		if (e) { __sync_fetch_and_add(&b200_aln_hits, 1); return e->r; }
	}
	/* miss: one job on the GPU */
	__sync_fetch_and_add(&b200_aln_misses, 1);
	if (b200_n_gpus < 0) b200_n_gpus = b200_gpu_count();
	if (b200_n_gpus < 1) err_fatal(__func__, "no usable CUDA device: the B200 path has no CPU fallback");
	job.q_off = 0; job.t_off = 0; job.qlen = qlen; job.tlen = tlen; job.xtra = xtra; job.reserved = 0;
	b200_aln_run(0, opt, 1, &job, query, target, &ar);
	r.score = ar.score; r.te = ar.te; r.qe = ar.qe; r.score2 = ar.score2; r.te2 = ar.te2; r.tb = ar.tb; r.qb = ar.qb;
	return r;
}

typedef struct { int64_t rb; int read, is_rev; } b200_aln_meta_t;        /* window start, index of the mate in seqs[], strand */
typedef struct { ksw_b200_ajob_t *jobs; b200_aln_meta_t *meta; int64_t n, m; uint8_t *q, *t; size_t nq, mq, nt, mt; } b200_aln_batch_t;

/* the orientation of two hits and their distance (the reference's mem_infer_dir, bwamem_pair.c:27-35, is static there) */
static int b200_infer_dir(int64_t l_pac, int64_t b1, int64_t b2, int64_t *dist)
{
	const int r1 = b1 >= l_pac, r2 = b2 >= l_pac;
	const int64_t p2 = r1 == r2 ? b2 : (l_pac << 1) - 1 - b2;         /* hit 2 on the strand of hit 1 */
	*dist = p2 > b1 ? p2 - b1 : b1 - p2;
	return (r1 == r2 ? 0 : 1) ^ (p2 > b1 ? 0 : 3);
}

/* the jobs mem_matesw(opt, l_pac, pac, pes, a, l_ms, ms, ma) can run, with `ma` as it is before any rescue (bwamem_pair.c:112-150) */
static void b200_aln_push_matesw(b200_aln_batch_t *B, const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, const mem_pestat_t pes[4],
                                 const mem_alnreg_t *a, int l_ms, const uint8_t *ms, const mem_alnreg_v *ma, int mate_read)
{
	const int64_t l_pac = bns->l_pac;
	int skip[4], r;
	size_t i;
	if (l_ms < 1 || l_ms > 4096) return;
	for (r = 0; r < 4; ++r) skip[r] = pes[r].failed ? 1 : 0;
	for (i = 0; i < ma->n; ++i) {
		int64_t dist;
		r = b200_infer_dir(l_pac, a->rb, ma->a[i].rb, &dist);
		if (dist >= pes[r].low && dist <= pes[r].high) skip[r] = 1;
	}
	for (r = 0; r < 4; ++r) {
		const int is_rev = (r >> 1) != (r & 1), is_larger = !(r >> 1);
		int64_t rb, re, len;
		uint8_t *ref;
		int k;
		if (skip[r]) continue;
		if (!is_rev) {
			rb = is_larger ? a->rb + pes[r].low : a->rb - pes[r].high;
			re = (is_larger ? a->rb + pes[r].high : a->rb - pes[r].low) + l_ms;
		} else {
			rb = (is_larger ? a->rb + pes[r].low : a->rb - pes[r].high) - l_ms;
			re = is_larger ? a->rb + pes[r].high : a->rb - pes[r].low;
		}
		if (rb < 0) rb = 0;
		if (re > l_pac << 1) re = l_pac << 1;
		ref = bns_get_seq(l_pac, pac, rb, re, &len);
		if (len == re - rb && len >= 0) {
			if (B->n == B->m) {
				B->m = B->m ? B->m << 1 : 1024;
				B->jobs = (ksw_b200_ajob_t *)realloc(B->jobs, sizeof(ksw_b200_ajob_t) * B->m);
				B->meta = (b200_aln_meta_t *)realloc(B->meta, sizeof(b200_aln_meta_t) * B->m);
			}
			if (B->nq + l_ms > B->mq) { B->mq = (B->mq ? B->mq << 1 : 1 << 18) + l_ms; B->q = (uint8_t *)realloc(B->q, B->mq); }
			if (B->nt + len > B->mt) { B->mt = (B->mt ? B->mt << 1 : 1 << 20) + len; B->t = (uint8_t *)realloc(B->t, B->mt); }
			if (is_rev) for (k = 0; k < l_ms; ++k) B->q[B->nq + l_ms - 1 - k] = ms[k] < 4 ? 3 - ms[k] : 4;
			else memcpy(B->q + B->nq, ms, l_ms);
			memcpy(B->t + B->nt, ref, len);
			B->jobs[B->n].q_off = B->nq; B->jobs[B->n].t_off = B->nt; B->jobs[B->n].qlen = l_ms; B->jobs[B->n].tlen = (int)len;
			B->jobs[B->n].xtra = KSW_XSUBO | KSW_XSTART | (l_ms * opt->a < 250 ? KSW_XBYTE : 0) | (opt->min_seed_len * opt->a);
			B->jobs[B->n].reserved = 0;
			B->meta[B->n].rb = rb; B->meta[B->n].read = mate_read; B->meta[B->n].is_rev = is_rev;
			B->nq += l_ms; B->nt += len; ++B->n;
		}
		free(ref);
	}
}

/* kt_for_batch body over PAIRS [start, start + batch_size): collect, run on this worker's GPU, fill the table */
static void worker_rescue_b200(void *data, int start, int batch_size, int tid)
{
	worker_t *w = (worker_t *)data;
	const mem_opt_t *opt = w->opt;
	b200_aln_batch_t B;
	ksw_b200_ares_t *res;
	double t0 = realtime();
	int p, i;
	int64_t k;
	size_t j;
	memset(&B, 0, sizeof(B));
	if (b200_n_gpus < 0) b200_n_gpus = b200_gpu_count();
	for (p = start; p < start + batch_size; ++p) {
		const bseq1_t *s = &w->seqs[p << 1];
		const mem_alnreg_v *a = &w->regs[p << 1];
		for (i = 0; i < 2; ++i) {                                          /* mem_sam_pe, bwamem_pair.c:254-261 */
			int n_anchor = 0;
			for (j = 0; j < a[i].n && n_anchor < opt->max_matesw; ++j) {
				if (a[i].a[j].score < a[i].a[0].score - opt->pen_unpaired) continue;
				++n_anchor;
				b200_aln_push_matesw(&B, opt, w->bns, w->pac, w->pes, &a[i].a[j], s[!i].l_seq, (const uint8_t *)s[!i].seq, &a[!i], (p << 1) + !i);
			}
		}
	}
	if (B.n > 0) {
		res = (ksw_b200_ares_t *)malloc(sizeof(ksw_b200_ares_t) * B.n);
		b200_aln_run(tid % (b200_n_gpus < B200_MAX_GPUS ? b200_n_gpus : B200_MAX_GPUS), opt, B.n, B.jobs, B.q, B.t, res);
		for (k = 0; k < B.n; ++k)
			b200_aln_insert(tid, B.jobs[k].qlen, B.q + B.jobs[k].q_off, B.jobs[k].tlen, B.t + B.jobs[k].t_off, B.jobs[k].xtra, &res[k]);
		__sync_fetch_and_add(&b200_aln_jobs, B.n);
		if (b200_cigar_mode() == 1) {
			/* the regions these rescues can add (bwamem_pair.c:151-160: truesc = 0 and w = 0 after the memset) will ask pass 2
			 * for their CIGAR like any other: compute those ahead too, or every one of them is a single-job miss */
			b200_cig_batch_t C;
			const int64_t l_pac = w->bns->l_pac;
			double t1 = realtime();
			memset(&C, 0, sizeof(C));
			for (k = 0; k < B.n; ++k) {
				const ksw_b200_ares_t *x = &res[k];
				const b200_aln_meta_t *mt = &B.meta[k];
				const int l_ms = B.jobs[k].qlen;
				int qb, qe;
				int64_t rb, re;
				if (x->score < opt->min_seed_len || x->qb < 0) continue;
				qb = mt->is_rev ? l_ms - (x->qe + 1) : x->qb;
				qe = mt->is_rev ? l_ms - x->qb : x->qe + 1;
				rb = mt->is_rev ? (l_pac << 1) - (mt->rb + x->te + 1) : mt->rb + x->tb;
				re = mt->is_rev ? (l_pac << 1) - (mt->rb + x->tb) : mt->rb + x->te + 1;
				b200_cig_push_region(&C, w, mt->read, qb, qe, rb, re, 0, 0);
			}
			b200_cig_rounds(b200_thread_state(opt, w->bns, w->pac, tid), tid, w, &C);
			b200_add_time(&b200_t_cigar, realtime() - t1);
		}
		free(res);
	}
	free(B.jobs); free(B.meta); free(B.q); free(B.t);
	b200_add_time(&b200_t_rescue, realtime() - t0);
}

typedef struct { int handle; int short_ok; mem_alnreg_t short_reg; } b200_chain_state_t;

/* KSW_B200_SCHED=rounds selects the exact minimal-work scheduler (only the seeds the reference extends are extended, in
 * as many GPU rounds as the longest per-read dependency chain); the default extends every seed in two passes and replays */
static int b200_use_rounds(void)
{
	static int v = -1;
	if (v < 0) { const char *e = getenv("KSW_B200_SCHED"); v = (e && strcmp(e, "rounds") == 0) ? 1 : 0; }
	return v;
}

static void worker1_b200(void *data, int start, int batch_size, int tid)
{
	worker_t *w = (worker_t *)data;
	const mem_opt_t *opt = w->opt;
	double tinit = realtime();
	b200_thread_t *t = b200_thread_state(opt, w->bns, w->pac, tid);
	mem_chain_v *chn = (mem_chain_v *)malloc(sizeof(mem_chain_v) * batch_size);
	b200_chain_state_t **cst = (b200_chain_state_t **)malloc(sizeof(void *) * batch_size);
	int b, i, rc;
	double t0 = realtime(), t1, ts = 0, tp = 0;
	b200_add_time(&b200_t_init, t0 - tinit);

	const int rounds = b200_use_rounds();
	b200_ext_plan_reset(t->plan);
	for (b = 0; b < batch_size; ++b) {
		bseq1_t *s = &w->seqs[start + b];
		int rd;
		t1 = realtime();
		for (i = 0; i < s->l_seq; ++i)                     /* bwamem.c:1093-1094 */
			s->seq[i] = s->seq[i] < 4 ? s->seq[i] : nst_nt4_table[(int)s->seq[i]];
		chn[b] = mem_chain(opt, w->bwt, w->bns->l_pac, s->l_seq, (uint8_t *)s->seq);
		chn[b].n = mem_chain_flt(opt, chn[b].n, chn[b].a);
		ts += realtime() - t1;
		rd = b200_ext_plan_add_read(t->plan, s->l_seq, (uint8_t *)s->seq);
		cst[b] = (b200_chain_state_t *)calloc(chn[b].n ? chn[b].n : 1, sizeof(b200_chain_state_t));
		for (i = 0; i < (int)chn[b].n; ++i) {
			mem_alnreg_v tmp;
			int ret;
			kv_init(tmp);
			ret = mem_chain2aln_short(opt, w->bns->l_pac, w->pac, s->l_seq, (uint8_t *)s->seq, &chn[b].a[i], &tmp);
			cst[b][i].handle = -1;
			if (ret == 0) {
				cst[b][i].short_ok = 1; cst[b][i].short_reg = tmp.a[0];
				if (rounds) b200_ext_plan_add_region(t->plan, rd, (const b200_alnreg_t *)&tmp.a[0]);   /* keeps its place in the read's timeline */
			} else if (ret > 0) cst[b][i].handle = b200_ext_plan_add_chain(t->plan, rd, (const b200_chain_t *)&chn[b].a[i]);
			free(tmp.a);
		}
	}
	tp = realtime() - t0 - ts;
	t1 = realtime();
	rc = rounds ? b200_ext_plan_run_rounds(t->plan, t->ctx) : b200_ext_plan_run(t->plan, t->ctx);
	if (rc != 0) err_fatal(__func__, "GPU extension pass failed (%d): %s", rc, t->queue ? ksw_b200_queue_strerror(t->queue) : ksw_b200_strerror(t->ctx));
	b200_add_time(&b200_t_seed, ts); b200_add_time(&b200_t_plan, tp); b200_add_time(&b200_t_gpu, realtime() - t1);
	t1 = realtime();
	for (b = 0; b < batch_size; ++b) {
		bseq1_t *s = &w->seqs[start + b];
		mem_alnreg_v regs;
		kv_init(regs);
		if (rounds) b200_ext_plan_take_regions(t->plan, b, (b200_alnreg_v *)&regs);     /* read handles are 0..batch_size-1 */
		for (i = 0; i < (int)chn[b].n; ++i) {
			if (!rounds) {
				if (cst[b][i].short_ok) kv_push(mem_alnreg_t, regs, cst[b][i].short_reg);
				else b200_ext_replay_chain(t->plan, cst[b][i].handle, (b200_alnreg_v *)&regs);
			}
			free(chn[b].a[i].seeds);
		}
		free(chn[b].a); free(cst[b]);
		regs.n = mem_sort_and_dedup(regs.n, regs.a, opt->mask_level_redun);
		if (opt->flag & MEM_F_NO_EXACT) regs.n = mem_test_and_remove_exact(opt, regs.n, regs.a, s->l_seq);
		w->regs[start + b] = regs;
	}
	free(chn); free(cst);
	b200_add_time(&b200_t_replay, realtime() - t1);
	if (b200_cigar_mode() == 1) {
		t1 = realtime();
		b200_cig_lookahead(t, tid, w, start, batch_size);
		b200_add_time(&b200_t_cigar, realtime() - t1);
	}
}

void mem_process_seqs(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, const uint8_t *pac, int64_t n_processed,
                      int n, bseq1_t *seqs, const mem_pestat_t *pes0)
{
	extern void kt_for(int n_threads, void (*func)(void *, int, int), void *data, int n);
	extern void kt_for_batch(int n_threads, void (*func)(void *, int, int, int), void *data, int n, int batch_size);
	worker_t w;
	mem_alnreg_v *regs;
	mem_pestat_t pes[4];
	double ctime, rtime, t_pass1 = 0;
	int batch = opt->batch_size, min_batch = 4096;
	const char *e = getenv("KSW_B200_MIN_BATCH");

	if (e) min_batch = atoi(e);
	if (batch < min_batch) batch = min_batch;
	ctime = cputime(); rtime = realtime();
	regs = malloc(n * sizeof(mem_alnreg_v));
	w.opt = opt; w.bwt = bwt; w.bns = bns; w.pac = pac;
	w.seqs = seqs; w.regs = regs; w.n_processed = n_processed;
	w.pes = &pes[0];
	if (b200_cigar_on()) b200_cig_reset();
	kt_for_batch(opt->n_threads, worker1_b200, &w, n, batch);                       /* pass 1: extension on the GPU */
	t_pass1 = realtime() - rtime;
	if (opt->flag & MEM_F_PE) {
		if (pes0) memcpy(pes, pes0, 4 * sizeof(mem_pestat_t));
		else mem_pestat(opt, bns->l_pac, n, regs, pes);
	}
	b200_aln_opt = opt;
	if ((opt->flag & MEM_F_PE) && !(opt->flag & MEM_F_NO_RESCUE) && b200_rescue_on()) {           /* mate-rescue look-ahead */
		b200_aln_reset();
		/* few pairs need rescue, and a GPU batch should carry thousands of jobs: one slice of the chunk's pairs per thread */
		const int per = ((n >> 1) + opt->n_threads - 1) / opt->n_threads;
		kt_for_batch(opt->n_threads, worker_rescue_b200, &w, n >> 1, per > 2048 ? per : 2048);
	}
	kt_for(opt->n_threads, worker2, &w, (opt->flag & MEM_F_PE) ? n >> 1 : n);          /* pass 2: unchanged */
	free(regs);
	if (bwa_verbose >= 3)
		fprintf(stderr, "[M::%s] Processed %d reads in %.3f CPU sec, %.3f real sec (extension on B200; thread-seconds so far: "
		        "init %.2f, seed+chain %.2f, plan %.2f, gpu passes %.2f, replay %.2f, cigar look-ahead %.2f; pass1 %.3f s real; "
		        "global alignments so far: %lld computed ahead, %lld hits, %lld misses; mate-rescue alignments: look-ahead %.2f thread-s, "
		        "%lld computed ahead, %lld hits, %lld misses)\n", __func__, n,
		        cputime() - ctime, realtime() - rtime, b200_t_init, b200_t_seed, b200_t_plan, b200_t_gpu, b200_t_replay, b200_t_cigar,
		        t_pass1, b200_cig_jobs, b200_cig_hits, b200_cig_misses, b200_t_rescue, b200_aln_jobs, b200_aln_hits, b200_aln_misses);
	if (bwa_verbose >= 3 && b200_queue_on() && b200_queue[0]) {
		int64_t nb = 0, ns = 0;
		ksw_b200_queue_stats(b200_queue[0], &nb, &ns);
		fprintf(stderr, "[M::%s] GPU 0 queue so far: %lld submissions of the workers ran as %lld merged batches\n", __func__, (long long)ns, (long long)nb);
	}
}
