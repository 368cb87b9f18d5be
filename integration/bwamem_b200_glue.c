/*
 * integration/bwamem_b200_glue.c — the reference-side binding: `bwa mem` with its seed-extension
 * pass running on the B200 library.
 *
 * How it is built (oracle/Makefile, target _ref/bwa_b200): this one file is compiled INSTEAD of the
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
 * Everything else (seeding, chaining, pairing, mate rescue, CIGAR, SAM text) is the reference's own code,
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
	ksw_b200_ctx_t *ctx;
	b200_ext_plan_t *plan;
	const uint8_t *pac;
} b200_thread_t;

/* kt_for_batch creates fresh pthreads for every chunk (kthread_batch.c:54), so the per-worker state is kept
 * in a table indexed by the worker id, not in thread-local storage */
#define B200_MAX_WORKERS 1024
static b200_thread_t b200_workers[B200_MAX_WORKERS];
static int b200_n_gpus = -1;
/* wall-clock seconds summed over worker threads: seeding+chaining, planning, GPU passes, replay+dedup */
static double b200_t_seed, b200_t_plan, b200_t_gpu, b200_t_replay, b200_t_init;
static void b200_add_time(double *acc, double dt) { __sync_synchronize(); { static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER; pthread_mutex_lock(&mu); *acc += dt; pthread_mutex_unlock(&mu); } }

static b200_thread_t *b200_thread_state(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, int tid)
{
	b200_thread_t *t;
	if (tid < 0 || tid >= B200_MAX_WORKERS) err_fatal(__func__, "more than %d worker threads", B200_MAX_WORKERS);
	t = &b200_workers[tid];
	if (!t->ctx) {
		if (b200_n_gpus < 0) b200_n_gpus = ksw_b200_device_count();
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
	}
	return t;
}

/* CUDA context creation costs seconds on a large GPU; start it while `bwa mem` is still loading the index.
 * (glibc passes argc/argv to constructors.) */
static void *b200_warmup_thread(void *arg)
{
	int d, n = ksw_b200_device_count();
	(void)arg;
	for (d = 0; d < n; ++d) {
		ksw_b200_ctx_t *c = 0;
		if (ksw_b200_ctx_create(d, &c) == 0) ksw_b200_ctx_destroy(c);
	}
	return 0;
}
__attribute__((constructor)) static void b200_warmup(int argc, char **argv)
{
	pthread_t th;
	if (argc >= 2 && strcmp(argv[1], "mem") == 0 && !getenv("KSW_B200_NO_WARMUP"))
		if (pthread_create(&th, 0, b200_warmup_thread, 0) == 0) pthread_detach(th);
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
	if (rc != 0) err_fatal(__func__, "GPU extension pass failed (%d): %s", rc, ksw_b200_strerror(t->ctx));
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
	kt_for_batch(opt->n_threads, worker1_b200, &w, n, batch);                       /* pass 1: extension on the GPU */
	t_pass1 = realtime() - rtime;
	if (opt->flag & MEM_F_PE) {
		if (pes0) memcpy(pes, pes0, 4 * sizeof(mem_pestat_t));
		else mem_pestat(opt, bns->l_pac, n, regs, pes);
	}
	kt_for(opt->n_threads, worker2, &w, (opt->flag & MEM_F_PE) ? n >> 1 : n);          /* pass 2: unchanged */
	free(regs);
	if (bwa_verbose >= 3)
		fprintf(stderr, "[M::%s] Processed %d reads in %.3f CPU sec, %.3f real sec (extension on B200; thread-seconds so far: "
		        "init %.2f, seed+chain %.2f, plan %.2f, gpu passes %.2f, replay %.2f; pass1 %.3f s real)\n", __func__, n,
		        cputime() - ctime, realtime() - rtime, b200_t_init, b200_t_seed, b200_t_plan, b200_t_gpu, b200_t_replay, t_pass1);
}
