/*
 * ksw_b200.h — C ABI of the B200-native seed-extension path.
 *
 * This is the drop-in boundary for BWA-MEM 0.7.8's banded affine-gap extension
 * (reference: bwa-0.7.8/ksw.c:379-481, declared at bwa-0.7.8/ksw.h:107-108):
 *
 *   - ksw_extend / ksw_extend2 keep the reference's scalar signatures verbatim
 *     (ksw.h:107-108), but run the one job on the GPU.  No CPU fallback exists:
 *     if no CUDA device / kernel image is available the call aborts like the
 *     reference's err_fatal (bwa-0.7.8/utils.c:90).
 *   - ksw_b200_extend_batch is the batched entry that sits beside them.  It
 *     replaces a loop of ksw_extend2 calls (the two call sites in
 *     mem_chain2aln, bwa-0.7.8/bwamem.c:826 and :854).  The record layout
 *     follows the fork's own, unused, accelerator sketch ext_param_t /
 *     ext_res_t (bwamem.c:553-577) but is per *side*, because the right
 *     extension's h0 is the left extension's score (bwamem.c:842,854).
 *   - ksw_global / ksw_global2 (ksw.h:83-84, the CIGAR generator behind
 *     bwa_gen_cigar2, bwa.c:132) and their batched form ksw_b200_global_batch:
 *     banded global alignment with backtrace, SURVEY.md §8(f) rank 2.
 *
 * Plain C types only; no CUDA or torch types cross this boundary.
 * All functions return 0 on success and a non-zero code on a CUDA/usage error
 * (message via ksw_b200_strerror); the scalar reference API has no error path,
 * so the scalar wrappers treat any error as fatal.
 */
#ifndef KSW_B200_H
#define KSW_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- records ------------------------------------------------------------ */

/* Scoring shared by every job of a batch == the by-value arguments every
 * ksw_extend2 call in mem_chain2aln passes from mem_opt_t (bwamem.c:826,854):
 * mat (opt->mat, 5x5, indexed [target*5+query], bwa.c:77-86), gap costs,
 * zdrop, and end_bonus (pen_clip5 for a left pass, pen_clip3 for a right pass). */
typedef struct {
	int8_t  mat[25];
	int32_t m;                 /* alphabet size; must be 5 (every reference caller passes 5) */
	int32_t o_del, e_del, o_ins, e_ins;
	int32_t zdrop;             /* <= 0 disables z-drop (ksw.c:455) */
	int32_t end_bonus;
} ksw_b200_cfg_t;

/* One extension job (one side of one seed).  Sequences are byte codes 0..4
 * (A,C,G,T,N) exactly as the reference passes them; they live in two caller-owned
 * pools so that a batch is three flat arrays. */
typedef struct {
	uint64_t q_off, t_off;     /* byte offsets of query / target in the pools */
	int32_t  qlen, tlen;       /* qlen >= 1, tlen >= 0 */
	int32_t  h0;               /* score carried in (seed score or the left score) */
	int32_t  w;                /* band width before the reference's clamp (ksw.c:398-406) */
} ksw_b200_job_t;

/* The six outputs of ksw_extend2 (return value, *_qle, *_tle, *_gtle, *_gscore, *_max_off). */
typedef struct {
	int32_t score, qle, tle, gtle, gscore, max_off;
} ksw_b200_res_t;

typedef struct ksw_b200_ctx ksw_b200_ctx_t;       /* one per (host thread, GPU) */
typedef struct ksw_b200_batch ksw_b200_batch_t;   /* a packed batch resident in HBM */

/* ---- scalar entries: the reference signatures, unchanged (ksw.h:107-108) -- */
int ksw_extend(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
               int gapo, int gape, int w, int end_bonus, int zdrop, int h0,
               int *qle, int *tle, int *gtle, int *gscore, int *max_off);
int ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                int *qle, int *tle, int *gtle, int *gscore, int *max_off);

/* ---- contexts -------------------------------------------------------------- */
int  ksw_b200_device_count(void);
int  ksw_b200_ctx_create(int device, ksw_b200_ctx_t **out);
void ksw_b200_ctx_destroy(ksw_b200_ctx_t *ctx);
const char *ksw_b200_strerror(const ksw_b200_ctx_t *ctx);   /* last error text of this ctx ("" if none) */
/* number of host threads used to pack sequences (default: min(hardware threads, 32)) */
int  ksw_b200_ctx_set_pack_threads(ksw_b200_ctx_t *ctx, int n_threads);
/* jobs per pipeline chunk of ksw_b200_extend_batch (default 2^20): chunk c+1 is packed on the host while
 * chunk c is copied and computed on the GPU */
int  ksw_b200_ctx_set_chunk_jobs(ksw_b200_ctx_t *ctx, int64_t chunk_jobs);
/* cumulative number of kernels this ctx has launched (for bench accounting) */
int64_t ksw_b200_ctx_launch_count(const ksw_b200_ctx_t *ctx);

/* ---- batched entry: host buffers in, host results out ----------------------- */
/* Packs (2-bit + N masks), length-bins, copies H2D from pinned staging on the
 * ctx stream, runs the extension kernels, copies results D2H into res[0..n) in the
 * caller's job order, and returns when they are there.  res need not be pinned. */
int ksw_b200_extend_batch(ksw_b200_ctx_t *ctx, const ksw_b200_cfg_t *cfg, int64_t n,
                          const ksw_b200_job_t *jobs, const uint8_t *qpool, const uint8_t *tpool,
                          ksw_b200_res_t *res);

/* Multi-GPU form: the batch is cut into n_ctx contiguous ranges of (nearly) equal DP cells (sum of qlen*tlen), range r
 * runs on ctxs[r] (normally one context per GPU) on its own host thread; no exchange step, results land in res. */
int ksw_b200_extend_batch_multi(int n_ctx, ksw_b200_ctx_t **ctxs, const ksw_b200_cfg_t *cfg, int64_t n,
                                const ksw_b200_job_t *jobs, const uint8_t *qpool, const uint8_t *tpool,
                                ksw_b200_res_t *res);

/* ---- asynchronous batched entry for page-locked caller buffers (SURVEY.md 8(b) "batched entry to add") ---------- */
/* The caller owns pinned job / sequence / result arrays (allocated with ksw_b200_host_alloc, or its own memory pinned
 * with ksw_b200_host_register).  ksw_b200_extend_batch_async hands the batch to the context and returns at once; the
 * host does not touch the sequences: raw job records and raw byte-coded sequences are copied to the GPU as they are
 * (cudaMemcpyAsync, chunk by chunk, on the context's streams), 2-bit packing, binning and the extension kernels run
 * on the device, and the results are copied straight into res.  ksw_b200_wait returns when res[0..n) is complete.
 * One batch in flight per context; jobs, qpool, tpool and res must stay valid and unmodified until ksw_b200_wait
 * returns; no other call on this context in between.  qpool_bytes / tpool_bytes: sizes of the two pools (every job
 * must lie inside them).  Errors: 3 = a buffer is not page-locked, 4 = a batch is already in flight; CUDA and job
 * errors are reported by ksw_b200_wait.  Results are bit-identical to ksw_b200_extend_batch. */
void *ksw_b200_host_alloc(size_t bytes);
void  ksw_b200_host_free(void *p);
int   ksw_b200_host_register(void *p, size_t bytes);
int   ksw_b200_host_unregister(void *p);
int ksw_b200_extend_batch_async(ksw_b200_ctx_t *ctx, const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_job_t *jobs,
                                const uint8_t *qpool, size_t qpool_bytes, const uint8_t *tpool, size_t tpool_bytes,
                                ksw_b200_res_t *res);
int ksw_b200_wait(ksw_b200_ctx_t *ctx);

/* ---- extension against a reference kept on the device ------------------------------------- */
/* Replaces bns_get_seq (bwa-0.7.8/bntseq.c:355-376) + the window slicing of mem_chain2aln (bwamem.c:740-758, 813-817,
 * 844) on the host: the forward-strand 2-bit .pac (bwa_idx_load, bwa.c:291: l_pac/4+1 bytes, base k in byte k/4 at bits
 * 2*(~k&3), bntseq.c:191-192) is uploaded once per device (shared by all contexts of that device), and a job names its
 * target as a run of the doubled coordinate space [0, 2*l_pac): x < l_pac is the forward strand, x >= l_pac the
 * reverse strand = 3 - pac(2*l_pac - 1 - x).  A left extension (bwamem.c:813-817) reads both sequences downwards. */
typedef struct {
	uint64_t q_off;            /* byte offset in qpool of the FIRST query base (codes 0..4) */
	int64_t  t_pos;            /* coordinate of the FIRST target base in the doubled reference space */
	int32_t  qlen, tlen;       /* qlen >= 1, tlen >= 0; the target run must stay on one strand (bwamem.c:752-755) */
	int32_t  h0, w;
	int8_t   q_step, t_step;   /* +1: the run goes upwards from q_off / t_pos, -1: downwards */
	int8_t   reserved[6];
} ksw_b200_rjob_t;

/* pac is copied to the device; the host buffer is not needed afterwards.  Calling it again with the same pointer and
 * length on another context of the same device shares the copy. */
int ksw_b200_ref_set(ksw_b200_ctx_t *ctx, const uint8_t *pac, int64_t l_pac);
/* results as ksw_b200_extend_batch on the same sequences; qpool holds the reads (byte codes), qpool_bytes its size */
int ksw_b200_extend_batch_ref(ksw_b200_ctx_t *ctx, const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_rjob_t *jobs,
                              const uint8_t *qpool, size_t qpool_bytes, ksw_b200_res_t *res);

/* the same for a batch that arrives in several segments (one per submitting host thread): they run as ONE GPU batch */
typedef struct {
	int64_t n;
	const ksw_b200_rjob_t *jobs;
	const uint8_t *qpool;
	size_t qpool_bytes;
	ksw_b200_res_t *res;
} ksw_b200_rseg_t;
int ksw_b200_extend_batch_ref_segs(ksw_b200_ctx_t *ctx, const ksw_b200_cfg_t *cfg, int n_segs, const ksw_b200_rseg_t *segs);

/* bytes the last ksw_b200_extend_batch / ksw_b200_extend_batch_async call copied host->device and device->host */
int ksw_b200_ctx_last_transfer(const ksw_b200_ctx_t *ctx, int64_t *h2d_bytes, int64_t *d2h_bytes);

/* ---- split form: upload once, run many times (bench / two-pass drivers) ----- */
int ksw_b200_batch_upload(ksw_b200_ctx_t *ctx, const ksw_b200_cfg_t *cfg, int64_t n,
                          const ksw_b200_job_t *jobs, const uint8_t *qpool, const uint8_t *tpool,
                          ksw_b200_batch_t **out);
/* enqueue the kernels for a resident batch on the ctx stream (asynchronous) */
int ksw_b200_batch_run(ksw_b200_ctx_t *ctx, ksw_b200_batch_t *b);
/* run `iters` times, each bracketed by CUDA events on the ctx stream; ms[i] = device time of run i.  A run is everything
 * the GPU does for a packed batch: the binning (key kernel + radix sort) and the extension kernels */
int ksw_b200_batch_run_timed(ksw_b200_ctx_t *ctx, ksw_b200_batch_t *b, int iters, float *ms);
/* the same, and ms_ext[i] (if not NULL) = the part of run i spent in the extension kernels alone (an event between the
 * binning and the first extension launch): the launch duration the roofline figure of the benchmark divides by */
int ksw_b200_batch_run_timed2(ksw_b200_ctx_t *ctx, ksw_b200_batch_t *b, int iters, float *ms, float *ms_ext);
/* wait for the stream and copy results (caller's job order) to host */
int ksw_b200_batch_download(ksw_b200_ctx_t *ctx, ksw_b200_batch_t *b, ksw_b200_res_t *res);
/* visited DP cells per job (caller's order): sum over the rows the reference loop executes of
 * (end - beg), ksw.c:418-421 — the unit of the GCUPS metric.  Valid after a run. */
int ksw_b200_batch_download_cells(ksw_b200_ctx_t *ctx, ksw_b200_batch_t *b, uint32_t *cells);
/* statistics of a resident batch: n_fast / n_generic jobs, packed bytes in HBM */
int ksw_b200_batch_info(const ksw_b200_batch_t *b, int64_t *n_fast, int64_t *n_generic, int64_t *packed_bytes);
void ksw_b200_batch_free(ksw_b200_ctx_t *ctx, ksw_b200_batch_t *b);
int ksw_b200_ctx_sync(ksw_b200_ctx_t *ctx);

/* ---- banded global alignment with backtrace (CIGAR generation) ---------------- */
/* SURVEY.md §8(f) rank 2.  Replaces a loop of ksw_global2 calls (bwa-0.7.8/ksw.c:501-584; the call site is
 * bwa_gen_cigar2, bwa.c:132, reached from mem_reg2aln, bwamem.c:1196).  Scalar signatures unchanged (ksw.h:83-84):
 * *cigar_ is malloc'd and owned by the caller, one uint32 per operation, len<<4|op with op 0=M 1=I 2=D. */
int ksw_global(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
               int gapo, int gape, int w, int *n_cigar, uint32_t **cigar);
int ksw_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                int o_del, int e_del, int o_ins, int e_ins, int w, int *n_cigar, uint32_t **cigar);

typedef struct {
	uint64_t q_off, t_off;     /* byte offsets of query / target (codes 0..4) in the pools */
	int32_t  qlen, tlen;       /* >= 0 */
	int32_t  w;                /* band width, >= 0; as in the reference the result is only defined if the band holds
	                            * the end cell, |tlen - qlen| <= w (bwa_gen_cigar2 guarantees it, bwa.c:124-126) */
	int32_t  reserved;
} ksw_b200_gjob_t;

typedef struct {
	int32_t score;             /* return value of ksw_global2 */
	int32_t n_cigar;           /* number of CIGAR operations */
	int64_t cigar_off;         /* index of the first one in the pool returned by the call */
} ksw_b200_gres_t;

/* cfg: mat, m (= 5), o_del, e_del, o_ins, e_ins are used; zdrop and end_bonus are ignored.  res[0..n) in the caller's
 * order; *cigar_pool points to memory owned by ctx, valid until the next ksw_b200_global_batch call on this ctx or its
 * destruction; *n_cigar_total = number of operations in it. */
int ksw_b200_global_batch(ksw_b200_ctx_t *ctx, const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_gjob_t *jobs,
                          const uint8_t *qpool, const uint8_t *tpool, ksw_b200_gres_t *res,
                          const uint32_t **cigar_pool, int64_t *n_cigar_total);

/* ---- local alignment with start positions and second-best score ----------------------------- */
/* Replaces the reference's ksw_align2 (bwa-0.7.8/ksw.h:62, ksw.c:329-354; the striped SSE2 kernels ksw_u8 / ksw_i16 of
 * ksw.c:110-320), which mem_matesw calls once per rescue window (bwamem_pair.c:150).  Results are those of the reference
 * bit for bit, including what its vector layout makes observable (see csrc/ksw_align.cu).  Sequences are byte codes 0..4.
 * Domain: qlen >= 1 (<= 4096), and a KSW_XBYTE job must not saturate the byte kernel (qlen * max score + shift < 255, which
 * mem_matesw guarantees, bwamem_pair.c:147): if it does, the job ends after the forward pass with score = 255, where the
 * reference goes on with qe = -1 (undefined). */
#define KSW_B200_XBYTE  0x10000    /* ksw.h:6-9: the flags of `xtra`; its low 16 bits are the score threshold */
#define KSW_B200_XSTOP  0x20000
#define KSW_B200_XSUBO  0x40000
#define KSW_B200_XSTART 0x80000
typedef struct {
	uint64_t q_off, t_off;     /* byte offsets into qpool / tpool */
	int32_t  qlen, tlen;
	int32_t  xtra;             /* as passed to ksw_align2 */
	int32_t  reserved;
} ksw_b200_ajob_t;

typedef struct {               /* kswr_t, ksw.h:30-36 (unset fields are -1 like g_defr, ksw.c:43) */
	int32_t score, te, qe, score2, te2, tb, qb;
	int32_t reserved;
} ksw_b200_ares_t;

/* Scalar drop-ins with the reference's signatures (ksw.h:62-63; kswr_t, ksw.h:30-36, by value).  One job = one GPU round
 * trip; the sequences are not modified (the reference reverses them in place and back).  qry (the reference's cached query
 * profile) is not used: if given, *qry is left as it is. */
typedef struct { int score; int te, qe; int score2, te2; int tb, qb; } ksw_b200_kswr_t;
#ifndef __AC_KSW_H              /* a unit that has included the reference's ksw.h already has these two, with kswr_t / kswq_t */
ksw_b200_kswr_t ksw_align2(int qlen, uint8_t *query, int tlen, uint8_t *target, int m, const int8_t *mat,
                           int o_del, int e_del, int o_ins, int e_ins, int xtra, void **qry);
ksw_b200_kswr_t ksw_align(int qlen, uint8_t *query, int tlen, uint8_t *target, int m, const int8_t *mat,
                          int gapo, int gape, int xtra, void **qry);
#endif

/* cfg: mat, m (= 5), o_del, e_del, o_ins, e_ins are used */
int ksw_b200_align_batch(ksw_b200_ctx_t *ctx, const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_ajob_t *jobs,
                         const uint8_t *qpool, const uint8_t *tpool, ksw_b200_ares_t *res);

/* ---- one submission queue per GPU, shared by all host threads (cross-thread batch coalescing) - */
/* The reference's workers (kt_for_batch, kthread_batch.c:18-56) each own a slice of the reads; with one private context
 * per worker every worker launches its own small batches.  A queue owns two contexts ("lanes") on its device and no thread:
 * any number of host threads submit (blocking, thread-safe); a submitter that finds a lane free leads a batch — whatever
 * has been submitted meanwhile, of the same kind and scoring, runs as ONE merged batch from its thread — and every submitter
 * gets exactly the results of its own jobs (same read -> same regs[i]).  GPU memory is that of the lanes, however many
 * threads submit (KSW_B200_QUEUE_LANES, default 2). */
typedef struct ksw_b200_queue ksw_b200_queue_t;
int  ksw_b200_queue_create(int device, ksw_b200_queue_t **out);
void ksw_b200_queue_destroy(ksw_b200_queue_t *q);
int  ksw_b200_queue_ref_set(ksw_b200_queue_t *q, const uint8_t *pac, int64_t l_pac);
/* as ksw_b200_extend_batch_ref */
int  ksw_b200_queue_extend_ref(ksw_b200_queue_t *q, const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_rjob_t *jobs,
                               const uint8_t *qpool, size_t qpool_bytes, ksw_b200_res_t *res);
/* as ksw_b200_global_batch, but the CIGAR pool is a malloc'd array the caller owns (free() it): *cigar_pool, *n_cigar_total */
int  ksw_b200_queue_global(ksw_b200_queue_t *q, const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_gjob_t *jobs,
                           const uint8_t *qpool, const uint8_t *tpool, ksw_b200_gres_t *res,
                           uint32_t **cigar_pool, int64_t *n_cigar_total);
/* as ksw_b200_align_batch */
int  ksw_b200_queue_align(ksw_b200_queue_t *q, const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_ajob_t *jobs,
                          const uint8_t *qpool, const uint8_t *tpool, ksw_b200_ares_t *res);
const char *ksw_b200_queue_strerror(const ksw_b200_queue_t *q);
/* batches the queue has run and submissions they carried (for the record: submissions / batches = coalescing factor) */
void ksw_b200_queue_stats(const ksw_b200_queue_t *q, int64_t *n_batches, int64_t *n_submissions);

/* ---- measurement helper ------------------------------------------------------ */
/* Issue-rate microbenchmark of the DPX family used by the kernels (VIADDMNMX.S16x2[.RELU],
 * VIMNMX.S16x2): returns warp-level lane-operations per second sustained by the whole GPU
 * (SURVEY.md §8d "peak to divide by").  which: 0 = s16x2 mix, 1 = s32 mix. */
int ksw_b200_dpx_peak(ksw_b200_ctx_t *ctx, int which, double *lane_ops_per_s, float *ms);

/* the reference's band clamp (ksw.c:398-406), exposed for tests of the host packer */
int ksw_b200_clamp_w(int qlen, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
                     int w, int end_bonus);

#ifdef __cplusplus
}
#endif
#endif /* KSW_B200_H */
