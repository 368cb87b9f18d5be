/*
 * bwamem_b200.h — C ABI of the batched seed-extension driver: BWA-MEM 0.7.8's mem_chain2aln
 * (bwa-0.7.8/bwamem.c:730-878) restructured into batched GPU passes.
 *
 * The reference extends one seed at a time: left ksw_extend2 (bwamem.c:826), then right
 * ksw_extend2 seeded with the left score (bwamem.c:842,854), inside a per-seed loop whose only
 * cross-seed dependency is WHETHER a seed is extended (the containment test, bwamem.c:769-802).
 * The inputs of every extension depend only on (seed, query, the chain's reference window, options)
 * (SURVEY.md §3.3), so this driver
 *
 *   1. plans   : registers the left and right job of EVERY seed of every chain of a read batch
 *                (window math bwamem.c:740-758, slices bwamem.c:813-817 and :844);
 *   2. runs    : pass L (all left extensions, end_bonus = pen_clip5), the rare band retries
 *                (MAX_BAND_TRY, bwamem.c:818-829), pass R (all right extensions with h0 := left score,
 *                end_bonus = pen_clip3), its retries — each pass one ksw_b200_extend_batch call;
 *   3. replays : walks the reference's per-chain loop (bwamem.c:765-876) on the host, consuming the
 *                cached results, so that skipped seeds stay skipped and mem_alnreg_t comes out
 *                field-for-field identical.
 *
 * The record types mirror the reference's layouts so that a binding is a pointer cast:
 * b200_seed_t == mem_seed_t and b200_chain_t == mem_chain_t (bwamem.c:168-177),
 * b200_alnreg_t == mem_alnreg_t and b200_alnreg_v == mem_alnreg_v (bwamem.h:50-64).
 * b200_alnreg_v.a is grown with realloc() exactly like kv_pushp (kvec.h:83) so the caller's free() works.
 */
#ifndef BWAMEM_B200_H
#define BWAMEM_B200_H

#include <stdint.h>
#include <stddef.h>
#include "ksw_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {            /* mem_seed_t, bwamem.c:168-171 */
	int64_t rbeg;
	int32_t qbeg, len;
} b200_seed_t;

typedef struct {            /* mem_chain_t, bwamem.c:173-177 */
	int n, m;
	int64_t pos;
	b200_seed_t *seeds;
} b200_chain_t;

typedef struct {            /* mem_alnreg_t, bwamem.h:50-62 */
	int64_t rb, re;
	int qb, qe;
	int score;
	int truesc;
	int sub;
	int csub;
	int sub_n;
	int w;
	int seedcov;
	int secondary;
	uint64_t hash;
} b200_alnreg_t;

typedef struct { size_t n, m; b200_alnreg_t *a; } b200_alnreg_v;   /* mem_alnreg_v, bwamem.h:64 */

/* the fields of mem_opt_t (bwamem.h:21-48) the extension path reads */
typedef struct {
	int a, b;
	int o_del, e_del, o_ins, e_ins;
	int pen_clip5, pen_clip3;
	int w;
	int zdrop;
	int8_t mat[25];
} b200_ext_opt_t;

typedef struct b200_ext_plan b200_ext_plan_t;

/* pac: the 2-bit forward-strand reference (bntseq.c:191-192), l_pac bases; coordinates in
 * [l_pac, 2*l_pac) address the reverse strand (bntseq.c:355-376).  pac is not copied. */
b200_ext_plan_t *b200_ext_plan_create(const b200_ext_opt_t *opt, int64_t l_pac, const uint8_t *pac);
void b200_ext_plan_destroy(b200_ext_plan_t *p);
void b200_ext_plan_reset(b200_ext_plan_t *p);                 /* forget reads/chains, keep the memory */

/* registers a read (query = codes 0..4, as mem_align1_core encodes them, bwamem.c:1093-1094); returns its handle */
/* device-reference mode: on != 0 -> the plan no longer unpacks and copies the chains' reference
 * windows (bns_get_seq, bwamem.c:757); its jobs name their targets by coordinate and the GPU slices the .pac it keeps
 * (ksw_b200_ref_set is called for the run's context on first use; ksw_b200_extend_batch_ref runs the passes).  Same
 * regions, bit for bit.  Only between batches (after create or reset).  Default: off, or KSW_B200_REF=1 in the
 * environment. */
void b200_ext_plan_set_device_ref(b200_ext_plan_t *p, int on);
/* device-reference mode only: the passes are submitted to the GPU's shared queue (ksw_b200_queue_t, whose owner has
 * called ksw_b200_queue_ref_set) instead of the context given to the run functions, which may then be NULL */
void b200_ext_plan_set_queue(b200_ext_plan_t *p, ksw_b200_queue_t *q);
int b200_ext_plan_add_read(b200_ext_plan_t *p, int l_query, const uint8_t *query);
/* registers one chain of that read and the left/right job of every one of its seeds; returns the chain
 * handle (>= 0), or -1 for an empty chain (mem_chain2aln returns at once, bwamem.c:738) */
int b200_ext_plan_add_chain(b200_ext_plan_t *p, int read, const b200_chain_t *c);

/* runs the passes on the GPU; returns 0 or the ksw_b200 error code */
int b200_ext_plan_run(b200_ext_plan_t *p, ksw_b200_ctx_t *ctx);

/* same effect on *av as mem_chain2aln(opt, l_pac, pac, l_query, query, c, av) for the registered chain */
void b200_ext_replay_chain(const b200_ext_plan_t *p, int chain, b200_alnreg_v *av);

/* ---- rounds mode: exact, minimal-work scheduling (strategy B) ----
 * Instead of extending every seed and replaying, every read walks its timeline (its chains, and the regions other code
 * inserted between them — mem_chain2aln_short in the reference flow — registered with b200_ext_plan_add_region in
 * order) like the reference's sequential code and stops whenever it needs a DP result; one round = one batched GPU
 * call with the pending job of every active read.  Only the seeds the reference would extend are extended.
 * Usage: add_read, then add_chain / add_region in order, for every read; run_rounds; take_regions per read. */
int b200_ext_plan_add_region(b200_ext_plan_t *p, int read, const b200_alnreg_t *reg);
int b200_ext_plan_run_rounds(b200_ext_plan_t *p, ksw_b200_ctx_t *ctx);
void b200_ext_plan_take_regions(b200_ext_plan_t *p, int read, b200_alnreg_v *out);   /* caller frees out->a with free() */
int64_t b200_ext_plan_rounds(const b200_ext_plan_t *p);                                /* GPU rounds of the last run_rounds */

/* counters of the last run: seeds registered, left / right jobs computed, band retries */
void b200_ext_plan_stats(const b200_ext_plan_t *p, int64_t *n_seeds, int64_t *n_left, int64_t *n_right, int64_t *n_retry);

/* Convenience for callers that do not interleave anything between chains: for every read r, append to
 * av[r] what the reference loop "for each chain: mem_chain2aln(..., &av[r])" would append. */
typedef struct {
	int l_query;
	const uint8_t *query;
	int n_chains;
	const b200_chain_t *chains;
} b200_read_t;
int b200_chain2aln_batch(ksw_b200_ctx_t *ctx, const b200_ext_opt_t *opt, int64_t l_pac, const uint8_t *pac,
                         int n_reads, const b200_read_t *reads, b200_alnreg_v *av);

/* Flat-array form of b200_chain2aln_batch for bindings (ctypes / cgo slices): reads are (read_off, read_len)
 * into qpool; chains are (chain_read, chain_seed0, chain_nseeds) into seeds[], listed read by read.  Regions are
 * written read by read into out[0..*n_out), out_read[k] = read of region k.  Returns 0, a ksw_b200 error code,
 * or -1 if out_cap is too small. */
int b200_chain2aln_flat(ksw_b200_ctx_t *ctx, const b200_ext_opt_t *opt, int64_t l_pac, const uint8_t *pac,
                        int n_reads, const int64_t *read_off, const int32_t *read_len, const uint8_t *qpool,
                        int n_chains, const int32_t *chain_read, const int64_t *chain_seed0, const int32_t *chain_nseeds,
                        const b200_seed_t *seeds, int64_t out_cap, b200_alnreg_t *out, int32_t *out_read, int64_t *n_out);

/* the same with the rounds scheduler; *n_jobs_out (may be NULL) = DP jobs actually run */
int b200_chain2aln_flat_rounds(ksw_b200_ctx_t *ctx, const b200_ext_opt_t *opt, int64_t l_pac, const uint8_t *pac,
                               int n_reads, const int64_t *read_off, const int32_t *read_len, const uint8_t *qpool,
                               int n_chains, const int32_t *chain_read, const int64_t *chain_seed0, const int32_t *chain_nseeds,
                               const b200_seed_t *seeds, int64_t out_cap, b200_alnreg_t *out, int32_t *out_read, int64_t *n_out,
                               int64_t *n_jobs_out);

/* the reference's reference-slice fetch (bns_get_seq, bntseq.c:355-376) as the driver uses it; returns the
 * number of bases written (0 when [beg,end) bridges the forward/reverse boundary) */
int64_t b200_get_ref_slice(int64_t l_pac, const uint8_t *pac, int64_t beg, int64_t end, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif /* BWAMEM_B200_H */
