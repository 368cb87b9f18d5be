// ksw_dev.cuh — device-side records shared by the kernels and the host runtime.
//
// HBM layout of one packed batch (built by ksw_pack.cpp, consumed by the kernels):
//
//   d_jobs : DevJob[n]      32 B each, in the CALLER's order (the host packer only streams)
//   d_order: uint32[n]      the *binned* order, built on the device (ksw_bin.cu): fast classes first, then
//                           generic; inside a class sorted by (rows, carried-in score) so that the jobs a
//                           warp draws together sweep bands of similar width and length
//   d_pool : uint32[]       per job, 16-byte aligned:  query 2-bit words | target 2-bit words
//                           base k of a sequence sits in word k/16 at bits 2*(k%16)
//   d_npool: uint32[]       only for the (rare) jobs whose query or target holds an N (code 4):
//                           [query N-mask words][target N-mask words], 1 bit per base; the base
//                           itself is stored as 0 in the 2-bit stream
//   d_res  : DevRes[n]      24 B each, indexed by the CALLER's job index (DevJob::idx)
//
// Algorithmic HBM bytes per job = 32 (record) + 16*ceil((ceil(qlen/16)+ceil(tlen/16))*4/16) (sequences)
//                                 + 24 (result); for 101x101: 32 + 64 + 24 = 120 B.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KSW_HD __host__ __device__ __forceinline__
#else
#define KSW_HD inline
#endif

struct DevJob {
	uint32_t seq_off;   // offset of the job's packed sequences in d_pool, in 16-byte units
	uint32_t idx;       // caller's index of this job
	int32_t  qlen, tlen;
	int32_t  h0;        // already max(h0,0)              (ksw.c:384)
	int32_t  w;         // already clamped by the reference rule (ksw.c:398-406), done on the host
	uint32_t flags;     // bit0: query N-mask present, bit1: target N-mask present, bits 8-11: kernel class (KSW_CLASS_*)
	uint32_t nmask_off; // word offset of the job's N masks in d_npool (valid iff flags != 0)
};
static_assert(sizeof(DevJob) == 32, "DevJob must be 32 bytes");

struct DevRes {
	int32_t score, qle, tle, gtle, gscore, max_off;
};
static_assert(sizeof(DevRes) == 24, "DevRes must be 24 bytes");

// banded global alignment with backtrace (ksw_global.cu): sequences travel as byte codes, query then target
struct DevGJob {
	uint64_t seq_off;   // byte offset of the job's query in the chunk's sequence buffer (the target follows it)
	int32_t  qlen, tlen, w;
	uint32_t idx;       // index of this job in the chunk
};
static_assert(sizeof(DevGJob) == 24, "DevGJob must be 24 bytes");

// fast global-alignment kernel (ksw_gfast.cu): 32 jobs of similar size per warp; their H values go to one lane-interleaved slab
struct DevGGroup {
	long long z_off;    // first uint2 of the group's slab: element (row i, band quad r, lane l) at z_off + (i * nqb + r) * 32 + l
	int32_t rows, nqb;  // longest target / widest band (in quads) of the group
	int32_t first, n;   // the group's jobs are gorder[first .. first + n), n <= 32
};
static_assert(sizeof(DevGGroup) == 24, "DevGGroup must be 24 bytes");

// local alignment (ksw_align.cu): byte codes, query then target, like DevGJob
struct DevAJob {
	uint64_t seq_off;
	int32_t  qlen, tlen, xtra;   // xtra: the reference's KSW_X* flags | threshold (ksw.h:6-9)
	uint32_t idx;
};
static_assert(sizeof(DevAJob) == 24, "DevAJob must be 24 bytes");
struct DevARes { int32_t score, te, qe, score2, te2, tb, qb, pad; };      // kswr_t, ksw.h:30-36
static_assert(sizeof(DevARes) == 32, "DevARes must be 32 bytes");
struct KswAlignParams {
	int8_t  mat[25];
	int8_t  pad[3];
	int32_t o_del, e_del, o_ins, e_ins;
	int32_t shift, qmax;         // ksw_qinit's q->shift and q->max (ksw.c:77-83)
};

struct DevGRes {
	int32_t   score, n_cigar;
	long long cigar_off;  // first of the job's n_cigar operations in the chunk's CIGAR pool
};
static_assert(sizeof(DevGRes) == 16, "DevGRes must be 16 bytes");

// chunk totals written by the device packer's prep kernel (ksw_devpack.cu) and read back by the host before it sizes
// the chunk's 2-bit pool; nmask_used is the pack kernel's allocation cursor in the N side pool
struct DevPackStats {
	unsigned long long units;          // 16-byte units of the 2-bit pool
	unsigned long long nmask_words;    // words of the N side pool if every job held an N in both sequences (capacity)
	unsigned long long q_hi, t_hi;     // one past the last query / target byte any job of the chunk reads
	unsigned long long q_lo_inv, t_lo_inv;   // ~(first query / target byte any job reads): a maximum, so that 0 initialises it
	unsigned int class_n[6];           // jobs per kernel class (KSW_N_CLASSES) before the N demotion (class 0 -> 1)
	int class_qmax[6];                 // longest query per class, before the N demotion
	unsigned int bad;                  // a job with qlen < 1 or tlen < 0
	unsigned int nmask_used;
};

struct KswParams {          // passed by value as a kernel parameter (constant bank)
	int8_t  mat[25];
	int8_t  pad[3];
	int32_t o_del, e_del, o_ins, e_ins;
	int32_t zdrop;
};

#define KSW_FAST_QUADS(qlen) (((qlen) >> 2) + 1)   /* 4-column quads of the fast kernel that cover columns 0..qlen */

#define KSW_CLASS_SHIFT 8
#define KSW_CLASS_MASK 0xfu
#define KSW_CLASS_GENERIC 4u          /* classes 0..3: fast s16x2 kernel (0 = keyed); >= 4: the int32 kernels */
#define KSW_CLASS_WARP 4u             /* one job per warp, columns in shared memory (ksw_warp.cu) */
#define KSW_CLASS_THREAD 5u           /* one job per thread, columns in HBM (ksw_generic.cu): everything else */
#define KSW_N_CLASSES 6

#define KSW_FLAG_QN 1u
#define KSW_FLAG_TN 2u

static KSW_HD uint32_t ksw_words2(int len) { return (uint32_t)((len + 15) >> 4); }
static KSW_HD uint32_t ksw_words1(int len) { return (uint32_t)((len + 31) >> 5); }
