// ksw_class.h — per-job decisions shared by the host packer (ksw_pack.cpp) and the device packer (ksw_devpack.cu):
// the reference's band clamp (ksw.c:398-406) and the routing of a job to a kernel class.  One source, compiled for
// both sides, so that a batch packed on the GPU is routed exactly like the same batch packed on the host.
#pragma once
#include <stdint.h>
#include "ksw_dev.cuh"

// Fast-kernel job classes (one launch each): class 0 = "keyed" jobs (qlen <= 124 and every score < 512, so the
// row arg-max can be tracked as h*128+column in 16 bits, and no N in either sequence); classes 1..3 by query length.
#define KSW_FAST_CLASSES 4
#define KSW_FAST_KEYED_MAXSCORE 511
static KSW_HD int ksw_fast_class_qmax(int c) { return c == 0 ? 124 : (c == 1 ? 128 : (c == 2 ? 256 : 512)); }

// what the routing needs to know about the scoring of a batch
struct KswScoring {
	int32_t maxsc, minsc;              // largest (floored at 0, like ksw.c:399) / smallest matrix entry
	int32_t o_del, e_del, o_ins, e_ins, end_bonus;
	int32_t fast_qmax;                 // largest qlen the fast kernel accepts (0: fast kernel disabled)
	int32_t warp_qmax;                 // largest qlen the warp-cooperative kernel accepts (0: disabled)
};

// the reference's band clamp (ksw.c:398-406), evaluated with the identical C expression (IEEE double division on
// both the host and the device)
static KSW_HD int ksw_clamp_w_expr(int qlen, int maxsc, int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus)
{
	int max_ins = (int)((double)(qlen * maxsc + end_bonus - o_ins) / e_ins + 1.);
	max_ins = max_ins > 1 ? max_ins : 1;
	w = w < max_ins ? w : max_ins;
	int max_del = (int)((double)(qlen * maxsc + end_bonus - o_del) / e_del + 1.);
	max_del = max_del > 1 ? max_del : 1;
	w = w < max_del ? w : max_del;
	return w;
}

// A job may take the fast s16x2 kernel iff every value its DP can hold stays far inside int16
// and its columns fit the kernel's shared-memory budget; everything else goes to the generic
// int32 kernel (still on the GPU).
static KSW_HD bool ksw_fast_eligible(const KswScoring &S, int qlen, int h0)
{
	if (qlen > S.fast_qmax) return false;
	if (S.o_ins < 0 || S.o_del < 0 || S.e_ins < 1 || S.e_del < 1) return false;
	if (S.o_ins + S.e_ins > 4000 || S.o_del + S.e_del > 4000) return false;
	if (S.minsc < -120 || S.maxsc > 120) return false;
	if ((long long)h0 + (long long)qlen * S.maxsc > 20000) return false;
	return true;
}

// The warp-cooperative int32 kernel (ksw_warp.cu) takes what the s16x2 kernel cannot hold if the row is long enough to
// occupy 32 lanes and the query's columns fit an SM's shared memory; its F scan relies on o_ins >= 0 like the fast
// kernel's, and its offsets on scores far inside int32.
#define KSW_WARP_MIN_QLEN 129
#define KSW_WARP_MAX_QLEN 24000
static KSW_HD bool ksw_warp_eligible(const KswScoring &S, int qlen, int h0)
{
	if (qlen < KSW_WARP_MIN_QLEN || qlen > S.warp_qmax) return false;
	if (S.o_ins < 0 || S.e_ins < 0 || S.o_ins > (1 << 20) || S.e_ins > (1 << 20) || S.o_del > (1 << 20) || S.e_del > (1 << 20)) return false;
	if (S.o_del < -(1 << 20) || S.e_del < -(1 << 20)) return false;
	if ((long long)h0 + (long long)qlen * S.maxsc > (1ll << 27)) return false;
	return true;
}

// kernel class of a job before its sequences have been looked at (a class-0 job that holds an N moves to class 1)
static KSW_HD uint32_t ksw_job_class(const KswScoring &S, int qlen, int h0)
{
	if (!ksw_fast_eligible(S, qlen, h0)) return ksw_warp_eligible(S, qlen, h0) ? KSW_CLASS_WARP : KSW_CLASS_THREAD;
	if (qlen <= ksw_fast_class_qmax(0) && (long long)h0 + (long long)qlen * S.maxsc + S.o_del + S.e_del <= KSW_FAST_KEYED_MAXSCORE) return 0;
	uint32_t qc = 1;
	while (qc + 1 < KSW_FAST_CLASSES && qlen > ksw_fast_class_qmax((int)qc)) ++qc;
	return qc;
}

// 16-byte units of a job's packed sequences in the 2-bit pool: query words | target words, padded
static KSW_HD uint32_t ksw_job_units(int qlen, int tlen) { return (ksw_words2(qlen) + ksw_words2(tlen) + 3) >> 2; }

// ---- banded global alignment (ksw_global2): routing between the s16x2 kernel (ksw_gfast.cu) and the int32 kernel
// what the batch's scoring costs per cell at most (0: the s16x2 form cannot hold this scoring at all)
static inline int ksw_gfast_cell_cost(const int8_t *mat25, int o_del, int e_del, int o_ins, int e_ins)
{
	int amax = 0;
	for (int i = 0; i < 25; ++i) { const int a = mat25[i] < 0 ? -mat25[i] : mat25[i]; amax = a > amax ? a : amax; }
	if (amax > 100 || e_del < 0 || e_ins < 0 || o_del < 0 || o_ins < 0) return 0;
	if (o_del + e_del > 800 || o_ins + e_ins > 800) return 0;
	int c = amax > e_del ? amax : e_del;
	c = c > e_ins ? c : e_ins;
	return c > 1 ? c : 1;
}
// the band holds the end cell (outside that the reference reads cells it never wrote), the query's columns fit shared
// memory, and every H / E / F value of the band stays within +-7000 of zero: inside int16 with the kernel's bias
#define KSW_GFAST_MAX_QLEN 1000
static inline bool ksw_gfast_eligible(int cell_cost, int o_del, int e_del, int o_ins, int e_ins, int qlen, int tlen, int w)
{
	if (!cell_cost || qlen < 1 || tlen < 1 || qlen > KSW_GFAST_MAX_QLEN || tlen > 30000) return false;
	if ((tlen > qlen ? tlen - qlen : qlen - tlen) > w) return false;
	return ((long long)qlen + tlen) * cell_cost + o_del + o_ins + e_del + e_ins < 7000;
}
// quads per band row, at most
static inline int ksw_gfast_nqb(int qlen, int w) { const long long b = 2LL * w + 1; return (int)(((qlen < b ? qlen : b) + 6) / 4 + 1); }
