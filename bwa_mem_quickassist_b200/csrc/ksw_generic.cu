// ksw_generic.cu — the *generic* extension kernel (sm_100a): one job per thread, int32
// arithmetic, per-thread H/E columns in an HBM scratch slab.  It accepts every job the
// reference accepts (any qlen/tlen, any score range, N in query or target) and is the
// exact-by-construction fallback for jobs the fast s16x2 kernel (ksw_fast.cu) cannot take.
//
// Semantics follow bwa-0.7.8/ksw.c:379-476 row by row; every 0.7.8 quirk listed in
// SURVEY.md §7.3(2) is reproduced and tagged below.  The recurrences use the DPX integer
// instructions (VIADDMNMX / VIADDMNMX.RELU / VIMNMX3 via __viaddmax_s32 & friends).
#include <cuda_runtime.h>
#include "ksw_dev.cuh"
#include "ksw_launch.h"

namespace {

__device__ __forceinline__ int seq_code(const uint32_t *w2, const uint32_t *nmask, int k)
{
	int c = (int)((w2[k >> 4] >> ((k & 15) << 1)) & 3u);
	if (nmask && ((nmask[k >> 5] >> (k & 31)) & 1u)) c = 4;
	return c;
}

// scratch slab: column j of thread g lives at (j * n_threads + g) so that the threads of a
// warp, which sweep similar columns at the same time, touch neighbouring addresses.
__global__ void __launch_bounds__(128)
ksw_generic_kernel(const DevJob *__restrict__ jobs, int64_t n_jobs, const uint32_t *__restrict__ pool,
                   const uint32_t *__restrict__ npool, KswParams P, int2 *__restrict__ eh, uint8_t *__restrict__ qc,
                   const uint32_t *__restrict__ order, DevRes *__restrict__ res,
                   uint32_t *__restrict__ cells)
{
	const int64_t n_threads = (int64_t)gridDim.x * blockDim.x;
	const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;

	for (int64_t k = g; k < n_jobs; k += n_threads) {
		const DevJob jb = jobs[order[k]];
		const int qlen = jb.qlen, tlen = jb.tlen, h0 = jb.h0, w = jb.w;
		const uint32_t *q2 = pool + (size_t)jb.seq_off * 4;
		const uint32_t *t2 = q2 + ksw_words2(qlen);
		const uint32_t *qn = (jb.flags & KSW_FLAG_QN) ? npool + jb.nmask_off : nullptr;
		const uint32_t *tn = (jb.flags & KSW_FLAG_TN)
		                         ? npool + jb.nmask_off + ((jb.flags & KSW_FLAG_QN) ? ksw_words1(qlen) : 0)
		                         : nullptr;
#define EH(j) eh[(size_t)(j) * n_threads + g]
#define QC(j) qc[(size_t)(j) * n_threads + g]
		// row -1 (ksw.c:394-396) and the query codes
		{
			int v = h0;
			EH(0) = make_int2(h0, 0);
			for (int j = 1; j <= qlen; ++j) {
				if (j == 1) v = h0 > oe_ins ? h0 - oe_ins : 0;
				else v = v > P.e_ins ? v - P.e_ins : 0;      // once it reaches <= e_ins the tail stays 0
				EH(j) = make_int2(v, 0);
			}
			for (int j = 0; j < qlen; ++j) QC(j) = (uint8_t)seq_code(q2, qn, j);
		}
		int best = h0, best_i = -1, best_j = -1, end_i = -1, end_sc = -1, off = 0;   // ksw.c:408-410
		int lo = 0, hi = qlen;
		uint32_t ncell = 0;
		for (int i = 0; i < tlen; ++i) {
			const int t = seq_code(t2, tn, i);
			const int8_t *srow = P.mat + t * 5;
			int left = max(h0 - (P.o_del + P.e_del * (i + 1)), 0);   // used at column lo even if lo>0 (ksw.c:415-416)
			int f = 0, rmax = 0, rarg = -1;
			lo = max(lo, i - w);
			hi = min(min(hi, i + w + 1), qlen);
			if (hi > lo) ncell += (uint32_t)(hi - lo);
			for (int j = lo; j < hi; ++j) {
				int2 c = EH(j);
				int h = c.x + srow[QC(j)];                           // no zero guard (ksw.c:430)
				h = __vimax3_s32(h, c.y, f);
				if (h >= rmax) rarg = j;                             // ties -> last column (ksw.c:434)
				rmax = max(rmax, h);
				int e = __viaddmax_s32(c.y, -P.e_del, __viaddmax_s32_relu(h, -oe_del, 0));
				f = __viaddmax_s32(f, -P.e_ins, __viaddmax_s32_relu(h, -oe_ins, 0));
				EH(j) = make_int2(left, e);
				left = h;
			}
			EH(hi) = make_int2(left, 0);                             // ksw.c:446 — E right of the band restarts at 0
			if (max(lo, hi) == qlen) {                               // ksw.c:447 tests j after the loop
				if (left >= end_sc) end_i = i;                       // ties -> last row (ksw.c:448)
				end_sc = max(end_sc, left);
			}
			if (rmax == 0) break;
			if (rmax > best) {
				best = rmax; best_i = i; best_j = rarg;
				off = max(off, abs(rarg - i));
			} else if (P.zdrop > 0) {
				const int di = i - best_i, dj = rarg - best_j;
				if (di > dj) { if (best - rmax - (di - dj) * P.e_del > P.zdrop) break; }
				else         { if (best - rmax - (dj - di) * P.e_ins > P.zdrop) break; }
			}
			int j;                                                   // ksw.c:463-466
			for (j = rarg; j >= lo && EH(j).x; --j) ;
			lo = j + 1;
			for (j = rarg + 2; j <= hi && EH(j).x; ++j) ;
			hi = j;
		}
#undef EH
#undef QC
		DevRes r;
		r.score = best; r.qle = best_j + 1; r.tle = best_i + 1;
		r.gtle = end_i + 1; r.gscore = end_sc; r.max_off = off;
		res[jb.idx] = r;
		cells[jb.idx] = ncell;
	}
}

// ---------------------------------------------------------------- DPX issue-rate probe
// Independent dependency chains of the exact instruction mix the extension kernels use.
// Each thread runs CHAINS independent accumulators so that the pipe, not latency, binds.
template <int WHICH>
__global__ void __launch_bounds__(256)
dpx_peak_kernel(unsigned *out, int iters, unsigned seed)
{
	constexpr int CHAINS = 8;
	unsigned a[CHAINS], b[CHAINS];
#pragma unroll
	for (int c = 0; c < CHAINS; ++c) { a[c] = seed + threadIdx.x * 7u + c; b[c] = seed * 3u + c * 11u; }
	const unsigned k1 = seed | 0x00010001u, k2 = (seed >> 3) | 0x00020002u;
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int c = 0; c < CHAINS; ++c) {
			if (WHICH == 0) {
				a[c] = __viaddmax_s16x2(a[c], k1, b[c]);
				b[c] = __viaddmax_s16x2_relu(b[c], k2, a[c]);
				a[c] = __vmaxs2(a[c], k2);
				b[c] = __viaddmax_s16x2(b[c], k1, a[c]);
			} else {
				a[c] = (unsigned)__viaddmax_s32((int)a[c], (int)k1, (int)b[c]);
				b[c] = (unsigned)__viaddmax_s32_relu((int)b[c], (int)k2, (int)a[c]);
				a[c] = (unsigned)max((int)a[c], (int)k2);
				b[c] = (unsigned)__viaddmax_s32((int)b[c], (int)k1, (int)a[c]);
			}
		}
	}
	unsigned acc = 0;
#pragma unroll
	for (int c = 0; c < CHAINS; ++c) acc ^= a[c] + b[c];
	if (acc == 0x12345u) out[blockIdx.x * blockDim.x + threadIdx.x] = acc;   // keeps the chains alive
}

} // namespace

cudaError_t ksw_launch_generic(const DevJob *jobs, int64_t n_jobs, const uint32_t *pool, const uint32_t *npool,
                               const KswParams &P,
                               int2 *eh, uint8_t *qc, int n_blocks, const uint32_t *order, DevRes *res, uint32_t *cells, cudaStream_t st)
{
	if (n_jobs <= 0) return cudaSuccess;
	ksw_generic_kernel<<<n_blocks, KSW_GENERIC_THREADS, 0, st>>>(jobs, n_jobs, pool, npool, P, eh, qc, order, res, cells);
	return cudaGetLastError();
}

cudaError_t ksw_launch_dpx_peak(int which, unsigned *out, int n_blocks, int iters, cudaStream_t st)
{
	if (which == 0) dpx_peak_kernel<0><<<n_blocks, 256, 0, st>>>(out, iters, 12345u);
	else dpx_peak_kernel<1><<<n_blocks, 256, 0, st>>>(out, iters, 12345u);
	return cudaGetLastError();
}
