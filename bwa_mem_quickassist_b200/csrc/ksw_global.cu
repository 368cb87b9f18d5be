// ksw_global.cu — banded global alignment with backtrace (the CIGAR generator) on sm_100a: the reference's
// ksw_global2, bwa-0.7.8/ksw.c:501-584, called by bwa_gen_cigar2 (bwa.c:132) for every reported alignment.
// the next-largest DP after the extension path; same recurrence family, fixed band, plus a
// direction matrix z and a backtrack.
//
// First GPU version: one job per thread, int32 arithmetic, everything a job needs in HBM slabs that are interleaved
// over the threads (element c of thread g at [c * n_threads + g], so the lanes of a warp touch neighbouring bytes when
// they are at the same position): H/E columns, query codes, and the direction matrix z with the reference's exact
// indexing (row i, column j at i*n_col + (j - beg_i), n_col = min(qlen, 2w+1); ksw.c:509,531,563).  The backtrack
// (ksw.c:562-579) runs on the device in two passes over the same path — count the CIGAR runs, claim a slice of the
// output pool with one atomicAdd, write the runs back to front — so no per-job scratch for the reversed CIGAR is needed.
// Integer work: results are bit-exact (score and every CIGAR operation).
#include <cuda_runtime.h>
#include "ksw_dev.cuh"
#include "ksw_launch.h"

namespace {

constexpr int G_MINUS_INF = -0x40000000;           // ksw.c:36

struct GTrace {                                    // walks the path of ksw.c:566-573 and reports maximal runs
	const uint8_t *z; size_t stride; int n_col, w; long long zcap;
	template <class F>
	__device__ __forceinline__ int run(int tlen, int qlen, F &&emit) const
	{
		int i = tlen - 1, k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;        // the last cell (ksw.c:565)
		int which = 0, n = 0, cur_op = -1, cur_len = 0;
		auto push = [&](int op, int len) {                                       // push_cigar, ksw.c:486-499
			if (op == cur_op) { cur_len += len; return; }
			if (cur_op >= 0) emit(n++, cur_op, cur_len);
			cur_op = op; cur_len = len;
		};
		while (i >= 0 && k >= 0) {
			const int beg = i > w ? i - w : 0;
			// inside the band this is the cell the DP wrote; a path that leaves the band (only possible when the band
			// cannot hold the end cell, |tlen - qlen| > w: undefined in the reference too) is clamped into the slab
			long long c = (long long)i * n_col + (k - beg);
			c = c < 0 ? 0 : (c >= zcap ? zcap - 1 : c);
			which = (z[(size_t)c * stride] >> (which << 1)) & 3;
			if (which == 0) { push(0, 1); --i; --k; }
			else if (which == 1) { push(2, 1); --i; }
			else { push(1, 1); --k; }
		}
		if (i >= 0) push(2, i + 1);
		if (k >= 0) push(1, k + 1);
		if (cur_op >= 0) emit(n++, cur_op, cur_len);
		return n;
	}
};

__global__ void __launch_bounds__(KSW_GENERIC_THREADS)
ksw_global_kernel(const DevGJob *__restrict__ jobs, int64_t n_jobs, const uint8_t *__restrict__ seq, KswParams P,
                  int2 *__restrict__ eh, uint8_t *__restrict__ qc, uint8_t *__restrict__ zslab, const long long zcap,
                  unsigned long long *__restrict__ pool_used, uint32_t *__restrict__ cigar_pool, DevGRes *__restrict__ res)
{
	const int64_t n_threads = (int64_t)gridDim.x * blockDim.x;
	const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;

	for (int64_t kk = g; kk < n_jobs; kk += n_threads) {
		const DevGJob jb = jobs[kk];
		const int qlen = jb.qlen, tlen = jb.tlen, w = jb.w;
		const uint8_t *query = seq + jb.seq_off, *target = query + qlen;
		const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;                  // ksw.c:509
#define EH(j) eh[(size_t)(j) * n_threads + g]
#define QC(j) qc[(size_t)(j) * n_threads + g]
#define Z(c) zslab[(size_t)(c) * n_threads + g]
		// first row (ksw.c:520-523) and the query codes
		EH(0) = make_int2(0, G_MINUS_INF);
		{
			int j = 1;
			for (; j <= qlen && j <= w; ++j) EH(j) = make_int2(-(P.o_ins + P.e_ins * j), G_MINUS_INF);
			for (; j <= qlen; ++j) EH(j) = make_int2(G_MINUS_INF, G_MINUS_INF);
			for (j = 0; j < qlen; ++j) QC(j) = query[j];
		}
		for (int i = 0; i < tlen; ++i) {                                         // ksw.c:525-559
			const int8_t *srow = P.mat + (int)target[i] * 5;
			int f = G_MINUS_INF;
			const int beg = i > w ? i - w : 0;
			const int end = i + w + 1 < qlen ? i + w + 1 : qlen;
			int h1 = beg == 0 ? -(P.o_del + P.e_del * (i + 1)) : G_MINUS_INF;
			const size_t zrow = (size_t)i * n_col;
			for (int j = beg; j < end; ++j) {
				const int2 c = EH(j);
				const int m = c.x + srow[QC(j)];
				int e = c.y;
				uint32_t d = m >= e ? 0u : 1u;
				int h = m >= e ? m : e;
				d = h >= f ? d : 2u;
				h = h >= f ? h : f;
				int t = m - oe_del;
				e -= P.e_del;
				d |= e > t ? 1u << 2 : 0u;
				e = e > t ? e : t;
				t = m - oe_ins;
				f -= P.e_ins;
				d |= f > t ? 2u << 4 : 0u;
				f = f > t ? f : t;
				EH(j) = make_int2(h1, e);
				h1 = h;
				Z(zrow + (j - beg)) = (uint8_t)d;
			}
			EH(end) = make_int2(h1, G_MINUS_INF);                                // ksw.c:558
		}
		DevGRes r;
		r.score = EH(qlen).x;                                                     // ksw.c:560
		const GTrace tr{zslab + g, (size_t)n_threads, n_col, w, zcap};
		r.n_cigar = tr.run(tlen, qlen, [](int, int, int) {});
		const unsigned long long off = atomicAdd(pool_used, (unsigned long long)r.n_cigar);
		uint32_t *out = cigar_pool + off;
		const int n = r.n_cigar;
		tr.run(tlen, qlen, [&](int r_idx, int op, int len) { out[n - 1 - r_idx] = (uint32_t)len << 4 | (uint32_t)op; });
		r.cigar_off = (long long)off;
		res[jb.idx] = r;
#undef EH
#undef QC
#undef Z
	}
}

} // namespace

cudaError_t ksw_launch_global(const DevGJob *jobs, int64_t n_jobs, const uint8_t *seq, const KswParams &P, int2 *eh,
                              uint8_t *qc, uint8_t *z, long long zcap, int n_blocks, unsigned long long *pool_used,
                              uint32_t *cigar_pool, DevGRes *res, cudaStream_t st)
{
	if (n_jobs <= 0) return cudaSuccess;
	ksw_global_kernel<<<n_blocks, KSW_GENERIC_THREADS, 0, st>>>(jobs, n_jobs, seq, P, eh, qc, z, zcap, pool_used, cigar_pool, res);
	return cudaGetLastError();
}
