// ksw_warp.cu — the warp-cooperative extension kernel (sm_100a): ONE job per warp, int32 arithmetic, any score range,
// query lengths from 129 up to what fits the shared memory of an SM (KSW_WARP_MAX_QLEN).  It takes the jobs the
// one-job-per-lane s16x2 kernel cannot hold — long queries (the 512-column limit of ksw_fast.cu) and jobs whose scores
// could leave int16 — which used to fall to the thread-per-job generic kernel and its global-memory columns.
//
// Semantics of bwa-0.7.8/ksw.c:379-476, row by row (the band trim of ksw.c:463-466 makes row i+1 depend on the completed
// row i), each row swept by the 32 lanes together:
//
//   * a row is cut into passes of 128 columns; in a pass lane L owns the 4 consecutive columns lo + 128*pass + 4L .. +3;
//     H(i-1,j-1) / E(i,j) of the whole query (the reference's eh[], ksw.c:375-377) and the query codes live in shared memory;
//   * h'(j) = max(H(i-1,j-1) + S, E(i,j)) is independent per column (ksw.c:430-431);
//   * F is the one serial quantity of a row (ksw.c:442-444).  Without the reference's floor at 0 (which changes no H and
//     no E, see ksw_fast_core.h) it is F~(j+1) = max(F~(j) - e_ins, h'(j) - oe_ins), i.e. across a lane's 4 columns
//     F~_out = max(F~_in - 4 e_ins, g) with g = max_k (h'(k) - oe_ins - (3-k) e_ins): a max-plus scan over the lanes, done as
//     a prefix maximum of g_L + 4 e_ins L with five __shfl_up_sync steps; exact in integers;
//   * the row maximum and its LAST column (ksw.c:434: ties go to the larger j) by a lane-local scan and a shuffle
//     reduction on (value, column);
//   * the trim scans of ksw.c:463-466 look at 32 stored H values at a time (__ballot_sync of "is zero", then the nearest
//     set bit);
//   * band clamp, best cell, z-drop and the end-of-query score are warp-uniform scalars every lane computes alike.
//
// Every 0.7.8 quirk of SURVEY.md §7.3(2) is kept exactly as in ksw_generic.cu: first-column value used even when lo > 0,
// no zero guard on H + S, the tie rules, the `j == qlen` test after the loop, eh[end] written every row.
#include <cuda_runtime.h>
#include <climits>
#include "ksw_dev.cuh"
#include "ksw_class.h"
#include "ksw_launch.h"

namespace {

#define KSW_NEG_INF (-(1 << 29))            /* "no cell": far below any reachable score, far above INT_MIN after the scan's offsets */

__device__ __forceinline__ int seq_code_w(const uint32_t *w2, const uint32_t *nmask, int k)
{
	int c = (int)((w2[k >> 4] >> ((k & 15) << 1)) & 3u);
	if (nmask && ((nmask[k >> 5] >> (k & 31)) & 1u)) c = 4;
	return c;
}

__global__ void __launch_bounds__(32)
ksw_warp_kernel(const DevJob *__restrict__ jobs, long long n_jobs, const uint32_t *__restrict__ pool,
                const uint32_t *__restrict__ npool, const KswParams P, const int qcap, unsigned long long *__restrict__ counter,
                const uint32_t *__restrict__ order, DevRes *__restrict__ res, uint32_t *__restrict__ cells)
{
	extern __shared__ int2 smem_w[];
	int2 *eh = smem_w;                                              // eh[j] = {H(i-1, j-1), E(i, j)}, j = 0..qlen
	uint8_t *qc = reinterpret_cast<uint8_t *>(eh + qcap + 1);       // query codes 0..4
	const int lane = threadIdx.x;
	const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins, e_ins = P.e_ins;
	const unsigned FULL = 0xffffffffu;

	for (;;) {
		unsigned long long k = 0;
		if (lane == 0) k = atomicAdd(counter, 1ull);
		k = __shfl_sync(FULL, k, 0);
		if ((long long)k >= n_jobs) break;
		const DevJob jb = jobs[order[k]];
		const int qlen = jb.qlen, tlen = jb.tlen, h0 = jb.h0, w = jb.w;
		const uint32_t *q2 = pool + (size_t)jb.seq_off * 4;
		const uint32_t *t2 = q2 + ksw_words2(qlen);
		const uint32_t *qn = (jb.flags & KSW_FLAG_QN) ? npool + jb.nmask_off : nullptr;
		const uint32_t *tn = (jb.flags & KSW_FLAG_TN) ? npool + jb.nmask_off + ((jb.flags & KSW_FLAG_QN) ? ksw_words1(qlen) : 0) : nullptr;
		// row -1 (ksw.c:394-396): eh[0].h = h0, eh[1].h = max(h0 - oe_ins, 0), then minus e_ins per column down to 0
		for (int j = lane; j <= qlen; j += 32) {
			int v = j == 0 ? h0 : h0 - oe_ins - (j - 1) * e_ins;
			eh[j] = make_int2(v > 0 ? v : 0, 0);
			if (j < qlen) qc[j] = (uint8_t)seq_code_w(q2, qn, j);
		}
		__syncwarp();
		int best = h0, best_i = -1, best_j = -1, end_i = -1, end_sc = -1, off = 0;   // ksw.c:408-410
		int lo = 0, hi = qlen;
		uint32_t ncell = 0;
		for (int i = 0; i < tlen; ++i) {
			const int t = seq_code_w(t2, tn, i);
			// the target base's matrix row, one score per query code, in registers
			const int s0 = P.mat[t * 5 + 0], s1 = P.mat[t * 5 + 1], s2 = P.mat[t * 5 + 2], s3 = P.mat[t * 5 + 3], s4 = P.mat[t * 5 + 4];
			const int left0 = max(h0 - (P.o_del + P.e_del * (i + 1)), 0);   // used at column lo even if lo > 0 (ksw.c:415-416)
			lo = max(lo, i - w);                                             // ksw.c:418-420
			hi = min(min(hi, i + w + 1), qlen);
			if (hi > lo) ncell += (uint32_t)(hi - lo);
			int m = 0, mj = -1;                                              // lane-local row maximum and its last column
			int fcarry = 0;                                                  // F entering the pass's first column
			int hprev = left0;                                               // H(i, j-1) of the pass's first column
			int hlast = left0;                                               // H(i, hi-1): the reference's h1 after the loop
			for (int base = lo; base < hi; base += 128) {
				const int j0 = base + 4 * lane;
				int hp[4], e_in[4];
#pragma unroll
				for (int c = 0; c < 4; ++c) {
					const int j = j0 + c;
					if (j < hi) {
						const int2 v = eh[j];
						const int code = qc[j];
						const int s = code == 0 ? s0 : (code == 1 ? s1 : (code == 2 ? s2 : (code == 3 ? s3 : s4)));
						e_in[c] = v.y;
						hp[c] = max(v.x + s, v.y);                           // no zero guard (ksw.c:430)
					} else { e_in[c] = 0; hp[c] = KSW_NEG_INF; }
				}
				// the lane's contribution to the F scan: F~_out = max(F~_in - 4 e, g)
				int g = hp[0] - oe_ins;
#pragma unroll
				for (int c = 1; c < 4; ++c) g = __viaddmax_s32(g, -e_ins, hp[c] - oe_ins);
				// exclusive max-plus scan over the lanes: u = g + 4 e L, inclusive prefix maximum, shifted back
				const int step = 4 * e_ins;
				int u = g + step * lane;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) {
					const int v = __shfl_up_sync(FULL, u, o);
					if (lane >= o) u = max(u, v);
				}
				int f = __shfl_up_sync(FULL, u, 1) - step * (lane - 1);      // max over the lanes before this one
				f = lane == 0 ? fcarry : max(fcarry - step * lane, f);
				const int f_next = __shfl_sync(FULL, max(f - step, g), 31);  // F~ leaving the last lane
				// the lane's 4 cells
				int hc[4];
#pragma unroll
				for (int c = 0; c < 4; ++c) {
					const int j = j0 + c;
					const int h = max(hp[c], f);                             // ksw.c:432 (max with F~ equals max with F: h' >= 0)
					hc[c] = h;
					if (j < hi) {
						if (h >= m) mj = j;                                  // ties -> last column (ksw.c:434)
						m = max(m, h);
					}
					f = __viaddmax_s32(f, -e_ins, hp[c] - oe_ins);
				}
				// eh[j] = {H(i, j-1), E(i+1, j)}: the left neighbour of the lane's first column sits in the lane before
				int hl = __shfl_up_sync(FULL, hc[3], 1);
				if (lane == 0) hl = hprev;
#pragma unroll
				for (int c = 0; c < 4; ++c) {
					const int j = j0 + c;
					if (j < hi) {
						const int e = __vimax3_s32(e_in[c] - P.e_del, hc[c] - oe_del, 0);   // ksw.c:436-439
						eh[j] = make_int2(hl, e);
						if (j == hi - 1) hlast = hc[c];
					}
					hl = hc[c];
				}
				hprev = __shfl_sync(FULL, hc[3], 31);
				fcarry = f_next;
			}
			// H(i, hi-1) lives in one lane; the row maximum and its last column come from all of them
			{
				const unsigned own = __ballot_sync(FULL, hi > lo && ((hi - 1 - lo) & 127) >> 2 == lane);
				if (own) hlast = __shfl_sync(FULL, hlast, __ffs(own) - 1);
			}
#pragma unroll
			for (int o = 16; o > 0; o >>= 1) {
				const int om = __shfl_xor_sync(FULL, m, o), oj = __shfl_xor_sync(FULL, mj, o);
				if (om > m || (om == m && oj > mj)) { m = om; mj = oj; }
			}
			__syncwarp();
			if (lane == 0) eh[hi] = make_int2(hlast, 0);                     // ksw.c:446 — E right of the band restarts at 0
			__syncwarp();
			if (max(lo, hi) == qlen) {                                       // ksw.c:447 tests j after the loop
				if (hlast >= end_sc) end_i = i;                              // ties -> last row (ksw.c:448)
				end_sc = max(end_sc, hlast);
			}
			if (m == 0) break;                                               // ksw.c:451
			if (m > best) {                                                  // ksw.c:452-461
				best = m; best_i = i; best_j = mj;
				off = max(off, abs(mj - i));
			} else if (P.zdrop > 0) {
				const int di = i - best_i, dj = mj - best_j;
				if (di > dj) { if (best - m - (di - dj) * P.e_del > P.zdrop) break; }
				else         { if (best - m - (dj - di) * P.e_ins > P.zdrop) break; }
			}
			// band trim (ksw.c:463-466), 32 stored H values per step
			{
				int nl = lo;                                                 // no zero found: beg stays
				for (int b = mj; b >= lo; b -= 32) {
					const int j = b - lane;
					const unsigned z = __ballot_sync(FULL, j >= lo && eh[j].x == 0);
					if (z) { nl = b - (__ffs(z) - 1) + 1; break; }
				}
				int nh = hi + 1;                                             // no zero found: the scan runs off the end
				for (int b = mj + 2; b <= hi; b += 32) {
					const int j = b + lane;
					const unsigned z = __ballot_sync(FULL, j <= hi && eh[j].x == 0);
					if (z) { nh = b + (__ffs(z) - 1); break; }
				}
				lo = nl; hi = nh;
			}
		}
		if (lane == 0) {
			DevRes r;
			r.score = best; r.qle = best_j + 1; r.tle = best_i + 1;
			r.gtle = end_i + 1; r.gscore = end_sc; r.max_off = off;
			res[jb.idx] = r;
			cells[jb.idx] = ncell;
		}
		__syncwarp();
	}
}

} // namespace

size_t ksw_warp_smem_bytes(int qmax) { return (size_t)(qmax + 1) * sizeof(int2) + (((size_t)qmax + 15) & ~(size_t)15); }

cudaError_t ksw_launch_warp(const DevJob *jobs, int64_t n_jobs, const uint32_t *pool, const uint32_t *npool, const KswParams &P,
                            int qmax, int sm_count, unsigned long long *counter, const uint32_t *order, DevRes *res,
                            uint32_t *cells, cudaStream_t st)
{
	if (n_jobs <= 0) return cudaSuccess;
	const size_t smem = ksw_warp_smem_bytes(qmax);
	static unsigned long long raised = 0;                            // bit d: dynamic shared-memory ceiling raised on device d
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	if (dev < 0 || dev >= 64 || !((__atomic_load_n(&raised, __ATOMIC_ACQUIRE) >> dev) & 1ull)) {
		int optin = 0;
		e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
		if (e != cudaSuccess) return e;
		e = cudaFuncSetAttribute(ksw_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
		if (e != cudaSuccess) return e;
		if (dev >= 0 && dev < 64) __atomic_fetch_or(&raised, 1ull << dev, __ATOMIC_RELEASE);
	}
	int per_sm = 0;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ksw_warp_kernel, 32, smem);
	if (e != cudaSuccess) return e;
	if (per_sm < 1) return cudaErrorLaunchOutOfResources;
	long long blocks = (long long)sm_count * per_sm;
	if (blocks > n_jobs) blocks = n_jobs;
	e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st);
	if (e != cudaSuccess) return e;
	ksw_warp_kernel<<<(unsigned)blocks, 32, smem, st>>>(jobs, (long long)n_jobs, pool, npool, P, qmax, counter, order, res, cells);
	return cudaGetLastError();
}
