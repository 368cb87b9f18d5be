// ksw_gfast.cu — the fast form of the banded global alignment with backtrace (the reference's ksw_global2,
// bwa-0.7.8/ksw.c:501-584, called by bwa_gen_cigar2, bwa.c:132): s16x2 DPX rows like the extension kernel (ksw_fast_core.h),
// and NO direction matrix.
//
//   DP kernel (ksw_gfast_dp_kernel): one job per lane, 32 jobs of similar size per warp, rows in lockstep.  Per lane the
//   reference's eh[] (H(i-1,j-1), E(i,j)) lives in shared memory as lane-interleaved quads {Hd A, E A, Hd B, E B} with the
//   extension kernel's layout (pair A = columns c0,c2; pair B = c1,c3), so the serial F chain F(j+1) = max(F(j)-e_ins,
//   M(j)-oe_ins) (ksw.c:554-557: it depends on M = H(i-1,j-1)+S only, not on H) is four DPX instructions and two FMA-pipe
//   half moves per four cells; E(i+1,j) = max(E-e_del, M-oe_del) (ksw.c:549-552) is one DPX instruction and two plain adds.
//   Everything is kept with a bias of 16384 so that "minus infinity" (ksw.c:36) is a small positive number and the adds
//   cannot borrow between the halves.  Per four cells the kernel stores the four H values (8 bytes) to HBM — 2 bytes per
//   band cell instead of the reference's direction byte — and nothing else.
//
//   Backtrack kernel (ksw_gfast_trace_kernel): one job per thread.  The reference's walk (ksw.c:562-579) reads, per step,
//   two bits of the direction byte; here the same bits are RECOMPUTED from the stored H values and the sequences:
//     in state H at (i,k):  M = H(i-1,k-1)+S.  H == M                    -> diagonal (ksw.c:545: m >= e, then h >= f)
//                           else E(i,k) == H (a short walk up column k)  -> deletion  (E beats F on ties, ksw.c:547)
//                           else                                         -> insertion
//     in state E at (i,k), knowing E(i+1,k):  M(i,k)-oe_del == E(i+1,k)  -> the gap was opened here (ksw.c:551: e > t is false)
//                           else it was extended: E(i,k) = E(i+1,k)+e_del
//     in state F at (i,k), knowing F(i,k+1):  likewise with oe_ins / e_ins along the row (ksw.c:555).
//   All in exact integers, so the CIGAR is the reference's, operation for operation (tests/test_global.py: oracle, golden
//   vectors, compiled reference).
//
// Jobs the s16x2 form cannot hold (band that misses the end cell, very long or very high-scoring pairs) stay with the int32
// kernel of ksw_global.cu.
#include <cuda_runtime.h>
#include "ksw_dev.cuh"
#include "ksw_fast_core.h"
#include "ksw_launch.h"

namespace {

constexpr int GT = 32;
constexpr int G_BIAS = 16384;                  // added to every H / E / F value
constexpr int G_NEG = 1024;                    // biased "minus infinity": real values stay above 16384 - 7000
constexpr int G_SC = 128;                      // added to every matrix score so that it is a non-negative byte
constexpr int G_MINF = -0x40000000;            // the reference's MINUS_INF (ksw.c:36), used by the backtrack

__device__ __forceinline__ uint32_t pk2(int v) { return ((uint32_t)v & 0xffffu) | ((uint32_t)v << 16); }

// one quad of a row.  EDGE: out-of-band columns of the quad are masked to "minus infinity" on load
template <bool EDGE>
__device__ __forceinline__ void gquad(uint32_t &X, uint32_t &Hc, const uint4 v_in, const uint32_t sw, const uint2 mr, uint4 *dst, uint2 *zdst,
                                      const uint32_t neg_ei, const uint32_t neg_oei, const uint32_t ed32, const uint32_t oed32,
                                      const uint32_t keepA, const uint32_t keepB, const uint32_t endA, const uint32_t endB)
{
	using namespace kswdpx;
	uint4 v = v_in;
	const uint32_t negH = pk2(G_NEG - G_SC), negE = pk2(G_NEG);
	if (EDGE) {
		v.x = (v.x & keepA) | (negH & ~keepA); v.y = (v.y & keepA) | (negE & ~keepA);
		v.z = (v.z & keepB) | (negH & ~keepB); v.w = (v.w & keepB) | (negE & ~keepB);
	}
	// M = H(i-1,j-1) + S: the stored diagonal value carries -G_SC, the looked-up score +G_SC: a plain 32-bit add, no borrow
	const uint32_t MA = v.x + prmt(mr.x, mr.y, sw), MB = v.z + prmt(mr.x, mr.y, ksw_hi16_of(sw));
	// F chain on Fs = F + oe_ins (ksw_fast_core.h): lo half c0 -> c1 -> c2, hi half c2 -> c3 -> c0'
	const uint32_t u1 = addmax2(X, neg_ei, MA);
	const uint32_t u2 = addmax2(u1, neg_ei, MB);
	const uint32_t FA = ksw_lo_to_hi_add(u2, X);           // (Fs(c0), Fs(c2))
	const uint32_t FB = addmax2(FA, neg_ei, MA);           // (Fs(c1), Fs(c3))
	const uint32_t u4 = addmax2(FB, neg_ei, MB);
	X = ksw_hi16_of(u4);
	// H = max(M, E, F) (ksw.c:545-548)
	const uint32_t hA = addmax2(FA, neg_oei, max2(MA, v.y)), hB = addmax2(FB, neg_oei, max2(MB, v.w));
	// E(i+1,j) = max(E - e_del, M - oe_del) (ksw.c:549-552)
	uint32_t eA = max2(v.y - ed32, MA - oed32), eB = max2(v.w - ed32, MB - oed32);
	if (EDGE) { eA = (eA & ~endA) | (negE & endA); eB = (eB & ~endB) | (negE & endB); }   // eh[end].e = -inf (ksw.c:558)
	*zdst = make_uint2(hA, hB);
	const uint32_t sA = hA - pk2(G_SC), sB = hB - pk2(G_SC);
	uint4 o;
	o.x = prmt(Hc, sB, 0x5432u);                           // (H(c0-1), H(c1)) for columns c0, c2
	o.y = eA;
	o.z = sA;                                              // (H(c0), H(c2)) for columns c1, c3
	o.w = eB;
	*dst = o;
	Hc = sB;
}

__global__ void __launch_bounds__(GT)
ksw_gfast_dp_kernel(const DevGJob *__restrict__ jobs, const uint8_t *__restrict__ seq, const KswParams P, const uint32_t *__restrict__ gorder,
                    const DevGGroup *__restrict__ groups, const int n_groups, const int nq_cap, uint2 *__restrict__ z,
                    unsigned *__restrict__ counter, DevGRes *__restrict__ res)
{
	extern __shared__ uint4 gsm[];
	const int lane = threadIdx.x;
	uint4 *hq = gsm + lane;
	uint32_t *sq = reinterpret_cast<uint32_t *>(gsm + (size_t)nq_cap * GT) + lane;
	uint2 *mrow = reinterpret_cast<uint2 *>(reinterpret_cast<uint32_t *>(gsm + (size_t)nq_cap * GT) + (size_t)nq_cap * GT);
	KswFastEdge *edge = reinterpret_cast<KswFastEdge *>(mrow + 6);
	if (lane < 5) {
		uint2 r;
		r.x = ((uint32_t)(P.mat[lane * 5 + 0] + G_SC)) | ((uint32_t)(P.mat[lane * 5 + 1] + G_SC) << 8) |
		      ((uint32_t)(P.mat[lane * 5 + 2] + G_SC) << 16) | ((uint32_t)(P.mat[lane * 5 + 3] + G_SC) << 24);
		r.y = (uint32_t)(P.mat[lane * 5 + 4] + G_SC);      // byte 4; bytes 5..7 are the zero the selectors' high nibbles pick
		mrow[lane] = r;
		KswFastEdge e;
		ksw_fast_edge_entry(lane, e);
		edge[lane] = e;
	}
	__syncwarp();
	const int oe_ins = P.o_ins + P.e_ins, oe_del = P.o_del + P.e_del;
	const uint32_t neg_ei = pk2(-P.e_ins), neg_oei = pk2(-oe_ins), ed32 = pk2(P.e_del), oed32 = pk2(oe_del);

	for (;;) {
		unsigned g = 0;
		if (lane == 0) g = atomicAdd(counter, 1u);
		g = __shfl_sync(0xffffffffu, g, 0);
		if (g >= (unsigned)n_groups) break;
		const DevGGroup grp = groups[g];
		const bool live = lane < grp.n;
		DevGJob jb;
		jb.seq_off = 0; jb.qlen = 1; jb.tlen = 0; jb.w = 0; jb.idx = 0;
		if (live) jb = jobs[gorder[grp.first + lane]];
		const int qlen = jb.qlen, tlen = jb.tlen, w = jb.w;
		const uint8_t *query = seq + jb.seq_off, *target = query + qlen;
		uint2 *zl = z + (size_t)grp.z_off + lane;
		// row -1 (ksw.c:520-523): eh[0].h = 0, eh[j].h = -(o_ins + e_ins j) for j <= w, -inf beyond; every e = -inf
		if (live) {
			const int nq = (qlen >> 2) + 1;
			for (int q = 0; q < nq; ++q) {
				int hv[4];
				uint32_t sb[4];
#pragma unroll
				for (int k = 0; k < 4; ++k) {
					const int c = (q << 2) + k;
					int v = c == 0 ? 0 : (c <= w ? -(P.o_ins + P.e_ins * c) : G_NEG - G_BIAS);
					if (c > qlen) v = G_NEG - G_BIAS;
					hv[k] = v + G_BIAS - G_SC;
					const uint32_t code = c < qlen ? (uint32_t)query[c] : 0u;
					sb[k] = (code > 4u ? 4u : code) | 0x50u;                  // byte look-up, zero above it
				}
				uint4 v4;
				v4.x = (uint32_t)hv[0] | ((uint32_t)hv[2] << 16);
				v4.y = pk2(G_NEG);
				v4.z = (uint32_t)hv[1] | ((uint32_t)hv[3] << 16);
				v4.w = pk2(G_NEG);
				hq[q * GT] = v4;
				sq[q * GT] = (sb[0] | (sb[2] << 8)) | ((sb[1] | (sb[3] << 8)) << 16);
			}
		}
		__syncwarp();
		for (int i = 0; i < grp.rows; ++i) {
			if (i >= tlen) continue;                                          // (the lane's job is done or absent)
			uint32_t t = target[i];
			t = t > 4u ? 4u : t;
			const uint2 mr = mrow[t];
			const int beg = i > w ? i - w : 0;                                // ksw.c:529-530
			const int end = i + w + 1 < qlen ? i + w + 1 : qlen;
			const int q0 = beg >> 2, q1 = (end - 1) >> 2, hi_rel = end - (q1 << 2);
			const KswFastEdge eL = edge[beg & 3], eR = edge[hi_rel];
			uint32_t X = (uint32_t)(G_NEG + oe_ins);                          // f = -inf entering the band (ksw.c:527)
			uint32_t Hc = 0;
			uint4 *ph = hq + q0 * GT;
			const uint32_t *ps = sq + q0 * GT;
			uint2 *pz = zl + (size_t)i * grp.nqb * GT;
			if (q1 == q0) {
				gquad<true>(X, Hc, *ph, *ps, mr, ph, pz, neg_ei, neg_oei, ed32, oed32, eL.geA & eR.ltA, eL.geB & eR.ltB, eR.onlyA, eR.onlyB);
			} else {
				uint4 vn = ph[GT];
				uint32_t swn = ps[GT];
				gquad<true>(X, Hc, *ph, *ps, mr, ph, pz, neg_ei, neg_oei, ed32, oed32, eL.geA, eL.geB, 0u, 0u);
				uint4 *const pl = hq + q1 * GT;
				ph += GT; ps += GT; pz += GT;
#pragma unroll 2
				for (; ph != pl; ph += GT, ps += GT, pz += GT) {
					const uint4 v = vn;
					const uint32_t sw = swn;
					vn = ph[GT]; swn = ps[GT];
					gquad<false>(X, Hc, v, sw, mr, ph, pz, neg_ei, neg_oei, ed32, oed32, 0u, 0u, 0u, 0u);
				}
				gquad<true>(X, Hc, vn, swn, mr, pl, pz, neg_ei, neg_oei, ed32, oed32, eR.ltA, eR.ltB, eR.onlyA, eR.onlyB);
			}
			// eh[beg].h = h1 (ksw.c:542): H(i,-1) = -(o_del + e_del (i+1)) while the band starts at column 0, never read otherwise
			if (beg == 0) *reinterpret_cast<uint16_t *>(&hq[0]) = (uint16_t)(G_BIAS - G_SC - (P.o_del + P.e_del * (i + 1)));
			if (hi_rel == 4) {
				// column `end` opens the next quad: eh[end] = {H(i,end-1), -inf} (ksw.c:558); the other half-words of the 64-bit
				// slot belong to column end+2, which is rewritten before it is read (the band end moves one column per row)
				uint2 w2;
				w2.x = Hc >> 16;                                              // H(c3) of the last quad, already minus G_SC
				w2.y = pk2(G_NEG);
				*reinterpret_cast<uint2 *>(&hq[(q1 + 1) * GT]) = w2;
			}
		}
		if (live) {
			// score = eh[qlen].h after the last row (ksw.c:560)
			const uint16_t hv = *(reinterpret_cast<const uint16_t *>(&hq[(qlen >> 2) * GT]) + ksw_fast_hslot(qlen));
			res[jb.idx].score = (int)hv + G_SC - G_BIAS;
		}
		__syncwarp();
	}
}

// ---------------------------------------------------------------- backtrack by recomputation
struct GWalk {
	const uint2 *z;                // the job's group slab + lane
	const uint8_t *query, *target;
	const int8_t *mat;
	int qlen, tlen, w, nqb;
	int o_del, e_del, o_ins, e_ins;

	__device__ __forceinline__ int hcell(int i, int k) const       // H(i,k) incl. the boundary row / column and "outside the band"
	{
		if (i < 0) return k < 0 ? 0 : ((k + 1 <= w && k + 1 <= qlen) ? -(o_ins + e_ins * (k + 1)) : G_MINF);
		if (k < 0) return i <= w ? -(o_del + e_del * (i + 1)) : G_MINF;
		const int beg = i > w ? i - w : 0, end = i + w + 1 < qlen ? i + w + 1 : qlen;
		if (k < beg || k >= end) return G_MINF;
		const uint2 v = z[((size_t)i * nqb + ((k >> 2) - (beg >> 2))) * GT];
		const int c = k & 3;
		const uint32_t word = (c & 1) ? v.y : v.x;                 // pair B holds c1, c3; pair A c0, c2
		return (int)((c & 2) ? (word >> 16) : (word & 0xffffu)) - G_BIAS;
	}
	__device__ __forceinline__ int mcell(int i, int k) const       // M(i,k) = H(i-1,k-1) + S(i,k)
	{
		const int d = hcell(i - 1, k - 1);
		int t = target[i], q = query[k];
		t = t > 4 ? 4 : t; q = q > 4 ? 4 : q;
		return d <= G_MINF / 2 ? G_MINF : d + mat[t * 5 + q];
	}
	template <class F>
	__device__ __forceinline__ int run(F &&emit) const
	{
		int i = tlen - 1, k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;        // the last cell (ksw.c:565)
		int which = 0, n = 0, cur_op = -1, cur_len = 0, gapv = 0;
		auto push = [&](int op, int len) {                                       // push_cigar, ksw.c:486-499
			if (op == cur_op) { cur_len += len; return; }
			if (cur_op >= 0) emit(n++, cur_op, cur_len);
			cur_op = op; cur_len = len;
		};
		const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
		while (i >= 0 && k >= 0) {
			const int m = mcell(i, k);
			if (which == 0) {
				const int h = hcell(i, k);
				if (h == m) which = 0;
				else {
					// E(i,k) = max over the rows above of M(i',k) - oe_del - (i-1-i') e_del: does it reach h?
					which = 2;
					const int top = k - w > 0 ? k - w : 0;
					for (int r = i - 1; r >= top; --r) {
						const int mm = mcell(r, k);
						if (mm > G_MINF / 2 && mm - oe_del - (i - 1 - r) * e_del == h) { which = 1; break; }
					}
				}
				gapv = h;                                                          // E(i,k) resp. F(i,k) if the path turns into a gap here
			} else if (which == 1) {
				if (m > G_MINF / 2 && m - oe_del == gapv) which = 0;               // opened here
				else gapv += e_del;                                                // extended: E(i,k) = E(i+1,k) + e_del
			} else {
				if (m > G_MINF / 2 && m - oe_ins == gapv) which = 0;
				else gapv += e_ins;
			}
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

__global__ void __launch_bounds__(128)
ksw_gfast_trace_kernel(const DevGJob *__restrict__ jobs, const uint8_t *__restrict__ seq, const KswParams P, const uint32_t *__restrict__ gorder,
                       const DevGGroup *__restrict__ groups, const int n_groups, const uint2 *__restrict__ z,
                       unsigned long long *__restrict__ pool_used, uint32_t *__restrict__ cigar_pool, DevGRes *__restrict__ res)
{
	const int g = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
	if (g >= n_groups) return;
	const DevGGroup grp = groups[g];
	if (lane >= grp.n) return;
	const DevGJob jb = jobs[gorder[grp.first + lane]];
	GWalk wk;
	wk.z = z + (size_t)grp.z_off + lane;
	wk.query = seq + jb.seq_off; wk.target = wk.query + jb.qlen; wk.mat = P.mat;
	wk.qlen = jb.qlen; wk.tlen = jb.tlen; wk.w = jb.w; wk.nqb = grp.nqb;
	wk.o_del = P.o_del; wk.e_del = P.e_del; wk.o_ins = P.o_ins; wk.e_ins = P.e_ins;
	const int n = wk.run([](int, int, int) {});
	const unsigned long long off = atomicAdd(pool_used, (unsigned long long)n);
	uint32_t *out = cigar_pool + off;
	wk.run([&](int r_idx, int op, int len) { out[n - 1 - r_idx] = (uint32_t)len << 4 | (uint32_t)op; });
	res[jb.idx].n_cigar = n;
	res[jb.idx].cigar_off = (long long)off;
}

} // namespace

size_t ksw_gfast_smem_bytes(int qmax)
{
	return (size_t)((qmax >> 2) + 1) * GT * (sizeof(uint4) + sizeof(uint32_t)) + 6 * sizeof(uint2) + 5 * sizeof(KswFastEdge);
}

cudaError_t ksw_launch_gfast(const DevGJob *jobs, const uint8_t *seq, const KswParams &P, const uint32_t *gorder, const DevGGroup *groups,
                             int n_groups, int qmax, int sm_count, uint2 *z, unsigned *counter, unsigned long long *pool_used,
                             uint32_t *cigar_pool, DevGRes *res, cudaStream_t st)
{
	if (n_groups <= 0) return cudaSuccess;
	const size_t smem = ksw_gfast_smem_bytes(qmax);
	static unsigned long long raised = 0;
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	if (dev < 0 || dev >= 64 || !((__atomic_load_n(&raised, __ATOMIC_ACQUIRE) >> dev) & 1ull)) {
		int optin = 0;
		e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
		if (e != cudaSuccess) return e;
		e = cudaFuncSetAttribute(ksw_gfast_dp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
		if (e != cudaSuccess) return e;
		if (dev >= 0 && dev < 64) __atomic_fetch_or(&raised, 1ull << dev, __ATOMIC_RELEASE);
	}
	int per_sm = 0;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ksw_gfast_dp_kernel, GT, smem);
	if (e != cudaSuccess) return e;
	if (per_sm < 1) return cudaErrorLaunchOutOfResources;
	int blocks = sm_count * per_sm;
	if (blocks > n_groups) blocks = n_groups;
	e = cudaMemsetAsync(counter, 0, sizeof(unsigned), st);
	if (e != cudaSuccess) return e;
	ksw_gfast_dp_kernel<<<blocks, GT, smem, st>>>(jobs, seq, P, gorder, groups, n_groups, (qmax >> 2) + 1, z, counter, res);
	e = cudaGetLastError();
	if (e != cudaSuccess) return e;
	const int tb = (n_groups * 32 + 127) / 128;
	ksw_gfast_trace_kernel<<<tb, 128, 0, st>>>(jobs, seq, P, gorder, groups, n_groups, z, pool_used, cigar_pool, res);
	return cudaGetLastError();
}
