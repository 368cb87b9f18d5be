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
#include "ksw_gfast_core.h"
#include "ksw_launch.h"

namespace {

constexpr int GT = 32;

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
		mrow[lane] = ksw_gfast_matrow(P, lane);
		KswFastEdge e;
		ksw_fast_edge_entry(lane, e);
		edge[lane] = e;
	}
	__syncwarp();
	KswGConst C;
	ksw_gfast_make_const(P, C);

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
		if (live) ksw_gfast_setup<GT>(hq, sq, qlen, w, query, C);
		for (int i = 0; i < tlen; ++i) {                                      // lanes of a group have similar tlen (sorted)
			uint32_t t = target[i];
			t = t > 4u ? 4u : t;
			ksw_gfast_row<GT>(hq, sq, edge, mrow[t], zl + (size_t)i * grp.nqb * GT, i, qlen, w, C);
		}
		if (live) res[jb.idx].score = ksw_gfast_score<GT>(hq, qlen);
		__syncwarp();
	}
}

__global__ void __launch_bounds__(128)
ksw_gfast_trace_kernel(const DevGJob *__restrict__ jobs, const uint8_t *__restrict__ seq, const KswParams P, const uint32_t *__restrict__ gorder,
                       const DevGGroup *__restrict__ groups, const int n_groups, const uint2 *__restrict__ z,
                       uint32_t *__restrict__ scratch, unsigned long long *__restrict__ pool_used, uint32_t *__restrict__ cigar_pool,
                       DevGRes *__restrict__ res)
{
	const int g = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
	if (g >= n_groups) return;
	const DevGGroup grp = groups[g];
	if (lane >= grp.n) return;
	const DevGJob jb = jobs[gorder[grp.first + lane]];
	KswGWalk<GT> wk;
	wk.z = z + (size_t)grp.z_off + lane;
	wk.query = seq + jb.seq_off; wk.target = wk.query + jb.qlen; wk.mat = P.mat;
	wk.qlen = jb.qlen; wk.tlen = jb.tlen; wk.w = jb.w; wk.nqb = grp.nqb;
	wk.o_del = P.o_del; wk.e_del = P.e_del; wk.o_ins = P.o_ins; wk.e_ins = P.e_ins;
	// one walk: the operations come out last first, into the job's worst-case slot (qlen + tlen + 2 words at seq_off + 2 idx);
	// then they are reversed into the dense pool
	uint32_t *mine = scratch + jb.seq_off + 2ull * jb.idx;
	const int n = wk.run([&](int r_idx, int op, int len) { mine[r_idx] = (uint32_t)len << 4 | (uint32_t)op; });
	const unsigned long long off = atomicAdd(pool_used, (unsigned long long)n);
	uint32_t *out = cigar_pool + off;
	for (int r = 0; r < n; ++r) out[n - 1 - r] = mine[r];
	res[jb.idx].n_cigar = n;
	res[jb.idx].cigar_off = (long long)off;
}

} // namespace
size_t ksw_gfast_smem_bytes(int qmax)
{
	return (size_t)((qmax >> 2) + 1) * GT * (sizeof(uint4) + sizeof(uint32_t)) + 6 * sizeof(uint2) + 5 * sizeof(KswFastEdge);
}

cudaError_t ksw_launch_gfast(const DevGJob *jobs, const uint8_t *seq, const KswParams &P, const uint32_t *gorder, const DevGGroup *groups,
                             int n_groups, int qmax, int sm_count, uint2 *z, unsigned *counter, uint32_t *scratch,
                             unsigned long long *pool_used, uint32_t *cigar_pool, DevGRes *res, cudaStream_t st)
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
	ksw_gfast_trace_kernel<<<tb, 128, 0, st>>>(jobs, seq, P, gorder, groups, n_groups, z, scratch, pool_used, cigar_pool, res);
	return cudaGetLastError();
}
