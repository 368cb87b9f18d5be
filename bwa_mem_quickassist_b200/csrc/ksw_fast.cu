// ksw_fast.cu — the fast extension kernel for sm_100a: one job per lane, one warp per CTA,
// H/E columns and PRMT score selectors of the 32 jobs interleaved in shared memory
// (quad q of lane l at hq[q*32+l]: every LDS.128/STS.128 of a warp is conflict-free whatever
// column each lane is at), DPX s16x2 arithmetic (see ksw_fast_core.h for the per-lane logic).
//
// Scheduling: persistent CTAs (SM count x CTAs that fit by shared memory); each lane pulls its
// next job as soon as its previous job ends, so lanes only wait for each other inside one DP row, never for a
// whole job.  Jobs are taken in the binned order built on the device (ksw_bin.cu), so that the lanes of a warp
// sweep bands of similar width.
#include <cuda_runtime.h>
#include <atomic>
#include <cstdlib>
#include "ksw_dev.cuh"
#include "ksw_fast_core.h"
#include "ksw_launch.h"

namespace {

constexpr int T = KSW_FAST_THREADS;   // 32: one warp per CTA

template <bool KEYED>
__global__ void __launch_bounds__(T, 12)
ksw_fast_kernel(const DevJob *__restrict__ jobs, long long n_jobs, const uint32_t *__restrict__ pool,
                const uint32_t *__restrict__ npool, const KswParams P, const int nq_cap, const int chunk,
                unsigned long long *__restrict__ counter, const uint32_t *__restrict__ order,
                DevRes *__restrict__ res, uint32_t *__restrict__ cells, const uint32_t *__restrict__ drange, const int c_lo, const int c_hi)
{
	// device-packed batches (ksw_devpack.cu): the launch covers the kernel classes [c_lo, c_hi) of the binned order,
	// whose bounds only the device knows (n_jobs is then the host's upper bound, used for the grid size)
	if (drange) {
		const uint32_t f = drange[c_lo];
		n_jobs = (long long)drange[c_hi] - (long long)f;
		order += f;
	}
	extern __shared__ uint4 smem[];
	const int lane = threadIdx.x;
	uint4 *hq = smem;
	uint32_t *sq = reinterpret_cast<uint32_t *>(hq + (size_t)nq_cap * T);
	uint2 *mrow = reinterpret_cast<uint2 *>(sq + (size_t)nq_cap * T);
	KswFastEdge *edge = reinterpret_cast<KswFastEdge *>(mrow + 6);
	if (lane < 5) {
		mrow[lane] = ksw_fast_matrow(P, lane);
		KswFastEdge e;
		ksw_fast_edge_entry(lane, e);
		edge[lane] = e;
	}
	__syncwarp();

	KswFastConst K;
	ksw_fast_make_const(P, K);
	const KswFastMem<T> M{hq + lane, sq + lane, edge};
	KswFastLane L;
	L.tlen = 0; L.i = 0;
	enum { IDLE = 0, RUN = 1, DONE = 2 };
	int state = IDLE;

	// Job supply: the warp takes CHUNK (32 by default, see the launcher) consecutive jobs of the binned list at a time (one atomicAdd per chunk) and its
	// lanes draw from that chunk, so the 32 jobs a warp works on at any moment are neighbours in the binned order
	// (similar band width and length) even after the lanes have drifted apart in time.
	const long long CHUNK = chunk;
	long long wcur = 0, wend = 0;                                  // warp-uniform cursor into the current chunk
	bool exhausted = false;

	while (true) {
		unsigned need = __ballot_sync(0xffffffffu, state == IDLE);
		if (need) {
			DevJob jb;
			jb.seq_off = 0; jb.idx = 0; jb.qlen = 0; jb.tlen = 0; jb.h0 = 0; jb.w = 0; jb.flags = 0; jb.nmask_off = 0;
			bool got = false;
			for (int round = 0; round < 2 && need; ++round) {
				if (wcur >= wend && !exhausted) {
					unsigned long long base = 0;
					if (lane == 0) base = atomicAdd(counter, (unsigned long long)CHUNK);
					base = __shfl_sync(0xffffffffu, base, 0);
					wcur = (long long)base;
					wend = wcur + CHUNK < n_jobs ? wcur + CHUNK : n_jobs;
					if (wcur >= n_jobs) { exhausted = true; wend = wcur; }
				}
				const long long avail = wend - wcur;
				const int rank = __popc(need & ((1u << lane) - 1u));
				if (state == IDLE && !got && (long long)rank < avail) { jb = jobs[order[wcur + rank]]; got = true; }
				const int served = (long long)__popc(need) < avail ? __popc(need) : (int)avail;
				wcur += served;
				need = __ballot_sync(0xffffffffu, state == IDLE && !got);
			}
			if (state == IDLE && !got && exhausted) state = DONE;
			const unsigned fetched = __ballot_sync(0xffffffffu, got);
			if (__popc(fetched) >= 8) {
				// many lanes start together (e.g. at launch, or equal-length jobs): each builds its own state
				if (got) ksw_fast_setup_quads<T>(hq, sq, lane, 0, 1, K, jb.seq_off, jb.qlen, jb.h0, jb.flags, jb.nmask_off, pool, npool);
			} else {
				// a few stragglers: all 32 lanes build the state of each newly fetched job, one job after the other,
				// instead of 31 lanes idling through a one-lane setup
				for (unsigned todo = fetched; todo; todo &= todo - 1) {
					const int owner = __ffs(todo) - 1;
					const uint32_t o_seq = __shfl_sync(0xffffffffu, jb.seq_off, owner);
					const int o_qlen = __shfl_sync(0xffffffffu, jb.qlen, owner);
					const int o_h0 = __shfl_sync(0xffffffffu, jb.h0, owner);
					const uint32_t o_flags = __shfl_sync(0xffffffffu, jb.flags, owner);
					const uint32_t o_nmask = __shfl_sync(0xffffffffu, jb.nmask_off, owner);
					ksw_fast_setup_quads<T>(hq, sq, owner, lane, T, K, o_seq, o_qlen, o_h0, o_flags, o_nmask, pool, npool);
				}
			}
			__syncwarp();
			if (got) { ksw_fast_init_lane(L, jb, pool, npool); state = RUN; }
		}
		if (__all_sync(0xffffffffu, state == DONE)) break;
		// rows, until a lane runs out of work: one vote per row instead of the two of the supply logic above
		do {
			if (state == RUN) {
				if (ksw_fast_row<T, KEYED>(L, M, K, mrow)) {
					DevRes r;
					ksw_fast_result(L, r);
					res[L.idx] = r;
					cells[L.idx] = L.cells;
					state = IDLE;
				}
			}
		} while (!__any_sync(0xffffffffu, state == IDLE));
	}
}

} // namespace

size_t ksw_fast_smem_bytes(int qmax)
{
	return (size_t)KSW_FAST_QUADS(qmax) * T * (sizeof(uint4) + sizeof(uint32_t)) + 6 * sizeof(uint2) + 5 * sizeof(KswFastEdge);
}

template <bool KEYED>
static cudaError_t launch_fast_t(const DevJob *jobs, int64_t n_jobs, const uint32_t *pool, const uint32_t *npool,
                                 const KswParams &P, int qmax, int sm_count, unsigned long long *counter,
                                 const uint32_t *order, DevRes *res, uint32_t *cells, cudaStream_t st,
                                 const uint32_t *drange, int c_lo, int c_hi, int ctas_cap)
{
	const size_t smem = ksw_fast_smem_bytes(qmax);
	// The dynamic shared-memory ceiling of the kernel is raised ONCE per device to the opt-in maximum and never lowered:
	// several host threads (one context each) launch this kernel concurrently with different sizes, and a per-launch
	// cudaFuncSetAttribute would let one thread shrink the ceiling under another thread's larger launch.
	static std::atomic<unsigned long long> raised[2] = {{0ull}, {0ull}};     // bit d: done for device d (KEYED / not)
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	if (dev < 0 || dev >= 64 || !((raised[KEYED ? 1 : 0].load() >> dev) & 1ull)) {
		int optin = 0;
		e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
		if (e != cudaSuccess) return e;
		e = cudaFuncSetAttribute(ksw_fast_kernel<KEYED>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
		if (e != cudaSuccess) return e;
		if (dev >= 0 && dev < 64) raised[KEYED ? 1 : 0].fetch_or(1ull << dev);
	}
	int per_sm = 0;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ksw_fast_kernel<KEYED>, T, smem);
	if (e != cudaSuccess) return e;
	if (per_sm < 1) return cudaErrorLaunchOutOfResources;
	if (ctas_cap > 0 && ctas_cap < per_sm) per_sm = ctas_cap;
	else if (ctas_cap < 0 && per_sm + ctas_cap >= 6) per_sm += ctas_cap;      // leave -ctas_cap CTAs' worth of room per SM
	if (const char *ev = getenv("KSW_B200_FAST_CTAS")) per_sm = atoi(ev) > 0 && atoi(ev) < per_sm ? atoi(ev) : per_sm;   // tuning knob
	long long blocks = (long long)sm_count * per_sm;
	const long long need = (n_jobs + T - 1) / T;
	if (blocks > need) blocks = need;
	e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st);
	if (e != cudaSuccess) return e;
	// consecutive jobs of the binned list a warp claims at once (one atomicAdd per claim).  Measured on config 2 with the
	// exact-h0 binning key: 32 -> 14.33 ms, 128 -> 14.47, 256 -> 14.68, 1024 -> 15.65 per 4 M jobs: neighbours in the
	// binned order are alike anyway, so the smallest claim (best load balance) wins.
	long long chunk = 32;
	if (const char *ev = getenv("KSW_B200_FAST_CHUNK")) chunk = atoi(ev) > 0 ? atoi(ev) : chunk;      // tuning knob
	ksw_fast_kernel<KEYED><<<(unsigned)blocks, T, smem, st>>>(jobs, (long long)n_jobs, pool, npool, P,
	                                                           KSW_FAST_QUADS(qmax), (int)chunk, counter, order, res, cells,
	                                                           drange, c_lo, c_hi);
	return cudaGetLastError();
}

cudaError_t ksw_launch_fast(const DevJob *jobs, int64_t n_jobs, const uint32_t *pool, const uint32_t *npool,
                            const KswParams &P, int qmax, bool keyed, int sm_count, unsigned long long *counter,
                            const uint32_t *order, DevRes *res, uint32_t *cells, cudaStream_t st,
                            const uint32_t *drange, int c_lo, int c_hi, int ctas_per_sm_cap)
{
	if (n_jobs <= 0) return cudaSuccess;
	return keyed ? launch_fast_t<true>(jobs, n_jobs, pool, npool, P, qmax, sm_count, counter, order, res, cells, st, drange, c_lo, c_hi, ctas_per_sm_cap)
	             : launch_fast_t<false>(jobs, n_jobs, pool, npool, P, qmax, sm_count, counter, order, res, cells, st, drange, c_lo, c_hi, ctas_per_sm_cap);
}
