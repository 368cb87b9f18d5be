// ksw_pair.cu — the pair extension kernel for sm_100a: TWO jobs per lane (one in each half of the s16x2 registers),
// so a warp sweeps 64 extensions at once and every DPX instruction updates 64 DP cells.  One persistent CTA per SM with
// as many warps as fit into the 227 KB of shared memory (7 for 101 bp queries); the warps are independent (each owns
// a slice of the shared memory and synchronises only with __syncwarp).
// Per-lane logic: ksw_pair_core.h.  Shared memory: H/E columns and PRMT selectors of the 32 lanes interleaved
// (column pair p of lane l at he[p*32+l]: every LDS.128/STS.128 of a warp is conflict-free whatever column each lane
// is at); 20 bytes per column pair and lane, i.e. 10 bytes per job and column.
//
// Scheduling: each of the 64 job slots of a warp pulls its
// next job as soon as its previous job ends.  Jobs are taken in the binned order built on the device (ksw_bin.cu), in
// chunks of consecutive entries per warp, so that the two jobs of a lane — and the 64 of a warp — sweep bands of similar
// position and width.
#include <cuda_runtime.h>
#include <atomic>
#include "ksw_dev.cuh"
#include "ksw_pair_core.h"
#include "ksw_launch.h"

namespace {

constexpr int T = KSW_FAST_THREADS;   // 32 lanes share one interleaved slice
constexpr int MAX_WARPS = 8;          // warps per CTA (256 threads: the full register file stays available per thread)

__global__ void __launch_bounds__(T * MAX_WARPS, 1)
ksw_pair_kernel(const DevJob *__restrict__ jobs, long long n_jobs, const uint32_t *__restrict__ pool,
                const uint32_t *__restrict__ npool, const KswParams P, const int np_cap, const int chunk,
                unsigned long long *__restrict__ counter, const uint32_t *__restrict__ order,
                DevRes *__restrict__ res, uint32_t *__restrict__ cells)
{
	extern __shared__ uint4 smem[];
	const int lane = threadIdx.x & (T - 1), warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
	// per warp: np_cap*T uint4 of H/E, then np_cap*T words of selectors (20*T bytes per column pair: a multiple of 16)
	uint4 *he = smem + (size_t)warp * np_cap * (T + T / 4);
	uint32_t *sq = reinterpret_cast<uint32_t *>(he + (size_t)np_cap * T);
	uint2 *mrow = reinterpret_cast<uint2 *>(smem + (size_t)n_warps * np_cap * (T + T / 4));
	if (threadIdx.x < 5) mrow[threadIdx.x] = ksw_fast_matrow(P, threadIdx.x);
	__syncthreads();

	KswFastConst K;
	ksw_fast_make_const(P, K);
	const KswPairMem<T> M{he + lane, sq + lane};
	KswFastLane L[2] = {};                                          // (the row code reads both slots' scalars before it masks)
	unsigned run = 0;                                               // bit X: slot X holds a running job
	unsigned dead = 0;                                              // bit X: slot X will get no more jobs

	const long long CHUNK = chunk;
	long long wcur = 0, wend = 0;                                  // warp-uniform cursor into the current chunk
	bool exhausted = false;

	while (true) {
#pragma unroll
		for (int X = 0; X < 2; ++X) {
			const bool idle = !((run | dead) & (1u << X));
			unsigned need = __ballot_sync(0xffffffffu, idle);
			if (!need) continue;
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
				if (idle && !got && (long long)rank < avail) { jb = jobs[order[wcur + rank]]; got = true; }
				const int served = (long long)__popc(need) < avail ? __popc(need) : (int)avail;
				wcur += served;
				need = __ballot_sync(0xffffffffu, idle && !got);
			}
			if (idle && !got && exhausted) dead |= 1u << X;
			const unsigned fetched = __ballot_sync(0xffffffffu, got);
			if (__popc(fetched) >= 8) {
				// many slots start together (at launch, or equal-length jobs): each lane builds its own state
				if (got) ksw_pair_setup<T>(he, sq, lane, X, 0, 1, K, jb.seq_off, jb.qlen, jb.h0, pool);
			} else {
				// a few stragglers: all 32 lanes build the state of each newly fetched job, one job after the other
				for (unsigned todo = fetched; todo; todo &= todo - 1) {
					const int owner = __ffs(todo) - 1;
					const uint32_t o_seq = __shfl_sync(0xffffffffu, jb.seq_off, owner);
					const int o_qlen = __shfl_sync(0xffffffffu, jb.qlen, owner);
					const int o_h0 = __shfl_sync(0xffffffffu, jb.h0, owner);
					ksw_pair_setup<T>(he, sq, owner, X, lane, T, K, o_seq, o_qlen, o_h0, pool);
				}
			}
			__syncwarp();
			if (got) { ksw_fast_init_lane(L[X], jb, pool, npool); run |= 1u << X; }
		}
		if (__all_sync(0xffffffffu, run == 0u && dead == 3u)) break;
		if (run) {
			const unsigned fin = ksw_pair_row<T>(L, run, M, K, mrow);
#pragma unroll
			for (int X = 0; X < 2; ++X) {
				if (!((fin >> X) & 1u)) continue;
				DevRes r;
				ksw_fast_result(L[X], r);
				res[L[X].idx] = r;
				cells[L[X].idx] = L[X].cells;
			}
			run &= ~fin;
		}
	}
}

} // namespace

size_t ksw_pair_smem_bytes(int qmax, int n_warps)
{
	return (size_t)n_warps * KSW_PAIR_COLPAIRS(qmax) * T * (sizeof(uint4) + sizeof(uint32_t)) + 6 * sizeof(uint2);
}

cudaError_t ksw_launch_pair(const DevJob *jobs, int64_t n_jobs, const uint32_t *pool, const uint32_t *npool,
                            const KswParams &P, int qmax, int sm_count, unsigned long long *counter,
                            const uint32_t *order, DevRes *res, uint32_t *cells, cudaStream_t st)
{
	if (n_jobs <= 0) return cudaSuccess;
	// raised ONCE per device and never lowered (see ksw_fast.cu: concurrent launches from several host threads)
	static std::atomic<unsigned long long> raised{0ull};
	static std::atomic<int> optin_of[64];
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	int optin = (dev >= 0 && dev < 64 && ((raised.load() >> dev) & 1ull)) ? optin_of[dev].load() : 0;
	if (!optin) {
		e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
		if (e != cudaSuccess) return e;
		e = cudaFuncSetAttribute(ksw_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
		if (e != cudaSuccess) return e;
		if (dev >= 0 && dev < 64) { optin_of[dev].store(optin); raised.fetch_or(1ull << dev); }
	}
	// one CTA per SM; as many warps as the shared memory holds (at most MAX_WARPS), fewer when the batch is small
	const size_t per_warp = (size_t)KSW_PAIR_COLPAIRS(qmax) * T * (sizeof(uint4) + sizeof(uint32_t));
	int n_warps = (int)(((size_t)optin - 6 * sizeof(uint2)) / per_warp);
	if (n_warps < 1) return cudaErrorLaunchOutOfResources;
	if (n_warps > MAX_WARPS) n_warps = MAX_WARPS;
	const long long need = (n_jobs + 2 * T - 1) / (2 * T);                     // warps the batch can occupy
	long long blocks = sm_count;
	if (need < blocks * n_warps) {
		n_warps = (int)((need + blocks - 1) / blocks);
		blocks = (need + n_warps - 1) / n_warps;
	}
	const size_t smem = ksw_pair_smem_bytes(qmax, n_warps);
	e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st);
	if (e != cudaSuccess) return e;
	// consecutive jobs a warp claims at once: one round of its 64 slots (see ksw_fast.cu: small claims balance best)
	long long chunk = 64;
	ksw_pair_kernel<<<(unsigned)blocks, T * n_warps, smem, st>>>(jobs, (long long)n_jobs, pool, npool, P, KSW_PAIR_COLPAIRS(qmax),
	                                                    (int)chunk, counter, order, res, cells);
	return cudaGetLastError();
}
