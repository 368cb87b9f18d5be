// ksw_bin.cu — binning on the device.  The host packer only streams (DevJob[k] describes the caller's job k); here
// a key is computed per job and (key, k) pairs are radix-sorted (cub::DeviceRadixSort, 16-bit keys), which yields the
// order in which the extension kernels take the jobs:
//     key = kernel class | 63 - g(qlen) | 127 - f(h0)   (f exact below 96, then in steps of 16)
// fast classes first (one contiguous range per class, so a launch is a sub-range), then the int32 kernels' jobs; inside a
// class long jobs first (short tail at the end of a launch), then by carried-in score, which sets the band width.
// A warp claims chunks of consecutive entries of this order, so the jobs it works on at any moment are alike.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include "ksw_dev.cuh"
#include "ksw_launch.h"

namespace {

__device__ __forceinline__ uint32_t ksw_bin_key(const DevJob &jb)
{
	const uint32_t cls = (jb.flags >> KSW_CLASS_SHIFT) & KSW_CLASS_MASK;
	// query length (6 bits; steps of 2 / 2 / 4 / 8 columns for the four classes): the band of a row is capped by it, and
	// the lanes of a warp wait for the widest band in every row.  (It replaced the row count tlen/16: how long a job
	// lasts does not matter, a lane draws its next job as soon as it is done.  Real bwa mem mix, PE150: lanes busy in
	// the interior loop 62 % -> see DESIGN.md §7.)
	const uint32_t tl = 63u - min((uint32_t)jb.qlen >> (cls >= 3u ? 3 : (cls == 2u ? 2 : 1)), 63u);
	// carried-in score: exact below 96 (seed scores and most left-extension scores), 16 per step above — it sets the
	// band a job starts with, and the jobs of a warp finish their rows together only if their bands are alike
	const uint32_t h0 = (uint32_t)jb.h0;
	const uint32_t hb = 127u - (h0 < 96u ? h0 : 96u + min((h0 - 96u) >> 4, 31u));   // 7 bits
	// fast classes: cls << 13; warp-cooperative jobs 0x8000 | ..., thread-per-job jobs 0xC000 | ... (13 key bits below)
	return (cls >= KSW_CLASS_GENERIC ? (cls == KSW_CLASS_WARP ? 0x8000u : 0xC000u) : (cls << 13)) | (tl << 7) | hb;
}

__device__ __forceinline__ uint32_t ksw_class_lowest_key(int c)
{
	return c >= (int)KSW_CLASS_GENERIC ? (c == (int)KSW_CLASS_WARP ? 0x8000u : 0xC000u) : ((uint32_t)c << 13);
}

__global__ void __launch_bounds__(256)
ksw_bin_keys_kernel(const DevJob *__restrict__ jobs, long long n, uint16_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
	const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	keys[k] = (uint16_t)ksw_bin_key(jobs[k]);
	vals[k] = (uint32_t)k;
}

// ---- the same binning as a counting sort WITHOUT shared memory (three kernels): the pinned-caller pipeline launches it
// while the previous chunk's extension kernels fill every SM's shared memory, and a kernel that needs none still finds
// room beside them (ksw_runtime.cu); the radix sort would wait for them to drain, and the next chunk's extension kernels
// with it.  Equal keys come out in the order their jobs got there, which the kernels do not care about.
__global__ void __launch_bounds__(256)
ksw_bin_hist_kernel(const DevJob *__restrict__ jobs, long long n, uint16_t *__restrict__ keys, uint32_t *__restrict__ hist)
{
	const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const bool live = k < n;
	uint32_t key = 0xffffffffu;
	if (live) { key = ksw_bin_key(jobs[k]); keys[k] = (uint16_t)key; }
	// one atomic per distinct key of the warp
	const unsigned act = __ballot_sync(0xffffffffu, live);
	if (!live) return;
	const unsigned same = __match_any_sync(act, key);
	if ((threadIdx.x & 31) == __ffs(same) - 1) atomicAdd(&hist[key], (uint32_t)__popc(same));
}

#define KSW_BIN_SCAN_THREADS 512
#define KSW_BIN_KEYS 65536
// Exclusive scan of the 65536-entry histogram by ONE block without shared memory.  Each of the 16 warps owns 4096
// consecutive keys and walks them 32 at a time, lane l on key base + 32 * step + l: every load and store of a warp is one
// 128-byte line (the first form of this kernel gave each thread 128 consecutive keys: 32 lines per warp access, 0.11 ms —
// the longest kernel of a chunk's packing / binning chain, which sits on the critical path of the pinned-caller pipeline).
__global__ void __launch_bounds__(KSW_BIN_SCAN_THREADS)
ksw_bin_scan_kernel(const uint32_t *__restrict__ hist, uint32_t *__restrict__ cursor, uint32_t *__restrict__ scratch,
                    uint32_t *__restrict__ range, long long n)
{
	constexpr int WARPS = KSW_BIN_SCAN_THREADS / 32;                // 16
	constexpr int PER_WARP = KSW_BIN_KEYS / WARPS;                  // 4096 keys per warp
	constexpr int STEPS = PER_WARP / 32;                            // 128
	const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
	const uint32_t *h = hist + (size_t)wid * PER_WARP + lane;
	// pass 1: the warp's total
	uint32_t sum = 0;
#pragma unroll 8
	for (int i = 0; i < STEPS; ++i) sum += h[i * 32];
	for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
	if (lane == 0) scratch[wid] = sum;                              // warp totals through global memory: no shared memory here
	__threadfence_block();
	__syncthreads();
	if (wid == 0) {
		uint32_t w = lane < WARPS ? scratch[lane] : 0u, wi = w;
		for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += v; }
		if (lane < WARPS) scratch[32 + lane] = wi - w;
	}
	__threadfence_block();
	__syncthreads();
	// pass 2: exclusive prefix of every key; the first entry of every kernel class in the binned order (class bounds sit on
	// multiples of 0x2000 / 0x4000)
	uint32_t run = scratch[32 + wid];                               // keys before this warp's first key
	uint32_t *c = cursor + (size_t)wid * PER_WARP + lane;
#pragma unroll 4
	for (int i = 0; i < STEPS; ++i) {
		const uint32_t v = h[i * 32];
		uint32_t incl = v;
		for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
		const uint32_t excl = run + incl - v;
		c[i * 32] = excl;
		const uint32_t key = (uint32_t)(wid * PER_WARP + i * 32 + lane);
		if ((key & 0x1fffu) == 0u)
			for (int cl = 0; cl < KSW_N_CLASSES; ++cl) if (key == ksw_class_lowest_key(cl)) range[cl] = excl;
		run += __shfl_sync(0xffffffffu, incl, 31);
	}
	if (t == 0) range[KSW_N_CLASSES] = (uint32_t)n;
}

__global__ void __launch_bounds__(256)
ksw_bin_scatter_kernel(const uint16_t *__restrict__ keys, long long n, uint32_t *__restrict__ cursor, uint32_t *__restrict__ order)
{
	const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const bool live = k < n;
	const unsigned act = __ballot_sync(0xffffffffu, live);
	if (!live) return;
	const uint32_t key = keys[k];
	const unsigned same = __match_any_sync(act, key);
	const int leader = __ffs(same) - 1, lane = threadIdx.x & 31;
	uint32_t base = 0;
	if (lane == leader) base = atomicAdd(&cursor[key], (uint32_t)__popc(same));
	base = __shfl_sync(same, base, leader);
	order[base + __popc(same & ((1u << lane) - 1u))] = (uint32_t)k;
}

} // namespace

size_t ksw_bin_temp_bytes(int64_t n)
{
	size_t bytes = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint16_t *)nullptr, (uint16_t *)nullptr,
	                                (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)(n > 0 ? n : 1), 0, 16);
	return bytes;
}

cudaError_t ksw_launch_bin(const DevJob *jobs, int64_t n, uint16_t *keys_in, uint16_t *keys_out, uint32_t *vals_in,
                           uint32_t *order, void *temp, size_t temp_bytes, cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	const unsigned blocks = (unsigned)((n + 255) / 256);
	ksw_bin_keys_kernel<<<blocks, 256, 0, st>>>(jobs, (long long)n, keys_in, vals_in);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return e;
	return cub::DeviceRadixSort::SortPairs(temp, temp_bytes, (const uint16_t *)keys_in, keys_out, (const uint32_t *)vals_in, order,
	                                       (int)n, 0, 16, st);
}

size_t ksw_bin_counting_bytes(void) { return sizeof(uint32_t) * (2 * (size_t)KSW_BIN_KEYS + 64); }

cudaError_t ksw_launch_bin_counting(const DevJob *jobs, int64_t n, uint16_t *keys, void *work, uint32_t *order, uint32_t *range,
                                    cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	uint32_t *hist = (uint32_t *)work, *cursor = hist + KSW_BIN_KEYS, *scratch = cursor + KSW_BIN_KEYS;
	cudaError_t e = cudaMemsetAsync(hist, 0, sizeof(uint32_t) * KSW_BIN_KEYS, st);
	if (e != cudaSuccess) return e;
	const unsigned blocks = (unsigned)((n + 255) / 256);
	ksw_bin_hist_kernel<<<blocks, 256, 0, st>>>(jobs, (long long)n, keys, hist);
	ksw_bin_scan_kernel<<<1, KSW_BIN_SCAN_THREADS, 0, st>>>(hist, cursor, scratch, range, (long long)n);
	ksw_bin_scatter_kernel<<<blocks, 256, 0, st>>>(keys, (long long)n, cursor, order);
	return cudaGetLastError();
}
