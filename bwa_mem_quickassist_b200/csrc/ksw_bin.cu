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

__global__ void __launch_bounds__(256)
ksw_bin_keys_kernel(const DevJob *__restrict__ jobs, long long n, uint16_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
	const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	const DevJob jb = jobs[k];
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
	keys[k] = (uint16_t)((cls >= KSW_CLASS_GENERIC ? (cls == KSW_CLASS_WARP ? 0x8000u : 0xC000u) : (cls << 13)) | (tl << 7) | hb);
	vals[k] = (uint32_t)k;
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
