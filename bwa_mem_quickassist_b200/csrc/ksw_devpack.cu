// ksw_devpack.cu — packing on the device.  When the caller's job / sequence / result arrays are page-locked, the host
// does not touch the sequences at all: the raw job records (ksw_b200_job_t, 32 B) and the raw byte-coded sequences are
// copied to HBM as they are (a few large cudaMemcpyAsync calls per chunk), and three small kernels turn them into what
// the extension kernels read (ksw_dev.cuh):
//
//   ksw_prep_kernel   one thread per job: band clamp (ksw.c:398-406, same double expression as the host), kernel class,
//                     DevJob record, the job's slice of the 2-bit pool; chunk totals for the host (DevPackStats)
//   ksw_pack_kernel   sixteen lanes per job: 16 bases -> one 32-bit word per lane (base k in word k/16 at bits 2(k%16)),
//                     N masks for the rare jobs that hold an N (allocated with one atomicAdd per such job); a class-0
//                     job that holds an N moves to class 1, exactly as in the host packer (ksw_pack.cpp)
//   ksw_range_kernel  after the binning sort: first entry of each kernel class in the binned order, read by the
//                     extension kernels from device memory (the host never learns the post-demotion class sizes)
//
// Same routing source as the host packer (ksw_class.h), so a batch packed here runs through the same kernels in the
// same classes as the same batch packed on the host; tests compare the two paths bit for bit.
#include <cuda_runtime.h>
#include "../../include/ksw_b200.h"
#include "ksw_class.h"
#include "ksw_dev.cuh"
#include "ksw_launch.h"

namespace {

// No shared memory and no separate scan kernel: the kernel is launched while the extension kernels of the previous chunk
// fill every SM's shared memory, and must still find room beside them (a block without shared memory fits into the
// 1 KB that 13 extension CTAs leave free).  Offsets in the 2-bit pool come from a warp-level exclusive scan plus one
// atomicAdd per warp on the chunk total, so the order of the jobs' slices in the pool follows the order in which the
// warps get there — it has no meaning, every job finds its slice through DevJob::seq_off.
__global__ void __launch_bounds__(256)
ksw_prep_kernel(const ksw_b200_job_t *__restrict__ raw, long long n, const KswScoring S, DevJob *__restrict__ jobs,
                uint32_t *__restrict__ offs, DevPackStats *__restrict__ stats)
{
	const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const int lane = threadIdx.x & 31;
	unsigned long long my_nmask = 0, my_qhi = 0, my_thi = 0, my_qlo = 0, my_tlo = 0;   // *_lo hold ~offset
	uint32_t my_units = 0, cls = 0xffu;
	int my_qlen = 0;
	bool bad = false;
	if (k < n) {
		const ksw_b200_job_t j = raw[k];
		DevJob d;
		d.seq_off = 0; d.idx = (uint32_t)k; d.qlen = j.qlen; d.tlen = j.tlen;
		d.h0 = j.h0 < 0 ? 0 : j.h0;                                                    // ksw.c:384
		d.nmask_off = 0;
		if (j.qlen < 1 || j.tlen < 0) {
			bad = true;
			d.qlen = 1; d.tlen = 0; d.w = 0; d.flags = KSW_CLASS_THREAD << KSW_CLASS_SHIFT;
		} else {
			d.w = ksw_clamp_w_expr(j.qlen, S.maxsc, S.o_del, S.e_del, S.o_ins, S.e_ins, j.w, S.end_bonus);
			cls = ksw_job_class(S, j.qlen, d.h0);
			d.flags = cls << KSW_CLASS_SHIFT;
			my_units = ksw_job_units(j.qlen, j.tlen);
			my_nmask = ksw_words1(j.qlen) + ksw_words1(j.tlen);
			my_qhi = j.q_off + (unsigned long long)j.qlen;
			my_thi = j.t_off + (unsigned long long)j.tlen;
			my_qlo = ~(unsigned long long)j.q_off;
			if (j.tlen > 0) my_tlo = ~(unsigned long long)j.t_off;
			my_qlen = j.qlen;
		}
		jobs[k] = d;
	}
	// this job's slice: exclusive scan of the sizes inside the warp, one atomicAdd per warp
	uint32_t incl = my_units;
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl += v;
	}
	const uint32_t warp_units = __shfl_sync(0xffffffffu, incl, 31);
	unsigned long long base = 0;
	if (lane == 0 && warp_units) base = atomicAdd(&stats->units, (unsigned long long)warp_units);
	base = __shfl_sync(0xffffffffu, base, 0);
	if (k < n) offs[k] = (uint32_t)(base + incl - my_units);       // the host rejects chunks whose total exceeds 32 bits
	// chunk totals: warp-level reductions, then one global atomic per warp and quantity
	for (int o = 16; o > 0; o >>= 1) {
		my_nmask += __shfl_down_sync(0xffffffffu, my_nmask, o);
		const unsigned long long a = __shfl_down_sync(0xffffffffu, my_qhi, o), b = __shfl_down_sync(0xffffffffu, my_thi, o);
		const unsigned long long c = __shfl_down_sync(0xffffffffu, my_qlo, o), d = __shfl_down_sync(0xffffffffu, my_tlo, o);
		my_qhi = my_qhi > a ? my_qhi : a; my_thi = my_thi > b ? my_thi : b;
		my_qlo = my_qlo > c ? my_qlo : c; my_tlo = my_tlo > d ? my_tlo : d;
	}
	if (lane == 0) {
		if (my_nmask) atomicAdd(&stats->nmask_words, my_nmask);
		atomicMax(&stats->q_hi, my_qhi); atomicMax(&stats->t_hi, my_thi);
		atomicMax(&stats->q_lo_inv, my_qlo); atomicMax(&stats->t_lo_inv, my_tlo);
	}
	if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&stats->bad, 1u);
	for (uint32_t c = 0; c < KSW_N_CLASSES; ++c) {
		const unsigned m = __ballot_sync(0xffffffffu, cls == c);
		if (!m) continue;
		int q = cls == c ? my_qlen : 0;
		for (int o = 16; o > 0; o >>= 1) { const int v = __shfl_down_sync(0xffffffffu, q, o); q = q > v ? q : v; }
		if (lane == 0) { atomicAdd(&stats->class_n[c], (unsigned)__popc(m)); atomicMax(&stats->class_qmax[c], q); }
	}
}

// 16 byte codes starting at s (only the first `lim` exist) -> one 2-bit word; *n_bits gets bit x set where code x > 3.
// A full group of 16 is read as five aligned 32-bit words and funnel-shifted into place (the job slices start wherever
// the caller put them); that touches up to 3 bytes before s and 7 after s+16, inside the buffer's alignment / slack.
__device__ __forceinline__ uint32_t squeeze16_dev(const uint8_t *__restrict__ s, int lim, uint32_t *n_bits)
{
	uint32_t word = 0, nb = 0;
	if (lim >= 16) {
		const uint32_t *w = reinterpret_cast<const uint32_t *>(reinterpret_cast<uintptr_t>(s) & ~(uintptr_t)3);
		const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(s) & 3) * 8u;
		uint32_t x0 = w[0], x1 = w[1], x2 = w[2], x3 = w[3], x4 = sh ? w[4] : 0u;
		uint32_t v[4] = {__funnelshift_r(x0, x1, sh), __funnelshift_r(x1, x2, sh), __funnelshift_r(x2, x3, sh), __funnelshift_r(x3, x4, sh)};
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			// bytes > 3 are N: flagged and stored as 0
			const uint32_t hi = v[q] & 0xfcfcfcfcu;
			if (hi) {
#pragma unroll
				for (int b = 0; b < 4; ++b)
					if ((hi >> (8 * b)) & 0xffu) { nb |= 1u << (4 * q + b); v[q] &= ~(0xffu << (8 * b)); }
			}
			const uint32_t x = v[q];                                                   // four codes 0..3, one per byte
			word |= ((x & 3u) | ((x >> 6) & 0xcu) | ((x >> 12) & 0x30u) | ((x >> 18) & 0xc0u)) << (8 * q);
		}
	} else {
		for (int x = 0; x < lim; ++x) {
			const uint32_t c = s[x];
			if (c > 3u) nb |= 1u << x;
			else word |= c << (2 * x);
		}
	}
	*n_bits = nb;
	return word;
}

#define KSW_PACK_THREADS 512
#define KSW_PACK_LANES 16            /* lanes per job: a 101 x 101 bp job is 7 + 7 words (+ 2 of padding) */

__global__ void __launch_bounds__(KSW_PACK_THREADS)
ksw_pack_kernel(const ksw_b200_job_t *__restrict__ raw, long long n, const uint8_t *__restrict__ qraw, const uint8_t *__restrict__ traw,
                const uint32_t *__restrict__ offs, DevJob *__restrict__ jobs, uint32_t *__restrict__ pool,
                uint32_t *__restrict__ npool, DevPackStats *__restrict__ stats)
{
	const int sub = threadIdx.x & (KSW_PACK_LANES - 1);                                 // lane inside the job's half-warp
	const unsigned grp = 0xffffu << (threadIdx.x & 16);                                 // the half-warp's lanes
	const long long k = ((long long)blockIdx.x * KSW_PACK_THREADS + threadIdx.x) / KSW_PACK_LANES;
	if (k >= n) return;                                                                 // whole half-warps leave together
	const ksw_b200_job_t j = raw[k];
	if (j.qlen < 1 || j.tlen < 0) return;                                              // reported by the prep kernel
	const uint32_t qw = ksw_words2(j.qlen), tw = ksw_words2(j.tlen), total = ksw_job_units(j.qlen, j.tlen) * 4u;
	const uint32_t off = offs[k];
	uint32_t *out = pool + (size_t)off * 4;
	const uint8_t *q = qraw + j.q_off, *t = traw + j.t_off;
	bool qn = false, tn = false;
	for (uint32_t wi = sub; wi < total; wi += KSW_PACK_LANES) {
		uint32_t word = 0, nb = 0;
		if (wi < qw) { word = squeeze16_dev(q + 16 * wi, j.qlen - 16 * (int)wi, &nb); qn |= nb != 0; }
		else if (wi < qw + tw) { word = squeeze16_dev(t + 16 * (wi - qw), j.tlen - 16 * (int)(wi - qw), &nb); tn |= nb != 0; }
		out[wi] = word;
	}
	const unsigned bq = __ballot_sync(grp, qn), bt = __ballot_sync(grp, tn);
	qn = (bq & grp) != 0; tn = (bt & grp) != 0;
	if (!qn && !tn) {
		if (sub == 0) jobs[k].seq_off = off;
		return;
	}
	// rare: the job holds an N.  Masks (1 bit per base) go to the side pool: [query mask words][target mask words]
	const uint32_t qmw = ksw_words1(j.qlen), tmw = ksw_words1(j.tlen);
	uint32_t base = 0;
	if (sub == 0) base = atomicAdd(&stats->nmask_used, (qn ? qmw : 0u) + (tn ? tmw : 0u));
	base = __shfl_sync(grp, base, threadIdx.x & 16);
	uint32_t at = base;
	if (qn) {
		for (uint32_t mi = sub; mi < qmw; mi += KSW_PACK_LANES) {
			uint32_t m = 0;
			const int lim = j.qlen - 32 * (int)mi < 32 ? j.qlen - 32 * (int)mi : 32;
			for (int x = 0; x < lim; ++x) if (q[32 * mi + x] > 3) m |= 1u << x;
			npool[at + mi] = m;
		}
		at += qmw;
	}
	if (tn) {
		for (uint32_t mi = sub; mi < tmw; mi += KSW_PACK_LANES) {
			uint32_t m = 0;
			const int lim = j.tlen - 32 * (int)mi < 32 ? j.tlen - 32 * (int)mi : 32;
			for (int x = 0; x < lim; ++x) if (t[32 * mi + x] > 3) m |= 1u << x;
			npool[at + mi] = m;
		}
	}
	if (sub == 0) {
		DevJob d = jobs[k];
		d.seq_off = off;
		d.nmask_off = base;
		d.flags |= (qn ? KSW_FLAG_QN : 0u) | (tn ? KSW_FLAG_TN : 0u);
		if (((d.flags >> KSW_CLASS_SHIFT) & KSW_CLASS_MASK) == 0u)                      // class 0 is N-free (ksw_pack.cpp)
			d.flags = (d.flags & ~(KSW_CLASS_MASK << KSW_CLASS_SHIFT)) | (1u << KSW_CLASS_SHIFT);
		jobs[k] = d;
	}
}

// ---- jobs against a reference kept on the device (ksw_b200_extend_batch_ref) -------------------
// The target of a job is a run of the doubled reference coordinate space [0, 2 l_pac) read upwards or downwards from
// t_pos; base x of that space is what bns_get_seq produces (bwa-0.7.8/bntseq.c:355-376): the 2-bit .pac base for
// x < l_pac (_get_pac, bntseq.c:192), and 3 - pac(2 l_pac - 1 - x) for the reverse strand.  The query is a run of the
// byte-coded read pool, upwards or downwards from q_off.  No window is materialised on the host.
__device__ __forceinline__ uint32_t ref_base(const uint8_t *__restrict__ pac, long long l_pac, long long x)
{
	const bool rev = x >= l_pac;
	const long long k = rev ? (l_pac << 1) - 1 - x : x;
	const uint32_t b = (pac[k >> 2] >> ((~k & 3) << 1)) & 3u;
	return rev ? 3u - b : b;
}

__global__ void __launch_bounds__(256)
ksw_prep_ref_kernel(const ksw_b200_rjob_t *__restrict__ raw, long long n, const KswScoring S, long long l_pac, unsigned long long qbytes,
                    DevJob *__restrict__ jobs, uint32_t *__restrict__ offs, DevPackStats *__restrict__ stats)
{
	const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const int lane = threadIdx.x & 31;
	uint32_t my_units = 0, cls = 0xffu;
	unsigned long long my_nmask = 0;
	int my_qlen = 0;
	bool bad = false;
	if (k < n) {
		const ksw_b200_rjob_t j = raw[k];
		DevJob d;
		d.seq_off = 0; d.idx = (uint32_t)k; d.qlen = j.qlen; d.tlen = j.tlen;
		d.h0 = j.h0 < 0 ? 0 : j.h0;                                                    // ksw.c:384
		d.nmask_off = 0;
		bool ok = j.qlen >= 1 && j.tlen >= 0 && (j.q_step == 1 || j.q_step == -1) && (j.t_step == 1 || j.t_step == -1);
		if (ok) {
			// both ends of both runs must exist, and the target must stay on one strand (bwamem.c:752-755)
			const long long q_last = (long long)j.q_off + (long long)j.q_step * (j.qlen - 1);
			const long long t_last = j.t_pos + (long long)j.t_step * (j.tlen - 1);
			ok = q_last >= 0 && (unsigned long long)q_last < qbytes && j.q_off < qbytes;
			if (j.tlen > 0) ok = ok && j.t_pos >= 0 && t_last >= 0 && j.t_pos < (l_pac << 1) && t_last < (l_pac << 1) &&
			                     ((j.t_pos < l_pac) == (t_last < l_pac));
		}
		if (!ok) {
			bad = true;
			d.qlen = 1; d.tlen = 0; d.w = 0; d.flags = KSW_CLASS_THREAD << KSW_CLASS_SHIFT;
		} else {
			d.w = ksw_clamp_w_expr(j.qlen, S.maxsc, S.o_del, S.e_del, S.o_ins, S.e_ins, j.w, S.end_bonus);
			cls = ksw_job_class(S, j.qlen, d.h0);
			d.flags = cls << KSW_CLASS_SHIFT;
			my_units = ksw_job_units(j.qlen, j.tlen);
			my_nmask = ksw_words1(j.qlen);                                             // the reference holds no N (bntseq.c:228)
			my_qlen = j.qlen;
		}
		jobs[k] = d;
	}
	uint32_t incl = my_units;
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= o) incl += v;
	}
	const uint32_t warp_units = __shfl_sync(0xffffffffu, incl, 31);
	unsigned long long base = 0;
	if (lane == 0 && warp_units) base = atomicAdd(&stats->units, (unsigned long long)warp_units);
	base = __shfl_sync(0xffffffffu, base, 0);
	if (k < n) offs[k] = (uint32_t)(base + incl - my_units);
	for (int o = 16; o > 0; o >>= 1) my_nmask += __shfl_down_sync(0xffffffffu, my_nmask, o);
	if (lane == 0 && my_nmask) atomicAdd(&stats->nmask_words, my_nmask);
	if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&stats->bad, 1u);
	for (uint32_t c = 0; c < KSW_N_CLASSES; ++c) {
		const unsigned m = __ballot_sync(0xffffffffu, cls == c);
		if (!m) continue;
		int q = cls == c ? my_qlen : 0;
		for (int o = 16; o > 0; o >>= 1) { const int v = __shfl_down_sync(0xffffffffu, q, o); q = q > v ? q : v; }
		if (lane == 0) { atomicAdd(&stats->class_n[c], (unsigned)__popc(m)); atomicMax(&stats->class_qmax[c], q); }
	}
}

__global__ void __launch_bounds__(KSW_PACK_THREADS)
ksw_pack_ref_kernel(const ksw_b200_rjob_t *__restrict__ raw, long long n, const uint8_t *__restrict__ qraw, const uint8_t *__restrict__ pac,
                    long long l_pac, const uint32_t *__restrict__ offs, DevJob *__restrict__ jobs, uint32_t *__restrict__ pool,
                    uint32_t *__restrict__ npool, DevPackStats *__restrict__ stats)
{
	const int sub = threadIdx.x & (KSW_PACK_LANES - 1);
	const unsigned grp = 0xffffu << (threadIdx.x & 16);
	const long long k = ((long long)blockIdx.x * KSW_PACK_THREADS + threadIdx.x) / KSW_PACK_LANES;
	if (k >= n) return;
	const ksw_b200_rjob_t j = raw[k];
	if (j.qlen < 1 || j.tlen < 0) return;                                              // (a batch with such a job is rejected after prep)
	const uint32_t qw = ksw_words2(j.qlen), tw = ksw_words2(j.tlen), total = ksw_job_units(j.qlen, j.tlen) * 4u;
	const uint32_t off = offs[k];
	uint32_t *out = pool + (size_t)off * 4;
	bool qn = false;
	for (uint32_t wi = sub; wi < total; wi += KSW_PACK_LANES) {
		uint32_t word = 0;
		if (wi < qw) {
			const int lim = j.qlen - 16 * (int)wi < 16 ? j.qlen - 16 * (int)wi : 16;
			const long long q0 = (long long)j.q_off + (long long)j.q_step * 16 * (long long)wi;
			for (int x = 0; x < lim; ++x) {
				const uint32_t c = qraw[q0 + (long long)j.q_step * x];
				if (c > 3u) qn = true; else word |= c << (2 * x);
			}
		} else if (wi < qw + tw) {
			const uint32_t ti = wi - qw;
			const int lim = j.tlen - 16 * (int)ti < 16 ? j.tlen - 16 * (int)ti : 16;
			const long long t0 = j.t_pos + (long long)j.t_step * 16 * (long long)ti;
			for (int x = 0; x < lim; ++x) word |= ref_base(pac, l_pac, t0 + (long long)j.t_step * x) << (2 * x);
		}
		out[wi] = word;
	}
	qn = (__ballot_sync(grp, qn) & grp) != 0;
	if (!qn) {
		if (sub == 0) jobs[k].seq_off = off;
		return;
	}
	const uint32_t qmw = ksw_words1(j.qlen);
	uint32_t base = 0;
	if (sub == 0) base = atomicAdd(&stats->nmask_used, qmw);
	base = __shfl_sync(grp, base, threadIdx.x & 16);
	for (uint32_t mi = sub; mi < qmw; mi += KSW_PACK_LANES) {
		uint32_t m = 0;
		const int lim = j.qlen - 32 * (int)mi < 32 ? j.qlen - 32 * (int)mi : 32;
		for (int x = 0; x < lim; ++x) if (qraw[(long long)j.q_off + (long long)j.q_step * (32 * (long long)mi + x)] > 3) m |= 1u << x;
		npool[base + mi] = m;
	}
	if (sub == 0) {
		DevJob d = jobs[k];
		d.seq_off = off;
		d.nmask_off = base;
		d.flags |= KSW_FLAG_QN;
		if (((d.flags >> KSW_CLASS_SHIFT) & KSW_CLASS_MASK) == 0u)
			d.flags = (d.flags & ~(KSW_CLASS_MASK << KSW_CLASS_SHIFT)) | (1u << KSW_CLASS_SHIFT);
		jobs[k] = d;
	}
}

// sorted_keys: the binning keys in ascending order (ksw_bin.cu): class c occupies [range[c], range[c+1])
__global__ void ksw_range_kernel(const uint16_t *__restrict__ sorted_keys, long long n, uint32_t *__restrict__ range)
{
	const int c = threadIdx.x;                                                          // 0 .. KSW_N_CLASSES
	if (c > KSW_N_CLASSES) return;
	if (c == KSW_N_CLASSES) { range[c] = (uint32_t)n; return; }
	const uint32_t lowest = c >= (int)KSW_CLASS_GENERIC ? (c == (int)KSW_CLASS_WARP ? 0x8000u : 0xC000u) : ((uint32_t)c << 13);  // smallest key of class c
	long long lo = 0, hi = n;
	while (lo < hi) {
		const long long mid = (lo + hi) >> 1;
		if ((uint32_t)sorted_keys[mid] < lowest) lo = mid + 1; else hi = mid;
	}
	range[c] = (uint32_t)lo;
}

} // namespace

cudaError_t ksw_launch_prep(const void *raw_jobs, int64_t n, const KswScoring &S, DevJob *jobs, uint32_t *offs,
                            DevPackStats *stats, cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(DevPackStats), st);
	if (e != cudaSuccess) return e;
	ksw_prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const ksw_b200_job_t *)raw_jobs, (long long)n, S, jobs, offs, stats);
	return cudaGetLastError();
}

cudaError_t ksw_launch_pack(const void *raw_jobs, int64_t n, const uint8_t *qraw, const uint8_t *traw, const uint32_t *offs,
                            DevJob *jobs, uint32_t *pool, uint32_t *npool, DevPackStats *stats, cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	const long long per_block = KSW_PACK_THREADS / KSW_PACK_LANES;
	ksw_pack_kernel<<<(unsigned)((n + per_block - 1) / per_block), KSW_PACK_THREADS, 0, st>>>(
	    (const ksw_b200_job_t *)raw_jobs, (long long)n, qraw, traw, offs, jobs, pool, npool, stats);
	return cudaGetLastError();
}

cudaError_t ksw_launch_ranges(const uint16_t *sorted_keys, int64_t n, uint32_t *range, cudaStream_t st)
{
	ksw_range_kernel<<<1, 32, 0, st>>>(sorted_keys, (long long)n, range);
	return cudaGetLastError();
}

cudaError_t ksw_launch_prep_ref(const void *raw_jobs, int64_t n, const KswScoring &S, int64_t l_pac, uint64_t qbytes, DevJob *jobs,
                                uint32_t *offs, DevPackStats *stats, cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(DevPackStats), st);
	if (e != cudaSuccess) return e;
	ksw_prep_ref_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const ksw_b200_rjob_t *)raw_jobs, (long long)n, S, (long long)l_pac,
	                                                               (unsigned long long)qbytes, jobs, offs, stats);
	return cudaGetLastError();
}

cudaError_t ksw_launch_pack_ref(const void *raw_jobs, int64_t n, const uint8_t *qraw, const uint8_t *pac, int64_t l_pac, const uint32_t *offs,
                                DevJob *jobs, uint32_t *pool, uint32_t *npool, DevPackStats *stats, cudaStream_t st)
{
	if (n <= 0) return cudaSuccess;
	const long long per_block = KSW_PACK_THREADS / KSW_PACK_LANES;
	ksw_pack_ref_kernel<<<(unsigned)((n + per_block - 1) / per_block), KSW_PACK_THREADS, 0, st>>>(
	    (const ksw_b200_rjob_t *)raw_jobs, (long long)n, qraw, pac, (long long)l_pac, offs, jobs, pool, npool, stats);
	return cudaGetLastError();
}
