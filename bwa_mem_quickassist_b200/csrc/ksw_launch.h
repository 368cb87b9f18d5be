// ksw_launch.h — host-callable launchers of the kernels (defined in the .cu files).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "ksw_dev.cuh"
#include "ksw_class.h"

#define KSW_GENERIC_THREADS 128

// generic int32 kernel: eh/qc are scratch slabs of (qmax+1) * n_blocks*KSW_GENERIC_THREADS entries
cudaError_t ksw_launch_generic(const DevJob *jobs, int64_t n_jobs, const uint32_t *pool, const uint32_t *npool,
                               const KswParams &P,
                               int2 *eh, uint8_t *qc, int n_blocks, const uint32_t *order, DevRes *res, uint32_t *cells, cudaStream_t st);

// banded global alignment + backtrace (ksw_global.cu), one job per thread: eh/qc as for the generic kernel, z = direction
// matrix slab of zcap cells per thread; cigar_pool must hold the sum of (qlen + tlen) operations in the worst case
cudaError_t ksw_launch_global(const DevGJob *jobs, int64_t n_jobs, const uint8_t *seq, const KswParams &P, int2 *eh,
                              uint8_t *qc, uint8_t *z, long long zcap, int n_blocks, unsigned long long *pool_used,
                              uint32_t *cigar_pool, DevGRes *res, cudaStream_t st);

// warp-cooperative int32 kernel (ksw_warp.cu) over the jobs jobs[order[0..n_jobs)]: one job per warp, qlen <= qmax
cudaError_t ksw_launch_warp(const DevJob *jobs, int64_t n_jobs, const uint32_t *pool, const uint32_t *npool, const KswParams &P,
                            int qmax, int sm_count, unsigned long long *counter, const uint32_t *order, DevRes *res,
                            uint32_t *cells, cudaStream_t st);
size_t ksw_warp_smem_bytes(int qmax);

// DPX issue-rate probe; each thread issues iters*32 DPX instructions
cudaError_t ksw_launch_dpx_peak(int which, unsigned *out, int n_blocks, int iters, cudaStream_t st);

#define KSW_FAST_THREADS 32

// fast s16x2 kernel over the jobs jobs[order[0..n_jobs)] whose qlen <= qmax; keyed: every job satisfies the class-0 bounds of
// ksw_class.h; counter: one device uint64 scratch word.  drange != nullptr (device-packed batches): the launch covers
// order[drange[c_lo] .. drange[c_hi]) instead, n_jobs is only the host's upper bound of that count.  ctas_per_sm_cap > 0:
// at most that many CTAs per SM; < 0: that many fewer than fit, if at least six remain (the pinned-caller pipeline leaves
// room for the next chunks' packing / binning kernels)
cudaError_t ksw_launch_fast(const DevJob *jobs, int64_t n_jobs, const uint32_t *pool, const uint32_t *npool,
                            const KswParams &P, int qmax, bool keyed, int sm_count, unsigned long long *counter, const uint32_t *order,
                            DevRes *res, uint32_t *cells, cudaStream_t st,
                            const uint32_t *drange = nullptr, int c_lo = 0, int c_hi = 0, int ctas_per_sm_cap = 0);
size_t ksw_fast_smem_bytes(int qmax);

// pair kernel (two jobs per lane, ksw_pair.cu) over the class-0 jobs jobs[order[0..n_jobs)]: qlen <= qmax <= 124, biased
// scores < 512, no N in the query
cudaError_t ksw_launch_pair(const DevJob *jobs, int64_t n_jobs, const uint32_t *pool, const uint32_t *npool,
                            const KswParams &P, int qmax, int sm_count, unsigned long long *counter, const uint32_t *order,
                            DevRes *res, uint32_t *cells, cudaStream_t st);
size_t ksw_pair_smem_bytes(int qmax, int n_warps);

// device-side binning (ksw_bin.cu): fills order[0..n) with the job indices sorted by the bin key
size_t ksw_bin_temp_bytes(int64_t n);
cudaError_t ksw_launch_bin(const DevJob *jobs, int64_t n, uint16_t *keys_in, uint16_t *keys_out, uint32_t *vals_in,
                           uint32_t *order, void *temp, size_t temp_bytes, cudaStream_t st);

// the same order (up to the order of equal keys) by a counting sort that needs no shared memory, plus range[c] = first entry
// of kernel class c (c = 0..KSW_N_CLASSES; the last one is n); work: ksw_bin_counting_bytes() bytes of device scratch
size_t ksw_bin_counting_bytes(void);
cudaError_t ksw_launch_bin_counting(const DevJob *jobs, int64_t n, uint16_t *keys, void *work, uint32_t *order, uint32_t *range,
                                    cudaStream_t st);

// device-side packing (ksw_devpack.cu): raw ksw_b200_job_t records + raw byte-coded sequences in HBM -> DevJob[] + 2-bit pool
// prep: DevJob records (seq_off left 0), each job's offset in the 2-bit pool (offs), chunk totals (stats; zeroed first)
cudaError_t ksw_launch_prep(const void *raw_jobs, int64_t n, const KswScoring &S, DevJob *jobs, uint32_t *offs,
                            DevPackStats *stats, cudaStream_t st);
cudaError_t ksw_launch_pack(const void *raw_jobs, int64_t n, const uint8_t *qraw, const uint8_t *traw, const uint32_t *offs,
                            DevJob *jobs, uint32_t *pool, uint32_t *npool, DevPackStats *stats, cudaStream_t st);
// range[c] = first entry of kernel class c in the binned order (c = 0..KSW_N_CLASSES; the last one is n)
cudaError_t ksw_launch_ranges(const uint16_t *sorted_keys, int64_t n, uint32_t *range, cudaStream_t st);

// jobs against a reference kept on the device (ksw_b200_rjob_t records, 2-bit .pac in HBM): same products as prep / pack
cudaError_t ksw_launch_prep_ref(const void *raw_jobs, int64_t n, const KswScoring &S, int64_t l_pac, uint64_t qbytes, DevJob *jobs,
                                uint32_t *offs, DevPackStats *stats, cudaStream_t st);
cudaError_t ksw_launch_pack_ref(const void *raw_jobs, int64_t n, const uint8_t *qraw, const uint8_t *pac, int64_t l_pac, const uint32_t *offs,
                                DevJob *jobs, uint32_t *pool, uint32_t *npool, DevPackStats *stats, cudaStream_t st);

// fast banded global alignment (ksw_gfast.cu): s16x2 DP kernel (stores H, 2 bytes per band cell) + backtrack kernel that recomputes
// the reference's direction bits from H; jobs[gorder[..]] in groups of 32; z: slab of the groups; counter: one device word;
// scratch: (sum of qlen + tlen + 2 over ALL jobs of the array) words
size_t ksw_gfast_smem_bytes(int qmax);
cudaError_t ksw_launch_gfast(const DevGJob *jobs, const uint8_t *seq, const KswParams &P, const uint32_t *gorder, const DevGGroup *groups,
                             int n_groups, int qmax, int sm_count, uint2 *z, unsigned *counter, uint32_t *scratch,
                             unsigned long long *pool_used, uint32_t *cigar_pool, DevGRes *res, cudaStream_t st);

// local alignment, the reference's ksw_align2 (ksw_align.cu): bytes_per_score 1 = the byte kernel (16 lanes per job), 2 = the
// 16-bit kernel (8 lanes per job); jobs[order[0..n_jobs)] sorted by size; *bscr / *bscr_cap: scratch the launcher (re)allocates
#define KSW_ALIGN_MAX_QLEN 4096
cudaError_t ksw_launch_align(int bytes_per_score, const DevAJob *jobs, const uint32_t *order, int n_jobs, const uint8_t *seq,
                             const KswAlignParams &A, int qmax, int tmax, int sm_count, void **bscr, size_t *bscr_cap,
                             unsigned *counter, DevARes *res, cudaStream_t st);
