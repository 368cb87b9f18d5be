// ksw_runtime.cu — host side of the C ABI declared in include/ksw_b200.h:
// contexts, H2D/D2H staging on per-context streams, device-side binning and kernel dispatch, the chunked
// pack/copy/compute pipeline of the one-shot batched entry, and the scalar
// ksw_extend/ksw_extend2 wrappers.  (The packer itself is ksw_pack.cpp.)
//
// There is deliberately NO CPU implementation of the DP in this file: every result comes
// from a kernel launch.  If CUDA is unusable the calls fail loudly.
#include <cuda_runtime.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ksw_b200.h"
#include "ksw_dev.cuh"
#include "ksw_launch.h"
#include "ksw_class.h"
#include "ksw_pack.h"

// ------------------------------------------------------------------ small helpers
namespace {

// Growth of the grow-only buffers: a (re)allocation of device or page-locked memory synchronises the whole device, and
// the batches a shared queue merges come in every size between one submission and sixteen, so small buffers double
// (a handful of steps instead of dozens); large ones (a 10 M-job batch is gigabytes) get an eighth of slack.
inline size_t grow_to(size_t bytes) { return bytes + std::max<size_t>(bytes / 8, std::min<size_t>(bytes, (size_t)32 << 20)) + 4096; }

struct PinnedBuf {            // grow-only pinned host buffer
	void *p = nullptr; size_t cap = 0;
	cudaError_t reserve(size_t bytes) {
		if (bytes <= cap) return cudaSuccess;
		if (p) cudaFreeHost(p);
		p = nullptr; cap = 0;
		size_t want = grow_to(bytes);
		cudaError_t e = cudaMallocHost(&p, want);
		if (e == cudaSuccess) cap = want;
		return e;
	}
	void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct DevBuf {               // grow-only device buffer
	void *p = nullptr; size_t cap = 0;
	cudaError_t reserve(size_t bytes) {
		if (bytes <= cap) return cudaSuccess;
		if (p) cudaFree(p);
		p = nullptr; cap = 0;
		size_t want = grow_to(bytes);
		cudaError_t e = cudaMalloc(&p, want);
		if (e == cudaSuccess) cap = want;
		return e;
	}
	void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

double now_ms()
{
	return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // namespace

#define KSW_N_SLOTS 6
#define KSW_N_HSLOTS 8         /* staging sets the host lane may own; it uses ctx->n_hslots of them */

// ------------------------------------------------------------------ opaque types
struct ksw_b200_batch {       // a packed batch in HBM + what the launcher needs to know about it
	int64_t n = 0, n_fast = 0, n_warp = 0, n_generic = 0;   // n_generic: the thread-per-job kernel's jobs
	int64_t fast_class_n[KSW_FAST_CLASSES] = {0, 0, 0, 0};
	int fast_class_qmax[KSW_FAST_CLASSES] = {0, 0, 0, 0};
	KswParams P;
	DevBuf d_jobs, d_pool, d_npool, d_res, d_cells;
	DevBuf d_order, d_keys, d_vals, d_sort_tmp;                 // device-side binning (ksw_bin.cu)
	size_t pool_bytes = 0, npool_bytes = 0;
	int qmax_warp = 0, qmax_generic = 0;
	// device-packed batches (ksw_devpack.cu): the class sizes above are the host's view BEFORE class-0 jobs that hold an
	// N moved to class 1; the kernels read the true class bounds of the binned order from d_range
	bool dev_ranges = false;
	DevBuf d_range;
};

namespace {

// One pipeline stage set: pinned staging, a reusable plan, a device batch, a stream, kernel scratch.
struct Slot {
	cudaStream_t stream = nullptr;
	PinnedBuf h_jobs, h_pool, h_npool, h_res;
	KswPackStats stats;
	std::vector<uint32_t> nmask;
	ksw_b200_batch batch;
	DevBuf d_eh, d_qc, d_counter;          // scratch of the generic kernel / job counters of the fast kernel
	bool busy = false;                     // results of a chunk are in flight into h_res
	int64_t first = 0, n = 0;              // caller range of that chunk
	// device-side packing (pinned callers): raw job records, sizes / offsets of the 2-bit slices, chunk totals
	DevBuf d_rawjobs, d_offs, d_stats;
	PinnedBuf h_stats;
	cudaEvent_t ev_jobs = nullptr, ev_stats = nullptr, ev_up = nullptr, ev_packed = nullptr, ev_ext = nullptr, ev_done = nullptr;
	bool host_packed = false;              // pinned-caller pipeline: this chunk was packed on the host threads
};

// a reference (.pac) resident on a device, shared by the contexts of that device
struct RefEntry {
	int device = -1;
	const uint8_t *host = nullptr;
	int64_t l_pac = 0;
	void *dev = nullptr;
	int refs = 0;
};

// one chunk of ksw_b200_global_batch: page-locked staging, device buffers, completion event
struct GStage {
	PinnedBuf hjobs, hseq, hres, hcig, horder, hgroups, hused;
	DevBuf djobs, dseq, dres, dcig, dused, dorder, dgroups;
	cudaEvent_t e_start = nullptr, e_up = nullptr, e_kstart = nullptr, e_kern = nullptr, e_done = nullptr;
	int64_t first = 0, m = 0;
	bool busy = false;
	void release()
	{
		hjobs.release(); hseq.release(); hres.release(); hcig.release(); horder.release(); hgroups.release(); hused.release();
		djobs.release(); dseq.release(); dres.release(); dcig.release(); dused.release(); dorder.release(); dgroups.release();
		for (cudaEvent_t *e : {&e_start, &e_up, &e_kstart, &e_kern, &e_done}) if (*e) { cudaEventDestroy(*e); *e = nullptr; }
	}
};

struct AsyncReq {
	ksw_b200_cfg_t cfg;
	int64_t n = 0;
	const ksw_b200_job_t *jobs = nullptr;
	const uint8_t *qpool = nullptr, *tpool = nullptr;
	size_t qbytes = 0, tbytes = 0;
	ksw_b200_res_t *res = nullptr;
};

} // namespace

struct ksw_b200_ctx {
	int device = 0;
	int sm_count = 148;
	cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_mid = nullptr;
	std::string err;
	int pack_threads = 8;
	KswPool *pool = nullptr;               // (re)created lazily with pack_threads workers
	int64_t chunk_jobs = 1 << 20;
	// pinned-caller pipeline: shorter chunks, the two lanes share them one by one.  Config 2, 10 M jobs, 16 host threads, with
	// the 64-byte host packer: 2^18 -> 40.4 ms per call, 393216 -> 40.8, 2^19 -> 42.3, 786432 -> 44.6, 2^20 -> 47.1
	// (profiles/r2_e2e_knobs_sweep.txt): short chunks let the host lane take 17 of 40 instead of 7 of 21
	int64_t async_chunk_jobs = 1 << 18;
	int trace = 0;
	std::atomic<long long> launches{0};
	int64_t last_h2d = 0, last_d2h = 0;    // bytes moved by the last ksw_b200_extend_batch call
	std::mutex err_mu;                     // the two lanes of the pinned-caller pipeline may both report an error
	// Pipeline slots of the one-shot entry; slot[0] also serves upload / run / download of resident batches.  Three, not
	// two: a slot is reusable only after its chunk's results are back, and pack (3 ms) + H2D (2.5 ms) of the next chunk
	// for that slot take longer than the other slot's kernel (3.7 ms per 2^20 config-2 jobs), so with two slots the GPU
	// idled 2.4 ms per chunk.
	Slot slot[KSW_N_SLOTS];
	// pinned callers (ksw_b200_extend_batch_async): the raw byte-coded pools in HBM, the stream their uploads are
	// ordered on, and the worker thread that feeds the pipeline while the caller goes on
	DevBuf d_qraw, d_traw;
	cudaStream_t up_stream = nullptr, pre_stream = nullptr, main_stream = nullptr, ext2_stream = nullptr, host_stream = nullptr, down_stream = nullptr, hdown_stream = nullptr;
	int hybrid = 1;                        // a second lane packs chunks on the host threads (KSW_B200_HYBRID=0: device packing only)
	Slot hslot[KSW_N_HSLOTS];                         // that lane's staging / device buffers (events only; it uses the shared streams)
	int n_hslots = 3;                                 // how many of them it cycles through (KSW_B200_HSLOTS)
	// The extension kernel's CTAs per SM in this pipeline: negative = that many fewer than fit.  The launches of the chunks
	// run while the next chunks are packed and binned; beside 13 extension CTAs an SM has room for ONE more CTA (1 KB of
	// shared memory, 22 k registers), the packing / binning chain of a chunk then takes as long as the chunk's extension
	// kernel, ends when that kernel's CTAs start to leave, and the next launch is never ready to fill the half-empty last
	// wave.  Two CTAs fewer: 40.2 -> 38.2 ms per 10 M config-2 jobs (13: 40.2, 12: 39.6, 11: 38.2, 10: 38.6, 9: 39.0,
	// 8: 39.8; profiles/r2_e2e_knobs_sweep.txt).  KSW_B200_ASYNC_CTAS=<n> sets an absolute number, 0 = as many as fit.
	int async_fast_ctas = -2;
	cudaStream_t hup_stream = nullptr;                // the host lane's own upload stream (KSW_B200_HUP=1; default: up_stream)
	int hup = 0;
	std::vector<int64_t> lead_jobs;                   // sizes of the call's first chunks (KSW_B200_LEAD=a,b,..; default chunk/4, chunk/2)
	double host_ms_per_job = 0, dev_ms_per_job = 0;   // measured pace of the two lanes (0 = not measured yet)
	RefEntry *ref = nullptr;               // ksw_b200_ref_set
	PinnedBuf h_rq;                        // ksw_b200_extend_batch_ref: staging of a pageable read pool
	std::thread worker;
	std::mutex mu;
	std::condition_variable cv;
	bool worker_started = false, req_pending = false, req_running = false, stop = false;
	int async_rc = 0;
	AsyncReq req;
	// banded global alignment (ksw_b200_global_batch): staging, device buffers, the CIGAR pool handed to the caller
	GStage g_st[2];                        // two chunks in flight: the host prepares one while the GPU works on the other
	DevBuf g_deh, g_dqc, g_dz, g_dzfast, g_dcounter, g_dscratch;      // kernel scratch: used in stream order, one copy
	std::vector<uint32_t> g_cigar, g_key, g_tmp, g_bucket;
	// local alignment (ksw_b200_align_batch)
	PinnedBuf a_hjobs, a_hseq, a_hres, a_horder;
	DevBuf a_djobs, a_dseq, a_dres, a_dorder, a_dcounter;
	void *a_bscr = nullptr;
	size_t a_bscr_cap = 0;
};

namespace {

int fail(ksw_b200_ctx *ctx, int code, const std::string &msg)
{
	if (ctx) {
		std::lock_guard<std::mutex> lk(ctx->err_mu);
		ctx->err = msg;
	}
	return code;
}
#define CU(call)                                                                                   \
	do {                                                                                           \
		cudaError_t e__ = (call);                                                                  \
		if (e__ != cudaSuccess)                                                                    \
			return fail(ctx, 100 + (int)e__, std::string(#call) + ": " + cudaGetErrorString(e__)); \
	} while (0)

KswPool *pool_of(ksw_b200_ctx *ctx)
{
	if (ctx->pool && ctx->pool->size() != ctx->pack_threads) { delete ctx->pool; ctx->pool = nullptr; }
	if (!ctx->pool) ctx->pool = new KswPool(ctx->pack_threads);
	return ctx->pool;
}

std::mutex g_ref_mu;
std::vector<RefEntry *> g_refs;

void ref_release(ksw_b200_ctx *ctx)
{
	if (!ctx->ref) return;
	std::lock_guard<std::mutex> lk(g_ref_mu);
	if (--ctx->ref->refs == 0) {
		cudaFree(ctx->ref->dev);
		g_refs.erase(std::remove(g_refs.begin(), g_refs.end(), ctx->ref), g_refs.end());
		delete ctx->ref;
	}
	ctx->ref = nullptr;
}

int fast_qmax_enabled()
{
	static const int v = [] {
		const char *s = getenv("KSW_B200_DISABLE_FAST");
		return (s && *s && *s != '0') ? 0 : KSW_FAST_CLASS_QMAX[KSW_FAST_CLASSES - 1];
	}();
	return v;
}

// KSW_B200_PAIR=1 sends class 0 to the pair kernel (two jobs per lane, ksw_pair.cu) instead of the one-job-per-lane
// kernel.  Off by default: on config 2 the pair kernel needs 40 % fewer ALU-pipe instructions per cell but, with the
// same shared memory per job, runs half as many warps per SM and ends up latency-bound (1.72 vs 2.03 TCUPS, DESIGN.md
// §5.2).  Read per launch: the tests and the A/B bench flip it inside one process.
bool pair_enabled()
{
	const char *s = getenv("KSW_B200_PAIR");
	return s && *s && *s != '0';
}

void batch_release_buffers(ksw_b200_batch *b)
{
	b->d_jobs.release(); b->d_pool.release(); b->d_npool.release(); b->d_res.release(); b->d_cells.release();
	b->d_order.release(); b->d_keys.release(); b->d_vals.release(); b->d_sort_tmp.release();
	b->d_range.release();
}

// pack jobs[0..n) into the slot's pinned staging, start the H2D copies into `b` on the slot's stream and enqueue the
// device-side binning behind them
int pack_and_upload(ksw_b200_ctx *ctx, Slot &s, const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_job_t *jobs,
                    const uint8_t *qpool, const uint8_t *tpool, ksw_b200_batch *b, double *t_plan, double *t_fill,
                    cudaStream_t copy_st = nullptr, cudaStream_t comp_st = nullptr)
{
	if (!copy_st) copy_st = s.stream;
	if (!comp_st) comp_st = s.stream;
	std::string err;
	const double t0 = now_ms();
	KswPool *tp = pool_of(ctx);
	int rc = ksw_pack_sizes(cfg, n, jobs, fast_qmax_enabled(), tp, s.stats, err);
	if (rc) return fail(ctx, rc, err);
	KswPackStats &st = s.stats;
	const double t1 = now_ms();
	CU(s.h_jobs.reserve(sizeof(DevJob) * (size_t)std::max<int64_t>(n, 1)));
	CU(s.h_pool.reserve(std::max<size_t>(st.pool_bytes, 16)));
	rc = ksw_pack_stream(st, cfg, jobs, fast_qmax_enabled(), qpool, tpool, (DevJob *)s.h_jobs.p, (uint32_t *)s.h_pool.p, s.nmask, tp);
	if (rc) return fail(ctx, rc, "ksw_b200: packing failed");
	const size_t npool_bytes = s.nmask.size() * 4;
	CU(s.h_npool.reserve(std::max<size_t>(npool_bytes, 16)));
	if (npool_bytes) memcpy(s.h_npool.p, s.nmask.data(), npool_bytes);
	const double t2 = now_ms();
	if (t_plan) *t_plan += t1 - t0;
	if (t_fill) *t_fill += t2 - t1;

	b->n = n; b->n_fast = 0;
	for (int c = 0; c < KSW_FAST_CLASSES; ++c) {
		b->fast_class_n[c] = st.class_n[c];
		b->fast_class_qmax[c] = st.class_qmax[c];
		b->n_fast += st.class_n[c];
	}
	b->n_warp = st.class_n[KSW_CLASS_WARP]; b->qmax_warp = st.class_qmax[KSW_CLASS_WARP];
	b->n_generic = st.class_n[KSW_CLASS_THREAD];
	b->qmax_generic = st.class_qmax[KSW_CLASS_THREAD];
	b->pool_bytes = st.pool_bytes; b->npool_bytes = npool_bytes;
	b->dev_ranges = false;                                       // packed on the host: the class sizes are exact
	ksw_params_from_cfg(cfg, b->P);
	const size_t n1 = (size_t)std::max<int64_t>(n, 1);
	const size_t tmp_bytes = ksw_bin_temp_bytes(n);
	CU(b->d_jobs.reserve(sizeof(DevJob) * n1));
	CU(b->d_pool.reserve(std::max<size_t>(st.pool_bytes, 16)));
	CU(b->d_npool.reserve(std::max<size_t>(npool_bytes, 16)));
	CU(b->d_res.reserve(sizeof(DevRes) * n1));
	CU(b->d_cells.reserve(sizeof(uint32_t) * n1));
	CU(b->d_order.reserve(sizeof(uint32_t) * n1));
	CU(b->d_keys.reserve(sizeof(uint16_t) * 2 * n1));
	CU(b->d_vals.reserve(sizeof(uint32_t) * n1));
	CU(b->d_sort_tmp.reserve(std::max<size_t>(tmp_bytes, 16)));
	if (n > 0) {
		CU(cudaMemcpyAsync(b->d_jobs.p, s.h_jobs.p, sizeof(DevJob) * (size_t)n, cudaMemcpyHostToDevice, copy_st));
		if (st.pool_bytes)
			CU(cudaMemcpyAsync(b->d_pool.p, s.h_pool.p, st.pool_bytes, cudaMemcpyHostToDevice, copy_st));
		if (npool_bytes)
			CU(cudaMemcpyAsync(b->d_npool.p, s.h_npool.p, npool_bytes, cudaMemcpyHostToDevice, copy_st));
		if (copy_st != comp_st) {
			CU(cudaEventRecord(s.ev_up, copy_st));
			CU(cudaStreamWaitEvent(comp_st, s.ev_up, 0));
		}
		CU(ksw_launch_bin((const DevJob *)b->d_jobs.p, n, (uint16_t *)b->d_keys.p, (uint16_t *)b->d_keys.p + n1,
		                  (uint32_t *)b->d_vals.p, (uint32_t *)b->d_order.p, b->d_sort_tmp.p, b->d_sort_tmp.cap, comp_st));
		ctx->launches += 2;                                      // key kernel + the radix sort (counted as one more)
	}
	return 0;
}

int ensure_generic_scratch(ksw_b200_ctx *ctx, Slot &s, int qmax, int &n_blocks)
{
	// one column slab per resident thread; bound the slab to ~1 GiB
	const size_t per_thread = (size_t)(qmax + 1) * (sizeof(int2) + 1);
	size_t threads = (size_t)ctx->sm_count * 8 * KSW_GENERIC_THREADS;
	const size_t budget = (size_t)1 << 30;
	while (threads > KSW_GENERIC_THREADS && threads * per_thread > budget) threads >>= 1;
	n_blocks = (int)(threads / KSW_GENERIC_THREADS);
	CU(s.d_eh.reserve(threads * (size_t)(qmax + 1) * sizeof(int2)));
	CU(s.d_qc.reserve(threads * (size_t)(qmax + 1)));
	return 0;
}

// launches the kernels of batch b on slot s's stream: the fast classes (contiguous in the binned order) are grouped
// into as few launches as is free — a class is folded into the next one when it is small or when the next class's
// actual longest query needs (almost) the same shared memory — then the generic kernel.
// Device-packed batches (b->dev_ranges): the class sizes are upper bounds (a class-0 job that holds an N has moved to
// class 1 on the device), so class 1 counts as populated whenever class 0 is, and every launch takes its bounds from
// b->d_range.
int enqueue_kernels(ksw_b200_ctx *ctx, Slot &s, ksw_b200_batch *b, cudaStream_t st = nullptr, int fast_ctas_cap = 0)
{
	if (!st) st = s.stream;
	int64_t cn[KSW_FAST_CLASSES];
	int cq[KSW_FAST_CLASSES];
	for (int x = 0; x < KSW_FAST_CLASSES; ++x) { cn[x] = b->fast_class_n[x]; cq[x] = b->fast_class_qmax[x]; }
	if (b->dev_ranges && cn[0] > 0) { cn[1] += cn[0]; cq[1] = std::max(cq[1], cq[0]); }
	const uint32_t *drange = b->dev_ranges ? (const uint32_t *)b->d_range.p : nullptr;
	int64_t first = 0;
	int c = 0;
	CU(s.d_counter.reserve(sizeof(unsigned long long) * (KSW_FAST_CLASSES + 2)));
	if (cn[0] > 0 && pair_enabled() && !b->dev_ranges) {
		// class 0 goes to the pair kernel (two jobs per lane), unless it is a small minority next to other fast classes:
		// then one launch of the one-job-per-lane kernel over all of them fills the GPU better than two thin launches
		bool others = false;
		for (int x = 1; x < KSW_FAST_CLASSES; ++x) others |= cn[x] > 0;
		if (!others || cn[0] >= (int64_t)ctx->sm_count * 6 * 64 * 2) {
			CU(ksw_launch_pair((const DevJob *)b->d_jobs.p, cn[0], (const uint32_t *)b->d_pool.p,
			                   (const uint32_t *)b->d_npool.p, b->P, cq[0], ctx->sm_count,
			                   (unsigned long long *)s.d_counter.p + KSW_FAST_CLASSES, (const uint32_t *)b->d_order.p,
			                   (DevRes *)b->d_res.p, (uint32_t *)b->d_cells.p, st));
			ctx->launches++;
			first = cn[0];
			c = 1;
		}
	}
	const int64_t thin = (int64_t)ctx->sm_count * 13 * 32 * 4;
	while (c < KSW_FAST_CLASSES) {
		if (cn[c] <= 0) { ++c; continue; }
		int64_t n_grp = cn[c];
		int qmax = cq[c];
		bool keyed = c == 0;
		int e = c + 1;
		while (e < KSW_FAST_CLASSES) {
			if (cn[e] <= 0) { ++e; continue; }
			// the folding test looks at the host's own counts: class 1's upper bound of a device-packed batch would
			// never look small
			const int64_t ne = b->fast_class_n[e];
			const bool small = n_grp < thin || ne < thin;
			const bool same_smem = KSW_FAST_QUADS(cq[e]) <= KSW_FAST_QUADS(qmax) + KSW_FAST_QUADS(qmax) / 8;
			// never give up the keyed variant of a big class 0 for a small neighbour; fold class 0 only if it is small itself
			if (keyed && n_grp >= thin) break;
			if (!(small || same_smem)) break;
			n_grp += cn[e];
			qmax = std::max(qmax, cq[e]);
			keyed = false;
			++e;
		}
		while (e < KSW_FAST_CLASSES && cn[e] <= 0) ++e;          // empty classes in between belong to the range too
		CU(ksw_launch_fast((const DevJob *)b->d_jobs.p, n_grp, (const uint32_t *)b->d_pool.p,
		                   (const uint32_t *)b->d_npool.p, b->P, qmax, keyed, ctx->sm_count,
		                   (unsigned long long *)s.d_counter.p + c, (const uint32_t *)b->d_order.p + (drange ? 0 : first),
		                   (DevRes *)b->d_res.p, (uint32_t *)b->d_cells.p, st, drange, c, e, fast_ctas_cap));
		ctx->launches++;
		first += n_grp;
		c = e;
	}
	if (b->n_warp > 0) {
		CU(ksw_launch_warp((const DevJob *)b->d_jobs.p, b->n_warp, (const uint32_t *)b->d_pool.p, (const uint32_t *)b->d_npool.p,
		                   b->P, b->qmax_warp, ctx->sm_count, (unsigned long long *)s.d_counter.p + KSW_FAST_CLASSES + 1,
		                   (const uint32_t *)b->d_order.p + b->n_fast, (DevRes *)b->d_res.p, (uint32_t *)b->d_cells.p, st));
		ctx->launches++;
	}
	if (b->n_generic > 0) {
		int n_blocks = 0;
		int rc = ensure_generic_scratch(ctx, s, b->qmax_generic, n_blocks);
		if (rc) return rc;
		const int64_t need = (b->n_generic + KSW_GENERIC_THREADS - 1) / KSW_GENERIC_THREADS;
		if (need < n_blocks) n_blocks = (int)need;
		CU(ksw_launch_generic((const DevJob *)b->d_jobs.p, b->n_generic, (const uint32_t *)b->d_pool.p,
		                      (const uint32_t *)b->d_npool.p, b->P, (int2 *)s.d_eh.p, (uint8_t *)s.d_qc.p,
		                      n_blocks, (const uint32_t *)b->d_order.p + b->n_fast + b->n_warp, (DevRes *)b->d_res.p,
		                      (uint32_t *)b->d_cells.p, st));
		ctx->launches++;
	}
	return 0;
}

// waits for the chunk in flight on slot s and hands its results to the caller
int retire_slot(ksw_b200_ctx *ctx, Slot &s, ksw_b200_res_t *res, double *t_wait)
{
	if (!s.busy) return 0;
	const double t0 = now_ms();
	CU(cudaStreamSynchronize(s.stream));
	if (t_wait) *t_wait += now_ms() - t0;
	// pinned staging -> caller's array, split over the pack threads (a 100 MB memcpy is not free)
	{
		char *dst = (char *)(res + s.first);
		const char *src = (const char *)s.h_res.p;
		const size_t bytes = sizeof(DevRes) * (size_t)s.n;
		KswPool *tp = pool_of(ctx);
		const int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)tp->size(), bytes >> 20));
		const size_t per = (bytes + T - 1) / T;
		auto body = [&](int t) {
			const size_t b = std::min(bytes, (size_t)t * per), e = std::min(bytes, b + per);
			if (b < e) memcpy(dst + b, src + b, e - b);
		};
		if (T == 1) body(0); else tp->run(T, body);
	}
	s.busy = false;
	return 0;
}

} // namespace

// ------------------------------------------------------------------ C ABI
extern "C" {

int ksw_b200_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}

int ksw_b200_ctx_create(int device, ksw_b200_ctx_t **out)
{
	if (!out) return 1;
	*out = nullptr;
	ksw_b200_ctx *ctx = new ksw_b200_ctx();
	ctx->device = device;
	cudaError_t e = cudaSetDevice(device);
	// the events a feeder thread sleeps on (chunk totals back, a slot's results home).  Blocking-sync events make those
	// waits yield the core instead of spinning: the feeder threads of the pinned-caller pipeline wait most of a call, and
	// on a box with few cores per GPU a spinning feeder takes a core away from the host lane's pack threads
	// (measured with the process pinned to 4 cores: 43.8 -> 43.0 ms per 10 M config-2 jobs; with 16 cores 40.2 vs 40.5).
	// Default: on when the box has fewer than 8 hardware threads per GPU.
	unsigned wait_flags = cudaEventDisableTiming;
	{
		int n_dev = 1;
		if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev < 1) n_dev = 1;
		const unsigned hw0 = std::thread::hardware_concurrency();
		bool blocking = hw0 && hw0 / (unsigned)n_dev < 8u;
		if (const char *s = getenv("KSW_B200_BLOCKSYNC")) blocking = atoi(s) != 0;
		if (blocking) wait_flags |= cudaEventBlockingSync;
	}
	for (int i = 0; i < KSW_N_SLOTS && e == cudaSuccess; ++i) {
		Slot &s = ctx->slot[i];
		e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
		cudaEvent_t *evs[6] = {&s.ev_jobs, &s.ev_stats, &s.ev_up, &s.ev_packed, &s.ev_ext, &s.ev_done};
		for (cudaEvent_t *ev : evs) if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, (ev == &s.ev_stats || ev == &s.ev_done) ? wait_flags : cudaEventDisableTiming);
	}
	for (Slot &s : ctx->hslot) {
		cudaEvent_t *evs[3] = {&s.ev_up, &s.ev_ext, &s.ev_done};
		for (cudaEvent_t *ev : evs) if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, ev == &s.ev_done ? wait_flags : cudaEventDisableTiming);
	}
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->main_stream, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->host_stream, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->hdown_stream, cudaStreamNonBlocking);
	{
		// KSW_B200_PRE_PRIO=1 (A/B knob): the packing / binning stream at the highest priority, so that its CTAs are placed
		// before the next extension launch's when an SM frees up
		int lo = 0, hi = 0;
		const char *pp = getenv("KSW_B200_PRE_PRIO");
		if (e == cudaSuccess && pp && atoi(pp) && cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess)
			e = cudaStreamCreateWithPriority(&ctx->pre_stream, cudaStreamNonBlocking, hi);
		else if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->pre_stream, cudaStreamNonBlocking);
	}
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->ext2_stream, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->down_stream, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->hup_stream, cudaStreamNonBlocking);
	if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev0);
	if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev1);
	if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
	if (e != cudaSuccess) {
		fprintf(stderr, "[ksw_b200] cannot create context on device %d: %s\n", device, cudaGetErrorString(e));
		for (Slot &s : ctx->slot) {
			if (s.stream) cudaStreamDestroy(s.stream);
			for (cudaEvent_t ev : {s.ev_jobs, s.ev_stats, s.ev_up, s.ev_packed, s.ev_ext, s.ev_done}) if (ev) cudaEventDestroy(ev);
		}
		for (Slot &s : ctx->hslot) for (cudaEvent_t ev : {s.ev_up, s.ev_ext, s.ev_done}) if (ev) cudaEventDestroy(ev);
		for (cudaStream_t st : {ctx->up_stream, ctx->pre_stream, ctx->main_stream, ctx->ext2_stream, ctx->host_stream, ctx->down_stream, ctx->hdown_stream, ctx->hup_stream}) if (st) cudaStreamDestroy(st);
		if (ctx->ev0) cudaEventDestroy(ctx->ev0);
		if (ctx->ev1) cudaEventDestroy(ctx->ev1);
		delete ctx;
		return 100 + (int)e;
	}
	unsigned hw = std::thread::hardware_concurrency();
	ctx->pack_threads = (int)std::max(1u, std::min(hw ? hw : 8u, 32u));
	if (const char *s = getenv("KSW_B200_CHUNK")) ctx->chunk_jobs = ctx->async_chunk_jobs = std::max<int64_t>(1024, atoll(s));
	if (const char *s = getenv("KSW_B200_TRACE")) ctx->trace = atoi(s);
	if (const char *s = getenv("KSW_B200_HYBRID")) ctx->hybrid = atoi(s);
	if (const char *s = getenv("KSW_B200_HSLOTS")) ctx->n_hslots = std::max(1, std::min(KSW_N_HSLOTS, atoi(s)));
	if (const char *s = getenv("KSW_B200_HUP")) ctx->hup = atoi(s);
	if (const char *s = getenv("KSW_B200_ASYNC_CTAS")) ctx->async_fast_ctas = atoi(s);
	if (const char *s = getenv("KSW_B200_LEAD")) {
		for (const char *p = s; *p;) {
			char *end = nullptr;
			const long long v = strtoll(p, &end, 10);
			if (end == p) break;
			if (v > 0) ctx->lead_jobs.push_back(v);
			p = *end ? end + 1 : end;
		}
	}
	*out = ctx;
	return 0;
}

void ksw_b200_ctx_destroy(ksw_b200_ctx_t *ctx)
{
	if (!ctx) return;
	if (ctx->worker_started) {
		{
			std::unique_lock<std::mutex> lk(ctx->mu);
			ctx->cv.wait(lk, [&] { return !ctx->req_pending && !ctx->req_running; });   // a batch in flight finishes first
			ctx->stop = true;
		}
		ctx->cv.notify_all();
		ctx->worker.join();
	}
	cudaSetDevice(ctx->device);
	for (Slot &s : ctx->slot) {
		if (s.stream) cudaStreamSynchronize(s.stream);
		batch_release_buffers(&s.batch);
		s.h_jobs.release(); s.h_pool.release(); s.h_npool.release(); s.h_res.release();
		s.d_eh.release(); s.d_qc.release(); s.d_counter.release();
		s.d_rawjobs.release(); s.d_offs.release(); s.d_stats.release();
		s.h_stats.release();
		for (cudaEvent_t ev : {s.ev_jobs, s.ev_stats, s.ev_up, s.ev_packed, s.ev_ext, s.ev_done}) if (ev) cudaEventDestroy(ev);
		if (s.stream) cudaStreamDestroy(s.stream);
	}
	ctx->d_qraw.release(); ctx->d_traw.release();
	ctx->h_rq.release();
	ref_release(ctx);
	for (cudaStream_t st : {ctx->up_stream, ctx->pre_stream, ctx->main_stream, ctx->ext2_stream, ctx->host_stream, ctx->down_stream, ctx->hdown_stream, ctx->hup_stream}) if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
	for (Slot &s : ctx->hslot) {
		batch_release_buffers(&s.batch);
		s.h_jobs.release(); s.h_pool.release(); s.h_npool.release();
		s.d_eh.release(); s.d_qc.release(); s.d_counter.release();
		for (cudaEvent_t ev : {s.ev_up, s.ev_ext, s.ev_done}) if (ev) cudaEventDestroy(ev);
	}
	if (ctx->ev0) cudaEventDestroy(ctx->ev0);
	if (ctx->ev1) cudaEventDestroy(ctx->ev1);
	if (ctx->ev_mid) cudaEventDestroy(ctx->ev_mid);
	ctx->g_st[0].release(); ctx->g_st[1].release();
	ctx->a_hjobs.release(); ctx->a_hseq.release(); ctx->a_hres.release(); ctx->a_horder.release();
	ctx->a_djobs.release(); ctx->a_dseq.release(); ctx->a_dres.release(); ctx->a_dorder.release(); ctx->a_dcounter.release();
	if (ctx->a_bscr) { cudaFree(ctx->a_bscr); ctx->a_bscr = nullptr; ctx->a_bscr_cap = 0; }
	ctx->g_deh.release(); ctx->g_dqc.release(); ctx->g_dz.release(); ctx->g_dzfast.release(); ctx->g_dcounter.release(); ctx->g_dscratch.release();
	delete ctx->pool;
	delete ctx;
}

const char *ksw_b200_strerror(const ksw_b200_ctx_t *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int ksw_b200_ctx_set_pack_threads(ksw_b200_ctx_t *ctx, int n_threads)
{
	if (!ctx || n_threads < 1) return 1;
	ctx->pack_threads = n_threads;
	return 0;
}

int ksw_b200_ctx_set_chunk_jobs(ksw_b200_ctx_t *ctx, int64_t chunk_jobs)
{
	if (!ctx || chunk_jobs < 1) return 1;
	ctx->chunk_jobs = ctx->async_chunk_jobs = chunk_jobs;
	return 0;
}

int64_t ksw_b200_ctx_launch_count(const ksw_b200_ctx_t *ctx) { return ctx ? (int64_t)ctx->launches.load() : 0; }

int ksw_b200_ctx_sync(ksw_b200_ctx_t *ctx)
{
	if (!ctx) return 1;
	CU(cudaSetDevice(ctx->device));
	for (Slot &s : ctx->slot) CU(cudaStreamSynchronize(s.stream));
	return 0;
}

int ksw_b200_batch_upload(ksw_b200_ctx_t *ctx, const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_job_t *jobs,
                          const uint8_t *qpool, const uint8_t *tpool, ksw_b200_batch_t **out)
{
	if (!ctx || !cfg || !out || n < 0) return 1;
	*out = nullptr;
	CU(cudaSetDevice(ctx->device));
	Slot &s = ctx->slot[0];
	ksw_b200_batch *b = new ksw_b200_batch();
	int rc = pack_and_upload(ctx, s, cfg, n, jobs, qpool, tpool, b, nullptr, nullptr);
	if (rc == 0) {
		cudaError_t e = cudaStreamSynchronize(s.stream);       // the staging is reused by the next upload
		if (e != cudaSuccess) rc = fail(ctx, 100 + (int)e, std::string("upload sync: ") + cudaGetErrorString(e));
	}
	if (rc) { batch_release_buffers(b); delete b; return rc; }
	*out = b;
	return 0;
}

int ksw_b200_batch_run(ksw_b200_ctx_t *ctx, ksw_b200_batch_t *b)
{
	if (!ctx || !b) return 1;
	CU(cudaSetDevice(ctx->device));
	return enqueue_kernels(ctx, ctx->slot[0], b);
}

int ksw_b200_batch_run_timed2(ksw_b200_ctx_t *ctx, ksw_b200_batch_t *b, int iters, float *ms, float *ms_ext)
{
	if (!ctx || !b || iters < 1 || !ms) return 1;
	CU(cudaSetDevice(ctx->device));
	Slot &s = ctx->slot[0];
	const size_t n1 = (size_t)std::max<int64_t>(b->n, 1);
	if (ms_ext && !ctx->ev_mid) CU(cudaEventCreate(&ctx->ev_mid));
	for (int i = 0; i < iters; ++i) {
		CU(cudaEventRecord(ctx->ev0, s.stream));
		// a step is everything the GPU does for a packed batch: binning (key kernel + radix sort) and the extension kernels
		if (b->n > 0 && !b->dev_ranges) {
			CU(ksw_launch_bin((const DevJob *)b->d_jobs.p, b->n, (uint16_t *)b->d_keys.p, (uint16_t *)b->d_keys.p + n1,
			                  (uint32_t *)b->d_vals.p, (uint32_t *)b->d_order.p, b->d_sort_tmp.p, b->d_sort_tmp.cap, s.stream));
			ctx->launches += 2;
		}
		if (ms_ext) CU(cudaEventRecord(ctx->ev_mid, s.stream));
		int rc = enqueue_kernels(ctx, s, b);
		if (rc) return rc;
		CU(cudaEventRecord(ctx->ev1, s.stream));
		CU(cudaEventSynchronize(ctx->ev1));
		CU(cudaEventElapsedTime(&ms[i], ctx->ev0, ctx->ev1));
		if (ms_ext) CU(cudaEventElapsedTime(&ms_ext[i], ctx->ev_mid, ctx->ev1));
	}
	return 0;
}

int ksw_b200_batch_run_timed(ksw_b200_ctx_t *ctx, ksw_b200_batch_t *b, int iters, float *ms)
{
	return ksw_b200_batch_run_timed2(ctx, b, iters, ms, nullptr);
}

int ksw_b200_batch_download(ksw_b200_ctx_t *ctx, ksw_b200_batch_t *b, ksw_b200_res_t *res)
{
	if (!ctx || !b || (!res && b->n)) return 1;
	CU(cudaSetDevice(ctx->device));
	Slot &s = ctx->slot[0];
	if (b->n == 0) { CU(cudaStreamSynchronize(s.stream)); return 0; }
	const size_t bytes = sizeof(DevRes) * (size_t)b->n;
	CU(s.h_res.reserve(bytes));
	CU(cudaMemcpyAsync(s.h_res.p, b->d_res.p, bytes, cudaMemcpyDeviceToHost, s.stream));
	CU(cudaStreamSynchronize(s.stream));
	memcpy(res, s.h_res.p, bytes);
	return 0;
}

int ksw_b200_batch_download_cells(ksw_b200_ctx_t *ctx, ksw_b200_batch_t *b, uint32_t *cells)
{
	if (!ctx || !b || (!cells && b->n)) return 1;
	CU(cudaSetDevice(ctx->device));
	CU(cudaStreamSynchronize(ctx->slot[0].stream));
	if (b->n) CU(cudaMemcpy(cells, b->d_cells.p, sizeof(uint32_t) * (size_t)b->n, cudaMemcpyDeviceToHost));
	return 0;
}

int ksw_b200_batch_info(const ksw_b200_batch_t *b, int64_t *n_fast, int64_t *n_generic, int64_t *packed_bytes)
{
	if (!b) return 1;
	if (n_fast) *n_fast = b->n_fast;
	if (n_generic) *n_generic = b->n_warp + b->n_generic;       // everything outside the s16x2 kernel
	if (packed_bytes) *packed_bytes = (int64_t)(b->pool_bytes + b->npool_bytes + sizeof(DevJob) * (size_t)b->n);
	return 0;
}

void ksw_b200_batch_free(ksw_b200_ctx_t *ctx, ksw_b200_batch_t *b)
{
	if (ctx) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->slot[0].stream); }
	if (b) { batch_release_buffers(b); delete b; }
}

static int extend_batch_pipelined(ksw_b200_ctx_t *ctx, const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_job_t *jobs,
                                  const uint8_t *qpool, const uint8_t *tpool, ksw_b200_res_t *res)
{
	const double t_begin = now_ms();
	double t_plan = 0, t_fill = 0, t_wait = 0;
	int64_t h2d = 0;
	const int64_t chunk = ctx->chunk_jobs;
	int which = 0;
	for (int64_t first = 0; first < n; first += chunk, which = (which + 1) % KSW_N_SLOTS) {
		const int64_t nc = std::min(chunk, n - first);
		Slot &s = ctx->slot[which];
		int rc = retire_slot(ctx, s, res, &t_wait);              // its staging and device buffers are about to be reused
		if (rc) return rc;
		rc = pack_and_upload(ctx, s, cfg, nc, jobs + first, qpool, tpool, &s.batch, &t_plan, &t_fill);
		if (rc) return rc;
		rc = enqueue_kernels(ctx, s, &s.batch);
		if (rc) return rc;
		CU(s.h_res.reserve(sizeof(DevRes) * (size_t)nc));
		CU(cudaMemcpyAsync(s.h_res.p, s.batch.d_res.p, sizeof(DevRes) * (size_t)nc, cudaMemcpyDeviceToHost, s.stream));
		s.busy = true; s.first = first; s.n = nc;
		h2d += (int64_t)(sizeof(DevJob) * (size_t)nc + s.batch.pool_bytes + s.batch.npool_bytes);
	}
	// retire in submission order (the oldest chunk in flight sits in the slot that would be used next)
	for (int k = 0; k < KSW_N_SLOTS; ++k) {
		int rc = retire_slot(ctx, ctx->slot[(which + k) % KSW_N_SLOTS], res, &t_wait);
		if (rc) return rc;
	}
	ctx->last_h2d = h2d;
	ctx->last_d2h = (int64_t)(sizeof(DevRes) * (size_t)n);
	if (ctx->trace)
		fprintf(stderr, "[ksw_b200] extend_batch n=%lld chunk=%lld: total %.2f ms (plan %.2f, fill %.2f, wait-gpu %.2f)\n",
		        (long long)n, (long long)chunk, now_ms() - t_begin, t_plan, t_fill, t_wait);
	return 0;
}

// One-shot batched entry.  The batch is cut into chunks of ctx->chunk_jobs jobs (caller order); chunk c+1 is
// packed on the host threads while chunks c and c-1 are being copied / computed on the GPU (three slots, three streams).
int ksw_b200_extend_batch(ksw_b200_ctx_t *ctx, const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_job_t *jobs,
                          const uint8_t *qpool, const uint8_t *tpool, ksw_b200_res_t *res)
{
	if (!ctx || !cfg || n < 0) return 1;
	if (n == 0) return 0;
	if (!jobs || !res) return 1;
	CU(cudaSetDevice(ctx->device));
	const int rc = extend_batch_pipelined(ctx, cfg, n, jobs, qpool, tpool, res);
	if (rc) {
		// leave the context reusable: nothing may stay in flight or marked busy after a failed call
		for (Slot &s : ctx->slot) { cudaStreamSynchronize(s.stream); s.busy = false; }
	}
	return rc;
}

// ---- pinned callers (ksw_b200_extend_batch_async).  The host does not have to touch the sequences: the raw job
// records and the raw byte-coded sequences go to HBM as they are, packing runs on the device (ksw_devpack.cu), results
// are copied straight into the caller's array.  Three streams: `up` (all host->device copies), `main` (all kernels),
// `down` (results).  Device lane, per chunk c (c' = the lane's next chunk):
//     up:    job records of c' | the bytes of the two sequence pools chunk c reads and no earlier chunk brought
//     main:  prep(c') + its totals D2H -> pack(c) -> bin(c) -> extension kernels(c)
//     down:  results(c) into res
// prep(c') sits BEFORE the kernels of chunk c: the extension kernels fill every SM's shared memory, so nothing launched
// later runs beside them, and the host needs the totals of c' (pool sizes, which pool bytes to upload) while the
// extension kernels of c are running, so that the uploads of c' overlap them.  The lane only ever waits for those
// totals and for a free slot.
// Host lane (hybrid): raw bytes are 2.4x the packed bytes, so with idle host threads the link is the bottleneck.  A
// second lane therefore packs chunks on the host threads (ksw_pack.cpp) and sends only their packed form.  Both lanes
// draw chunks from one counter, each at its own pace (the host lane stops early enough not to become the tail), so the
// split follows the hardware: with 16 host threads per GPU about half of the chunks, with 4 hardly any.
static bool is_pinned(const void *p)
{
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
	return a.type == cudaMemoryTypeHost;
}

namespace {

// byte ranges of a pool that are already in HBM (disjoint, sorted); add() returns the missing pieces of [lo, hi)
struct Uploaded {
	std::vector<std::pair<size_t, size_t>> iv;
	void add(size_t lo, size_t hi, std::vector<std::pair<size_t, size_t>> &missing)
	{
		missing.clear();
		if (lo >= hi) return;
		size_t cur = lo;
		for (const auto &x : iv) {
			if (x.second <= cur) continue;
			if (x.first >= hi) break;
			if (x.first > cur) missing.emplace_back(cur, x.first);
			cur = std::max(cur, x.second);
			if (cur >= hi) break;
		}
		if (cur < hi) missing.emplace_back(cur, hi);
		// merge [lo, hi) into the set
		std::vector<std::pair<size_t, size_t>> out;
		size_t a = lo, b = hi;
		bool placed = false;
		for (const auto &x : iv) {
			if (x.second < a) out.push_back(x);
			else if (x.first > b) { if (!placed) { out.emplace_back(a, b); placed = true; } out.push_back(x); }
			else { a = std::min(a, x.first); b = std::max(b, x.second); }
		}
		if (!placed) out.emplace_back(a, b);
		iv.swap(out);
	}
};

struct ChunkTimes {                       // KSW_B200_TRACE >= 2: where each chunk's time went on the GPU
	cudaEvent_t up = nullptr, k0 = nullptr, k1 = nullptr, done = nullptr;
	int lane = 0;
	double host_enq = 0;
};

struct DevpackShared {                    // what the two lanes of one call share
	std::atomic<long long> next{0};       // next unclaimed chunk
	std::atomic<int> rc{0};               // first error of either lane
	long long n_chunks = 0;
	std::vector<int64_t> start;           // chunk ci = jobs [start[ci], start[ci+1]); the first two chunks are short so
	                                      // that the GPU gets something to do early
	std::atomic<long long> h2d{0}, host_chunks{0};
	std::vector<ChunkTimes> times;        // empty unless tracing
	cudaEvent_t t0 = nullptr;
	double host_t0 = 0;
};

} // namespace

// finishing touches both lanes share: extension kernels, results home, slot marked busy
static int devpack_finish_chunk(ksw_b200_ctx_t *ctx, Slot &s, int64_t first, int64_t nc, ksw_b200_res_t *res, cudaStream_t lane,
                                cudaStream_t down, ChunkTimes *tm = nullptr)
{
	if (tm) { cudaEventCreate(&tm->k0); cudaEventCreate(&tm->k1); cudaEventCreate(&tm->done); cudaEventRecord(tm->k0, lane); }
	int rc = enqueue_kernels(ctx, s, &s.batch, lane, ctx->async_fast_ctas);
	if (rc) return rc;
	if (tm) cudaEventRecord(tm->k1, lane);
	CU(cudaEventRecord(s.ev_ext, lane));
	CU(cudaStreamWaitEvent(down, s.ev_ext, 0));
	CU(cudaMemcpyAsync(res + first, s.batch.d_res.p, sizeof(DevRes) * (size_t)nc, cudaMemcpyDeviceToHost, down));
	CU(cudaEventRecord(s.ev_done, down));
	if (tm) cudaEventRecord(tm->done, down);
	s.busy = true; s.first = first; s.n = nc;
	return 0;
}

static int devpack_wait_slot(ksw_b200_ctx_t *ctx, Slot &s)
{
	if (!s.busy) return 0;
	CU(cudaEventSynchronize(s.ev_done));
	s.busy = false;
	return 0;
}

// the host lane: packs whole chunks on the pack threads while the device lane works on others
static void devpack_host_lane(ksw_b200_ctx_t *ctx, DevpackShared *sh, const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_job_t *jobs,
                              const uint8_t *qpool, const uint8_t *tpool, ksw_b200_res_t *res)
{
	if (cudaSetDevice(ctx->device) != cudaSuccess) return;
	int k = 0;
	for (;;) {
		if (sh->rc.load()) break;
		// take a chunk only if the device lane will still be busy when it is packed: never become the tail
		const long long remaining = sh->n_chunks - sh->next.load();
		long long need = 4;
		if (ctx->host_ms_per_job > 0 && ctx->dev_ms_per_job > 0)
			need = (long long)(ctx->host_ms_per_job / ctx->dev_ms_per_job + 0.999) + 1;
		if (remaining < need) break;
		const long long ci = sh->next.fetch_add(1);
		if (ci >= sh->n_chunks) break;
		const int64_t first = sh->start[(size_t)ci], nc = sh->start[(size_t)ci + 1] - first;
		Slot &s = ctx->hslot[k++ % ctx->n_hslots];
		int rc = devpack_wait_slot(ctx, s);
		double t_plan = 0, t_fill = 0;
		cudaStream_t hup = ctx->hup ? ctx->hup_stream : ctx->up_stream;
		if (!rc) rc = pack_and_upload(ctx, s, cfg, nc, jobs + first, qpool, tpool, &s.batch, &t_plan, &t_fill, hup, ctx->host_stream);
		const double dt = t_plan + t_fill;                           // packing proper: buffer growth (first calls) is not the lane's pace
		ChunkTimes *tm = sh->times.empty() ? nullptr : &sh->times[(size_t)ci];
		if (tm) { tm->lane = 1; tm->host_enq = now_ms() - sh->host_t0; cudaEventCreate(&tm->up); cudaEventRecord(tm->up, hup); }
		if (!rc) rc = devpack_finish_chunk(ctx, s, first, nc, res, ctx->host_stream, ctx->hdown_stream, tm);
		if (rc) { int z = 0; sh->rc.compare_exchange_strong(z, rc); break; }
		ctx->host_ms_per_job = ctx->host_ms_per_job <= 0 ? dt / (double)nc : 0.5 * ctx->host_ms_per_job + 0.5 * dt / (double)nc;
		sh->h2d += (long long)(sizeof(DevJob) * (size_t)nc + s.batch.pool_bytes + s.batch.npool_bytes);
		sh->host_chunks++;
	}
}

static int extend_batch_devpack(ksw_b200_ctx_t *ctx, const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_job_t *jobs,
                                const uint8_t *qpool, size_t qbytes, const uint8_t *tpool, size_t tbytes, ksw_b200_res_t *res)
{
	if (cfg->m != 5) return fail(ctx, 2, "ksw_b200: only m == 5 is supported (every reference caller passes 5)");
	if (n > 0x7fffffffLL) return fail(ctx, 2, "ksw_b200: more than 2^31-1 jobs in one batch");
	const double t_begin = now_ms();
	double t_wait = 0;
	KswScoring S;
	ksw_scoring_from_cfg(cfg, fast_qmax_enabled(), S);
	KswParams P;
	ksw_params_from_cfg(cfg, P);
	CU(ctx->d_qraw.reserve(qbytes + 64));
	CU(ctx->d_traw.reserve(tbytes + 64));
	cudaStream_t up = ctx->up_stream, pre = ctx->pre_stream;
	const int64_t chunk = ctx->async_chunk_jobs;
	DevpackShared sh;
	{
		int64_t at = 0;
		std::vector<int64_t> lead = ctx->lead_jobs;
		if (lead.empty()) lead = {std::max<int64_t>(chunk / 4, 1024), std::max<int64_t>(chunk / 2, 1024)};
		for (size_t i = 0; at < n; ++i) {
			sh.start.push_back(at);
			at += std::min<int64_t>(n - at, (i < lead.size() && n > 2 * chunk) ? std::max<int64_t>(lead[i], 1024) : chunk);
		}
		sh.start.push_back(n);
		sh.n_chunks = (long long)sh.start.size() - 1;
	}
	Uploaded q_up, t_up;
	std::vector<std::pair<size_t, size_t>> missing;
	if (ctx->trace >= 2) {
		sh.times.resize((size_t)sh.n_chunks);
		sh.host_t0 = now_ms();
		cudaEventCreate(&sh.t0);
		cudaEventRecord(sh.t0, up);
	}
	std::thread host_lane;
	if (ctx->hybrid && sh.n_chunks >= 4) {
		pool_of(ctx);                                               // created here, used by the host lane only
		// the lane's pinned staging is sized here, from the average job, so that its first chunk does not pay for it
		const size_t per_job = (size_t)((qbytes + tbytes) / (size_t)n) / 4 + 24;
		for (int i = 0; i < ctx->n_hslots; ++i) {
			Slot &hs = ctx->hslot[i];
			CU(hs.h_jobs.reserve(sizeof(DevJob) * (size_t)chunk));
			CU(hs.h_pool.reserve(per_job * (size_t)chunk));
		}
		host_lane = std::thread(devpack_host_lane, ctx, &sh, cfg, n, jobs, qpool, tpool, res);
	}
	auto claim = [&]() -> long long { const long long c = sh.next.fetch_add(1); return c < sh.n_chunks ? c : -1; };
	// job records of chunk ci to the device (upload stream), then its prep kernel and its totals home (main stream)
	std::vector<int> slot_of((size_t)sh.n_chunks, 0);              // the device lane's chunks take its slots in turn
	long long dev_seq = 0;
	auto start_chunk = [&](long long ci) -> int {
		slot_of[(size_t)ci] = (int)(dev_seq++ % KSW_N_SLOTS);
		Slot &s = ctx->slot[slot_of[(size_t)ci]];
		const int64_t first = sh.start[(size_t)ci], nc = sh.start[(size_t)ci + 1] - first;
		const size_t n1 = (size_t)nc;
		{
			const double t0 = now_ms();
			int rc = devpack_wait_slot(ctx, s);
			t_wait += now_ms() - t0;
			if (rc) return rc;
		}
		CU(s.d_rawjobs.reserve(sizeof(ksw_b200_job_t) * n1));
		CU(s.d_offs.reserve(sizeof(uint32_t) * n1));
		CU(s.d_stats.reserve(sizeof(DevPackStats)));
		CU(s.h_stats.reserve(sizeof(DevPackStats)));
		CU(s.batch.d_jobs.reserve(sizeof(DevJob) * n1));
		CU(cudaMemcpyAsync(s.d_rawjobs.p, jobs + first, sizeof(ksw_b200_job_t) * n1, cudaMemcpyHostToDevice, up));
		CU(cudaEventRecord(s.ev_jobs, up));
		sh.h2d += (long long)(sizeof(ksw_b200_job_t) * n1);
		CU(cudaStreamWaitEvent(pre, s.ev_jobs, 0));
		CU(ksw_launch_prep(s.d_rawjobs.p, nc, S, (DevJob *)s.batch.d_jobs.p, (uint32_t *)s.d_offs.p, (DevPackStats *)s.d_stats.p, pre));
		CU(cudaMemcpyAsync(s.h_stats.p, s.d_stats.p, sizeof(DevPackStats), cudaMemcpyDeviceToHost, pre));
		CU(cudaEventRecord(s.ev_stats, pre));
		ctx->launches += 1;
		return 0;
	};
	auto device_lane = [&]() -> int {
		// the lane works two chunks ahead: while chunk `cur` is being uploaded and run, the records of the next two are
		// already on the device and their prep kernels sit ahead of cur's kernels in the stream, so that the totals of
		// chunk cur+1 are back before cur's upload ends and the link never waits for the host
		long long cur = claim(), nxt = -1;
		if (cur >= 0) { int rc = start_chunk(cur); if (rc) return rc; nxt = claim(); }
		if (nxt >= 0) { int rc = start_chunk(nxt); if (rc) return rc; }
		long long dev_jobs_done = 0, ext_seq = 0;
		while (cur >= 0) {
			if (sh.rc.load()) return 0;
			const int64_t first = sh.start[(size_t)cur], nc = sh.start[(size_t)cur + 1] - first;
			const size_t n1 = (size_t)nc;
			Slot &s = ctx->slot[slot_of[(size_t)cur]];
			ksw_b200_batch *b = &s.batch;
			{
				const double t0 = now_ms();
				CU(cudaEventSynchronize(s.ev_stats));
				t_wait += now_ms() - t0;
			}
			const DevPackStats st = *(const DevPackStats *)s.h_stats.p;
			if (st.bad) return fail(ctx, 2, "ksw_b200: job with qlen < 1 or tlen < 0");
			if (st.q_hi > qbytes || st.t_hi > tbytes) return fail(ctx, 2, "ksw_b200: a job reads past the end of its sequence pool");
			if (st.units > 0xffffffffull) return fail(ctx, 2, "ksw_b200: packed pool exceeds 64 GiB");
			b->n = nc; b->n_fast = 0;
			for (int c = 0; c < KSW_FAST_CLASSES; ++c) {
				b->fast_class_n[c] = st.class_n[c];
				b->fast_class_qmax[c] = st.class_qmax[c];
				b->n_fast += st.class_n[c];
			}
			b->n_warp = st.class_n[KSW_CLASS_WARP]; b->qmax_warp = st.class_qmax[KSW_CLASS_WARP];
			b->n_generic = st.class_n[KSW_CLASS_THREAD];
			b->qmax_generic = st.class_qmax[KSW_CLASS_THREAD];
			b->pool_bytes = (size_t)st.units * 16; b->npool_bytes = (size_t)st.nmask_words * 4;
			b->P = P;
			b->dev_ranges = true;
			CU(b->d_pool.reserve(std::max<size_t>(b->pool_bytes, 16)));
			CU(b->d_npool.reserve(std::max<size_t>(b->npool_bytes, 16)));
			CU(b->d_res.reserve(sizeof(DevRes) * n1));
			CU(b->d_cells.reserve(sizeof(uint32_t) * n1));
			CU(b->d_order.reserve(sizeof(uint32_t) * n1));
			CU(b->d_keys.reserve(sizeof(uint16_t) * 2 * n1));
			CU(b->d_vals.reserve(sizeof(uint32_t) * n1));
			CU(b->d_sort_tmp.reserve(std::max<size_t>(std::max(ksw_bin_temp_bytes(nc), ksw_bin_counting_bytes()), 16)));
			CU(b->d_range.reserve(sizeof(uint32_t) * (KSW_N_CLASSES + 1)));
			// the chunk after next: its records travel ahead of this chunk's sequences, its prep kernel runs ahead of
			// this chunk's kernels
			const long long nxt2 = nxt >= 0 ? claim() : -1;
			if (nxt2 >= 0) { int rc = start_chunk(nxt2); if (rc) return rc; }
			// the bytes of the two pools this chunk reads and no earlier chunk has brought over
			q_up.add((size_t)~st.q_lo_inv, (size_t)st.q_hi, missing);
			for (const auto &m : missing) {
				CU(cudaMemcpyAsync((uint8_t *)ctx->d_qraw.p + m.first, qpool + m.first, m.second - m.first, cudaMemcpyHostToDevice, up));
				sh.h2d += (long long)(m.second - m.first);
			}
			t_up.add((size_t)~st.t_lo_inv, (size_t)st.t_hi, missing);
			for (const auto &m : missing) {
				CU(cudaMemcpyAsync((uint8_t *)ctx->d_traw.p + m.first, tpool + m.first, m.second - m.first, cudaMemcpyHostToDevice, up));
				sh.h2d += (long long)(m.second - m.first);
			}
			CU(cudaEventRecord(s.ev_up, up));
			CU(cudaStreamWaitEvent(pre, s.ev_up, 0));
			ChunkTimes *tm = sh.times.empty() ? nullptr : &sh.times[(size_t)cur];
			if (tm) { tm->lane = 0; tm->host_enq = now_ms() - sh.host_t0; cudaEventCreate(&tm->up); cudaEventRecord(tm->up, up); }
			// pack / key kernels need no shared memory and run beside the previous chunk's extension kernels; the sort
			// does and runs when those end
			CU(ksw_launch_pack(s.d_rawjobs.p, nc, (const uint8_t *)ctx->d_qraw.p, (const uint8_t *)ctx->d_traw.p, (const uint32_t *)s.d_offs.p,
			                   (DevJob *)b->d_jobs.p, (uint32_t *)b->d_pool.p, (uint32_t *)b->d_npool.p, (DevPackStats *)s.d_stats.p, pre));
			// binning as a counting sort without shared memory: it, too, runs beside the previous chunk's extension kernels,
			// so this chunk's are ready to take over the SMs as those run out of jobs
			CU(ksw_launch_bin_counting((const DevJob *)b->d_jobs.p, nc, (uint16_t *)b->d_keys.p, b->d_sort_tmp.p,
			                           (uint32_t *)b->d_order.p, (uint32_t *)b->d_range.p, pre));
			CU(cudaEventRecord(s.ev_packed, pre));
			ctx->launches += 4;                                    // pack, histogram, scan, scatter
			// extension kernels of consecutive chunks alternate between two streams: the next launch fills the SMs as
			// the warps of the previous one run out of jobs
			cudaStream_t ext = (ext_seq++ & 1) ? ctx->ext2_stream : ctx->main_stream;
			CU(cudaStreamWaitEvent(ext, s.ev_packed, 0));
			int rc = devpack_finish_chunk(ctx, s, first, nc, res, ext, ctx->down_stream, tm);
			if (rc) return rc;
			dev_jobs_done += nc;
			ctx->dev_ms_per_job = (now_ms() - t_begin) / (double)dev_jobs_done;   // the lane's pace so far, throttled by the GPU
			cur = nxt; nxt = nxt2;
		}
		return 0;
	};
	int rc = device_lane();
	if (rc) { int z = 0; sh.rc.compare_exchange_strong(z, rc); }
	if (host_lane.joinable()) host_lane.join();
	// a host lane that sat a call out (its last measured pace looked too slow, e.g. a cold first call) gets another try
	if (ctx->hybrid && sh.n_chunks >= 4 && sh.host_chunks.load() == 0) ctx->host_ms_per_job *= 0.5;
	if (sh.rc.load()) return sh.rc.load();
	for (Slot &s : ctx->slot) { rc = devpack_wait_slot(ctx, s); if (rc) return rc; }
	for (Slot &s : ctx->hslot) { rc = devpack_wait_slot(ctx, s); if (rc) return rc; }
	ctx->last_h2d = (int64_t)sh.h2d.load();
	ctx->last_d2h = (int64_t)(sizeof(DevRes) * (size_t)n + sizeof(DevPackStats) * (size_t)(sh.n_chunks - sh.host_chunks.load()));
	if (!sh.times.empty()) {
		fprintf(stderr, "[ksw_b200] chunk lane host-enqueue | upload-done kernels-start kernels-end results-home  (ms since the call began)\n");
		for (size_t c = 0; c < sh.times.size(); ++c) {
			ChunkTimes &t = sh.times[c];
			float a = 0, b0 = 0, b1 = 0, d = 0;
			if (t.up) cudaEventElapsedTime(&a, sh.t0, t.up);
			if (t.k0) cudaEventElapsedTime(&b0, sh.t0, t.k0);
			if (t.k1) cudaEventElapsedTime(&b1, sh.t0, t.k1);
			if (t.done) cudaEventElapsedTime(&d, sh.t0, t.done);
			fprintf(stderr, "[ksw_b200] %3zu %s %7.2f | %7.2f %7.2f %7.2f %7.2f\n", c, t.lane ? "host" : "dev ", t.host_enq, a, b0, b1, d);
			for (cudaEvent_t ev : {t.up, t.k0, t.k1, t.done}) if (ev) cudaEventDestroy(ev);
		}
		cudaEventDestroy(sh.t0);
	}
	if (ctx->trace)
		fprintf(stderr, "[ksw_b200] extend_batch_async n=%lld chunk=%lld: total %.2f ms (device lane waited %.2f; host lane packed %lld of %lld chunks; "
		                "pace %.2f / %.2f ns per job)\n", (long long)n, (long long)chunk, now_ms() - t_begin, t_wait,
		        (long long)sh.host_chunks.load(), (long long)sh.n_chunks, ctx->dev_ms_per_job * 1e6, ctx->host_ms_per_job * 1e6);
	return 0;
}

static int run_devpack(ksw_b200_ctx_t *ctx, const AsyncReq &r)
{
	cudaError_t e = cudaSetDevice(ctx->device);
	if (e != cudaSuccess) return fail(ctx, 100 + (int)e, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
	const int rc = extend_batch_devpack(ctx, &r.cfg, r.n, r.jobs, r.qpool, r.qbytes, r.tpool, r.tbytes, r.res);
	if (rc) {
		// leave the context reusable: nothing may stay in flight or marked busy after a failed call
		for (cudaStream_t st : {ctx->up_stream, ctx->pre_stream, ctx->main_stream, ctx->ext2_stream, ctx->host_stream, ctx->down_stream, ctx->hdown_stream, ctx->hup_stream}) cudaStreamSynchronize(st);
		for (Slot &s : ctx->slot) { cudaStreamSynchronize(s.stream); s.busy = false; }
		for (Slot &s : ctx->hslot) s.busy = false;
	}
	return rc;
}

static void worker_main(ksw_b200_ctx_t *ctx)
{
	for (;;) {
		AsyncReq r;
		{
			std::unique_lock<std::mutex> lk(ctx->mu);
			ctx->cv.wait(lk, [&] { return ctx->stop || ctx->req_pending; });
			if (ctx->stop) return;
			r = ctx->req;
			ctx->req_pending = false; ctx->req_running = true;
		}
		const int rc = run_devpack(ctx, r);
		{
			std::unique_lock<std::mutex> lk(ctx->mu);
			ctx->async_rc = rc;
			ctx->req_running = false;
		}
		ctx->cv.notify_all();
	}
}

void *ksw_b200_host_alloc(size_t bytes)
{
	void *p = nullptr;
	if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	return p;
}

void ksw_b200_host_free(void *p) { if (p) cudaFreeHost(p); }

int ksw_b200_host_register(void *p, size_t bytes)
{
	if (!p || !bytes) return 1;
	cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
	if (e != cudaSuccess) { cudaGetLastError(); return 100 + (int)e; }
	return 0;
}

int ksw_b200_host_unregister(void *p)
{
	if (!p) return 1;
	cudaError_t e = cudaHostUnregister(p);
	if (e != cudaSuccess) { cudaGetLastError(); return 100 + (int)e; }
	return 0;
}

// Asynchronous batched entry for page-locked caller buffers: returns as soon as the batch is handed to
// the context's feeder thread; ksw_b200_wait returns when res[0..n) is complete.  One batch in flight per context.
int ksw_b200_extend_batch_async(ksw_b200_ctx_t *ctx, const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_job_t *jobs,
                                const uint8_t *qpool, size_t qpool_bytes, const uint8_t *tpool, size_t tpool_bytes,
                                ksw_b200_res_t *res)
{
	if (!ctx || !cfg || n < 0) return 1;
	if (n > 0 && (!jobs || !res)) return 1;
	{
		std::unique_lock<std::mutex> lk(ctx->mu);
		if (ctx->req_pending || ctx->req_running) return fail(ctx, 4, "ksw_b200_extend_batch_async: a batch is already in flight on this context (call ksw_b200_wait first)");
	}
	if (n > 0) {
		CU(cudaSetDevice(ctx->device));
		if (!is_pinned(jobs) || !is_pinned(res) || (qpool_bytes && !is_pinned(qpool)) || (tpool_bytes && !is_pinned(tpool)))
			return fail(ctx, 3, "ksw_b200_extend_batch_async: jobs, qpool, tpool and res must be page-locked "
			                    "(ksw_b200_host_alloc / ksw_b200_host_register); pageable callers use ksw_b200_extend_batch");
	}
	std::unique_lock<std::mutex> lk(ctx->mu);
	if (!ctx->worker_started) {
		ctx->worker = std::thread(worker_main, ctx);
		ctx->worker_started = true;
	}
	ctx->req.cfg = *cfg; ctx->req.n = n; ctx->req.jobs = jobs; ctx->req.qpool = qpool; ctx->req.tpool = tpool;
	ctx->req.qbytes = qpool_bytes; ctx->req.tbytes = tpool_bytes; ctx->req.res = res;
	ctx->async_rc = 0;
	ctx->req_pending = true;
	lk.unlock();
	ctx->cv.notify_all();
	return 0;
}

int ksw_b200_wait(ksw_b200_ctx_t *ctx)
{
	if (!ctx) return 1;
	std::unique_lock<std::mutex> lk(ctx->mu);
	ctx->cv.wait(lk, [&] { return !ctx->req_pending && !ctx->req_running; });
	return ctx->async_rc;
}

// ---- reference resident on the device ----
int ksw_b200_ref_set(ksw_b200_ctx_t *ctx, const uint8_t *pac, int64_t l_pac)
{
	if (!ctx || !pac || l_pac < 1) return 1;
	CU(cudaSetDevice(ctx->device));
	ref_release(ctx);
	std::lock_guard<std::mutex> lk(g_ref_mu);
	for (RefEntry *r : g_refs)
		if (r->device == ctx->device && r->host == pac && r->l_pac == l_pac) { ++r->refs; ctx->ref = r; return 0; }
	RefEntry *r = new RefEntry();
	r->device = ctx->device; r->host = pac; r->l_pac = l_pac; r->refs = 1;
	const size_t bytes = (size_t)(l_pac / 4 + 1);
	cudaError_t e = cudaMalloc(&r->dev, bytes + 16);
	if (e == cudaSuccess) e = cudaMemcpy(r->dev, pac, bytes, cudaMemcpyHostToDevice);
	if (e != cudaSuccess) {
		if (r->dev) cudaFree(r->dev);
		delete r;
		return fail(ctx, 100 + (int)e, std::string("ksw_b200_ref_set: ") + cudaGetErrorString(e));
	}
	g_refs.push_back(r);
	ctx->ref = r;
	return 0;
}

// Synchronous; the callers' arrays may be pageable (they are small: 40 B per job plus the reads).  The batch may come
// in several SEGMENTS (one per submitting host thread, ksw_queue.cpp) that run as one GPU batch: job records and read
// pools are gathered into the slot's pinned staging (q_off rebased), and the results are scattered back.  Per round
// (<= chunk_jobs jobs) on slot 0's stream: records + read pools H2D, prep, totals home, pack from the resident .pac,
// bin, kernels, results home.
int ksw_b200_extend_batch_ref_segs(ksw_b200_ctx_t *ctx, const ksw_b200_cfg_t *cfg, int n_segs, const ksw_b200_rseg_t *segs)
{
	if (!ctx || !cfg || n_segs < 0 || (n_segs > 0 && !segs)) return 1;
	for (int g = 0; g < n_segs; ++g)
		if (segs[g].n < 0 || (segs[g].n > 0 && (!segs[g].jobs || !segs[g].res || !segs[g].qpool))) return 1;
	if (!ctx->ref) return fail(ctx, 5, "ksw_b200_extend_batch_ref: no reference on the device (call ksw_b200_ref_set first)");
	if (cfg->m != 5) return fail(ctx, 2, "ksw_b200: only m == 5 is supported (every reference caller passes 5)");
	CU(cudaSetDevice(ctx->device));
	Slot &s = ctx->slot[0];
	ksw_b200_batch *b = &s.batch;
	KswScoring S;
	ksw_scoring_from_cfg(cfg, fast_qmax_enabled(), S);
	KswParams P;
	ksw_params_from_cfg(cfg, P);
	int64_t h2d = 0, d2h = 0;
	const int64_t chunk = ctx->chunk_jobs;
	struct Piece { int seg; int64_t first, n; size_t qbase; };
	std::vector<Piece> round;
	int g = 0;
	int64_t at = 0;                                                  // next job of segment g
	while (g < n_segs) {
		// one round: pieces of consecutive segments, <= chunk jobs in all; the read pool of every segment it touches
		round.clear();
		int64_t nc = 0;
		size_t qbytes = 0;
		while (g < n_segs && nc < chunk) {
			const int64_t take = std::min(segs[g].n - at, chunk - nc);
			if (take > 0) {
				round.push_back(Piece{g, at, take, qbytes});
				qbytes += (segs[g].qpool_bytes + 15) & ~(size_t)15;
				nc += take; at += take;
			}
			if (at >= segs[g].n) { ++g; at = 0; }
		}
		if (nc == 0) break;
		const size_t n1 = (size_t)nc;
		CU(ctx->d_qraw.reserve(qbytes + 64));
		CU(ctx->h_rq.reserve(std::max<size_t>(qbytes, 16)));
		CU(s.d_rawjobs.reserve(sizeof(ksw_b200_rjob_t) * n1));
		CU(s.d_offs.reserve(sizeof(uint32_t) * n1));
		CU(s.d_stats.reserve(sizeof(DevPackStats)));
		CU(s.h_stats.reserve(sizeof(DevPackStats)));
		CU(s.h_jobs.reserve(sizeof(ksw_b200_rjob_t) * n1));
		CU(s.h_res.reserve(sizeof(DevRes) * n1));
		CU(b->d_jobs.reserve(sizeof(DevJob) * n1));
		{
			ksw_b200_rjob_t *hj = (ksw_b200_rjob_t *)s.h_jobs.p;
			int64_t k = 0;
			for (const Piece &pc : round) {
				memcpy((uint8_t *)ctx->h_rq.p + pc.qbase, segs[pc.seg].qpool, segs[pc.seg].qpool_bytes);
				memcpy(hj + k, segs[pc.seg].jobs + pc.first, sizeof(ksw_b200_rjob_t) * (size_t)pc.n);
				if (pc.qbase)
					for (int64_t x = 0; x < pc.n; ++x) hj[k + x].q_off += pc.qbase;
				k += pc.n;
			}
		}
		CU(cudaMemcpyAsync(ctx->d_qraw.p, ctx->h_rq.p, qbytes, cudaMemcpyHostToDevice, s.stream));
		CU(cudaMemcpyAsync(s.d_rawjobs.p, s.h_jobs.p, sizeof(ksw_b200_rjob_t) * n1, cudaMemcpyHostToDevice, s.stream));
		h2d += (int64_t)(qbytes + sizeof(ksw_b200_rjob_t) * n1);
		CU(ksw_launch_prep_ref(s.d_rawjobs.p, nc, S, ctx->ref->l_pac, qbytes, (DevJob *)b->d_jobs.p, (uint32_t *)s.d_offs.p,
		                       (DevPackStats *)s.d_stats.p, s.stream));
		CU(cudaMemcpyAsync(s.h_stats.p, s.d_stats.p, sizeof(DevPackStats), cudaMemcpyDeviceToHost, s.stream));
		CU(cudaStreamSynchronize(s.stream));
		const DevPackStats st = *(const DevPackStats *)s.h_stats.p;
		if (st.bad) return fail(ctx, 2, "ksw_b200_extend_batch_ref: job with bad lengths / steps, a run outside its pool, or a target bridging the strands");
		if (st.units > 0xffffffffull) return fail(ctx, 2, "ksw_b200: packed pool exceeds 64 GiB");
		b->n = nc; b->n_fast = 0;
		for (int c = 0; c < KSW_FAST_CLASSES; ++c) {
			b->fast_class_n[c] = st.class_n[c];
			b->fast_class_qmax[c] = st.class_qmax[c];
			b->n_fast += st.class_n[c];
		}
		b->n_warp = st.class_n[KSW_CLASS_WARP]; b->qmax_warp = st.class_qmax[KSW_CLASS_WARP];
		b->n_generic = st.class_n[KSW_CLASS_THREAD];
		b->qmax_generic = st.class_qmax[KSW_CLASS_THREAD];
		b->pool_bytes = (size_t)st.units * 16; b->npool_bytes = (size_t)st.nmask_words * 4;
		b->P = P;
		b->dev_ranges = true;
		CU(b->d_pool.reserve(std::max<size_t>(b->pool_bytes, 16)));
		CU(b->d_npool.reserve(std::max<size_t>(b->npool_bytes, 16)));
		CU(b->d_res.reserve(sizeof(DevRes) * n1));
		CU(b->d_cells.reserve(sizeof(uint32_t) * n1));
		CU(b->d_order.reserve(sizeof(uint32_t) * n1));
		CU(b->d_keys.reserve(sizeof(uint16_t) * 2 * n1));
		CU(b->d_vals.reserve(sizeof(uint32_t) * n1));
		CU(b->d_sort_tmp.reserve(std::max<size_t>(ksw_bin_temp_bytes(nc), 16)));
		CU(b->d_range.reserve(sizeof(uint32_t) * (KSW_N_CLASSES + 1)));
		CU(ksw_launch_pack_ref(s.d_rawjobs.p, nc, (const uint8_t *)ctx->d_qraw.p, (const uint8_t *)ctx->ref->dev, ctx->ref->l_pac,
		                       (const uint32_t *)s.d_offs.p, (DevJob *)b->d_jobs.p, (uint32_t *)b->d_pool.p, (uint32_t *)b->d_npool.p,
		                       (DevPackStats *)s.d_stats.p, s.stream));
		CU(ksw_launch_bin((const DevJob *)b->d_jobs.p, nc, (uint16_t *)b->d_keys.p, (uint16_t *)b->d_keys.p + n1,
		                  (uint32_t *)b->d_vals.p, (uint32_t *)b->d_order.p, b->d_sort_tmp.p, b->d_sort_tmp.cap, s.stream));
		CU(ksw_launch_ranges((const uint16_t *)b->d_keys.p + n1, nc, (uint32_t *)b->d_range.p, s.stream));
		ctx->launches += 5;
		int rc = enqueue_kernels(ctx, s, b);
		if (rc) return rc;
		CU(cudaMemcpyAsync(s.h_res.p, b->d_res.p, sizeof(DevRes) * n1, cudaMemcpyDeviceToHost, s.stream));
		CU(cudaStreamSynchronize(s.stream));
		d2h += (int64_t)(sizeof(DevRes) * n1);
		{
			const DevRes *hr = (const DevRes *)s.h_res.p;
			int64_t k = 0;
			for (const Piece &pc : round) {
				memcpy(segs[pc.seg].res + pc.first, hr + k, sizeof(DevRes) * (size_t)pc.n);
				k += pc.n;
			}
		}
	}
	ctx->last_h2d = h2d;
	ctx->last_d2h = d2h;
	return 0;
}

int ksw_b200_extend_batch_ref(ksw_b200_ctx_t *ctx, const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_rjob_t *jobs,
                              const uint8_t *qpool, size_t qbytes, ksw_b200_res_t *res)
{
	if (!ctx || !cfg || n < 0) return 1;
	if (n == 0) return 0;
	ksw_b200_rseg_t seg;
	seg.n = n; seg.jobs = jobs; seg.qpool = qpool; seg.qpool_bytes = qbytes; seg.res = res;
	return ksw_b200_extend_batch_ref_segs(ctx, cfg, 1, &seg);
}

// Multi-GPU form: jobs are independent, so the batch is cut into n_ctx contiguous ranges of (nearly)
// equal DP cells (sum of qlen*tlen); range r runs on ctxs[r] (one host thread per context / GPU, no exchange step) and
// writes its results straight into res.  Returns the first non-zero code of any shard.
int ksw_b200_extend_batch_multi(int n_ctx, ksw_b200_ctx_t **ctxs, const ksw_b200_cfg_t *cfg, int64_t n,
                                const ksw_b200_job_t *jobs, const uint8_t *qpool, const uint8_t *tpool, ksw_b200_res_t *res)
{
	if (n_ctx < 1 || !ctxs || !cfg || n < 0) return 1;
	if (n > 0 && (!jobs || !res)) return 1;
	for (int r = 0; r < n_ctx; ++r) if (!ctxs[r]) return 1;
	if (n_ctx == 1) return ksw_b200_extend_batch(ctxs[0], cfg, n, jobs, qpool, tpool, res);
	std::vector<int64_t> cut(n_ctx + 1, n);
	{
		double total = 0;
		for (int64_t k = 0; k < n; ++k) total += (double)jobs[k].qlen * (double)std::max(jobs[k].tlen, 1);
		double acc = 0;
		int r = 1;
		cut[0] = 0;
		for (int64_t k = 0; k < n && r < n_ctx; ++k) {
			acc += (double)jobs[k].qlen * (double)std::max(jobs[k].tlen, 1);
			while (r < n_ctx && acc >= total * r / n_ctx) cut[r++] = k + 1;
		}
	}
	std::vector<int> rc(n_ctx, 0);
	std::vector<std::thread> th;
	for (int r = 0; r < n_ctx; ++r)
		th.emplace_back([&, r] {
			const int64_t b = cut[r], e = cut[r + 1];
			if (e > b) rc[r] = ksw_b200_extend_batch(ctxs[r], cfg, e - b, jobs + b, qpool, tpool, res + b);
		});
	for (auto &t : th) t.join();
	for (int r = 0; r < n_ctx; ++r) if (rc[r]) return rc[r];
	return 0;
}

int ksw_b200_ctx_last_transfer(const ksw_b200_ctx_t *ctx, int64_t *h2d_bytes, int64_t *d2h_bytes)
{
	if (!ctx) return 1;
	if (h2d_bytes) *h2d_bytes = ctx->last_h2d;
	if (d2h_bytes) *d2h_bytes = ctx->last_d2h;
	return 0;
}

int ksw_b200_dpx_peak(ksw_b200_ctx_t *ctx, int which, double *lane_ops_per_s, float *ms_out)
{
	if (!ctx || !lane_ops_per_s) return 1;
	CU(cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->slot[0].stream;
	const int blocks = ctx->sm_count * 8, iters = 4096;
	unsigned *d = nullptr;
	CU(cudaMalloc(&d, (size_t)blocks * 256 * 4));
	float best = 1e30f;
	for (int rep = 0; rep < 4; ++rep) {
		CU(cudaEventRecord(ctx->ev0, st));
		CU(ksw_launch_dpx_peak(which, d, blocks, iters, st));
		CU(cudaEventRecord(ctx->ev1, st));
		CU(cudaEventSynchronize(ctx->ev1));
		float ms = 0;
		CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
		if (rep > 0 && ms < best) best = ms;
		ctx->launches++;
	}
	cudaFree(d);
	const double ops = (double)blocks * 256.0 * iters * 32.0;   // 8 chains x 4 DPX instructions per iteration
	*lane_ops_per_s = ops / (best * 1e-3);
	if (ms_out) *ms_out = best;
	return 0;
}

int ksw_b200_clamp_w(int qlen, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus)
{
	return ksw_clamp_w(qlen, ksw_mat_max(mat), o_del, e_del, o_ins, e_ins, w, end_bonus);
}

// ---- banded global alignment with backtrace.  Synchronous, chunk by chunk on slot 0's stream: a chunk is bounded by its
// sequence bytes, its worst-case CIGAR pool and the memory the kernels need per job.  Jobs the s16x2 form can hold
// (ksw_gfast.cu: the band holds the end cell, values far inside int16) are sorted by target length and run in groups of 32
// per warp; the rest go to the int32 thread-per-job kernel (ksw_global.cu).  KSW_B200_GLOBAL_FAST=0: everything int32.
namespace {

inline int gfast_nqb(int qlen, int w) { return ksw_gfast_nqb(qlen, w); }

int gfast_cell_cost(const ksw_b200_cfg_t *cfg)
{
	const char *e = getenv("KSW_B200_GLOBAL_FAST");
	if (e && e[0] == '0') return 0;
	return ksw_gfast_cell_cost(cfg->mat, cfg->o_del, cfg->e_del, cfg->o_ins, cfg->e_ins);
}

inline bool gfast_eligible(const ksw_b200_cfg_t *cfg, int cell_cost, const ksw_b200_gjob_t &j)
{
	return ksw_gfast_eligible(cell_cost, cfg->o_del, cfg->e_del, cfg->o_ins, cfg->e_ins, j.qlen, j.tlen, j.w);
}

} // namespace

int ksw_b200_global_batch(ksw_b200_ctx_t *ctx, const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_gjob_t *jobs,
                          const uint8_t *qpool, const uint8_t *tpool, ksw_b200_gres_t *res,
                          const uint32_t **cigar_pool, int64_t *n_cigar_total)
{
	if (!ctx || !cfg || n < 0 || (n > 0 && (!jobs || !res))) return fail(ctx, 1, "ksw_b200_global_batch: bad argument");
	if (cfg->m != 5) return fail(ctx, 2, "ksw_b200: only m == 5 is supported (every reference caller passes 5)");
	CU(cudaSetDevice(ctx->device));
	Slot &s = ctx->slot[0];
	ctx->g_cigar.clear();
	KswParams P;
	ksw_params_from_cfg(cfg, P);
	const int64_t max_chunk_jobs = 1 << 18;
	const size_t max_seq = (size_t)256 << 20, max_ops = (size_t)128 << 20;
	// per-job device memory (H slab of the fast kernel / direction-matrix slab of the int32 kernel): at most 8 GiB, and never
	// more than a quarter of what the device has free right now (several contexts may share a GPU; what this context
	// already holds counts as available to it)
	size_t z_budget = (size_t)8 << 30;
	{
		size_t free_b = 0, total_b = 0;
		if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
			z_budget = std::min(z_budget, std::max<size_t>((free_b + ctx->g_dz.cap + ctx->g_dzfast.cap) / 4, (size_t)256 << 20));
	}
	for (GStage &st : ctx->g_st)
		for (cudaEvent_t *e : {&st.e_start, &st.e_up, &st.e_kstart, &st.e_kern, &st.e_done}) if (!*e) CU(cudaEventCreate(e));
	// whatever happens, no chunk stays in flight behind this call (its staging belongs to the context)
	struct Drain {
		ksw_b200_ctx *c; cudaStream_t a, b;
		~Drain() { if (c->g_st[0].busy || c->g_st[1].busy) { cudaStreamSynchronize(c->up_stream); cudaStreamSynchronize(a); cudaStreamSynchronize(b); c->g_st[0].busy = c->g_st[1].busy = false; } }
	} drain{ctx, s.stream, ctx->down_stream};
	const int cell_cost = gfast_cell_cost(cfg);
	const int NB = 4096, NQ = KSW_GFAST_MAX_QLEN / 4 + 4;                          // sort buckets: target length, band width in quads
	const uint32_t KEY_SLOW = 0xffffffffu;
	std::vector<uint32_t> &gkey = ctx->g_key, &tmp = ctx->g_tmp, &bucket = ctx->g_bucket;
	if ((int64_t)gkey.size() < std::min<int64_t>(n, max_chunk_jobs)) gkey.resize((size_t)std::min<int64_t>(n, max_chunk_jobs));
	double tr_plan = 0, tr_pack = 0, tr_sort = 0, tr_wait = 0, tr_out = 0, tr_up = 0, tr_kern = 0, tr_down = 0;   // KSW_B200_TRACE
	auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };

	// results of a finished chunk: CIGAR pool (its size is known only now) on the second stream, then the records
	auto finish = [&](GStage &st) -> int {
		double t0 = now();
		CU(cudaEventSynchronize(st.e_done));
		tr_wait += now() - t0; t0 = now();
		if (ctx->trace) {
			float a = 0, b = 0, c = 0;
			cudaEventElapsedTime(&a, st.e_start, st.e_up); cudaEventElapsedTime(&b, st.e_kstart, st.e_kern); cudaEventElapsedTime(&c, st.e_kern, st.e_done);
			tr_up += a; tr_kern += b; tr_down += c;
		}
		const unsigned long long used = *(const unsigned long long *)st.hused.p;
		const size_t base = ctx->g_cigar.size();
		if (used) {
			CU(st.hcig.reserve(sizeof(uint32_t) * (size_t)used));
			CU(cudaMemcpyAsync(st.hcig.p, st.dcig.p, sizeof(uint32_t) * (size_t)used, cudaMemcpyDeviceToHost, ctx->down_stream));
			CU(cudaStreamSynchronize(ctx->down_stream));
			ctx->g_cigar.insert(ctx->g_cigar.end(), (const uint32_t *)st.hcig.p, (const uint32_t *)st.hcig.p + used);
		}
		const DevGRes *hr = (const DevGRes *)st.hres.p;
		for (int64_t k = 0; k < st.m; ++k) {
			res[st.first + k].score = hr[k].score;
			res[st.first + k].n_cigar = hr[k].n_cigar;
			res[st.first + k].cigar_off = (int64_t)base + hr[k].cigar_off;
		}
		st.busy = false;
		tr_out += now() - t0;
		return 0;
	};

	int64_t first = 0;
	int cur = 0;
	while (first < n) {
		GStage &st = ctx->g_st[cur];
		if (st.busy) { const int rc = finish(st); if (rc) return rc; }
		double t0 = now();
		// chunk [first, last): as many jobs as the budgets allow (at least one).  gkey: KEY_SLOW = int32 kernel, else the
		// fast kernel's sort key (band width in quads, then target length; both descending)
		int64_t last = first, n_fast = 0;
		size_t seq_bytes = 0, ops = 0, zfast = 0;
		int qmax = 0, qmax_fast = 0;
		long long zmax = 0;
		while (last < n && last - first < max_chunk_jobs) {
			const ksw_b200_gjob_t &j = jobs[last];
			if (j.qlen < 0 || j.tlen < 0 || j.w < 0) return fail(ctx, 2, "ksw_b200_global_batch: job with qlen < 0, tlen < 0 or w < 0");
			const size_t sb = (size_t)j.qlen + (size_t)j.tlen, so = sb + 2;
			const bool fastj = gfast_eligible(cfg, cell_cost, j);
			const int nqb = fastj ? gfast_nqb(j.qlen, j.w) : 0;
			// the group slab is sized by the group's longest target and widest band: 15 % of slack on the job's own need
			const size_t zj = fastj ? (size_t)((double)j.tlen * nqb * 8.0 * 1.15) + 64 : 0;
			if (last > first && (seq_bytes + sb > max_seq || ops + so > max_ops || zfast + zj > z_budget / 2)) break;
			seq_bytes += sb; ops += so; zfast += zj;
			if (fastj) {
				++n_fast; qmax_fast = std::max(qmax_fast, j.qlen);
				gkey[(size_t)(last - first)] = (uint32_t)(NQ - 1 - std::min(nqb, NQ - 1)) << 12 | (uint32_t)(NB - 1 - std::min(j.tlen, NB - 1));
			} else {
				qmax = std::max(qmax, j.qlen);
				const long long n_col = std::min<long long>(j.qlen, 2LL * j.w + 1);
				zmax = std::max(zmax, n_col * j.tlen);
				gkey[(size_t)(last - first)] = KEY_SLOW;
			}
			++last;
		}
		const int64_t m = last - first, n_slow = m - n_fast;
		st.first = first; st.m = m;
		tr_plan += now() - t0; t0 = now();
		CU(st.hjobs.reserve(sizeof(DevGJob) * (size_t)(m + n_slow)));
		CU(st.hseq.reserve(std::max<size_t>(seq_bytes, 16)));
		CU(st.hres.reserve(sizeof(DevGRes) * (size_t)m));
		CU(st.hused.reserve(sizeof(unsigned long long)));
		CU(st.djobs.reserve(sizeof(DevGJob) * (size_t)(m + n_slow)));
		CU(st.dseq.reserve(std::max<size_t>(seq_bytes, 16)));
		CU(st.dres.reserve(sizeof(DevGRes) * (size_t)m));
		CU(st.dcig.reserve(sizeof(uint32_t) * ops));
		CU(st.dused.reserve(sizeof(unsigned long long)));
		DevGJob *hj = (DevGJob *)st.hjobs.p;
		uint8_t *hs = (uint8_t *)st.hseq.p;
		{
			// records (a running offset: serial, cheap), then the sequence bytes on the pack threads
			size_t off = 0;
			for (int64_t k = 0; k < m; ++k) {
				const ksw_b200_gjob_t &j = jobs[first + k];
				hj[k].seq_off = off; hj[k].qlen = j.qlen; hj[k].tlen = j.tlen; hj[k].w = j.w; hj[k].idx = (uint32_t)k;
				off += (size_t)j.qlen + (size_t)j.tlen;
			}
			KswPool *tp = pool_of(ctx);
			const int T = (int)std::max<int64_t>(1, std::min<int64_t>(tp->size(), m / 4096));
			const int64_t per = (m + T - 1) / T;
			auto body = [&](int t) {
				const int64_t b = std::min<int64_t>(m, t * per), e = std::min<int64_t>(m, b + per);
				for (int64_t k = b; k < e; ++k) {
					const ksw_b200_gjob_t &j = jobs[first + k];
					if (j.qlen) memcpy(hs + hj[k].seq_off, qpool + j.q_off, (size_t)j.qlen);
					if (j.tlen) memcpy(hs + hj[k].seq_off + j.qlen, tpool + j.t_off, (size_t)j.tlen);
				}
			};
			if (T == 1) body(0); else tp->run(T, body);
		}
		tr_pack += now() - t0; t0 = now();
		// the fast kernel's jobs: two counting-sort passes over the keys (least significant field first), groups of 32, one
		// slab per group: the 32 lanes of a warp walk their rows in lockstep, so a group costs (longest target) x (widest band)
		int n_groups = 0;
		size_t zfast_units = 0;
		if (n_fast > 0) {
			CU(st.horder.reserve(sizeof(uint32_t) * (size_t)n_fast));
			CU(st.hgroups.reserve(sizeof(DevGGroup) * (size_t)((n_fast + 31) / 32)));
			uint32_t *ho = (uint32_t *)st.horder.p;
			DevGGroup *hg = (DevGGroup *)st.hgroups.p;
			tmp.resize((size_t)n_fast);
			bucket.assign(NB + 1, 0);
			for (int64_t k = 0; k < m; ++k) if (gkey[k] != KEY_SLOW) ++bucket[(gkey[k] & 0xfffu) + 1];
			for (int x = 0; x < NB; ++x) bucket[x + 1] += bucket[x];
			for (int64_t k = 0; k < m; ++k) if (gkey[k] != KEY_SLOW) tmp[bucket[gkey[k] & 0xfffu]++] = (uint32_t)k;
			bucket.assign(NQ + 1, 0);
			for (int64_t x = 0; x < n_fast; ++x) ++bucket[(gkey[tmp[x]] >> 12) + 1];
			for (int x = 0; x < NQ; ++x) bucket[x + 1] += bucket[x];
			for (int64_t x = 0; x < n_fast; ++x) ho[bucket[gkey[tmp[x]] >> 12]++] = tmp[x];
			for (int64_t f = 0; f < n_fast; f += 32) {
				DevGGroup g;
				g.first = (int32_t)f; g.n = (int32_t)std::min<int64_t>(32, n_fast - f); g.rows = 0; g.nqb = 0;
				for (int x = 0; x < g.n; ++x) {
					const DevGJob &j = hj[ho[f + x]];
					g.rows = std::max(g.rows, j.tlen);
					g.nqb = std::max(g.nqb, gfast_nqb(j.qlen, j.w));
				}
				g.z_off = (long long)zfast_units;
				zfast_units += (size_t)g.rows * (size_t)g.nqb * 32;
				hg[n_groups++] = g;
			}
			CU(st.dorder.reserve(sizeof(uint32_t) * (size_t)n_fast));
			CU(st.dgroups.reserve(sizeof(DevGGroup) * (size_t)n_groups));
			CU(ctx->g_dzfast.reserve(zfast_units * sizeof(uint2) + 256));
			CU(ctx->g_dcounter.reserve(sizeof(unsigned)));
			CU(ctx->g_dscratch.reserve(sizeof(uint32_t) * ops));
		}
		// the int32 kernel's jobs: a compacted copy of their records behind the chunk's
		if (n_slow > 0) {
			int64_t w2 = m;
			for (int64_t k = 0; k < m; ++k) if (gkey[k] == KEY_SLOW) hj[w2++] = hj[k];
		}
		tr_sort += now() - t0;
		// uploads on their own stream: they overlap the previous chunk's kernels (this stage's buffers are free: its last
		// chunk was finished above)
		cudaStream_t up = ctx->up_stream;
		CU(cudaEventRecord(st.e_start, up));
		CU(cudaMemcpyAsync(st.djobs.p, hj, sizeof(DevGJob) * (size_t)(m + n_slow), cudaMemcpyHostToDevice, up));
		if (seq_bytes) CU(cudaMemcpyAsync(st.dseq.p, hs, seq_bytes, cudaMemcpyHostToDevice, up));
		CU(cudaMemsetAsync(st.dused.p, 0, sizeof(unsigned long long), up));
		if (n_fast > 0) {
			CU(cudaMemcpyAsync(st.dorder.p, st.horder.p, sizeof(uint32_t) * (size_t)n_fast, cudaMemcpyHostToDevice, up));
			CU(cudaMemcpyAsync(st.dgroups.p, st.hgroups.p, sizeof(DevGGroup) * (size_t)n_groups, cudaMemcpyHostToDevice, up));
		}
		CU(cudaEventRecord(st.e_up, up));
		CU(cudaStreamWaitEvent(s.stream, st.e_up, 0));
		CU(cudaEventRecord(st.e_kstart, s.stream));
		if (n_fast > 0) {
			CU(ksw_launch_gfast((const DevGJob *)st.djobs.p, (const uint8_t *)st.dseq.p, P, (const uint32_t *)st.dorder.p,
			                    (const DevGGroup *)st.dgroups.p, n_groups, qmax_fast, ctx->sm_count, (uint2 *)ctx->g_dzfast.p,
			                    (unsigned *)ctx->g_dcounter.p, (uint32_t *)ctx->g_dscratch.p, (unsigned long long *)st.dused.p, (uint32_t *)st.dcig.p,
			                    (DevGRes *)st.dres.p, s.stream));
			ctx->launches += 2;
		}
		if (n_slow > 0) {
			const long long zcap = zmax + qmax + 2;
			// threads: bounded by the direction-matrix slab; every thread owns one column set and one z slab
			int bps = 4;                                        // blocks per SM: 4 keeps the H/E slabs of a 150 bp batch inside L2
			if (const char *ev = getenv("KSW_B200_GLOBAL_BPS")) bps = std::max(1, atoi(ev));     // tuning knob
			size_t threads = (size_t)ctx->sm_count * bps * KSW_GENERIC_THREADS;
			while (threads > KSW_GENERIC_THREADS && threads * (size_t)zcap > z_budget / 2) threads >>= 1;
			threads = std::min<size_t>(threads, (size_t)((n_slow + KSW_GENERIC_THREADS - 1) / KSW_GENERIC_THREADS) * KSW_GENERIC_THREADS);
			const int n_blocks = (int)(threads / KSW_GENERIC_THREADS);
			CU(ctx->g_deh.reserve(threads * (size_t)(qmax + 1) * sizeof(int2)));
			CU(ctx->g_dqc.reserve(threads * (size_t)(qmax + 1)));
			CU(ctx->g_dz.reserve(threads * (size_t)zcap));
			CU(ksw_launch_global((const DevGJob *)st.djobs.p + m, n_slow, (const uint8_t *)st.dseq.p, P, (int2 *)ctx->g_deh.p,
			                     (uint8_t *)ctx->g_dqc.p, (uint8_t *)ctx->g_dz.p, zcap, n_blocks,
			                     (unsigned long long *)st.dused.p, (uint32_t *)st.dcig.p, (DevGRes *)st.dres.p, s.stream));
			ctx->launches++;
		}
		CU(cudaEventRecord(st.e_kern, s.stream));
		CU(cudaMemcpyAsync(st.hres.p, st.dres.p, sizeof(DevGRes) * (size_t)m, cudaMemcpyDeviceToHost, s.stream));
		CU(cudaMemcpyAsync(st.hused.p, st.dused.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s.stream));
		CU(cudaEventRecord(st.e_done, s.stream));
		st.busy = true;
		cur ^= 1;
		first = last;
	}
	for (int k = 0; k < 2; ++k, cur ^= 1)                                             // the older chunk first
		if (ctx->g_st[cur].busy) { const int rc = finish(ctx->g_st[cur]); if (rc) return rc; }
	if (ctx->trace)
		fprintf(stderr, "[ksw_b200] global_batch n=%lld: host plan %.2f  pack %.2f  sort %.2f  wait for GPU %.2f  results out %.2f ms | "
		                "GPU h2d %.2f  kernels %.2f  d2h %.2f ms\n",
		        (long long)n, tr_plan, tr_pack, tr_sort, tr_wait, tr_out, tr_up, tr_kern, tr_down);
	if (cigar_pool) *cigar_pool = ctx->g_cigar.data();
	if (n_cigar_total) *n_cigar_total = (int64_t)ctx->g_cigar.size();
	return 0;
}

// ---- local alignment with start positions and second-best score (ksw_align.cu): the reference's ksw_align2.  Synchronous,
// chunk by chunk on slot 0's stream.  Within a chunk the byte-kernel jobs (KSW_XBYTE) and the 16-bit jobs are each sorted
// by (query length, target length): the 2 resp. 4 jobs that share a warp walk their rows in lockstep.
int ksw_b200_align_batch(ksw_b200_ctx_t *ctx, const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_ajob_t *jobs,
                         const uint8_t *qpool, const uint8_t *tpool, ksw_b200_ares_t *res)
{
	if (!ctx || !cfg || n < 0 || (n > 0 && (!jobs || !res || !qpool || !tpool))) return fail(ctx, 1, "ksw_b200_align_batch: bad argument");
	if (cfg->m != 5) return fail(ctx, 2, "ksw_b200: only m == 5 is supported (every reference caller passes 5)");
	CU(cudaSetDevice(ctx->device));
	Slot &s = ctx->slot[0];
	KswAlignParams A;
	memset(&A, 0, sizeof(A));
	memcpy(A.mat, cfg->mat, 25);
	A.o_del = cfg->o_del; A.e_del = cfg->e_del; A.o_ins = cfg->o_ins; A.e_ins = cfg->e_ins;
	{
		// ksw_qinit (ksw.c:77-83): shift = 256 - (most negative entry, at most 127) as a byte, max = largest entry (at least 0)
		uint8_t shift = 127, mdiff = 0;
		for (int a = 0; a < 25; ++a) {
			if (cfg->mat[a] < (int8_t)shift) shift = (uint8_t)cfg->mat[a];
			if (cfg->mat[a] > (int8_t)mdiff) mdiff = (uint8_t)cfg->mat[a];
		}
		A.qmax = mdiff;
		A.shift = (uint8_t)(256 - shift);
		if (A.qmax <= 0) return fail(ctx, 2, "ksw_b200_align_batch: the scoring matrix has no positive entry (the reference divides by it, ksw.c:224)");
	}
	const int64_t max_chunk_jobs = 1 << 20;
	const size_t max_seq = (size_t)256 << 20;
	std::vector<uint32_t> &tmp = ctx->g_tmp, &bucket = ctx->g_bucket;
	int64_t first = 0;
	while (first < n) {
		int64_t last = first;
		size_t seq_bytes = 0;
		int64_t n_byte = 0;
		int qmax[2] = {0, 0}, tmax[2] = {0, 0};
		while (last < n && last - first < max_chunk_jobs) {
			const ksw_b200_ajob_t &j = jobs[last];
			if (j.qlen < 1 || j.tlen < 0) return fail(ctx, 2, "ksw_b200_align_batch: job with qlen < 1 or tlen < 0");
			if (j.qlen > KSW_ALIGN_MAX_QLEN) return fail(ctx, 2, "ksw_b200_align_batch: query longer than 4096 (not supported by the mate-rescue kernel)");
			const size_t sb = (size_t)j.qlen + (size_t)j.tlen;
			if (last > first && seq_bytes + sb > max_seq) break;
			seq_bytes += sb;
			const int c = (j.xtra & 0x10000) ? 0 : 1;
			n_byte += c == 0;
			qmax[c] = std::max(qmax[c], j.qlen); tmax[c] = std::max(tmax[c], j.tlen);
			++last;
		}
		const int64_t m = last - first;
		CU(ctx->a_hjobs.reserve(sizeof(DevAJob) * (size_t)m));
		CU(ctx->a_hseq.reserve(seq_bytes + 16));
		CU(ctx->a_hres.reserve(sizeof(DevARes) * (size_t)m));
		CU(ctx->a_horder.reserve(sizeof(uint32_t) * (size_t)m));
		CU(ctx->a_djobs.reserve(sizeof(DevAJob) * (size_t)m));
		CU(ctx->a_dseq.reserve(seq_bytes + 16));
		CU(ctx->a_dres.reserve(sizeof(DevARes) * (size_t)m));
		CU(ctx->a_dorder.reserve(sizeof(uint32_t) * (size_t)m));
		CU(ctx->a_dcounter.reserve(sizeof(unsigned)));
		DevAJob *hj = (DevAJob *)ctx->a_hjobs.p;
		uint8_t *hs = (uint8_t *)ctx->a_hseq.p;
		uint32_t *ho = (uint32_t *)ctx->a_horder.p;
		{
			size_t off = 0;
			for (int64_t k = 0; k < m; ++k) {
				const ksw_b200_ajob_t &j = jobs[first + k];
				hj[k].seq_off = off; hj[k].qlen = j.qlen; hj[k].tlen = j.tlen; hj[k].xtra = j.xtra; hj[k].idx = (uint32_t)k;
				off += (size_t)j.qlen + (size_t)j.tlen;
			}
			KswPool *tp = pool_of(ctx);
			const int T = (int)std::max<int64_t>(1, std::min<int64_t>(tp->size(), m / 1024));
			const int64_t per = (m + T - 1) / T;
			auto body = [&](int t) {
				const int64_t b = std::min<int64_t>(m, t * per), e = std::min<int64_t>(m, b + per);
				for (int64_t k = b; k < e; ++k) {
					const ksw_b200_ajob_t &j = jobs[first + k];
					memcpy(hs + hj[k].seq_off, qpool + j.q_off, (size_t)j.qlen);
					if (j.tlen) memcpy(hs + hj[k].seq_off + j.qlen, tpool + j.t_off, (size_t)j.tlen);
				}
			};
			if (T == 1) body(0); else tp->run(T, body);
		}
		{
			// order: byte-kernel jobs first, then the 16-bit ones; inside each by query length, then target length (descending):
			// three stable counting-sort passes, least significant key first
			const int NB = 4097;
			tmp.resize((size_t)m);
			auto pass = [&](const uint32_t *in, uint32_t *out, int nb, auto key) {
				bucket.assign((size_t)nb + 1, 0);
				for (int64_t x = 0; x < m; ++x) ++bucket[key(hj[in ? in[x] : (uint32_t)x]) + 1];
				for (int x = 0; x < nb; ++x) bucket[x + 1] += bucket[x];
				for (int64_t x = 0; x < m; ++x) { const uint32_t k = in ? in[x] : (uint32_t)x; out[bucket[key(hj[k])]++] = k; }
			};
			pass(nullptr, ho, NB, [&](const DevAJob &j) { return NB - 1 - std::min(j.tlen, NB - 1); });
			pass(ho, tmp.data(), NB, [&](const DevAJob &j) { return NB - 1 - std::min(j.qlen, NB - 1); });
			pass(tmp.data(), ho, 2, [&](const DevAJob &j) { return (j.xtra & 0x10000) ? 0 : 1; });
		}
		CU(cudaMemcpyAsync(ctx->a_djobs.p, hj, sizeof(DevAJob) * (size_t)m, cudaMemcpyHostToDevice, s.stream));
		CU(cudaMemcpyAsync(ctx->a_dseq.p, hs, seq_bytes, cudaMemcpyHostToDevice, s.stream));
		CU(cudaMemcpyAsync(ctx->a_dorder.p, ho, sizeof(uint32_t) * (size_t)m, cudaMemcpyHostToDevice, s.stream));
		for (int c = 0; c < 2; ++c) {
			const int64_t cnt = c == 0 ? n_byte : m - n_byte;
			if (cnt <= 0) continue;
			CU(ksw_launch_align(c == 0 ? 1 : 2, (const DevAJob *)ctx->a_djobs.p, (const uint32_t *)ctx->a_dorder.p + (c == 0 ? 0 : n_byte),
			                    (int)cnt, (const uint8_t *)ctx->a_dseq.p, A, qmax[c], tmax[c], ctx->sm_count, &ctx->a_bscr, &ctx->a_bscr_cap,
			                    (unsigned *)ctx->a_dcounter.p, (DevARes *)ctx->a_dres.p, s.stream));
			ctx->launches++;
		}
		CU(cudaMemcpyAsync(ctx->a_hres.p, ctx->a_dres.p, sizeof(DevARes) * (size_t)m, cudaMemcpyDeviceToHost, s.stream));
		CU(cudaStreamSynchronize(s.stream));
		const DevARes *hr = (const DevARes *)ctx->a_hres.p;
		for (int64_t k = 0; k < m; ++k) {
			ksw_b200_ares_t &o = res[first + k];
			o.score = hr[k].score; o.te = hr[k].te; o.qe = hr[k].qe; o.score2 = hr[k].score2; o.te2 = hr[k].te2; o.tb = hr[k].tb; o.qb = hr[k].qb;
			o.reserved = 0;
		}
		first = last;
	}
	return 0;
}

// ---- scalar drop-ins (ksw.h:107-108).  One lazily created context per host thread.
static ksw_b200_ctx *scalar_ctx()
{
	static thread_local struct Holder {
		ksw_b200_ctx *c = nullptr;
		~Holder() { if (c) ksw_b200_ctx_destroy(c); }
	} h;
	if (!h.c) {
		int dev = 0;
		if (const char *s = getenv("KSW_B200_DEVICE")) dev = atoi(s);
		if (ksw_b200_ctx_create(dev, &h.c) != 0) {
			fprintf(stderr, "[ksw_b200] fatal: no usable CUDA device for ksw_extend (there is no CPU fallback)\n");
			abort();
		}
		h.c->pack_threads = 1;
	}
	return h.c;
}

int ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                int *qle, int *tle, int *gtle, int *gscore, int *max_off)
{
	ksw_b200_ctx *ctx = scalar_ctx();
	ksw_b200_cfg_t cfg;
	memcpy(cfg.mat, mat, 25);
	cfg.m = m; cfg.o_del = o_del; cfg.e_del = e_del; cfg.o_ins = o_ins; cfg.e_ins = e_ins;
	cfg.zdrop = zdrop; cfg.end_bonus = end_bonus;
	ksw_b200_job_t job;
	job.q_off = 0; job.t_off = 0; job.qlen = qlen; job.tlen = tlen; job.h0 = h0; job.w = w;
	ksw_b200_res_t r;
	int rc = ksw_b200_extend_batch(ctx, &cfg, 1, &job, query, target, &r);
	if (rc) {
		fprintf(stderr, "[ksw_b200] fatal: ksw_extend2 failed on the GPU (%d): %s\n", rc, ksw_b200_strerror(ctx));
		abort();
	}
	if (qle) *qle = r.qle;
	if (tle) *tle = r.tle;
	if (gtle) *gtle = r.gtle;
	if (gscore) *gscore = r.gscore;
	if (max_off) *max_off = r.max_off;
	return r.score;
}

int ksw_extend(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
               int gapo, int gape, int w, int end_bonus, int zdrop, int h0,
               int *qle, int *tle, int *gtle, int *gscore, int *max_off)
{
	return ksw_extend2(qlen, query, tlen, target, m, mat, gapo, gape, gapo, gape, w, end_bonus, zdrop, h0,
	                   qle, tle, gtle, gscore, max_off);
}

// ksw.h:83-84.  As in the reference the CIGAR is only produced when both out-pointers are given, *n_cigar_ is zeroed
// first (ksw.c:507), and the array is malloc'd for the caller.
int ksw_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                int o_del, int e_del, int o_ins, int e_ins, int w, int *n_cigar_, uint32_t **cigar_)
{
	ksw_b200_ctx *ctx = scalar_ctx();
	ksw_b200_cfg_t cfg;
	memcpy(cfg.mat, mat, 25);
	cfg.m = m; cfg.o_del = o_del; cfg.e_del = e_del; cfg.o_ins = o_ins; cfg.e_ins = e_ins;
	cfg.zdrop = 0; cfg.end_bonus = 0;
	ksw_b200_gjob_t job;
	job.q_off = 0; job.t_off = 0; job.qlen = qlen; job.tlen = tlen; job.w = w; job.reserved = 0;
	ksw_b200_gres_t r;
	const uint32_t *pool = nullptr;
	int64_t total = 0;
	if (n_cigar_) *n_cigar_ = 0;
	int rc = ksw_b200_global_batch(ctx, &cfg, 1, &job, query, target, &r, &pool, &total);
	if (rc) {
		fprintf(stderr, "[ksw_b200] fatal: ksw_global2 failed on the GPU (%d): %s\n", rc, ksw_b200_strerror(ctx));
		abort();
	}
	if (n_cigar_ && cigar_) {
		uint32_t *c = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)std::max(r.n_cigar, 1));
		if (!c) { fprintf(stderr, "[ksw_b200] fatal: out of memory\n"); abort(); }
		for (int k = 0; k < r.n_cigar; ++k) c[k] = pool[r.cigar_off + k];
		*n_cigar_ = r.n_cigar; *cigar_ = c;
	}
	return r.score;
}

int ksw_global(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
               int gapo, int gape, int w, int *n_cigar_, uint32_t **cigar_)
{
	return ksw_global2(qlen, query, tlen, target, m, mat, gapo, gape, gapo, gape, w, n_cigar_, cigar_);
}

// ksw.h:62-63.  The reference's kswr_t by value; qry (its cached query profile) is not used.
ksw_b200_kswr_t ksw_align2(int qlen, uint8_t *query, int tlen, uint8_t *target, int m, const int8_t *mat,
                           int o_del, int e_del, int o_ins, int e_ins, int xtra, void **qry)
{
	(void)qry;
	ksw_b200_ctx *ctx = scalar_ctx();
	ksw_b200_cfg_t cfg;
	memcpy(cfg.mat, mat, 25);
	cfg.m = m; cfg.o_del = o_del; cfg.e_del = e_del; cfg.o_ins = o_ins; cfg.e_ins = e_ins; cfg.zdrop = 0; cfg.end_bonus = 0;
	ksw_b200_ajob_t job;
	job.q_off = 0; job.t_off = 0; job.qlen = qlen; job.tlen = tlen; job.xtra = xtra; job.reserved = 0;
	ksw_b200_ares_t r;
	const uint8_t none = 0;
	const int rc = ksw_b200_align_batch(ctx, &cfg, 1, &job, query, tlen > 0 ? target : &none, &r);
	if (rc) {
		fprintf(stderr, "[ksw_b200] fatal: ksw_align2 failed on the GPU (%d): %s\n", rc, ksw_b200_strerror(ctx));
		abort();
	}
	ksw_b200_kswr_t o;
	o.score = r.score; o.te = r.te; o.qe = r.qe; o.score2 = r.score2; o.te2 = r.te2; o.tb = r.tb; o.qb = r.qb;
	return o;
}

ksw_b200_kswr_t ksw_align(int qlen, uint8_t *query, int tlen, uint8_t *target, int m, const int8_t *mat,
                          int gapo, int gape, int xtra, void **qry)
{
	return ksw_align2(qlen, query, tlen, target, m, mat, gapo, gape, gapo, gape, xtra, qry);
}

} // extern "C"
