// ksw_pack.h — host packer: caller-order byte-coded jobs -> binned DevJob[] + 2-bit pool + N side pool.
// Plain C++ (no CUDA calls) so that the same packer feeds the GPU runtime and the CPU emulation used
// by the tests.  Two phases so the caller can size its (pinned) buffers in between.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>
#include <type_traits>
#include "../../include/ksw_b200.h"
#include "ksw_dev.cuh"

// Persistent fork-join worker pool (thread creation per call costs more than packing a small chunk).
class KswPool {
public:
	explicit KswPool(int n_threads);
	~KswPool();
	int size() const { return n_; }
	// runs fn(t) for t in [0, n_tasks) (n_tasks <= size()), task 0 on the calling thread; returns when all are done
	void run(int n_tasks, void (*fn)(void *, int), void *arg);
	template <class F> void run(int n_tasks, F &&f)
	{
		run(n_tasks, [](void *a, int t) { (*static_cast<typename std::remove_reference<F>::type *>(a))(t); }, (void *)&f);
	}
private:
	struct Impl;
	Impl *impl_;
	int n_;
};

#include "ksw_class.h"               // job classes, band clamp: shared with the device packer
static const int KSW_FAST_CLASS_QMAX[KSW_FAST_CLASSES] = {124, 128, 256, 512};

// What the launcher needs to know about a packed batch.  Classes 0..KSW_FAST_CLASSES-1 are the fast kernel's,
// KSW_CLASS_WARP the warp-cooperative int32 kernel's, KSW_CLASS_THREAD the thread-per-job generic kernel's.
struct KswPackStats {
	int64_t n = 0;
	int64_t class_n[KSW_N_CLASSES] = {0, 0, 0, 0, 0, 0};
	int class_qmax[KSW_N_CLASSES] = {0, 0, 0, 0, 0, 0};
	size_t pool_bytes = 0;                 // bytes of the 2-bit pool (multiple of 16)
	std::vector<uint64_t> range_base;      // scratch: pool offset (16-byte units) at which each host thread's range starts
};

// Pass 1 (lengths only): classes, per-class counts / longest query, pool size.  fast_qmax: largest qlen the fast kernel
// accepts (0 = fast kernel disabled: everything is generic).
int ksw_pack_sizes(const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_job_t *jobs, int fast_qmax, KswPool *tp,
                   KswPackStats &st, std::string &err);

// Pass 2: streams the caller's byte-coded sequences once, in the caller's order: dj[k] describes jobs[k] (idx = k), its
// sequences are 2-bit packed at consecutive pool offsets.  Binning (the order in which the kernels take the jobs) is
// done on the device.  N masks of the rare jobs that have them are appended to nmask.  Class-0 jobs whose query holds
// an N are moved to class 1 here (st.class_n / st.class_qmax are adjusted; class_qmax[0] stays an upper bound).
int ksw_pack_stream(KswPackStats &st, const ksw_b200_cfg_t *cfg, const ksw_b200_job_t *jobs, int fast_qmax,
                    const uint8_t *qpool, const uint8_t *tpool, DevJob *dj, uint32_t *pool,
                    std::vector<uint32_t> &nmask, KswPool *tp);

// A/B and test switch: non-zero keeps the packer on its word-at-a-time path where the 64-byte SIMD path would run
// (also KSW_B200_PACK_WORDS=1 in the environment at load time); the two paths write identical bytes
extern int ksw_pack_force_words;

void ksw_params_from_cfg(const ksw_b200_cfg_t *cfg, KswParams &P);
void ksw_scoring_from_cfg(const ksw_b200_cfg_t *cfg, int fast_qmax, KswScoring &S);

// the reference's band clamp (ksw.c:398-406), evaluated with the identical C expression
int ksw_clamp_w(int qlen, int maxsc, int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus);
int ksw_mat_max(const int8_t *mat);
