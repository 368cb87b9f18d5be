// ksw_pack.cpp — see ksw_pack.h.  Host C++ only.
#include "ksw_pack.h"
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>

namespace {

// splits [0,n) into contiguous ranges, one per pool thread (serial when pool is null or n is small)
template <class F>
void parallel_ranges(KswPool *pool, int64_t n, F &&fn)
{
	if (n <= 0) return;
	int T = pool ? pool->size() : 1;
	T = (int)std::max<int64_t>(1, std::min<int64_t>(T, (n + 16383) / 16384));
	const int64_t per = (n + T - 1) / T;
	auto body = [&](int t) {
		const int64_t b = std::min<int64_t>(n, t * per), e = std::min<int64_t>(n, b + per);
		if (b < e) fn(t, b, e);
	};
	if (T == 1) body(0); else pool->run(T, body);
}

#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("bmi2"))) inline uint32_t squeeze16_pext(uint64_t a, uint64_t b)
{
	return (uint32_t)__builtin_ia32_pext_di(a, 0x0303030303030303ull) |
	       ((uint32_t)__builtin_ia32_pext_di(b, 0x0303030303030303ull) << 16);
}
const bool g_has_bmi2 = __builtin_cpu_supports("bmi2");
#define KSW_HAVE_PEXT 1
#endif

// gather the low 2 bits of each of 16 byte codes into one word (base k at bits 2k)
inline uint32_t squeeze16(uint64_t a, uint64_t b)
{
#ifdef KSW_HAVE_PEXT
	if (g_has_bmi2) return squeeze16_pext(a, b);
#endif
	auto squeeze = [](uint64_t v) -> uint32_t {
		v = (v | (v >> 6)) & 0x000F000F000F000Full;
		v = (v | (v >> 12)) & 0x000000FF000000FFull;
		v = (v | (v >> 24)) & 0xFFFFull;
		return (uint32_t)v;
	};
	return squeeze(a) | (squeeze(b) << 16);
}

// 2-bit packing of `len` byte codes into ceil(len/16) words.  Returns true if a code > 3 (N) was seen; those bases
// are stored as 0 and flagged in nmask (ceil(len/32) words), which is zeroed here on the first N and otherwise untouched.
inline bool pack2(const uint8_t *s, int len, uint32_t *out, uint32_t *nmask)
{
	bool has_n = false;
	const int full = len >> 4;
	auto slow_word = [&](const uint8_t *p, int k0, int lim) -> uint32_t {
		uint32_t word = 0;
		for (int x = 0; x < lim; ++x) {
			uint32_t c = p[x];
			if (c > 3) {
				if (!has_n) { has_n = true; std::fill_n(nmask, (len + 31) >> 5, 0u); }
				nmask[(k0 + x) >> 5] |= 1u << ((k0 + x) & 31);
				c = 0;
			}
			word |= c << (2 * x);
		}
		return word;
	};
	for (int wi = 0; wi < full; ++wi) {
		uint64_t a, b;
		memcpy(&a, s + 16 * wi, 8); memcpy(&b, s + 16 * wi + 8, 8);
		out[wi] = ((a | b) & 0xFCFCFCFCFCFCFCFCull) ? slow_word(s + 16 * wi, 16 * wi, 16) : squeeze16(a, b);
	}
	const int rem = len & 15;
	if (rem) {
		uint8_t tail[16] = {0};
		memcpy(tail, s + 16 * full, (size_t)rem);
		uint64_t a, b;
		memcpy(&a, tail, 8); memcpy(&b, tail + 8, 8);
		out[full] = ((a | b) & 0xFCFCFCFCFCFCFCFCull) ? slow_word(s + 16 * full, 16 * full, rem) : squeeze16(a, b);
	}
	return has_n;
}

// A job may take the fast s16x2 kernel iff every value its DP can hold stays far inside int16
// and its columns fit the kernel's shared-memory budget; everything else goes to the generic
// int32 kernel (still on the GPU).
bool fast_eligible(const ksw_b200_cfg_t *cfg, int fast_qmax, int maxsc, int minsc, int qlen, int h0)
{
	if (qlen > fast_qmax) return false;
	if (cfg->o_ins < 0 || cfg->o_del < 0 || cfg->e_ins < 1 || cfg->e_del < 1) return false;
	if (cfg->o_ins + cfg->e_ins > 4000 || cfg->o_del + cfg->e_del > 4000) return false;
	if (minsc < -120 || maxsc > 120) return false;
	if ((int64_t)h0 + (int64_t)qlen * maxsc > 20000) return false;
	return true;
}

} // namespace

int ksw_mat_max(const int8_t *mat)
{
	int mx = 0;                                   // the reference's scan starts from 0 (ksw.c:399)
	for (int i = 0; i < 25; ++i) mx = mx > mat[i] ? mx : mat[i];
	return mx;
}

int ksw_clamp_w(int qlen, int maxsc, int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus)
{
	int max_ins = (int)((double)(qlen * maxsc + end_bonus - o_ins) / e_ins + 1.);
	max_ins = max_ins > 1 ? max_ins : 1;
	w = w < max_ins ? w : max_ins;
	int max_del = (int)((double)(qlen * maxsc + end_bonus - o_del) / e_del + 1.);
	max_del = max_del > 1 ? max_del : 1;
	w = w < max_del ? w : max_del;
	return w;
}

void ksw_params_from_cfg(const ksw_b200_cfg_t *cfg, KswParams &P)
{
	memset(&P, 0, sizeof(P));
	memcpy(P.mat, cfg->mat, 25);
	P.o_del = cfg->o_del; P.e_del = cfg->e_del; P.o_ins = cfg->o_ins; P.e_ins = cfg->e_ins;
	P.zdrop = cfg->zdrop;
}

int ksw_pack_plan(const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_job_t *jobs, int fast_qmax,
                  KswPool *pool, KswPackPlan &plan, std::string &err)
{
	if (cfg->m != 5) { err = "ksw_b200: only m == 5 is supported (every reference caller passes 5)"; return 2; }
	if (n > 0x7fffffffLL) { err = "ksw_b200: more than 2^31-1 jobs in one batch"; return 2; }
	const int maxsc = ksw_mat_max(cfg->mat);
	int minsc = 0;
	for (int i = 0; i < 25; ++i) minsc = std::min<int>(minsc, cfg->mat[i]);
	plan.n = n; plan.maxsc = maxsc;
	if ((int64_t)plan.pos_of.size() < n) { plan.pos_of.resize(n); plan.off_of.resize(n); plan.units_of.resize(n); plan.key.resize(n); }

	// Bin key from the lengths alone (no sequence bytes are touched in this phase):
	// generic bit | query-length class | rows (coarse) | carried-in score — the things that drive the band
	// a job sweeps.  Keys are inverted so that long jobs come first (short tail at the end of a launch).
	// Parallel counting sort: per-thread histograms over contiguous caller ranges, then a stable scatter.
	constexpr int NKEY = 1 << 16;
	const int T = (int)std::max<int64_t>(1, std::min<int64_t>(pool ? pool->size() : 1, (n + 65535) / 65536));
	const int64_t per = (n + T - 1) / T;
	std::vector<uint16_t> &key = plan.key;
	std::vector<std::vector<uint32_t>> &cnt = plan.cnt;
	std::vector<std::vector<uint64_t>> &usum = plan.usum;
	if ((int)cnt.size() < T) { cnt.resize(T); usum.resize(T); }
	std::vector<int> qmax_cls(T * (KSW_FAST_CLASSES + 1), 0);
	std::vector<int> krange(2 * T, 0);
	// threads beyond T keep stale histograms from a larger earlier call: clear the range they may hold
	for (int t = T; t < (int)cnt.size(); ++t)
		if ((int)cnt[t].size() == NKEY && plan.key_hi >= plan.key_lo) {
			std::fill(cnt[t].begin() + plan.key_lo, cnt[t].begin() + plan.key_hi + 1, 0u);
			std::fill(usum[t].begin() + plan.key_lo, usum[t].begin() + plan.key_hi + 1, (uint64_t)0);
		}
	std::atomic<int> bad{0};
	auto range_of = [&](int t, int64_t &b, int64_t &e) { b = std::min<int64_t>(n, t * per); e = std::min<int64_t>(n, b + per); };
	auto run_T = [&](auto &&fn) {
		if (T == 1) { fn(0); return; }
		pool->run(T, fn);
	};
	run_T([&](int t) {
		int64_t b, e; range_of(t, b, e);
		if ((int)cnt[t].size() != NKEY) { cnt[t].assign(NKEY, 0); usum[t].assign(NKEY, 0); }
		else if (plan.key_hi >= plan.key_lo) {
			std::fill(cnt[t].begin() + plan.key_lo, cnt[t].begin() + plan.key_hi + 1, 0u);
			std::fill(usum[t].begin() + plan.key_lo, usum[t].begin() + plan.key_hi + 1, (uint64_t)0);
		}
		int qm[KSW_FAST_CLASSES + 1] = {0};               // thread-local; written back once (no false sharing)
		int klo = NKEY, khi = -1;
		for (int64_t k = b; k < e; ++k) {
			const ksw_b200_job_t &j = jobs[k];
			if (j.qlen < 1 || j.tlen < 0) { bad = 1; key[k] = 0; plan.units_of[k] = 0; continue; }
			const int h0 = j.h0 < 0 ? 0 : j.h0;
			const bool fast = fast_eligible(cfg, fast_qmax, maxsc, minsc, j.qlen, h0);
			uint32_t qc = 0;
			if (!(j.qlen <= KSW_FAST_CLASS_QMAX[0] && (int64_t)h0 + (int64_t)j.qlen * maxsc + cfg->o_del + cfg->e_del <= KSW_FAST_KEYED_MAXSCORE)) {
				qc = 1;
				while (qc + 1 < KSW_FAST_CLASSES && j.qlen > KSW_FAST_CLASS_QMAX[qc]) ++qc;
			}
			// rows first (long jobs at the front keep the tail of a launch short), then the carried-in score, which sets
			// the band width; a warp claims chunks of consecutive jobs, so its lanes agree on both
			const uint32_t tl = 63u - ((uint32_t)std::min(j.tlen, 1008) >> 4);    // 6 bits
			const uint32_t hb = 127u - ((uint32_t)std::min(h0, 508) >> 2);        // 7 bits
			const uint16_t ky = (uint16_t)((fast ? (qc << 13) : 0x8000u) | (tl << 7) | hb);
			const uint32_t units = (ksw_words2(j.qlen) + ksw_words2(j.tlen) + 3) >> 2;
			key[k] = ky; plan.units_of[k] = units;
			cnt[t][ky]++; usum[t][ky] += units;
			klo = std::min<int>(klo, ky); khi = std::max<int>(khi, ky);
			int &slot = qm[fast ? (int)qc : KSW_FAST_CLASSES];
			slot = std::max(slot, j.qlen);
		}
		krange[2 * t] = klo; krange[2 * t + 1] = khi;
		for (int c = 0; c <= KSW_FAST_CLASSES; ++c) qmax_cls[t * (KSW_FAST_CLASSES + 1) + c] = qm[c];
	});
	if (bad) {
		plan.key_lo = 0; plan.key_hi = NKEY - 1;       // the histograms are dirty: have the next call clear all of them
		err = "ksw_b200: job with qlen < 1 or tlen < 0";
		return 2;
	}

	// exclusive prefix over (key, thread): start position / start pool offset of every (key, thread) run
	// (only the key range that occurs is walked; class boundaries are picked up on the way)
	int klo = NKEY, khi = -1;
	for (int t = 0; t < T; ++t) { klo = std::min(klo, krange[2 * t]); khi = std::max(khi, krange[2 * t + 1]); }
	plan.key_lo = klo; plan.key_hi = khi;
	uint64_t pos = 0, off = 0;
	int64_t below[KSW_FAST_CLASSES + 2];     // number of jobs with key < c<<13, and < 0x8000 in the last entry
	for (int c = 0; c <= KSW_FAST_CLASSES + 1; ++c) below[c] = -1;
	auto boundary_of = [](int c) { return c <= KSW_FAST_CLASSES ? (c << 13) : 0x8000; };
	for (int ky = klo; ky <= khi; ++ky) {
		for (int c = 0; c <= KSW_FAST_CLASSES + 1; ++c)
			if (below[c] < 0 && ky >= boundary_of(c)) below[c] = (int64_t)pos;
		for (int t = 0; t < T; ++t) {
			const uint32_t c = cnt[t][ky];
			const uint64_t u = usum[t][ky];
			cnt[t][ky] = (uint32_t)pos; usum[t][ky] = off;
			pos += c; off += u;
		}
	}
	for (int c = 0; c <= KSW_FAST_CLASSES + 1; ++c) if (below[c] < 0) below[c] = (int64_t)pos;
	if (off > 0xffffffffull) { plan.key_lo = 0; plan.key_hi = NKEY - 1; err = "ksw_b200: packed pool exceeds 64 GiB"; return 2; }
	// keys: class c spans [c<<13, (c+1)<<13) (0x6000-0x7fff unused), the generic block starts at 0x8000
	{
		plan.n_fast = below[KSW_FAST_CLASSES + 1];
		plan.n_generic = n - plan.n_fast;
		for (int c = 0; c < KSW_FAST_CLASSES; ++c) {
			plan.fast_class_n[c] = std::min(below[c + 1], plan.n_fast) - std::min(below[c], plan.n_fast);
			plan.fast_class_qmax[c] = 0;
			for (int t = 0; t < T; ++t)
				plan.fast_class_qmax[c] = std::max(plan.fast_class_qmax[c], qmax_cls[t * (KSW_FAST_CLASSES + 1) + c]);
		}
		plan.qmax_generic = 0;
		for (int t = 0; t < T; ++t)
			plan.qmax_generic = std::max(plan.qmax_generic, qmax_cls[t * (KSW_FAST_CLASSES + 1) + KSW_FAST_CLASSES]);
	}
	plan.pool_bytes = (size_t)off * 16;

	// stable scatter: each thread walks its caller range again
	run_T([&](int t) {
		int64_t b, e; range_of(t, b, e);
		uint32_t *c = cnt[t].data();
		uint64_t *u = usum[t].data();
		for (int64_t k = b; k < e; ++k) {
			const uint16_t ky = key[k];
			plan.pos_of[k] = c[ky]++;
			plan.off_of[k] = (uint32_t)u[ky];
			u[ky] += plan.units_of[k];
		}
	});
	return 0;
}

int ksw_pack_fill(const KswPackPlan &plan, const ksw_b200_cfg_t *cfg, const ksw_b200_job_t *jobs,
                  const uint8_t *qpool, const uint8_t *tpool, DevJob *dj, uint32_t *pool,
                  std::vector<uint32_t> &nmask, KswPool *tp)
{
	const int64_t n = plan.n;
	const int T = tp ? tp->size() : 1;
	struct NList { std::vector<uint32_t> words; std::vector<std::pair<int64_t, uint32_t>> where; };
	std::vector<NList> nl(T);
	// caller order: the byte-coded sequences (the bulk of the traffic) are streamed sequentially, the packed
	// records are written to their binned positions
	parallel_ranges(tp, n, [&](int t, int64_t b, int64_t e) {
		std::vector<uint32_t> qm, tm;
		int last_qlen = -1, last_w = 0, last_weff = 0;
		for (int64_t k = b; k < e; ++k) {
			const ksw_b200_job_t &j = jobs[k];
			const uint32_t p = plan.pos_of[k];
			DevJob d;
			d.seq_off = plan.off_of[k];
			d.idx = (uint32_t)k;
			d.qlen = j.qlen; d.tlen = j.tlen;
			d.h0 = j.h0 < 0 ? 0 : j.h0;                                                     // ksw.c:384
			if (j.qlen != last_qlen || j.w != last_w) {
				last_weff = ksw_clamp_w(j.qlen, plan.maxsc, cfg->o_del, cfg->e_del, cfg->o_ins, cfg->e_ins, j.w, cfg->end_bonus);
				last_qlen = j.qlen; last_w = j.w;
			}
			d.w = last_weff;
			d.flags = 0; d.nmask_off = 0;
			uint32_t *dst = pool + (size_t)d.seq_off * 4;
			const uint32_t qw = ksw_words2(j.qlen), tw = ksw_words2(j.tlen);
			const uint32_t qmw = ksw_words1(j.qlen), tmw = ksw_words1(j.tlen);
			if (qm.size() < qmw) qm.resize(qmw);
			if (tm.size() < tmw) tm.resize(tmw);
			const bool qn = pack2(qpool + j.q_off, j.qlen, dst, qm.data());
			const bool tn = pack2(tpool + j.t_off, j.tlen, dst + qw, tm.data());
			for (uint32_t x = qw + tw; x < plan.units_of[k] * 4u; ++x) dst[x] = 0;
			if (qn || tn) {
				d.flags = (qn ? KSW_FLAG_QN : 0u) | (tn ? KSW_FLAG_TN : 0u);
				nl[t].where.emplace_back((int64_t)p, (uint32_t)nl[t].words.size());
				if (qn) nl[t].words.insert(nl[t].words.end(), qm.begin(), qm.begin() + qmw);
				if (tn) nl[t].words.insert(nl[t].words.end(), tm.begin(), tm.begin() + tmw);
			}
			dj[p] = d;
		}
	});
	nmask.clear();
	for (auto &l : nl) {
		if (l.words.empty()) continue;
		const size_t base = nmask.size();
		nmask.insert(nmask.end(), l.words.begin(), l.words.end());
		for (auto &w : l.where) dj[w.first].nmask_off = (uint32_t)(base + w.second);
	}
	return 0;
}

// ------------------------------------------------------------------ KswPool
#include <condition_variable>
#include <mutex>

struct KswPool::Impl {
	std::vector<std::thread> th;
	std::mutex mu;
	std::condition_variable cv_go, cv_done;
	void (*fn)(void *, int) = nullptr;
	void *arg = nullptr;
	int n_tasks = 0;
	uint64_t gen = 0;
	int pending = 0;
	bool stop = false;
};

KswPool::KswPool(int n_threads) : impl_(new Impl()), n_(std::max(1, n_threads))
{
	for (int t = 1; t < n_; ++t)
		impl_->th.emplace_back([this, t] {
			uint64_t seen = 0;
			for (;;) {
				void (*fn)(void *, int); void *arg; bool mine;
				{
					std::unique_lock<std::mutex> lk(impl_->mu);
					impl_->cv_go.wait(lk, [&] { return impl_->stop || impl_->gen != seen; });
					if (impl_->stop) return;
					seen = impl_->gen;
					fn = impl_->fn; arg = impl_->arg; mine = t < impl_->n_tasks;
				}
				if (mine) {
					fn(arg, t);
					std::unique_lock<std::mutex> lk(impl_->mu);
					if (--impl_->pending == 0) impl_->cv_done.notify_one();
				}
			}
		});
}

KswPool::~KswPool()
{
	{
		std::unique_lock<std::mutex> lk(impl_->mu);
		impl_->stop = true;
	}
	impl_->cv_go.notify_all();
	for (auto &x : impl_->th) x.join();
	delete impl_;
}

void KswPool::run(int n_tasks, void (*fn)(void *, int), void *arg)
{
	n_tasks = std::max(1, std::min(n_tasks, n_));
	if (n_tasks > 1) {
		std::unique_lock<std::mutex> lk(impl_->mu);
		impl_->fn = fn; impl_->arg = arg; impl_->n_tasks = n_tasks;
		impl_->pending = n_tasks - 1;
		++impl_->gen;
		lk.unlock();
		impl_->cv_go.notify_all();
	}
	fn(arg, 0);
	if (n_tasks > 1) {
		std::unique_lock<std::mutex> lk(impl_->mu);
		impl_->cv_done.wait(lk, [&] { return impl_->pending == 0; });
	}
}
