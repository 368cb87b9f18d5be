// ksw_pack.cpp — see ksw_pack.h.  Host C++ only.
#include "ksw_pack.h"
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <thread>
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif

int ksw_pack_force_words = 0;

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
// The packed records go to pinned staging that the CPU never reads again: non-temporal 16-byte stores skip the
// read-for-ownership of every destination line (a quarter of the packer's memory traffic).
inline void stream16(void *dst, const void *src) { _mm_stream_si128((__m128i *)dst, _mm_loadu_si128((const __m128i *)src)); }
inline void stream_fence() { _mm_sfence(); }
#define KSW_HAVE_STREAM 1
__attribute__((target("bmi2"))) inline uint32_t squeeze16_pext(uint64_t a, uint64_t b)
{
	return (uint32_t)__builtin_ia32_pext_di(a, 0x0303030303030303ull) |
	       ((uint32_t)__builtin_ia32_pext_di(b, 0x0303030303030303ull) << 16);
}
const bool g_has_bmi2 = __builtin_cpu_supports("bmi2");
#define KSW_HAVE_PEXT 1
#endif

#ifndef KSW_HAVE_STREAM
inline void stream16(void *dst, const void *src) { memcpy(dst, src, 16); }
inline void stream_fence() {}
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
inline bool pack2_words(const uint8_t *s, int len, uint32_t *out, uint32_t *nmask)
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
		uint64_t a, b;
		if (len >= 16) {
			// the last `rem` bases without a variable-length copy: load the 16 bytes that END at the sequence end and
			// shift the (16 - rem) bytes that belong to the previous word out
			unsigned __int128 v;
			memcpy(&v, s + len - 16, 16);
			v >>= 8 * (16 - rem);
			a = (uint64_t)v; b = (uint64_t)(v >> 64);
		} else {
			uint8_t tail[16] = {0};
			for (int x = 0; x < rem; ++x) tail[x] = s[x];
			memcpy(&a, tail, 8); memcpy(&b, tail + 8, 8);
		}
		out[full] = ((a | b) & 0xFCFCFCFCFCFCFCFCull) ? slow_word(s + 16 * full, 16 * full, rem) : squeeze16(a, b);
	}
	return has_n;
}

#ifdef KSW_HAVE_PEXT
// 64 byte codes per step: two multiply-adds fold four codes into one byte (c0 + 4 c1, then + 16 (c2 + 4 c3)), a narrowing
// move gathers the sixteen bytes = four words of the 2-bit stream.  Bytes past the end are never touched (masked load)
// and pack as 0.  WRITES WHOLE 16-BYTE GROUPS: up to three words past ceil(len/16) are zeroed (the caller's line buffer
// has the room, and packs the query before the target).  An N anywhere sends the whole sequence down the word path.
__attribute__((target("avx512f,avx512bw"))) bool pack2_avx512(const uint8_t *s, int len, uint32_t *out, uint32_t *nmask)
{
	const __m512i not2 = _mm512_set1_epi8((char)0xFC), w14 = _mm512_set1_epi16(0x0401), w116 = _mm512_set1_epi32(0x00100001);
	__mmask64 bad = 0;
	int k = 0;
	for (; k + 64 <= len; k += 64) {
		const __m512i v = _mm512_loadu_si512((const void *)(s + k));
		bad |= _mm512_test_epi8_mask(v, not2);
		const __m512i y = _mm512_madd_epi16(_mm512_maddubs_epi16(v, w14), w116);
		_mm_storeu_si128((__m128i *)(out + (k >> 4)), _mm512_cvtepi32_epi8(y));
	}
	if (k < len) {
		const __m512i v = _mm512_maskz_loadu_epi8(~0ull >> (64 - (len - k)), (const void *)(s + k));
		bad |= _mm512_test_epi8_mask(v, not2);
		const __m512i y = _mm512_madd_epi16(_mm512_maddubs_epi16(v, w14), w116);
		const __m128i r = _mm512_cvtepi32_epi8(y);
		if (len - k > 32) _mm_storeu_si128((__m128i *)(out + (k >> 4)), r);
		else _mm_storel_epi64((__m128i *)(out + (k >> 4)), r);
	}
	return bad ? pack2_words(s, len, out, nmask) : false;
}
const bool g_has_avx512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && !getenv("KSW_B200_PACK_WORDS");
#endif

// how many words past ceil(len/16) pack2 may zero
#define KSW_PACK2_SLACK 3

inline bool pack2(const uint8_t *s, int len, uint32_t *out, uint32_t *nmask)
{
#ifdef KSW_HAVE_PEXT
	if (g_has_avx512 && !ksw_pack_force_words) return pack2_avx512(s, len, out, nmask);
#endif
	return pack2_words(s, len, out, nmask);
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
	return ksw_clamp_w_expr(qlen, maxsc, o_del, e_del, o_ins, e_ins, w, end_bonus);
}

void ksw_scoring_from_cfg(const ksw_b200_cfg_t *cfg, int fast_qmax, KswScoring &S)
{
	S.maxsc = ksw_mat_max(cfg->mat);
	S.minsc = 0;
	for (int i = 0; i < 25; ++i) S.minsc = std::min<int>(S.minsc, cfg->mat[i]);
	S.o_del = cfg->o_del; S.e_del = cfg->e_del; S.o_ins = cfg->o_ins; S.e_ins = cfg->e_ins;
	S.end_bonus = cfg->end_bonus; S.fast_qmax = fast_qmax;
	// KSW_B200_DISABLE_WARP=1 (A/B switch): what the warp-cooperative kernel would take goes to the thread-per-job kernel
	static const int warp_qmax = [] { const char *e = getenv("KSW_B200_DISABLE_WARP"); return (e && *e && *e != '0') ? 0 : KSW_WARP_MAX_QLEN; }();
	S.warp_qmax = warp_qmax;
}

void ksw_params_from_cfg(const ksw_b200_cfg_t *cfg, KswParams &P)
{
	memset(&P, 0, sizeof(P));
	memcpy(P.mat, cfg->mat, 25);
	P.o_del = cfg->o_del; P.e_del = cfg->e_del; P.o_ins = cfg->o_ins; P.e_ins = cfg->e_ins;
	P.zdrop = cfg->zdrop;
}

namespace {

// the same contiguous split of [0,n) in both passes
inline int pack_ranges(KswPool *tp, int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(tp ? tp->size() : 1, (n + 16383) / 16384)); }

template <class F>
void run_ranges(KswPool *tp, int T, F &&fn) { if (T == 1) fn(0); else tp->run(T, fn); }

} // namespace

int ksw_pack_sizes(const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_job_t *jobs, int fast_qmax, KswPool *tp,
                   KswPackStats &st, std::string &err)
{
	if (cfg->m != 5) { err = "ksw_b200: only m == 5 is supported (every reference caller passes 5)"; return 2; }
	if (n > 0x7fffffffLL) { err = "ksw_b200: more than 2^31-1 jobs in one batch"; return 2; }
	KswScoring S;
	ksw_scoring_from_cfg(cfg, fast_qmax, S);
	const int T = pack_ranges(tp, n);
	const int64_t per = (n + T - 1) / T;
	struct Local { int64_t cn[KSW_N_CLASSES]; int qm[KSW_N_CLASSES]; uint64_t units; int bad; char pad[64]; };
	std::vector<Local> loc(T);
	run_ranges(tp, T, [&](int t) {
		Local l;
		memset(&l, 0, sizeof(l));
		const int64_t b = std::min<int64_t>(n, t * per), e = std::min<int64_t>(n, b + per);
		for (int64_t k = b; k < e; ++k) {
			const ksw_b200_job_t &j = jobs[k];
			if (j.qlen < 1 || j.tlen < 0) { l.bad = 1; continue; }
			const uint32_t c = ksw_job_class(S, j.qlen, j.h0 < 0 ? 0 : j.h0);
			l.cn[c]++; l.qm[c] = std::max(l.qm[c], j.qlen);
			l.units += ksw_job_units(j.qlen, j.tlen);
		}
		loc[t] = l;
	});
	st.n = n;
	st.range_base.assign(T + 1, 0);
	for (int c = 0; c < KSW_N_CLASSES; ++c) { st.class_n[c] = 0; st.class_qmax[c] = 0; }
	uint64_t off = 0;
	for (int t = 0; t < T; ++t) {
		if (loc[t].bad) { err = "ksw_b200: job with qlen < 1 or tlen < 0"; return 2; }
		st.range_base[t] = off;
		off += loc[t].units;
		for (int c = 0; c < KSW_N_CLASSES; ++c) { st.class_n[c] += loc[t].cn[c]; st.class_qmax[c] = std::max(st.class_qmax[c], loc[t].qm[c]); }
	}
	st.range_base[T] = off;
	if (off > 0xffffffffull) { err = "ksw_b200: packed pool exceeds 64 GiB"; return 2; }
	st.pool_bytes = (size_t)off * 16;
	return 0;
}

int ksw_pack_stream(KswPackStats &st, const ksw_b200_cfg_t *cfg, const ksw_b200_job_t *jobs, int fast_qmax,
                    const uint8_t *qpool, const uint8_t *tpool, DevJob *dj, uint32_t *pool,
                    std::vector<uint32_t> &nmask, KswPool *tp)
{
	const int64_t n = st.n;
	KswScoring S;
	ksw_scoring_from_cfg(cfg, fast_qmax, S);
	const int maxsc = S.maxsc;
	const int T = pack_ranges(tp, n);
	const int64_t per = (n + T - 1) / T;
	struct NList { std::vector<uint32_t> words; std::vector<std::pair<int64_t, uint32_t>> where; };
	std::vector<NList> nl(T);
	run_ranges(tp, T, [&](int t) {
		const int64_t b = std::min<int64_t>(n, t * per), e = std::min<int64_t>(n, b + per);
		std::vector<uint32_t> qm, tm, line;                 // line: one job's packed words, built in cache and streamed out
		uint64_t off = st.range_base[t];
		int last_qlen = -1, last_w = 0, last_weff = 0;
		for (int64_t k = b; k < e; ++k) {
			const ksw_b200_job_t &j = jobs[k];
			DevJob d;
			d.seq_off = (uint32_t)off;
			d.idx = (uint32_t)k;
			d.qlen = j.qlen; d.tlen = j.tlen;
			d.h0 = j.h0 < 0 ? 0 : j.h0;                                                     // ksw.c:384
			if (j.qlen != last_qlen || j.w != last_w) {
				last_weff = ksw_clamp_w(j.qlen, maxsc, cfg->o_del, cfg->e_del, cfg->o_ins, cfg->e_ins, j.w, cfg->end_bonus);
				last_qlen = j.qlen; last_w = j.w;
			}
			d.w = last_weff;
			d.flags = ksw_job_class(S, j.qlen, d.h0) << KSW_CLASS_SHIFT;
			d.nmask_off = 0;
			const uint32_t qw = ksw_words2(j.qlen), tw = ksw_words2(j.tlen), units = (qw + tw + 3) >> 2;
			const uint32_t qmw = ksw_words1(j.qlen), tmw = ksw_words1(j.tlen);
			if (qm.size() < qmw) qm.resize(qmw);
			if (tm.size() < tmw) tm.resize(tmw);
			if (line.size() < units * 4u + KSW_PACK2_SLACK) line.resize(units * 4u + KSW_PACK2_SLACK);
			uint32_t *dst = line.data();
			const bool qn = pack2(qpool + j.q_off, j.qlen, dst, qm.data());
			const bool tn = pack2(tpool + j.t_off, j.tlen, dst + qw, tm.data());
			for (uint32_t x = qw + tw; x < units * 4u; ++x) dst[x] = 0;
			{
				uint32_t *out = pool + (size_t)off * 4;         // 16-byte aligned: the pool is, and off counts 16-byte units
				for (uint32_t u = 0; u < units; ++u) stream16(out + 4 * u, dst + 4 * u);
			}
			if (qn || tn) {
				d.flags |= (qn ? KSW_FLAG_QN : 0u) | (tn ? KSW_FLAG_TN : 0u);
				nl[t].where.emplace_back(k, (uint32_t)nl[t].words.size());
				if (qn) nl[t].words.insert(nl[t].words.end(), qm.begin(), qm.begin() + qmw);
				if (tn) nl[t].words.insert(nl[t].words.end(), tm.begin(), tm.begin() + tmw);
			}
			if (qn || tn) dj[k] = d;                          // its nmask_off is patched below: keep it an ordinary store
			else { stream16(&dj[k], &d); stream16(reinterpret_cast<char *>(&dj[k]) + 16, reinterpret_cast<const char *>(&d) + 16); }
			off += units;
		}
		stream_fence();
	});
	nmask.clear();
	for (auto &l : nl) {
		if (l.words.empty()) continue;
		const size_t base = nmask.size();
		nmask.insert(nmask.end(), l.words.begin(), l.words.end());
		for (auto &w : l.where) {
			DevJob &d = dj[w.first];
			d.nmask_off = (uint32_t)(base + w.second);
			// class 0 is N-free: the pair kernel's score look-up (ksw_pair_core.h) has no room for the query-N column of
			// the matrix, and the keyed one-job-per-lane kernel skips the N-mask look-ups of both sequences.  A class-0 job
			// that holds an N moves to class 1 (one job per lane, any base code)
			if (((d.flags >> KSW_CLASS_SHIFT) & KSW_CLASS_MASK) == 0u) {
				d.flags = (d.flags & ~(KSW_CLASS_MASK << KSW_CLASS_SHIFT)) | (1u << KSW_CLASS_SHIFT);
				st.class_n[0]--; st.class_n[1]++;
				st.class_qmax[1] = std::max(st.class_qmax[1], d.qlen);
			}
		}
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
