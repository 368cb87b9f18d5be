// ksw_pack.cpp — see ksw_pack.h.  Host C++ only.
#include "ksw_pack.h"
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>

namespace {

template <class F>
void parallel_for(int64_t n, int n_threads, F &&fn)
{
	if (n <= 0) return;
	n_threads = (int)std::max<int64_t>(1, std::min<int64_t>(n_threads, (n + 4095) / 4096));
	if (n_threads == 1) { fn(0, (int64_t)0, n); return; }
	std::vector<std::thread> th;
	th.reserve(n_threads);
	const int64_t per = (n + n_threads - 1) / n_threads;
	for (int t = 0; t < n_threads; ++t) {
		const int64_t b = t * per, e = std::min<int64_t>(n, b + per);
		if (b >= e) break;
		th.emplace_back([&fn, t, b, e] { fn(t, b, e); });
	}
	for (auto &x : th) x.join();
}

// 2-bit packing of `len` byte codes into ceil(len/16) words.  Returns true if a code > 3 (N) was
// seen; those bases are stored as 0 and flagged in nmask (ceil(len/32) words, caller-zeroed).
inline bool pack2(const uint8_t *s, int len, uint32_t *out, uint32_t *nmask)
{
	bool has_n = false;
	const int nw = (len + 15) >> 4;
	for (int wi = 0, k = 0; wi < nw; ++wi, k += 16) {
		uint32_t word = 0;
		const int lim = std::min(16, len - k);
		bool slow = lim < 16;
		if (!slow) {
			uint64_t a, b;
			memcpy(&a, s + k, 8); memcpy(&b, s + k + 8, 8);
			if ((a | b) & 0xFCFCFCFCFCFCFCFCull) slow = true;
			else {
				// gather the low 2 bits of each of 8 bytes into 16 contiguous bits
				auto squeeze = [](uint64_t v) -> uint32_t {
					v = (v | (v >> 6)) & 0x000F000F000F000Full;
					v = (v | (v >> 12)) & 0x000000FF000000FFull;
					v = (v | (v >> 24)) & 0xFFFFull;
					return (uint32_t)v;
				};
				word = squeeze(a) | (squeeze(b) << 16);
			}
		}
		if (slow) {
			for (int x = 0; x < lim; ++x) {
				uint32_t c = s[k + x];
				if (c > 3) { has_n = true; nmask[(k + x) >> 5] |= 1u << ((k + x) & 31); c = 0; }
				word |= c << (2 * x);
			}
		}
		out[wi] = word;
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
                  int n_threads, KswPackPlan &plan, std::string &err)
{
	if (cfg->m != 5) { err = "ksw_b200: only m == 5 is supported (every reference caller passes 5)"; return 2; }
	if (n > 0x7fffffffLL) { err = "ksw_b200: more than 2^31-1 jobs in one batch"; return 2; }
	const int maxsc = ksw_mat_max(cfg->mat);
	int minsc = 0;
	for (int i = 0; i < 25; ++i) minsc = std::min<int>(minsc, cfg->mat[i]);
	plan.n = n; plan.maxsc = maxsc;

	// bin key from the lengths alone (no sequence bytes are touched in this phase):
	// class bits | rows (coarse) | carried-in score — the two things that drive the band a job sweeps.
	// Keys are inverted so that long jobs come first (short tail at the end of the launch).
	constexpr int NKEY = 1 << 16;
	std::vector<uint16_t> key(n);
	std::atomic<int> bad{0};
	parallel_for(n, n_threads, [&](int, int64_t b, int64_t e) {
		for (int64_t k = b; k < e; ++k) {
			const ksw_b200_job_t &j = jobs[k];
			if (j.qlen < 1 || j.tlen < 0) { bad = 1; key[k] = 0; continue; }
			const int h0 = j.h0 < 0 ? 0 : j.h0;
			const bool fast = fast_eligible(cfg, fast_qmax, maxsc, minsc, j.qlen, h0);
			uint32_t qc = 0;
			while (qc + 1 < KSW_FAST_CLASSES && j.qlen > KSW_FAST_CLASS_QMAX[qc]) ++qc;
			const uint32_t tl = 127u - ((uint32_t)std::min(j.tlen, 2032) >> 4);   // 7 bits
			const uint32_t hb = 63u - ((uint32_t)std::min(h0, 504) >> 3);         // 6 bits
			key[k] = (uint16_t)((fast ? (qc << 13) : 0x8000u) | (tl << 6) | hb);
		}
	});
	if (bad) { err = "ksw_b200: job with qlen < 1 or tlen < 0"; return 2; }

	std::vector<int64_t> hist(NKEY + 1, 0);
	for (int64_t k = 0; k < n; ++k) hist[(size_t)key[k] + 1]++;
	for (int i = 0; i < NKEY; ++i) hist[i + 1] += hist[i];
	plan.n_fast = hist[0x8000];
	plan.n_generic = n - plan.n_fast;
	for (int c = 0; c < KSW_FAST_CLASSES; ++c) {
		const int hi_key = c + 1 < KSW_FAST_CLASSES ? ((c + 1) << 13) : 0x8000;
		plan.fast_class_n[c] = hist[hi_key] - hist[c << 13];
		plan.fast_class_qmax[c] = 0;
	}
	plan.order.resize(n);
	{
		std::vector<int64_t> cur(hist.begin(), hist.end() - 1);
		for (int64_t k = 0; k < n; ++k) plan.order[cur[key[k]]++] = (uint32_t)k;
	}
	plan.seq_off.resize(n + 1);
	uint64_t off = 0;
	int qmg = 0;
	int64_t class_end[KSW_FAST_CLASSES];
	{
		int64_t acc = 0;
		for (int c = 0; c < KSW_FAST_CLASSES; ++c) { acc += plan.fast_class_n[c]; class_end[c] = acc; }
	}
	int cls = 0;
	for (int64_t p = 0; p < n; ++p) {
		const ksw_b200_job_t &j = jobs[plan.order[p]];
		plan.seq_off[p] = (uint32_t)off;
		off += (ksw_words2(j.qlen) + ksw_words2(j.tlen) + 3) >> 2;
		if (off > 0xffffffffull) { err = "ksw_b200: packed pool exceeds 64 GiB"; return 2; }
		if (p < plan.n_fast) {
			while (p >= class_end[cls]) ++cls;
			plan.fast_class_qmax[cls] = std::max(plan.fast_class_qmax[cls], j.qlen);
		} else qmg = std::max(qmg, j.qlen);
	}
	plan.seq_off[n] = (uint32_t)off;
	plan.pool_bytes = (size_t)off * 16;
	plan.qmax_generic = qmg;
	return 0;
}

int ksw_pack_fill(const KswPackPlan &plan, const ksw_b200_cfg_t *cfg, const ksw_b200_job_t *jobs,
                  const uint8_t *qpool, const uint8_t *tpool, DevJob *dj, uint32_t *pool,
                  std::vector<uint32_t> &nmask, int n_threads)
{
	const int64_t n = plan.n;
	const int T = std::max(n_threads, 1);
	struct NList { std::vector<uint32_t> words; std::vector<std::pair<int64_t, uint32_t>> where; };
	std::vector<NList> nl(T);
	parallel_for(n, T, [&](int t, int64_t b, int64_t e) {
		std::vector<uint32_t> qm, tm;
		for (int64_t p = b; p < e; ++p) {
			const ksw_b200_job_t &j = jobs[plan.order[p]];
			DevJob d;
			d.seq_off = plan.seq_off[p];
			d.idx = plan.order[p];
			d.qlen = j.qlen; d.tlen = j.tlen;
			d.h0 = j.h0 < 0 ? 0 : j.h0;                                                     // ksw.c:384
			d.w = ksw_clamp_w(j.qlen, plan.maxsc, cfg->o_del, cfg->e_del, cfg->o_ins, cfg->e_ins, j.w, cfg->end_bonus);
			d.flags = 0; d.nmask_off = 0;
			uint32_t *dst = pool + (size_t)plan.seq_off[p] * 4;
			const uint32_t qw = ksw_words2(j.qlen), tw = ksw_words2(j.tlen);
			qm.assign(ksw_words1(j.qlen), 0); tm.assign(ksw_words1(j.tlen), 0);
			const bool qn = pack2(qpool + j.q_off, j.qlen, dst, qm.data());
			const bool tn = pack2(tpool + j.t_off, j.tlen, dst + qw, tm.data());
			for (uint32_t x = qw + tw; x < (plan.seq_off[p + 1] - plan.seq_off[p]) * 4u; ++x) dst[x] = 0;
			if (qn || tn) {
				d.flags = (qn ? KSW_FLAG_QN : 0u) | (tn ? KSW_FLAG_TN : 0u);
				nl[t].where.emplace_back(p, (uint32_t)nl[t].words.size());
				if (qn) nl[t].words.insert(nl[t].words.end(), qm.begin(), qm.end());
				if (tn) nl[t].words.insert(nl[t].words.end(), tm.begin(), tm.end());
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
