// tests/emu/fast_emu.cpp — TEST INFRASTRUCTURE.  Compiles the product's packer (ksw_pack.cpp) and
// the per-lane logic of the fast kernel (ksw_fast_core.h) as plain C++ with software DPX, so that the
// exact kernel source can be fuzzed against the oracle on a machine without a GPU.  Jobs the packer
// routes to the generic kernel are reported with score = INT32_MIN (the emulation covers only the
// fast path).  Never linked into the product library.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <climits>
#include "../../bwa_mem_quickassist_b200/csrc/ksw_pack.h"
#include "../../bwa_mem_quickassist_b200/csrc/ksw_fast_core.h"

extern "C" int ksw_fast_emu_batch(const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_job_t *jobs,
                                  const uint8_t *qpool, const uint8_t *tpool, ksw_b200_res_t *res,
                                  int64_t *n_fast_out, int threads, int64_t *n_keyed_out)
{
	KswPool tp(threads);
	KswPackStats st;
	std::string err;
	const int fast_qmax = KSW_FAST_CLASS_QMAX[KSW_FAST_CLASSES - 1];
	int rc = ksw_pack_sizes(cfg, n, jobs, fast_qmax, &tp, st, err);
	if (rc) return rc;
	std::vector<DevJob> dj(n ? n : 1);
	std::vector<uint32_t> pool(st.pool_bytes / 4 + 4), nmask;
	rc = ksw_pack_stream(st, cfg, jobs, fast_qmax, qpool, tpool, dj.data(), pool.data(), nmask, &tp);
	if (rc) return rc;
	if (nmask.empty()) nmask.push_back(0);
	KswParams P;
	ksw_params_from_cfg(cfg, P);
	KswFastConst K;
	ksw_fast_make_const(P, K);
	ksw_u2 mrow[5];
	for (int t = 0; t < 5; ++t) mrow[t] = ksw_fast_matrow(P, t);
	int64_t n_fast = 0;
	for (int c = 0; c < KSW_FAST_CLASSES; ++c) n_fast += st.class_n[c];
	if (n_fast_out) *n_fast_out = n_fast;
	for (int64_t k = 0; k < n; ++k) res[k].score = INT_MIN;
	KswFastEdge edge[5];
	for (int r = 0; r < 5; ++r) ksw_fast_edge_entry(r, edge[r]);
	for (int64_t p = 0; p < n; ++p) {                             // any order: the binned order only matters for speed
		const DevJob &jb = dj[p];
		const uint32_t cls = (jb.flags >> KSW_CLASS_SHIFT) & KSW_CLASS_MASK;
		if (cls >= KSW_CLASS_GENERIC) continue;
		const bool keyed = cls == 0;
		const int nq = KSW_FAST_QUADS(jb.qlen);
		std::vector<ksw_u4> hq(nq + 1);
		std::vector<uint32_t> sq(nq + 1);
		for (auto &v : hq) v.x = v.y = v.z = v.w = 0x5a5a5a5au;
		KswFastMem<1> M{hq.data(), sq.data(), edge};
		KswFastLane L;
		ksw_fast_setup_quads<1>(hq.data(), sq.data(), 0, 0, 1, K, jb.seq_off, jb.qlen, jb.h0, jb.flags, jb.nmask_off, pool.data(), nmask.data());
		ksw_fast_init_lane(L, jb, pool.data(), nmask.data());
		if (keyed) { while (!ksw_fast_row<1, true>(L, M, K, mrow)) {} }
		else { while (!ksw_fast_row<1, false>(L, M, K, mrow)) {} }
		DevRes r;
		ksw_fast_result(L, r);
		memcpy(&res[jb.idx], &r, sizeof(r));
	}
	if (n_keyed_out) *n_keyed_out = st.class_n[0];
	return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// The PAIR kernel's lane (ksw_pair_core.h: two jobs per lane) on the CPU: `lanes` emulated lanes, each with two job
// slots that are refilled from the class-0 job list as soon as a job ends — the same life cycle as on the GPU, so
// that jobs meet partners at every phase (fresh job next to a half-finished one, disjoint bands, one slot empty at
// the end).  order: 0 = caller order, 1 = the device's binning key order.  Jobs outside class 0 keep score = INT32_MIN.
#include <algorithm>
#include "../../bwa_mem_quickassist_b200/csrc/ksw_pair_core.h"

extern "C" int ksw_pair_emu_batch(const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_job_t *jobs,
                                  const uint8_t *qpool, const uint8_t *tpool, ksw_b200_res_t *res,
                                  uint32_t *cells_out, int lanes, int order, int64_t *n_pair_out, int64_t *steps_out)
{
	KswPool tp(2);
	KswPackStats st;
	std::string err;
	const int fast_qmax = KSW_FAST_CLASS_QMAX[KSW_FAST_CLASSES - 1];
	int rc = ksw_pack_sizes(cfg, n, jobs, fast_qmax, &tp, st, err);
	if (rc) return rc;
	std::vector<DevJob> dj(n ? n : 1);
	std::vector<uint32_t> pool(st.pool_bytes / 4 + 4), nmask;
	rc = ksw_pack_stream(st, cfg, jobs, fast_qmax, qpool, tpool, dj.data(), pool.data(), nmask, &tp);
	if (rc) return rc;
	if (nmask.empty()) nmask.push_back(0);
	KswParams P;
	ksw_params_from_cfg(cfg, P);
	KswFastConst K;
	ksw_fast_make_const(P, K);
	ksw_u2 mrow[5];
	for (int t = 0; t < 5; ++t) mrow[t] = ksw_fast_matrow(P, t);
	for (int64_t k = 0; k < n; ++k) res[k].score = INT_MIN;
	std::vector<uint32_t> list;
	int qmax = 1;
	for (int64_t k = 0; k < n; ++k)
		if (((dj[k].flags >> KSW_CLASS_SHIFT) & KSW_CLASS_MASK) == 0u) { list.push_back((uint32_t)k); qmax = std::max(qmax, dj[k].qlen); }
	if ((int64_t)list.size() != st.class_n[0]) return 77;            // the packer's class counts must match the flags
	if (order == 1)
		std::stable_sort(list.begin(), list.end(), [&](uint32_t a, uint32_t b) {
			auto key = [&](const DevJob &j) { return ((63u - std::min((uint32_t)j.qlen >> 1, 63u)) << 7) | (127u - ((uint32_t)j.h0 < 96u ? (uint32_t)j.h0 : 96u + std::min(((uint32_t)j.h0 - 96u) >> 4, 31u))); };
			return key(dj[a]) < key(dj[b]);
		});
	if (n_pair_out) *n_pair_out = (int64_t)list.size();
	const int np = KSW_PAIR_COLPAIRS(qmax);
	size_t next = 0;
	int64_t steps = 0;
	for (int ln = 0; ln < lanes; ++ln) {
		// lanes run one after the other here (they are independent); each takes every lanes-th slice of the list
		std::vector<ksw_u4> he(np);
		std::vector<uint32_t> sq(np);
		for (auto &v : he) v.x = v.y = v.z = v.w = 0x5a5a5a5au;  // stale garbage, as on the GPU
		for (auto &v : sq) v = 0xa5a5a5a5u;
		KswPairMem<1> M{he.data(), sq.data()};
		KswFastLane L[2] = {};
		unsigned run = 0;
		const size_t end = ln + 1 == lanes ? list.size() : std::min(list.size(), next + (list.size() + lanes - 1) / lanes);
		while (true) {
			for (int X = 0; X < 2; ++X) {
				if (((run >> X) & 1u) || next >= end) continue;
				const DevJob &jb = dj[list[next++]];
				ksw_pair_setup<1>(he.data(), sq.data(), 0, X, 0, 1, K, jb.seq_off, jb.qlen, jb.h0, pool.data());
				ksw_fast_init_lane(L[X], jb, pool.data(), nmask.data());
				run |= 1u << X;
			}
			if (!run) break;
			const unsigned fin = ksw_pair_row<1>(L, run, M, K, mrow);
			++steps;
			for (int X = 0; X < 2; ++X) {
				if (!((fin >> X) & 1u)) continue;
				DevRes r;
				ksw_fast_result(L[X], r);
				memcpy(&res[L[X].idx], &r, sizeof(r));
				if (cells_out) cells_out[L[X].idx] = L[X].cells;
				run &= ~(1u << X);
			}
		}
		next = end;
	}
	if (steps_out) *steps_out = steps;
	return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// The fast global-alignment kernels' per-lane source (ksw_gfast_core.h) on the CPU: the DP rows with software DPX into a
// per-job H slab, then the backtrack that recomputes the reference's direction bits from it.  res[k].cigar_off must be
// preset (capacity qlen + tlen + 2 each, like the oracle's batch entry).  Jobs the runtime would route to the int32
// kernel keep score = INT32_MIN.
#include "../../bwa_mem_quickassist_b200/csrc/ksw_class.h"
#include "../../bwa_mem_quickassist_b200/csrc/ksw_gfast_core.h"

extern "C" int ksw_gfast_emu_batch(const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_gjob_t *jobs, const uint8_t *qpool,
                                   const uint8_t *tpool, ksw_b200_gres_t *res, uint32_t *cigar_pool, int64_t *n_fast_out)
{
	KswParams P;
	ksw_params_from_cfg(cfg, P);
	KswGConst C;
	ksw_gfast_make_const(P, C);
	ksw_u2 mrow[5];
	for (int t = 0; t < 5; ++t) mrow[t] = ksw_gfast_matrow(P, t);
	KswFastEdge edge[5];
	for (int r = 0; r < 5; ++r) ksw_fast_edge_entry(r, edge[r]);
	const int cost = ksw_gfast_cell_cost(cfg->mat, cfg->o_del, cfg->e_del, cfg->o_ins, cfg->e_ins);
	int64_t n_fast = 0;
	for (int64_t k = 0; k < n; ++k) {
		const ksw_b200_gjob_t &j = jobs[k];
		res[k].score = INT_MIN; res[k].n_cigar = 0;
		if (!ksw_gfast_eligible(cost, cfg->o_del, cfg->e_del, cfg->o_ins, cfg->e_ins, j.qlen, j.tlen, j.w)) continue;
		++n_fast;
		const uint8_t *query = qpool + j.q_off, *target = tpool + j.t_off;
		const int nq = (j.qlen >> 2) + 1, nqb = ksw_gfast_nqb(j.qlen, j.w);
		std::vector<ksw_u4> hq(nq);
		std::vector<uint32_t> sq(nq + 1);                          // the row loop prefetches one quad ahead
		hq.reserve(nq + 1);
		std::vector<ksw_u2> z((size_t)j.tlen * nqb);
		for (auto &v : z) v.x = v.y = 0x5a5a5a5au;
		hq.resize(nq + 1);
		ksw_gfast_setup<1>(hq.data(), sq.data(), j.qlen, j.w, query, C);
		for (int i = 0; i < j.tlen; ++i) {
			const int t = target[i] > 4 ? 4 : target[i];
			ksw_gfast_row<1>(hq.data(), sq.data(), edge, mrow[t], z.data() + (size_t)i * nqb, i, j.qlen, j.w, C);
		}
		res[k].score = ksw_gfast_score<1>(hq.data(), j.qlen);
		KswGWalk<1> wk;
		wk.z = z.data(); wk.query = query; wk.target = target; wk.mat = P.mat;
		wk.qlen = j.qlen; wk.tlen = j.tlen; wk.w = j.w; wk.nqb = nqb;
		wk.o_del = P.o_del; wk.e_del = P.e_del; wk.o_ins = P.o_ins; wk.e_ins = P.e_ins;
		std::vector<uint32_t> rev((size_t)j.qlen + j.tlen + 2);
		const int nc = wk.run([&](int r, int op, int len) { rev[r] = (uint32_t)len << 4 | (uint32_t)op; });
		uint32_t *out = cigar_pool + res[k].cigar_off;
		for (int r = 0; r < nc; ++r) out[nc - 1 - r] = rev[r];
		res[k].n_cigar = nc;
	}
	if (n_fast_out) *n_fast_out = n_fast;
	return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// The host packer alone: what ksw_pack_sizes + ksw_pack_stream write (job records, 2-bit pool, N side pool), with the
// 64-byte SIMD path or the word-at-a-time path (force_words != 0), so that a test can compare the two byte for byte.
// Returns 0, a packer error code, or -1 when an output array is too small.
extern "C" int ksw_pack_emu(const ksw_b200_cfg_t *cfg, int64_t n, const ksw_b200_job_t *jobs, const uint8_t *qpool,
                            const uint8_t *tpool, int force_words, int threads, void *dj_out, uint32_t *pool_out,
                            int64_t pool_cap_words, int64_t *pool_words, uint32_t *nmask_out, int64_t nmask_cap_words,
                            int64_t *nmask_words, int64_t *class_n_out)
{
	KswPool tp(threads);
	KswPackStats st;
	std::string err;
	const int fast_qmax = KSW_FAST_CLASS_QMAX[KSW_FAST_CLASSES - 1];
	int rc = ksw_pack_sizes(cfg, n, jobs, fast_qmax, &tp, st, err);
	if (rc) return rc;
	if ((int64_t)(st.pool_bytes / 4) > pool_cap_words) return -1;
	std::vector<uint32_t> nmask;
	const int saved = ksw_pack_force_words;
	ksw_pack_force_words = force_words;
	rc = ksw_pack_stream(st, cfg, jobs, fast_qmax, qpool, tpool, (DevJob *)dj_out, pool_out, nmask, &tp);
	ksw_pack_force_words = saved;
	if (rc) return rc;
	if ((int64_t)nmask.size() > nmask_cap_words) return -1;
	if (!nmask.empty()) memcpy(nmask_out, nmask.data(), nmask.size() * 4);
	*pool_words = (int64_t)(st.pool_bytes / 4);
	*nmask_words = (int64_t)nmask.size();
	for (int c = 0; c < KSW_N_CLASSES; ++c) class_n_out[c] = st.class_n[c];
	return 0;
}
