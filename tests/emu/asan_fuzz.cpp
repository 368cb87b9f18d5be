// tests/emu/asan_fuzz.cpp — TEST INFRASTRUCTURE.  The packer and the per-lane sources of both s16x2 kernels (one job per lane, two jobs per lane) compiled for the
// CPU with AddressSanitizer + UBSan, every shared-memory array allocated at exactly the size the launcher gives the
// kernel, random jobs checked against the oracle.  (compute-sanitizer is not available on the GPU pool; this finds
// out-of-bounds accesses and undefined shifts in the same source on the host.)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "../../bwa_mem_quickassist_b200/csrc/ksw_pack.h"
#include "../../bwa_mem_quickassist_b200/csrc/ksw_fast_core.h"
#include "../../bwa_mem_quickassist_b200/csrc/ksw_pair_core.h"
#include "../../bwa_mem_quickassist_b200/csrc/ksw_class.h"
#include "../../bwa_mem_quickassist_b200/csrc/ksw_gfast_core.h"

extern "C" int ksw_oracle_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                                  int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                                  int *qle, int *tle, int *gtle, int *gscore, int *max_off, int64_t *cells, int32_t *rows);

extern "C" int ksw_oracle_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                                  int o_del, int e_del, int o_ins, int e_ins, int w, int *n_cigar, uint32_t *cigar);

// the fast global-alignment kernels' per-lane source (ksw_gfast_core.h): DP rows into an H slab of exactly the size the
// runtime reserves per job (tlen x nqb quads), backtrack by recomputation, against the oracle's score and CIGAR
static int gfast_fuzz(int n, std::mt19937 &rng, const ksw_b200_cfg_t &cfg, long long *n_run)
{
	auto U = [&](int lo, int hi) { return (int)(rng() % (unsigned)(hi - lo + 1)) + lo; };
	KswParams P; ksw_params_from_cfg(&cfg, P);
	KswGConst C; ksw_gfast_make_const(P, C);
	ksw_u2 mrow[5];
	for (int t = 0; t < 5; ++t) mrow[t] = ksw_gfast_matrow(P, t);
	KswFastEdge edge[5];
	for (int r = 0; r < 5; ++r) ksw_fast_edge_entry(r, edge[r]);
	const int cost = ksw_gfast_cell_cost(cfg.mat, cfg.o_del, cfg.e_del, cfg.o_ins, cfg.e_ins);
	int bad = 0;
	for (int k = 0; k < n; ++k) {
		const int ql = U(0, 3) == 0 ? U(1, 20) : U(1, 260);
		std::vector<uint8_t> q(ql), t;
		for (auto &x : q) x = (uint8_t)(U(0, 99) == 0 ? 4 : U(0, 3));
		for (int j = 0; j < ql; ++j) {                              // the target: the query with substitutions and short indels
			const int ev = U(0, 99);
			if (ev < 3) continue;                                   // base missing from the target
			if (ev < 6) t.push_back((uint8_t)U(0, 3));              // extra base
			t.push_back(ev < 14 ? (uint8_t)U(0, 3) : q[j]);
		}
		if (t.empty()) t.push_back(0);
		const int tl = (int)t.size(), w = abs(tl - ql) + U(0, 3) * U(0, 20);
		if (!ksw_gfast_eligible(cost, cfg.o_del, cfg.e_del, cfg.o_ins, cfg.e_ins, ql, tl, w)) continue;
		const int nq = (ql >> 2) + 1, nqb = ksw_gfast_nqb(ql, w);
		std::vector<ksw_u4> hq(nq);                                 // exactly what the kernel gives a lane
		std::vector<uint32_t> sq(nq);
		std::vector<ksw_u2> z((size_t)tl * nqb);
		ksw_gfast_setup<1>(hq.data(), sq.data(), ql, w, q.data(), C);
		for (int i = 0; i < tl; ++i) ksw_gfast_row<1>(hq.data(), sq.data(), edge, mrow[t[i] > 4 ? 4 : t[i]], z.data() + (size_t)i * nqb, i, ql, w, C);
		const int score = ksw_gfast_score<1>(hq.data(), ql);
		KswGWalk<1> wk;
		wk.z = z.data(); wk.query = q.data(); wk.target = t.data(); wk.mat = P.mat;
		wk.qlen = ql; wk.tlen = tl; wk.w = w; wk.nqb = nqb;
		wk.o_del = P.o_del; wk.e_del = P.e_del; wk.o_ins = P.o_ins; wk.e_ins = P.e_ins;
		std::vector<uint32_t> rev((size_t)ql + tl + 2);
		const int nc = wk.run([&](int r, int op, int len) { rev[r] = (uint32_t)len << 4 | (uint32_t)op; });
		int n_ref = 0;
		std::vector<uint32_t> cig((size_t)ql + tl + 2);
		const int want = ksw_oracle_global2(ql, q.data(), tl, t.data(), 5, cfg.mat, cfg.o_del, cfg.e_del, cfg.o_ins, cfg.e_ins, w, &n_ref, cig.data());
		bool ok = want == score && n_ref == nc;
		for (int r = 0; ok && r < nc; ++r) ok = cig[r] == rev[nc - 1 - r];
		if (!ok && bad++ < 5) fprintf(stderr, "global mismatch: qlen %d tlen %d w %d\n", ql, tl, w);
		++*n_run;
	}
	return bad;
}

int main(int argc, char **argv)
{
	const int n = argc > 1 ? atoi(argv[1]) : 20000;
	const unsigned seed = argc > 2 ? (unsigned)atoi(argv[2]) : 1u;
	std::mt19937 rng(seed);
	auto U = [&](int lo, int hi) { return (int)(rng() % (unsigned)(hi - lo + 1)) + lo; };
	ksw_b200_cfg_t cfg;
	memset(&cfg, 0, sizeof(cfg));
	const int a = U(1, 3), b = U(1, 6);
	for (int i = 0; i < 5; ++i) for (int j = 0; j < 5; ++j) cfg.mat[i * 5 + j] = (int8_t)((i < 4 && j < 4) ? (i == j ? a : -b) : -1);
	cfg.m = 5; cfg.o_del = U(0, 8); cfg.e_del = U(1, 3); cfg.o_ins = U(0, 8); cfg.e_ins = U(1, 3);
	cfg.zdrop = U(0, 3) ? U(5, 150) : -1; cfg.end_bonus = U(0, 10);
	std::vector<uint8_t> qpool, tpool;
	std::vector<ksw_b200_job_t> jobs(n);
	for (int k = 0; k < n; ++k) {
		const int ql = U(1, 3) == 1 ? U(1, 12) : U(1, 300), tl = U(0, 2 * ql + 40);
		jobs[k].q_off = qpool.size(); jobs[k].t_off = tpool.size(); jobs[k].qlen = ql; jobs[k].tlen = tl;
		jobs[k].h0 = U(0, 4) ? U(0, 260) : 0; jobs[k].w = U(0, 3) ? 100 : U(1, 60);
		std::vector<uint8_t> t(tl), q(ql);
		for (auto &x : t) x = (uint8_t)U(0, 3);
		const int shift = U(-3, 3), err = U(0, 30);
		for (int j = 0; j < ql; ++j) {
			const int p = j + shift;
			q[j] = (p >= 0 && p < tl && U(0, 99) >= err) ? t[p] : (uint8_t)U(0, 3);
			if (U(0, 199) == 0) q[j] = 4;
		}
		if (U(0, 29) == 0) for (auto &x : t) if (U(0, 19) == 0) x = 4;
		qpool.insert(qpool.end(), q.begin(), q.end());
		tpool.insert(tpool.end(), t.begin(), t.end());
	}
	qpool.push_back(0); tpool.push_back(0);
	KswPool tp(2);
	KswPackStats st;
	std::string err;
	const int fast_qmax = KSW_FAST_CLASS_QMAX[KSW_FAST_CLASSES - 1];
	if (ksw_pack_sizes(&cfg, n, jobs.data(), fast_qmax, &tp, st, err)) { fprintf(stderr, "%s\n", err.c_str()); return 2; }
	std::vector<DevJob> dj(n);
	std::vector<uint32_t> pool(st.pool_bytes / 4), nmask;          // exact sizes: ASan sees any overrun
	ksw_pack_stream(st, &cfg, jobs.data(), fast_qmax, qpool.data(), tpool.data(), dj.data(), pool.data(), nmask, &tp);
	KswParams P; ksw_params_from_cfg(&cfg, P);
	KswFastConst K; ksw_fast_make_const(P, K);
	ksw_u2 mrow[5]; for (int t = 0; t < 5; ++t) mrow[t] = ksw_fast_matrow(P, t);
	KswFastEdge edge[5]; for (int r = 0; r < 5; ++r) ksw_fast_edge_entry(r, edge[r]);
	int bad = 0;
	long long n_fast = 0, n_keyed = 0;
	for (int64_t p = 0; p < n; ++p) {
		const DevJob &jb = dj[p];
		const uint32_t cls = (jb.flags >> KSW_CLASS_SHIFT) & KSW_CLASS_MASK;
		if (cls >= KSW_CLASS_GENERIC) continue;
		++n_fast; n_keyed += cls == 0;
		const int nq = KSW_FAST_QUADS(jb.qlen);
		std::vector<ksw_u4> hq(nq);                               // exactly what ksw_fast_smem_bytes() reserves per lane
		std::vector<uint32_t> sq(nq);
		KswFastMem<1> M{hq.data(), sq.data(), edge};
		KswFastLane L;
		ksw_fast_setup_quads<1>(hq.data(), sq.data(), 0, 0, 1, K, jb.seq_off, jb.qlen, jb.h0, jb.flags, jb.nmask_off, pool.data(),
		                        nmask.empty() ? nullptr : nmask.data());
		ksw_fast_init_lane(L, jb, pool.data(), nmask.empty() ? nullptr : nmask.data());
		if (cls == 0) { while (!ksw_fast_row<1, true>(L, M, K, mrow)) {} }
		else { while (!ksw_fast_row<1, false>(L, M, K, mrow)) {} }
		DevRes r; ksw_fast_result(L, r);
		const ksw_b200_job_t &j = jobs[jb.idx];
		int qle, tle, gtle, gscore, max_off;
		const int sc = ksw_oracle_extend2(j.qlen, qpool.data() + j.q_off, j.tlen, tpool.data() + j.t_off, 5, cfg.mat, cfg.o_del, cfg.e_del,
		                                  cfg.o_ins, cfg.e_ins, j.w, cfg.end_bonus, cfg.zdrop, j.h0, &qle, &tle, &gtle, &gscore, &max_off, 0, 0);
		if (sc != r.score || qle != r.qle || tle != r.tle || gtle != r.gtle || gscore != r.gscore || max_off != r.max_off) {
			if (bad++ < 5) fprintf(stderr, "mismatch job %u: qlen %d tlen %d h0 %d w %d\n", jb.idx, j.qlen, j.tlen, j.h0, j.w);
		}
	}
	// the pair kernel's lane (two jobs per lane) over the class-0 jobs, arrays at exactly the launcher's size
	long long n_pair = 0;
	{
		std::vector<uint32_t> list;
		int qmax = 1;
		for (int64_t p = 0; p < n; ++p)
			if (((dj[p].flags >> KSW_CLASS_SHIFT) & KSW_CLASS_MASK) == 0u) { list.push_back((uint32_t)p); qmax = std::max(qmax, dj[p].qlen); }
		const int np = KSW_PAIR_COLPAIRS(qmax);                     // what ksw_pair_smem_bytes() reserves per lane
		std::vector<ksw_u4> he(np);
		std::vector<uint32_t> sq(np);
		KswPairMem<1> PM{he.data(), sq.data()};
		KswFastLane L[2] = {};
		unsigned run = 0;
		size_t next = 0;
		while (true) {
			for (int X = 0; X < 2; ++X) {
				if (((run >> X) & 1u) || next >= list.size()) continue;
				const DevJob &jb = dj[list[next++]];
				ksw_pair_setup<1>(he.data(), sq.data(), 0, X, 0, 1, K, jb.seq_off, jb.qlen, jb.h0, pool.data());
				ksw_fast_init_lane(L[X], jb, pool.data(), nmask.empty() ? nullptr : nmask.data());
				run |= 1u << X;
			}
			if (!run) break;
			const unsigned fin = ksw_pair_row<1>(L, run, PM, K, mrow);
			for (int X = 0; X < 2; ++X) {
				if (!((fin >> X) & 1u)) continue;
				DevRes r; ksw_fast_result(L[X], r);
				const ksw_b200_job_t &j = jobs[L[X].idx];
				int qle, tle, gtle, gscore, max_off;
				const int sc = ksw_oracle_extend2(j.qlen, qpool.data() + j.q_off, j.tlen, tpool.data() + j.t_off, 5, cfg.mat, cfg.o_del, cfg.e_del,
				                                  cfg.o_ins, cfg.e_ins, j.w, cfg.end_bonus, cfg.zdrop, j.h0, &qle, &tle, &gtle, &gscore, &max_off, nullptr, nullptr);
				if (sc != r.score || qle != r.qle || tle != r.tle || gtle != r.gtle || gscore != r.gscore || max_off != r.max_off) {
					if (bad++ < 5) fprintf(stderr, "pair mismatch job %u: qlen %d tlen %d h0 %d w %d\n", L[X].idx, j.qlen, j.tlen, j.h0, j.w);
				}
				++n_pair;
				run &= ~(1u << X);
			}
		}
	}
	long long n_glob = 0;
	bad += gfast_fuzz(n / 4, rng, cfg, &n_glob);
	printf("asan_fuzz: %d jobs, %lld on the fast path (%lld keyed), %lld through the pair lane, %lld global alignments, %d mismatches\n", n, n_fast, n_keyed, n_pair, n_glob, bad);
	return bad ? 1 : 0;
}
