// ksw_pair_core.h — per-lane logic of the PAIR extension kernel: TWO extension jobs per lane, one in each 16-bit
// half of every s16x2 register.  Written once and compiled twice, like ksw_fast_core.h: as sm_100a device code
// inside ksw_pair.cu and as plain C++ by the CPU emulation in tests/emu (test infrastructure only).
//
// Why pairs: with one job per lane (ksw_fast_core.h) the two halves of a register hold neighbouring columns of the
// SAME row, so the F recurrence F(j+1) = max(F(j)-e, ..) runs *inside* a register (4 dependent DPX ops + 2 PRMT per
// 4 cells) and the H store needs a one-column shift (2 PRMT).  With two independent jobs in the halves there is no
// dependency between the halves at all: every DPX instruction advances both jobs by one cell, F is one VIADDMNMX
// per column, and the "shift" of the H store is just last column's register.  Semantics per job: bwa-0.7.8/ksw.c:379-476.
//
//   per column (2 cells, one of each job):
//     score = PRMT(row of job A's target base | row of job B's target base, selector)     1 ALU   (ksw.c:430 q[j])
//     h'    = VIADDMNMX.S16x2(Hdiag, score, E)                                            1       (ksw.c:430-431)
//     g     = VIADDMNMX.S16x2(h', -oe_ins, 0)                                             1       (*)
//     h     = VIMNMX.S16x2(h', F)                                                         1       (ksw.c:432)
//     F'    = VIADDMNMX.S16x2(F, -e_ins, g)      the only loop-carried op                 1       (ksw.c:440-443)
//     E'    = max(E - e_del, h - oe_del, 0)                                               2 (or 1 + FMA-pipe adds)
//     (m,mj): key = h*128 + column (IMAD), VIMNMX3.U16x2 over two columns                 0.5     (ksw.c:433-435)
//     zero? : VIMNMX3.S16x2 min over two columns (feeds the band trim)                    0.5     (ksw.c:463-466)
//   (*) F(j+1) = max(F(j)-e_ins, h'(j)-oe_ins, 0): equal to the reference's expression because o_ins >= 0
//       (same identity as ksw_fast_core.h; jobs with o_ins < 0 never reach the s16 kernels).
//
// Every H/E/F value carries the constant bias B = o_del+e_del in both halves (B is the floor of all clamps), as in
// ksw_fast_core.h; a biased score of B means 0.
//
// The two jobs of a lane keep their own row counter, band [lo,hi), best cell, z-drop state ... (KswFastLane each);
// they share only the column sweep.  Row r of the lane = row iA of job A and row iB of job B, swept over the UNION
// of the two bands in absolute column numbers, two columns (one 128-bit shared-memory word) at a time:
//
//      masked pairs   up to the first column pair that lies inside BOTH bands
//      plain pairs    while both jobs are inside their bands (empty if the bands are disjoint)
//      masked pairs   from the first pair that is not inside both bands to the end of the union
//
// A masked pair forces the inputs of a job that is outside its band at a column to (H = -8192, E = 0): left of its
// band it then computes h = 0 and keeps F = 0, exactly the state the reference starts the band with; right of its
// band it only sees its own decaying F (< its row maximum), which can change neither (m, mj) nor any cell that is read
// later.  What such "phantom" cells store is harmless: row r+1 of a job reads eh[j] only for lo <= j <= hi of row r,
// the real cells write eh[lo..hi-1] (and eh[hi].h when column hi is swept), and the three slots the reference writes
// outside its inner loop — eh[lo].h = first-column value (ksw.c:429), eh[hi] = (h1, 0) (ksw.c:446) — are written
// after the sweep with 16-bit stores.  The zero detector ignores phantoms (a false zero would only cost the slow trim
// path, never exactness).
// Jobs whose query contains an N are not paired (a PRMT reaches 4 score bytes per job): the packer routes them to
// the one-job-per-lane kernel.  A target N simply selects row 4 of the matrix.
//
// Status: bit-exact (tests/test_pair_lane.py on the CPU, tests/test_gpu_parity.py on the GPU) but NOT the default:
// shared memory per job is the same as in the one-job-per-lane kernel, so an SM holds half as many warps (7 instead
// of 13 for 101 bp queries) and the kernel is latency-bound — see DESIGN.md §5.2 for the measurements.
#pragma once
#include <stdint.h>
#include "ksw_dev.cuh"
#include "ksw_fast_core.h"

#if defined(KSW_PAIR_STATS) && !defined(__CUDACC__)
extern long long ksw_pair_stats[32];      // test-only counters of the CPU emulation (tests/emu)
#define KSW_STAT(k, v) (ksw_pair_stats[k] += (v))
#else
#define KSW_STAT(k, v) ((void)0)
#endif

namespace kswdpx {
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ uint32_t min2(uint32_t a, uint32_t b) { return __vmins2(a, b); }
#else
KSW_EMU uint32_t min2(uint32_t a, uint32_t b) { return pk16(mn(lo16(a), lo16(b)), mn(hi16(a), hi16(b))); }
#endif
} // namespace kswdpx

// Shared-memory view of one lane.  Column pair p = columns 2p, 2p+1:
//   he[p*T] = uint4 { H(2p), E(2p), H(2p+1), E(2p+1) }, every word = (lo half: job 0, hi half: job 1)
//             (H(k) = eh[k].h, i.e. H(i-1, k-1) when row i starts; E(k) = eh[k].e)
//   sq[p*T] = PRMT selectors, bytes { job0(2p), job1(2p), job0(2p+1), job1(2p+1) }:
//             job 0: code | (8|code) << 4  (byte `code` of the low source word, then its sign)
//             job 1: (4+code) | (12+code) << 4  (byte `code` of the high source word, then its sign)
template <int T>
struct KswPairMem {
	ksw_u4 *he;
	uint32_t *sq;
	KSW_HD uint16_t *h16(int k, int X) const { return reinterpret_cast<uint16_t *>(&he[(k >> 1) * T]) + ((k & 1) << 2) + X; }
	KSW_HD uint16_t *e16(int k, int X) const { return h16(k, X) + 2; }
	KSW_HD uint8_t *s8(int k, int X) const { return reinterpret_cast<uint8_t *>(&sq[(k >> 1) * T]) + ((k & 1) << 1) + X; }
};

#define KSW_PAIR_COLPAIRS(qlen) (((qlen) >> 1) + 1)   /* column pairs that cover columns 0..qlen */

// Row -1 of a freshly fetched job (ksw.c:394-396), E = 0 and the selector bytes, written into half X of lane `owner`.
// COOPERATIVE like ksw_fast_setup_quads: columns are dealt round-robin to the helpers.
template <int T>
static KSW_HD void ksw_pair_setup(ksw_u4 *he, uint32_t *sq, const int owner, const int X, const int helper, const int n_helpers,
                                  const KswFastConst &K, const uint32_t seq_off, const int qlen, const int h0,
                                  const uint32_t *pool)
{
	const uint32_t *q2 = pool + (size_t)seq_off * 4;
	const KswPairMem<T> M{he + owner, sq + owner};
	for (int k = helper; k <= qlen; k += n_helpers) {
		int v = k == 0 ? h0 : h0 - K.oe_ins - (k - 1) * K.e_ins;          // closed form of ksw.c:394-396
		v = (v > 0 ? v : 0) + K.B;
		*M.h16(k, X) = (uint16_t)v;
		*M.e16(k, X) = (uint16_t)K.B;
		const uint32_t code = k < qlen ? (q2[k >> 4] >> ((k & 15) << 1)) & 3u : 0u;
		*M.s8(k, X) = (uint8_t)(X ? ((4u + code) | ((12u + code) << 4)) : (code | ((8u | code) << 4)));
	}
}

struct KswPairRowRegs {       // registers carried along a row, both jobs
	uint32_t F;               // F entering the next column
	uint32_t Hc;              // H(i, j-1): what the shifted store writes to eh[j].h
	uint32_t m, zmin;         // running max of the keys h*128+column / running min of h
};

// One column pair (4 cells: two columns of each job).  v = he[p], sw = sq[p]; colpk = column 2p in both halves;
// z0/z1: bits to OR into the zero detector's view of column 2p / 2p+1 (0 for a plain step).  Returns what is stored
// back: eh[j].h = H(i, j-1), eh[j].e = E(i+1, j).
static KSW_HD ksw_u4 ksw_pair_cells(KswPairRowRegs &R, const KswFastConst &K, const uint32_t mrx, const uint32_t mry,
                                    const ksw_u4 v, const uint32_t sw, const uint32_t colpk, const uint32_t z0, const uint32_t z1)
{
	using namespace kswdpx;
	const uint32_t sc0 = prmt(mrx, mry, sw), sc1 = prmt(mrx, mry, ksw_hi16_of(sw));
	const uint32_t hp0 = addmax2(v.x, sc0, v.y), hp1 = addmax2(v.z, sc1, v.w);
	const uint32_t g0 = addmax2(hp0, K.neg_oei, K.Bpk), g1 = addmax2(hp1, K.neg_oei, K.Bpk);
	const uint32_t h0 = max2(hp0, R.F);
	const uint32_t F1 = addmax2(R.F, K.neg_ei, g0);
	const uint32_t h1 = max2(hp1, F1);
	R.F = addmax2(F1, K.neg_ei, g1);
	ksw_u4 o;
	o.x = R.Hc;
	o.z = h0;
	o.y = max3_2(v.y - K.ed32, h0 - K.oed32, K.Bpk);           // every half is >= B = oe_del: the 32-bit subtractions cannot borrow
	o.w = max3_2(v.w - K.ed32, h1 - K.oed32, K.Bpk);
	R.Hc = h1;
	R.m = maxu2(R.m, maxu2(h0 * 128u + colpk, h1 * 128u + (colpk + 0x10001u)));
	R.zmin = min3_2(R.zmin, h0 | z0, h1 | z1);
	return o;
}

// A column pair at the edge of the bands: a job that is outside its own band [lo,hi) at a column gets the phantom
// inputs (H = -8192, E = 0) there and is hidden from the zero detector.
template <int T>
static KSW_HD void ksw_pair_step_masked(KswPairRowRegs &R, const KswPairMem<T> &M, const KswFastConst &K, const uint32_t mrx,
                                        const uint32_t mry, const int p, const int a0, const unsigned wa, const int b0,
                                        const unsigned wb)
{
	const int c0 = p << 1;
	// inside [lo, lo+width) as one unsigned comparison
	const uint32_t k0 = ((unsigned)(c0 - a0) < wa ? 0x0000ffffu : 0u) | ((unsigned)(c0 - b0) < wb ? 0xffff0000u : 0u);
	const uint32_t k1 = ((unsigned)(c0 + 1 - a0) < wa ? 0x0000ffffu : 0u) | ((unsigned)(c0 + 1 - b0) < wb ? 0xffff0000u : 0u);
	ksw_u4 v = M.he[p * T];
	const uint32_t sw = M.sq[p * T];
	v.x = (v.x & k0) | (KSW_NEGPK & ~k0); v.y = (v.y & k0) | (K.Bpk & ~k0);
	v.z = (v.z & k1) | (KSW_NEGPK & ~k1); v.w = (v.w & k1) | (K.Bpk & ~k1);
	M.he[p * T] = ksw_pair_cells(R, K, mrx, mry, v, sw, (uint32_t)c0 * 0x10001u, ~k0 & 0x7fff7fffu, ~k1 & 0x7fff7fffu);
}

// The reference's trim scans (ksw.c:463-466) over job X's stored eh[].h.  A zero is the biased value B.  The upward scan
// usually runs far (row 0 of a job starts with the whole query as its band), so it tests two columns per 128-bit load.
template <int T>
static KSW_HD void ksw_pair_trim_scan(const KswPairMem<T> &M, const KswFastConst &K, const int X, int rarg, int lo, int hi,
                                      int &new_lo, int &new_hi)
{
	const uint32_t Bh = (uint32_t)K.B;
	int j = rarg;
	while (j >= lo && *M.h16(j, X) != (uint16_t)Bh) --j;
	new_lo = j + 1;
	j = rarg + 2;
	if ((j & 1) && j <= hi && *M.h16(j, X) != (uint16_t)Bh) ++j;    // align to a column pair (or stop here)
	if (!(j & 1)) {
		while (j + 1 <= hi) {
			const ksw_u4 v = M.he[(j >> 1) * T];
			const uint32_t h0 = X ? v.x >> 16 : v.x & 0xffffu, h1 = X ? v.z >> 16 : v.z & 0xffffu;
			if (h0 == Bh || h1 == Bh) break;
			j += 2;
		}
		while (j <= hi && *M.h16(j, X) != (uint16_t)Bh) ++j;
	}
	new_hi = j;
}

// Processes the current row of every running job of the lane (bit X of `run`: job X is running).  Returns the mask of
// jobs that finished (their results are then in L[X]).  The bookkeeping of the two jobs is written as straight-line
// code with selects (rare cases are collected in masks and handled behind one branch), so that the compiler can
// interleave the two independent jobs: a warp of this kernel has few neighbours to hide latencies behind.
template <int T>
static KSW_HD unsigned ksw_pair_row(KswFastLane *L, const unsigned run, const KswPairMem<T> &M, const KswFastConst &K,
                                    const ksw_u2 *mrow)
{
	using namespace kswdpx;
	unsigned fin = 0, act = 0, odd = 0;
	int lo[2], hi[2], left0[2];
	uint32_t mr[2];
#ifdef __CUDACC__
#pragma unroll
#endif
	for (int X = 0; X < 2; ++X) {
		KswFastLane &J = L[X];
		const bool r = (run >> X) & 1u;
		const int i = J.i;
		// target base of this row
		if (r && (i & 15) == 0 && i) {
			J.tw = J.tw_next;
			const int nx = (i >> 4) + 1;
			if ((nx << 4) < J.tlen) J.tw_next = J.t2[nx];
		}
		int t = (int)((J.tw >> ((i & 15) << 1)) & 3u);
		if (r && J.tn && ((J.tn[i >> 5] >> (i & 31)) & 1u)) t = 4;
		mr[X] = mrow[t].x;
		const int l0 = J.h0 - (K.o_del + K.e_del * (i + 1));       // first-column value, used even when lo>0 (ksw.c:415-416)
		left0[X] = l0 > 0 ? l0 : 0;
		int a = J.lo, b = J.hi;
		a = a > i - J.w ? a : i - J.w;                              // ksw.c:418-420
		b = b < i + J.w + 1 ? b : i + J.w + 1;
		b = b < J.qlen ? b : J.qlen;
		const bool rows_left = i < J.tlen, nonempty = b > a;
		const bool on = r && rows_left && nonempty;
		lo[X] = on ? a : 0; hi[X] = on ? b : 0;
		act |= on ? 1u << X : 0u;
		odd |= (r && !on) ? 1u << X : 0u;
		J.cells += on ? (uint32_t)(b - a) : 0u;
	}
	if (odd) {
		// a running job without a row to sweep ends here: no rows at all (tlen == 0), or an empty band.  For the empty band
		// the reference's loop variable stays at beg, so "j == qlen" means lo == qlen (ksw.c:447); then m == 0 ends the job
		// (ksw.c:451)
#ifdef __CUDACC__
#pragma unroll
#endif
		for (int X = 0; X < 2; ++X) {
			if (!((odd >> X) & 1u)) continue;
			KswFastLane &J = L[X];
			const int i = J.i;
			if (i < J.tlen) {
				int a = J.lo;
				a = a > i - J.w ? a : i - J.w;
				if (a == J.qlen) {
					if (left0[X] >= J.end_sc) J.end_i = i;
					J.end_sc = J.end_sc > left0[X] ? J.end_sc : left0[X];
				}
			}
		}
		fin = odd;
	}
	if (!act) return fin;

	// The sweep, in column pairs: masked pairs up to the first pair that lies inside both bands, plain pairs, masked
	// pairs from the first pair that does not.  (Disjoint bands: everything masked, the gap included.)
	int a0 = lo[0], a1 = hi[0], b0 = lo[1], b1 = hi[1];
	if (!(act & 1u)) a0 = a1 = b0;
	if (!(act & 2u)) b0 = b1 = a0;
	const int lo_min = a0 < b0 ? a0 : b0, lo_max = a0 < b0 ? b0 : a0;
	const int hi_min = a1 < b1 ? a1 : b1, hi_max = a1 < b1 ? b1 : a1;
	const int P0 = lo_min >> 1, P1 = (hi_max + 1) >> 1;            // pairs [P0, P1) cover the union of the bands
	int M0 = P1, M1 = P1;                                          // plain pairs [M0, M1)
	if (lo_max < hi_min) { M0 = (lo_max + 1) >> 1; M1 = hi_min >> 1; }
	const unsigned wa = (unsigned)(a1 - a0), wb = (unsigned)(b1 - b0);
	const uint32_t mrx = mr[0], mry = mr[1];
	KSW_STAT(0, 1); KSW_STAT(1, M0 - P0); KSW_STAT(2, M1 - M0); KSW_STAT(3, P1 - M1); KSW_STAT(4, act == 3u);

	KswPairRowRegs R;
	R.F = K.Bpk; R.Hc = K.Bpk; R.m = 0; R.zmin = 0x7fff7fffu;
	for (int p = P0; p < M0; ++p) ksw_pair_step_masked<T>(R, M, K, mrx, mry, p, a0, wa, b0, wb);
	if (M0 < M1) {
		ksw_u4 vn = M.he[M0 * T];                                   // software prefetch, one column pair ahead
		uint32_t swn = M.sq[M0 * T];
		uint32_t colpk = (uint32_t)(M0 << 1) * 0x10001u;
#ifdef __CUDACC__
#pragma unroll 4
#endif
		for (int p = M0; p < M1; ++p) {
			const ksw_u4 v = vn;
			const uint32_t sw = swn;
			vn = M.he[(p + 1) * T]; swn = M.sq[(p + 1) * T];       // p+1 <= qlen/2 is always allocated
			M.he[p * T] = ksw_pair_cells(R, K, mrx, mry, v, sw, colpk, 0u, 0u);
			colpk += 0x20002u;
		}
	}
	for (int p = M1; p < P1; ++p) ksw_pair_step_masked<T>(R, M, K, mrx, mry, p, a0, wa, b0, wb);

	// After the sweep, both jobs in straight-line code.  What the sweep leaves to 16-bit stores: eh[beg].h = first-column
	// value (ksw.c:429: the sweep stored a phantom's h there), eh[end].h = h1 = H(i, hi-1), eh[end].e = 0 (ksw.c:446).
	// If column hi was swept (as a phantom of this job) its shifted store has already put h1 there; otherwise h1 is
	// still the carried register.  (A job that is not active has lo = hi = 0 and a dead half: the stores are harmless.)
	unsigned scan = 0;
	int rarg_[2];
#ifdef __CUDACC__
#pragma unroll
#endif
	for (int X = 0; X < 2; ++X) {
		KswFastLane &J = L[X];
		const bool on = (act >> X) & 1u;
		const int i = J.i;
		uint16_t *ph_lo = M.h16(lo[X], X), *ph_hi = M.h16(hi[X], X);
		const int swept_h1 = (int)*ph_hi;
		const int left_b = hi[X] < (P1 << 1) ? swept_h1 : (int)(X ? R.Hc >> 16 : R.Hc & 0xffffu);   // still biased
		const int left = left_b - K.B;
		*ph_lo = (uint16_t)(left0[X] + K.B);
		*ph_hi = (uint16_t)left_b;
		ph_hi[2] = (uint16_t)K.B;
		const bool at_end = on && hi[X] == J.qlen;                 // ksw.c:447-450, ties -> last row
		J.end_i = (at_end && left >= J.end_sc) ? i : J.end_i;
		J.end_sc = (at_end && left > J.end_sc) ? left : J.end_sc;
		const uint32_t kmax = X ? R.m >> 16 : R.m & 0xffffu;
		const int rmax = (int)(kmax >> 7) - K.B, rarg = (int)(kmax & 127u);
		rarg_[X] = rarg;
		const bool live = on && rmax > 0;                          // ksw.c:451 (a row of zeros has key max B*128+col)
		// ksw.c:452-461
		const bool better = live && rmax > J.best;
		const int di = i - J.best_i, dj = rarg - J.best_j;
		const int gap = di > dj ? (di - dj) * K.e_del : (dj - di) * K.e_ins;
		const bool drop = live && !better && K.zdrop > 0 && (J.best - rmax - gap > K.zdrop);
		const int d = rarg > i ? rarg - i : i - rarg;
		const int off2 = J.off > d ? J.off : d;
		J.best_i = better ? i : J.best_i;
		J.best_j = better ? rarg : J.best_j;
		J.off = better ? off2 : J.off;
		J.best = better ? rmax : J.best;
		const bool go_on = live && !drop;
		// band trim (ksw.c:463-466).  No zero in the row: every eh[j].h for j in (lo, hi] is non-zero, so only the
		// first-column slot eh[lo].h can stop the downward scan, and the upward scan runs off the end
		const uint32_t zx = X ? R.zmin >> 16 : R.zmin & 0xffffu;
		J.lo = go_on ? (left0[X] ? lo[X] : lo[X] + 1) : J.lo;
		J.hi = go_on ? hi[X] + 1 : J.hi;
		scan |= (go_on && zx == (uint32_t)K.B) ? 1u << X : 0u;
		J.i = go_on ? i + 1 : i;
		fin |= (on && (!go_on || i + 1 >= J.tlen)) ? 1u << X : 0u;
		KSW_STAT(7, go_on);
	}
	if (scan) {
#ifdef __CUDACC__
#pragma unroll
#endif
		for (int X = 0; X < 2; ++X) {
			if (!((scan >> X) & 1u)) continue;
			int nl, nh;
			ksw_pair_trim_scan<T>(M, K, X, rarg_[X], lo[X], hi[X], nl, nh);
			KSW_STAT(8, 1); KSW_STAT(9, rarg_[X] - nl + 1); KSW_STAT(10, nh - (rarg_[X] + 2));
			L[X].lo = nl; L[X].hi = nh;
		}
	}
	return fin;
}
