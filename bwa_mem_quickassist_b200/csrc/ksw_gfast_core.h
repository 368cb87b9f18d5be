// ksw_gfast_core.h — per-lane logic of the fast global-alignment kernels (ksw_gfast.cu), written once and compiled twice like
// ksw_fast_core.h: as sm_100a device code, and as plain C++ with software DPX by the CPU emulation in tests/emu (test
// infrastructure; the product always runs the CUDA build).  T = the stride between a lane's consecutive quads (32 on the
// device: lane-interleaved; 1 in the emulation).  See ksw_gfast.cu for the design.
#pragma once
#include "ksw_fast_core.h"

#define KSW_G_BIAS 16384                 /* added to every H / E / F value */
#define KSW_G_NEG 1024                   /* biased "minus infinity": real values stay above 16384 - 7000 */
#define KSW_G_SC 128                     /* added to every matrix score so that it is a non-negative byte */
#define KSW_G_MINF (-0x40000000)         /* the reference's MINUS_INF (ksw.c:36), used by the backtrack */

struct KswGConst {
	uint32_t neg_ei, neg_oei, ed32, oed32;
	int o_del, e_del, o_ins, e_ins;
};

static KSW_HD void ksw_gfast_make_const(const KswParams &P, KswGConst &C)
{
	C.o_del = P.o_del; C.e_del = P.e_del; C.o_ins = P.o_ins; C.e_ins = P.e_ins;
	C.neg_ei = ksw_pk2(-P.e_ins); C.neg_oei = ksw_pk2(-(P.o_ins + P.e_ins));
	C.ed32 = ksw_pk2(P.e_del); C.oed32 = ksw_pk2(P.o_del + P.e_del);
}

// scores of target base t against query codes 0..4, each plus KSW_G_SC, as the bytes a PRMT selector picks
static KSW_HD ksw_u2 ksw_gfast_matrow(const KswParams &P, int t)
{
	ksw_u2 r;
	r.x = ((uint32_t)(P.mat[t * 5 + 0] + KSW_G_SC)) | ((uint32_t)(P.mat[t * 5 + 1] + KSW_G_SC) << 8) |
	      ((uint32_t)(P.mat[t * 5 + 2] + KSW_G_SC) << 16) | ((uint32_t)(P.mat[t * 5 + 3] + KSW_G_SC) << 24);
	r.y = (uint32_t)(P.mat[t * 5 + 4] + KSW_G_SC);         // byte 4; bytes 5..7 are the zero the selectors' high nibbles pick
	return r;
}

// one quad of a row.  EDGE: out-of-band columns of the quad are masked to "minus infinity" on load
template <bool EDGE>
static KSW_HD void ksw_gfast_quad(uint32_t &X, uint32_t &Hc, const ksw_u4 v_in, const uint32_t sw, const ksw_u2 mr, ksw_u4 *dst, ksw_u2 *zdst,
                                  const KswGConst &C, const uint32_t keepA, const uint32_t keepB, const uint32_t endA, const uint32_t endB)
{
	using namespace kswdpx;
	ksw_u4 v = v_in;
	const uint32_t negH = ksw_pk2(KSW_G_NEG - KSW_G_SC), negE = ksw_pk2(KSW_G_NEG);
	if (EDGE) {
		v.x = (v.x & keepA) | (negH & ~keepA); v.y = (v.y & keepA) | (negE & ~keepA);
		v.z = (v.z & keepB) | (negH & ~keepB); v.w = (v.w & keepB) | (negE & ~keepB);
	}
	// M = H(i-1,j-1) + S: the stored diagonal value carries -KSW_G_SC, the looked-up score +KSW_G_SC: a plain 32-bit add, no carry
	const uint32_t MA = v.x + prmt(mr.x, mr.y, sw), MB = v.z + prmt(mr.x, mr.y, ksw_hi16_of(sw));
	// F chain on Fs = F + oe_ins (ksw_fast_core.h): lo half c0 -> c1 -> c2, hi half c2 -> c3 -> c0'
	const uint32_t u1 = addmax2(X, C.neg_ei, MA);
	const uint32_t u2 = addmax2(u1, C.neg_ei, MB);
	const uint32_t FA = ksw_lo_to_hi_add(u2, X);           // (Fs(c0), Fs(c2))
	const uint32_t FB = addmax2(FA, C.neg_ei, MA);         // (Fs(c1), Fs(c3))
	const uint32_t u4 = addmax2(FB, C.neg_ei, MB);
	X = ksw_hi16_of(u4);
	// H = max(M, E, F) (ksw.c:545-548)
	const uint32_t hA = addmax2(FA, C.neg_oei, max2(MA, v.y)), hB = addmax2(FB, C.neg_oei, max2(MB, v.w));
	// E(i+1,j) = max(E - e_del, M - oe_del) (ksw.c:549-552)
	uint32_t eA = max2(v.y - C.ed32, MA - C.oed32), eB = max2(v.w - C.ed32, MB - C.oed32);
	if (EDGE) { eA = (eA & ~endA) | (negE & endA); eB = (eB & ~endB) | (negE & endB); }   // eh[end].e = -inf (ksw.c:558)
	ksw_u2 zo;
	zo.x = hA; zo.y = hB;
	*zdst = zo;
	const uint32_t sA = hA - ksw_pk2(KSW_G_SC), sB = hB - ksw_pk2(KSW_G_SC);
	ksw_u4 o;
	o.x = prmt(Hc, sB, 0x5432u);                           // (H(c0-1), H(c1)) for columns c0, c2
	o.y = eA;
	o.z = sA;                                              // (H(c0), H(c2)) for columns c1, c3
	o.w = eB;
	*dst = o;
	Hc = sB;
}

// row -1 (ksw.c:520-523): eh[0].h = 0, eh[j].h = -(o_ins + e_ins j) for j <= w, -inf beyond; every e = -inf
template <int T>
static KSW_HD void ksw_gfast_setup(ksw_u4 *hq, uint32_t *sq, const int qlen, const int w, const uint8_t *query, const KswGConst &C)
{
	const int nq = (qlen >> 2) + 1;
	for (int q = 0; q < nq; ++q) {
		int hv[4];
		uint32_t sb[4];
#pragma unroll
		for (int k = 0; k < 4; ++k) {
			const int c = (q << 2) + k;
			int v = c == 0 ? 0 : (c <= w ? -(C.o_ins + C.e_ins * c) : KSW_G_NEG - KSW_G_BIAS);
			if (c > qlen) v = KSW_G_NEG - KSW_G_BIAS;
			hv[k] = v + KSW_G_BIAS - KSW_G_SC;
			const uint32_t code = c < qlen ? (uint32_t)query[c] : 0u;
			sb[k] = (code > 4u ? 4u : code) | 0x50u;                  // byte look-up, zero above it
		}
		ksw_u4 v4;
		v4.x = (uint32_t)hv[0] | ((uint32_t)hv[2] << 16);
		v4.y = ksw_pk2(KSW_G_NEG);
		v4.z = (uint32_t)hv[1] | ((uint32_t)hv[3] << 16);
		v4.w = ksw_pk2(KSW_G_NEG);
		hq[q * T] = v4;
		sq[q * T] = (sb[0] | (sb[2] << 8)) | ((sb[1] | (sb[3] << 8)) << 16);
	}
}

// row i of one job (ksw.c:525-559); zrow = where the row's H quads go (stride T)
template <int T>
static KSW_HD void ksw_gfast_row(ksw_u4 *hq, const uint32_t *sq, const KswFastEdge *edge, const ksw_u2 mr, ksw_u2 *zrow,
                                 const int i, const int qlen, const int w, const KswGConst &C)
{
	const int beg = i > w ? i - w : 0;                                // ksw.c:529-530
	const int end = i + w + 1 < qlen ? i + w + 1 : qlen;
	const int q0 = beg >> 2, q1 = (end - 1) >> 2, hi_rel = end - (q1 << 2);
	const KswFastEdge eL = edge[beg & 3], eR = edge[hi_rel];
	uint32_t X = (uint32_t)(KSW_G_NEG + C.o_ins + C.e_ins);           // f = -inf entering the band (ksw.c:527)
	uint32_t Hc = 0;
	ksw_u4 *ph = hq + q0 * T;
	const uint32_t *ps = sq + q0 * T;
	ksw_u2 *pz = zrow;
	if (q1 == q0) {
		ksw_gfast_quad<true>(X, Hc, *ph, *ps, mr, ph, pz, C, eL.geA & eR.ltA, eL.geB & eR.ltB, eR.onlyA, eR.onlyB);
	} else {
		ksw_u4 vn = ph[T];
		uint32_t swn = ps[T];
		ksw_gfast_quad<true>(X, Hc, *ph, *ps, mr, ph, pz, C, eL.geA, eL.geB, 0u, 0u);
		ksw_u4 *const pl = hq + q1 * T;
		ph += T; ps += T; pz += T;
#pragma unroll 2
		for (; ph != pl; ph += T, ps += T, pz += T) {
			const ksw_u4 v = vn;
			const uint32_t sw = swn;
			vn = ph[T]; swn = ps[T];
			ksw_gfast_quad<false>(X, Hc, v, sw, mr, ph, pz, C, 0u, 0u, 0u, 0u);
		}
		ksw_gfast_quad<true>(X, Hc, vn, swn, mr, pl, pz, C, eR.ltA, eR.ltB, eR.onlyA, eR.onlyB);
	}
	// eh[beg].h = h1 (ksw.c:542): H(i,-1) = -(o_del + e_del (i+1)) while the band starts at column 0, never read otherwise
	if (beg == 0) *reinterpret_cast<uint16_t *>(&hq[0]) = (uint16_t)(KSW_G_BIAS - KSW_G_SC - (C.o_del + C.e_del * (i + 1)));
	if (hi_rel == 4) {
		// column `end` opens the next quad: eh[end] = {H(i,end-1), -inf} (ksw.c:558); the other half-words of the 64-bit
		// slot belong to column end+2, which is rewritten before it is read (the band end moves one column per row)
		ksw_u2 w2;
		w2.x = Hc >> 16;                                              // H(c3) of the last quad, already minus KSW_G_SC
		w2.y = ksw_pk2(KSW_G_NEG);
		*reinterpret_cast<ksw_u2 *>(&hq[(q1 + 1) * T]) = w2;
	}
}

// score = eh[qlen].h after the last row (ksw.c:560)
template <int T>
static KSW_HD int ksw_gfast_score(const ksw_u4 *hq, const int qlen)
{
	const uint16_t hv = *(reinterpret_cast<const uint16_t *>(&hq[(qlen >> 2) * T]) + ksw_fast_hslot(qlen));
	return (int)hv + KSW_G_SC - KSW_G_BIAS;
}

// ---------------------------------------------------------------- backtrack by recomputation
template <int T>
struct KswGWalk {
	const ksw_u2 *z;               // the job's H quads: element (row i, band quad r) at z[(i * nqb + r) * T]
	const uint8_t *query, *target;
	const int8_t *mat;
	int qlen, tlen, w, nqb;
	int o_del, e_del, o_ins, e_ins;

	KSW_HD int hcell(int i, int k) const       // H(i,k) incl. the boundary row / column and "outside the band"
	{
		if (i < 0) return k < 0 ? 0 : ((k + 1 <= w && k + 1 <= qlen) ? -(o_ins + e_ins * (k + 1)) : KSW_G_MINF);
		if (k < 0) return i <= w ? -(o_del + e_del * (i + 1)) : KSW_G_MINF;
		const int beg = i > w ? i - w : 0, end = i + w + 1 < qlen ? i + w + 1 : qlen;
		if (k < beg || k >= end) return KSW_G_MINF;
		const ksw_u2 v = z[((size_t)i * nqb + ((k >> 2) - (beg >> 2))) * T];
		const int c = k & 3;
		const uint32_t word = (c & 1) ? v.y : v.x;                 // pair B holds c1, c3; pair A c0, c2
		return (int)((c & 2) ? (word >> 16) : (word & 0xffffu)) - KSW_G_BIAS;
	}
	KSW_HD int mcell(int i, int k) const       // M(i,k) = H(i-1,k-1) + S(i,k)
	{
		const int d = hcell(i - 1, k - 1);
		int t = target[i], q = query[k];
		t = t > 4 ? 4 : t; q = q > 4 ? 4 : q;
		return d <= KSW_G_MINF / 2 ? KSW_G_MINF : d + mat[t * 5 + q];
	}
	KSW_HD int hload(int i, int k) const       // H(i,k) of a cell that IS inside the band: no checks, one load
	{
		const int beg = i > w ? i - w : 0;
		const ksw_u2 v = z[((size_t)i * nqb + ((k >> 2) - (beg >> 2))) * T];
		const uint32_t word = (k & 1) ? v.y : v.x;
		return (int)((k & 2) ? (word >> 16) : (word & 0xffffu)) - KSW_G_BIAS;
	}
	// the reference's walk (ksw.c:562-579); emit(r, op, len): r-th operation counted from the END of the alignment.
	// Per step one H load (the diagonal neighbour, which is the next cell's own H if the step is diagonal) and the two bases,
	// all independent of each other: one memory latency per step.  (i-1,k-1) is always inside the band when i, k > 0:
	// beg(i-1) >= beg(i)-1 and end(i-1) >= end(i)-1.
	template <class F>
	KSW_HD int run(F &&emit) const
	{
		int i = tlen - 1, k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;        // the last cell (ksw.c:565)
		int which = 0, n = 0, cur_op = -1, cur_len = 0, gapv = 0;
		const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
		int hcur = hload(i, k);
		while (i >= 0 && k >= 0) {
			const bool inb = i > 0 && k > 0;
			const int dv = hload(inb ? i - 1 : 0, inb ? k - 1 : 0);
			int t = target[i], q = query[k];
			t = t > 4 ? 4 : t; q = q > 4 ? 4 : q;
			const int sc = mat[t * 5 + q];
			const int bnd = i == 0 ? (k == 0 ? 0 : ((k <= w && k <= qlen) ? -(o_ins + e_ins * k) : KSW_G_MINF))
			                       : (i - 1 <= w ? -(o_del + e_del * i) : KSW_G_MINF);      // k == 0: H(i-1,-1)
			const int d = inb ? dv : bnd;
			const int m = d <= KSW_G_MINF / 2 ? KSW_G_MINF : d + sc;                       // M(i,k) = H(i-1,k-1) + S(i,k)
			if (which == 0) {
				const int h = hcur;
				if (h != m) {
					// E(i,k) = max over the rows above of M(i',k) - oe_del - (i-1-i') e_del: does it reach h?  (E beats F on ties)
					which = 2;
					const int top = k - w > 0 ? k - w : 0;
					for (int r = i - 1; r >= top; --r) {
						const int mm = mcell(r, k);
						if (mm > KSW_G_MINF / 2 && mm - oe_del - (i - 1 - r) * e_del == h) { which = 1; break; }
					}
				}
				gapv = h;                                                          // E(i,k) resp. F(i,k) if the path turns into a gap here
			} else if (which == 1) {
				if (m > KSW_G_MINF / 2 && m - oe_del == gapv) which = 0;               // opened here
				else gapv += e_del;                                                // extended: E(i,k) = E(i+1,k) + e_del
			} else {
				if (m > KSW_G_MINF / 2 && m - oe_ins == gapv) which = 0;
				else gapv += e_ins;
			}
			const int op = which == 0 ? 0 : (which == 1 ? 2 : 1);                  // push_cigar, ksw.c:486-499
			if (op == cur_op) ++cur_len;
			else { if (cur_op >= 0) emit(n++, cur_op, cur_len); cur_op = op; cur_len = 1; }
			if (which == 0) hcur = d;                                              // the diagonal neighbour is the next cell
			if (which != 2) --i;
			if (which != 1) --k;
		}
		if (i >= 0) { if (cur_op == 2) cur_len += i + 1; else { if (cur_op >= 0) emit(n++, cur_op, cur_len); cur_op = 2; cur_len = i + 1; } }
		if (k >= 0) { if (cur_op == 1) cur_len += k + 1; else { if (cur_op >= 0) emit(n++, cur_op, cur_len); cur_op = 1; cur_len = k + 1; } }
		if (cur_op >= 0) emit(n++, cur_op, cur_len);
		return n;
	}
};
