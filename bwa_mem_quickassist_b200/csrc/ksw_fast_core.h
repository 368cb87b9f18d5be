// ksw_fast_core.h — per-lane logic of the fast extension kernel, written once and compiled
// twice: as sm_100a device code inside ksw_fast.cu, and as plain C++ by the CPU emulation that
// tests/ uses to fuzz this exact source against the oracle where no GPU exists.  (The emulation
// is test infrastructure; the product path always runs the CUDA build.)
//
// One extension job per lane, semantics of bwa-0.7.8/ksw.c:379-476 row by row, but the cells of
// a row are computed four columns at a time in two s16x2 registers with the DPX instructions:
//
//   quad q = columns 4q..4q+3 = pair A (lo half = c0, hi half = c1) and pair B (lo = c3, hi = c2).
//   The B pair is stored half-swapped so that the serial F chain  c0 -> c1 -> c2 -> c3 -> c0'
//   walks lo -> hi -> hi -> lo -> lo' and never needs a shift between pairs.
//
//   per pair:  score  = PRMT(matrow, selector)                       1 ALU   (sign-extending byte lookup)
//              h'     = VIADDMNMX.S16x2(Hdiag, score, E)             H(i-1,j-1)+S vs E(i,j)      ksw.c:430-431
//              g      = VIADDMNMX.S16x2.RELU(h', -oe_ins, 0)
//              F      : 2 x VIADDMNMX.S16x2 + 1 PRMT                 F(j+1) = max(F(j)-e_ins, g(j))   (*)
//              h      = VIMNMX.S16x2(h', F)                                                      ksw.c:432
//              E'     = VIADDMNMX.S16x2(E, -e_del, RELU(h - oe_del)) 2 DPX                       ksw.c:436-439
//              (m,mj) : VIMNMX.S16x2 with predicate outputs + 2 predicated index moves          ksw.c:434-435
//              zero?  : VIMNMX3.S16x2 min over the row (1 per quad)  feeds the band trim        ksw.c:463-466
//   (*) the reference computes F(j+1) = max(F(j)-e_ins, max(H(j)-oe_ins,0)) with H = max(h',F);
//       since o_ins >= 0 implies F-oe_ins <= F-e_ins this equals max(F(j)-e_ins, h'(j)-oe_ins, 0)
//       in exact integer arithmetic (jobs with o_ins < 0 are routed to the generic kernel).
//
// Band edges that are not multiples of four are handled by "phantom" columns: their inputs are
// forced to (H = -8192, E = 0) so that they produce h = 0 on the left of the band and a decaying F
// on the right; DESIGN.md §kernel shows why neither can change (m, mj), the trim, or any value
// that is read later.  After each row three 16-bit stores restore the reference's edge writes
// (eh[beg].h = first-column value, eh[end].h = h1, eh[end].e = 0; ksw.c:429,446).
#pragma once
#include <stdint.h>
#include "ksw_dev.cuh"

#ifndef __CUDACC__
struct ksw_u2 { uint32_t x, y; };
struct ksw_u4 { uint32_t x, y, z, w; };
#else
typedef uint2 ksw_u2;
typedef uint4 ksw_u4;
#endif

// ----------------------------------------------------------------------------- DPX wrappers
namespace kswdpx {

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ uint32_t addmax2(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }
__device__ __forceinline__ uint32_t addmax2_relu(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2_relu(a, b, c); }
__device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
__device__ __forceinline__ uint32_t min3_2(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_s16x2(a, b, c); }
__device__ __forceinline__ uint32_t bmax2(uint32_t a, uint32_t b, bool &ge_hi, bool &ge_lo) { return __vibmax_s16x2(a, b, &ge_hi, &ge_lo); }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
	uint32_t d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
	return d;
}
#else
#define KSW_EMU static KSW_HD
KSW_EMU int16_t lo16(uint32_t v) { return (int16_t)(v & 0xffffu); }
KSW_EMU int16_t hi16(uint32_t v) { return (int16_t)(v >> 16); }
KSW_EMU uint32_t pk16(int lo, int hi) { return ((uint32_t)(uint16_t)(int16_t)lo) | ((uint32_t)(uint16_t)(int16_t)hi << 16); }
KSW_EMU int mx(int a, int b) { return a > b ? a : b; }
KSW_EMU int mn(int a, int b) { return a < b ? a : b; }
KSW_EMU uint32_t addmax2(uint32_t a, uint32_t b, uint32_t c)
{
	return pk16(mx((int16_t)(lo16(a) + lo16(b)), lo16(c)), mx((int16_t)(hi16(a) + hi16(b)), hi16(c)));
}
KSW_EMU uint32_t addmax2_relu(uint32_t a, uint32_t b, uint32_t c)
{
	return pk16(mx(mx((int16_t)(lo16(a) + lo16(b)), lo16(c)), 0), mx(mx((int16_t)(hi16(a) + hi16(b)), hi16(c)), 0));
}
KSW_EMU uint32_t max2(uint32_t a, uint32_t b) { return pk16(mx(lo16(a), lo16(b)), mx(hi16(a), hi16(b))); }
KSW_EMU uint32_t min3_2(uint32_t a, uint32_t b, uint32_t c)
{
	return pk16(mn(mn(lo16(a), lo16(b)), lo16(c)), mn(mn(hi16(a), hi16(b)), hi16(c)));
}
KSW_EMU uint32_t bmax2(uint32_t a, uint32_t b, bool &ge_hi, bool &ge_lo)
{
	ge_lo = lo16(a) >= lo16(b); ge_hi = hi16(a) >= hi16(b);
	return max2(a, b);
}
KSW_EMU uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)   // PTX prmt.b32, generic mode
{
	const uint64_t src = (uint64_t)a | ((uint64_t)b << 32);
	uint32_t d = 0;
	for (int k = 0; k < 4; ++k) {
		const uint32_t c = (s >> (4 * k)) & 0xf;
		uint32_t byte = (uint32_t)(src >> (8 * (c & 7))) & 0xff;
		if (c & 8) byte = (byte & 0x80) ? 0xff : 0x00;
		d |= byte << (8 * k);
	}
	return d;
}
#endif

} // namespace kswdpx

// ----------------------------------------------------------------------------- lane state
struct KswFastConst {
	uint32_t neg_ei, neg_oei, neg_ed, neg_oed;   // both halves: -e_ins, -(o_ins+e_ins), -e_del, -(o_del+e_del)
	int32_t o_del, e_del, e_ins, oe_ins, zdrop;
};

struct KswFastLane {
	// the job
	const uint32_t *t2;      // packed target
	const uint32_t *tn;      // target N mask or nullptr
	int32_t qlen, tlen, h0, w;
	uint32_t idx;
	// running state (ksw.c:408-410)
	int32_t i, lo, hi;
	int32_t best, best_i, best_j, end_i, end_sc, off;
	uint32_t cells;          // visited DP cells: sum over executed rows of (hi - lo), the GCUPS unit
	uint32_t tw, tw_next;    // current / prefetched target word (16 bases each)
};

#define KSW_NEGPK 0xE000E000u          /* -8192 in both halves */
#define KSW_FAST_QUADS(qlen) (((qlen) >> 2) + 1)   /* quads that cover columns 0..qlen */

static KSW_HD uint32_t ksw_pk2(int v) { return ((uint32_t)v & 0xffffu) | ((uint32_t)v << 16); }

static KSW_HD void ksw_fast_make_const(const KswParams &P, KswFastConst &K)
{
	K.neg_ei = ksw_pk2(-P.e_ins); K.neg_oei = ksw_pk2(-(P.o_ins + P.e_ins));
	K.neg_ed = ksw_pk2(-P.e_del); K.neg_oed = ksw_pk2(-(P.o_del + P.e_del));
	K.o_del = P.o_del; K.e_del = P.e_del; K.e_ins = P.e_ins; K.oe_ins = P.o_ins + P.e_ins; K.zdrop = P.zdrop;
}

// row t of the scoring matrix as PRMT source bytes: x = mat[t][0..3], y = mat[t][4] (upper bytes 0)
static KSW_HD ksw_u2 ksw_fast_matrow(const KswParams &P, int t)
{
	ksw_u2 r;
	r.x = ((uint32_t)(uint8_t)P.mat[t * 5 + 0]) | ((uint32_t)(uint8_t)P.mat[t * 5 + 1] << 8) |
	      ((uint32_t)(uint8_t)P.mat[t * 5 + 2] << 16) | ((uint32_t)(uint8_t)P.mat[t * 5 + 3] << 24);
	r.y = (uint32_t)(uint8_t)P.mat[t * 5 + 4];
	return r;
}

// halfword index of column c's H inside its quad's uint4 {HA, EA, HB, EB}; E is at +2
static KSW_HD int ksw_fast_hslot(int c) { const int k = c & 3; return k < 2 ? k : 7 - k; }

// Shared-memory view of one lane: quad q lives at hq[q*T]; the PRMT selectors of its pair A / pair B
// at sa[q*T] / sb[q*T] (two separate 16-bit arrays so that each is one LDS.U16 and no ALU extract).
template <int T>
struct KswFastMem {
	ksw_u4 *hq;
	uint16_t *sa, *sb;
	KSW_HD uint16_t *h16(int c) const { return reinterpret_cast<uint16_t *>(&hq[(c >> 2) * T]) + ksw_fast_hslot(c); }
};

// ----------------------------------------------------------------------------- job setup
// Fills row -1 (ksw.c:394-396), clears E, builds the PRMT selectors from the 2-bit query.
template <int T>
static KSW_HD void ksw_fast_setup(KswFastLane &L, const KswFastMem<T> &M, const KswFastConst &K, const DevJob &jb,
                                  const uint32_t *pool, const uint32_t *npool)
{
	const uint32_t *q2 = pool + (size_t)jb.seq_off * 4;
	const uint32_t *qn = (jb.flags & KSW_FLAG_QN) ? npool + jb.nmask_off : nullptr;
	L.t2 = q2 + ksw_words2(jb.qlen);
	L.tn = (jb.flags & KSW_FLAG_TN) ? npool + jb.nmask_off + ((jb.flags & KSW_FLAG_QN) ? ksw_words1(jb.qlen) : 0) : nullptr;
	L.qlen = jb.qlen; L.tlen = jb.tlen; L.h0 = jb.h0; L.w = jb.w; L.idx = jb.idx;
	L.i = 0; L.lo = 0; L.hi = jb.qlen;
	L.best = jb.h0; L.best_i = -1; L.best_j = -1; L.end_i = -1; L.end_sc = -1; L.off = 0;
	L.cells = 0;
	L.tw = jb.tlen > 0 ? L.t2[0] : 0u;
	L.tw_next = jb.tlen > 16 ? L.t2[1] : 0u;

	const int nq = KSW_FAST_QUADS(jb.qlen);
	const int h0 = jb.h0;
	uint32_t qw = 0, nw = 0;
	for (int q = 0; q < nq; ++q) {
		const int c0 = q << 2;
		if ((q & 3) == 0) qw = (c0 < jb.qlen) ? q2[c0 >> 4] : 0u;           // 16 bases = 4 quads per word
		if (qn && (q & 7) == 0) nw = (c0 < jb.qlen) ? qn[c0 >> 5] : 0u;
		int hv[4];
		uint32_t sb[4];
#ifdef __CUDACC__
#pragma unroll
#endif
		for (int k = 0; k < 4; ++k) {
			const int c = c0 + k;
			int v = c == 0 ? h0 : h0 - K.oe_ins - (c - 1) * K.e_ins;         // closed form of ksw.c:394-396
			hv[k] = v > 0 ? v : 0;
			uint32_t code = (qw >> (((c & 15)) << 1)) & 3u;
			if (qn && ((nw >> (c & 31)) & 1u)) code = 4u;
			if (c >= jb.qlen) code = 0u;
			sb[k] = code | ((8u | code) << 4);                                // byte lookup + sign replicate
		}
		ksw_u4 v4;
		v4.x = (uint32_t)hv[0] | ((uint32_t)hv[1] << 16);   // pair A: lo c0, hi c1
		v4.y = 0u;
		v4.z = (uint32_t)hv[3] | ((uint32_t)hv[2] << 16);   // pair B: lo c3, hi c2
		v4.w = 0u;
		M.hq[q * T] = v4;
		M.sa[q * T] = (uint16_t)(sb[0] | (sb[1] << 8));
		M.sb[q * T] = (uint16_t)(sb[3] | (sb[2] << 8));
	}
}

// ----------------------------------------------------------------------------- one row
struct KswFastRowRegs {       // registers carried along a row
	uint32_t X;               // F entering the next column (lo half at quad entry)
	uint32_t Hc;              // lo half = H(i, c0-1): carry for the shifted H store
	uint32_t m, zmin;         // per-half running max / min of the row
	int mjl, mjh;             // last column where the lo / hi half reached its running max
	uint32_t hA, hB;          // H of the last processed quad
};

// One quad (4 cells).  EDGE quads (first / last of the band) mask their out-of-band columns.
template <int T, bool EDGE>
static KSW_HD void ksw_fast_quad(KswFastRowRegs &R, const KswFastMem<T> &M, const KswFastConst &K, const ksw_u2 mr,
                                 const int q, const int lo, const int hi)
{
	using namespace kswdpx;
	ksw_u4 v = M.hq[q * T];
	const uint32_t selA = M.sa[q * T], selB = M.sb[q * T];
	const int c0 = q << 2;
	uint32_t keepA = 0xffffffffu, keepB = 0xffffffffu;
	if (EDGE) {
		// in-band test per column; phantom columns get H = -8192, E = 0
		const uint32_t in0 = (uint32_t)(c0 >= lo && c0 < hi), in1 = (uint32_t)(c0 + 1 >= lo && c0 + 1 < hi);
		const uint32_t in2 = (uint32_t)(c0 + 2 >= lo && c0 + 2 < hi), in3 = (uint32_t)(c0 + 3 >= lo && c0 + 3 < hi);
		keepA = in0 * 0xffffu + in1 * 0xffff0000u;
		keepB = in3 * 0xffffu + in2 * 0xffff0000u;
		v.x = (v.x & keepA) | (KSW_NEGPK & ~keepA); v.y &= keepA;
		v.z = (v.z & keepB) | (KSW_NEGPK & ~keepB); v.w &= keepB;
	}
	const uint32_t scA = prmt(mr.x, mr.y, selA), scB = prmt(mr.x, mr.y, selB);
	const uint32_t hpA = addmax2(v.x, scA, v.y), hpB = addmax2(v.z, scB, v.w);
	// relu(h' - oe_ins): the third operand only has to be <= 0, a live constant saves a zero register
	const uint32_t gA = addmax2_relu(hpA, K.neg_oei, K.neg_oei), gB = addmax2_relu(hpB, K.neg_oei, K.neg_oei);
	// F chain: c0 (A.lo) -> c1 (A.hi) -> c2 (B.hi) -> c3 (B.lo) -> next quad
	const uint32_t t1 = addmax2(R.X, K.neg_ei, gA);        // lo = F(c1)
	const uint32_t FA = prmt(R.X, t1, 0x5410u);            // (F(c0), F(c1))
	const uint32_t t2 = addmax2(FA, K.neg_ei, gA);         // hi = F(c2)
	const uint32_t t3 = addmax2(t2, K.neg_ei, gB);         // hi = F(c3)
	const uint32_t FB = prmt(t2, t3, 0x3276u);             // (lo = F(c3), hi = F(c2))
	R.X = addmax2(FB, K.neg_ei, gB);                       // lo = F(c0 of the next quad)
	const uint32_t hA = max2(hpA, FA), hB = max2(hpB, FB);
	// E(i+1,j) = max(E - e_del, relu(H - oe_del))
	const uint32_t eA = addmax2(v.y, K.neg_ed, addmax2_relu(hA, K.neg_oed, K.neg_oed));
	const uint32_t eB = addmax2(v.w, K.neg_ed, addmax2_relu(hB, K.neg_oed, K.neg_oed));
	// row maximum with last-index-wins ties, tracked per half (each half sees its columns in rising order)
	bool ph, pl;
	R.m = bmax2(hA, R.m, ph, pl);
	if (pl) R.mjl = c0;
	if (ph) R.mjh = c0 + 1;
	R.m = bmax2(hB, R.m, ph, pl);
	if (ph) R.mjh = c0 + 2;
	if (pl) R.mjl = c0 + 3;
	// zero detector over the in-band cells
	if (EDGE) R.zmin = min3_2(R.zmin, hA | (~keepA & 0x7fff7fffu), hB | (~keepB & 0x7fff7fffu));
	else R.zmin = min3_2(R.zmin, hA, hB);
	// store: eh[j].h = H(i, j-1), eh[j].e = E(i+1, j)
	ksw_u4 o;
	o.x = prmt(R.Hc, hA, 0x5410u);                         // (H(c0-1), H(c0))
	o.y = eA;
	o.z = prmt(hA, hB, 0x3276u);                           // (lo: H(c2) for column c3, hi: H(c1) for column c2)
	o.w = eB;
	M.hq[q * T] = o;
	R.Hc = hB;                                             // lo half = H(c3)
	R.hA = hA; R.hB = hB;
}

// Processes row L.i.  Returns true when the job is finished (results are then in L).
template <int T>
static KSW_HD bool ksw_fast_row(KswFastLane &L, const KswFastMem<T> &M, const KswFastConst &K, const ksw_u2 *mrow)
{
	const int i = L.i;
	if (i >= L.tlen) return true;
	// target base of this row
	if ((i & 15) == 0 && i) {
		L.tw = L.tw_next;
		const int nx = (i >> 4) + 1;
		if ((nx << 4) < L.tlen) L.tw_next = L.t2[nx];
	}
	int t = (int)((L.tw >> ((i & 15) << 1)) & 3u);
	if (L.tn && ((L.tn[i >> 5] >> (i & 31)) & 1u)) t = 4;
	const ksw_u2 mr = mrow[t];

	int left0 = L.h0 - (K.o_del + K.e_del * (i + 1));          // first-column value, used even when lo>0 (ksw.c:415-416)
	left0 = left0 > 0 ? left0 : 0;
	int lo = L.lo, hi = L.hi;
	lo = lo > i - L.w ? lo : i - L.w;                           // ksw.c:418-420
	hi = hi < i + L.w + 1 ? hi : i + L.w + 1;
	hi = hi < L.qlen ? hi : L.qlen;
	if (hi <= lo) {
		// empty row: the reference's loop variable stays at beg, so "j == qlen" means lo == qlen (ksw.c:447);
		// then m == 0 ends the job (ksw.c:451)
		if (lo == L.qlen) {
			if (left0 >= L.end_sc) L.end_i = i;
			L.end_sc = L.end_sc > left0 ? L.end_sc : left0;
		}
		return true;
	}

	L.cells += (uint32_t)(hi - lo);
	const int q0 = lo >> 2, q1 = (hi - 1) >> 2;
	KswFastRowRegs R;
	R.X = 0; R.Hc = 0; R.m = 0; R.zmin = 0x7fff7fffu; R.mjl = -1; R.mjh = -1; R.hA = 0; R.hB = 0;
	ksw_fast_quad<T, true>(R, M, K, mr, q0, lo, hi);
	if (q1 > q0) {
#ifdef __CUDACC__
#pragma unroll 2
#endif
		for (int q = q0 + 1; q < q1; ++q) ksw_fast_quad<T, false>(R, M, K, mr, q, lo, hi);
		ksw_fast_quad<T, true>(R, M, K, mr, q1, lo, hi);
	}
	// H(i, hi-1): the reference's h1 after the loop
	int left;
	{
		const int k = (hi - 1) & 3;
		const uint32_t r = k < 2 ? R.hA : R.hB;
		left = (int)((k == 0 || k == 3) ? (r & 0xffffu) : (r >> 16));
	}
	// edge writes of the reference (ksw.c:429 for column lo, ksw.c:446 for column hi)
	*M.h16(lo) = (uint16_t)left0;
	{
		uint16_t *p = M.h16(hi);
		p[0] = (uint16_t)left;
		p[2] = 0;                                              // E slot is two halfwords after the H slot
	}
	if (hi == L.qlen) {                                        // ksw.c:447-450, ties -> last row
		if (left >= L.end_sc) L.end_i = i;
		L.end_sc = L.end_sc > left ? L.end_sc : left;
	}
	const int m_lo = (int)(int16_t)(R.m & 0xffffu), m_hi = (int)(int16_t)(R.m >> 16);
	const int rmax = m_lo > m_hi ? m_lo : m_hi;
	const int rarg = m_lo > m_hi ? R.mjl : (m_hi > m_lo ? R.mjh : (R.mjl > R.mjh ? R.mjl : R.mjh));
	if (rmax == 0) return true;                                // ksw.c:451
	if (rmax > L.best) {                                       // ksw.c:452-454
		L.best = rmax; L.best_i = i; L.best_j = rarg;
		const int d = rarg > i ? rarg - i : i - rarg;
		L.off = L.off > d ? L.off : d;
	} else if (K.zdrop > 0) {                                  // ksw.c:455-461
		const int di = i - L.best_i, dj = rarg - L.best_j;
		if (di > dj) { if (L.best - rmax - (di - dj) * K.e_del > K.zdrop) return true; }
		else         { if (L.best - rmax - (dj - di) * K.e_ins > K.zdrop) return true; }
	}
	// band trim (ksw.c:463-466)
	const bool any_zero = ((R.zmin & 0xffffu) == 0u) | ((R.zmin >> 16) == 0u);
	if (!any_zero) {
		// every eh[j].h for j in (lo, hi] is non-zero, so only the first-column slot eh[lo].h can stop the
		// downward scan, and the upward scan runs off the end
		L.lo = left0 ? lo : lo + 1;
		L.hi = hi + 1;
	} else {
		int j;
		for (j = rarg; j >= lo && *M.h16(j); --j) ;
		L.lo = j + 1;
		for (j = rarg + 2; j <= hi && *M.h16(j); ++j) ;
		L.hi = j;
	}
	L.i = i + 1;
	return L.i >= L.tlen;
}

static KSW_HD void ksw_fast_result(const KswFastLane &L, DevRes &r)
{
	r.score = L.best; r.qle = L.best_j + 1; r.tle = L.best_i + 1;
	r.gtle = L.end_i + 1; r.gscore = L.end_sc; r.max_off = L.off;
}
