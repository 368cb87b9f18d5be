// ksw_fast_core.h — per-lane logic of the fast extension kernel, written once and compiled
// twice: as sm_100a device code inside ksw_fast.cu, and as plain C++ by the CPU emulation that
// tests/ uses to fuzz this exact source against the oracle where no GPU exists.  (The emulation
// is test infrastructure; the product path always runs the CUDA build.)
//
// One extension job per lane, semantics of bwa-0.7.8/ksw.c:379-476 row by row, but the cells of
// a row are computed four columns at a time in two s16x2 registers with the DPX instructions:
//
//   quad q = columns c0..c3 = 4q..4q+3 = pair A (lo half = c0, hi half = c2) and pair B (lo = c1, hi = c3).
//   With this interleaving the serial F chain c0 -> c1 -> c2 -> c3 -> c0' stays inside one half for two steps
//   (lo: c0 -> c1 -> c2, hi: c2 -> c3 -> c0'), and the two moves between the halves are whole-half moves that
//   integer multiply-adds do on the FMA pipe (x * 65536 + y, x >> 16 as a multiply-high), so the chain costs
//   four DPX instructions per quad and no byte permutes on the ALU pipe.
//
//   per quad:  score   = PRMT(matrow, selector)                     2 ALU   (sign-extending byte lookup)
//              h'      = VIADDMNMX.S16x2(Hdiag, score, E)           2 ALU   H(i-1,j-1)+S vs E(i,j)      ksw.c:430-431
//              Fs      : 4 x VIADDMNMX.S16x2 + 2 IMAD               4 ALU   Fs(j+1) = max(Fs(j)-e_ins, h'(j))   (*)
//              h       = VIADDMNMX.S16x2(Fs, -oe_ins, h')           2 ALU                               ksw.c:432
//              E'      = VIMNMX3.S16x2(E - e_del, h - oe_del, 0)    2 ALU   the two subtractions are plain 32-bit IADDs
//                        that the compiler places on the FMA pipe.  They are borrow-free because every H/E/F
//                        value is kept with a constant bias B = o_del+e_del in both halves (B instead of 0 is
//                        the floor of all the max() clamps)                                     ksw.c:436-439
//              (m,mj)  : KEYED jobs (qlen <= 124, scores < 512): key = h*128 + column relative to the current quad
//                        (1 IMAD per pair on the FMA pipe), m = VIADDMNMX.U16x2(m, -4, keyA), m = VIMNMX.U16x2(m, keyB)
//                        — the running maximum is moved back by four columns per quad instead of the column counters
//                        being moved forward; the larger column wins ties, as in the reference;
//                        other jobs: VIMNMX.S16x2 with predicate outputs + 2 predicated index moves   ksw.c:434-435
//              zero?   : VIMNMX3.S16x2 min over the row (1 per quad)  feeds the band trim        ksw.c:463-466
//              store   : eh[j].h = H(i,j-1) is pair A' = (H(c0-1), H(c1)) = one funnel shift of (previous hB, hB), and
//                        pair B' = (H(c0), H(c2)) = hA as it is
//   (*) Fs = F + oe_ins, WITHOUT the reference's floor at 0: the reference computes
//       F(j+1) = max(F(j)-e_ins, max(H(j)-oe_ins,0)) with H = max(h',F).  Since o_ins >= 0 implies F-oe_ins <= F-e_ins,
//       that equals max(F(j)-e_ins, h'(j)-oe_ins, 0).  Dropping the floor gives F~ <= F with max(F~,0) = F by induction
//       (if F~(j) < 0 = F(j) then F~(j)-e_ins < 0, so max(F~(j+1),0) = max(h'(j)-oe_ins,0) = F(j+1)), and F is only ever
//       used through max(h', F) with h' >= E >= 0, so H and E are unchanged; F~ >= min h' - oe_ins is bounded below.
//       (jobs with o_ins < 0 are routed to the generic kernel).
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
__device__ __forceinline__ uint32_t maxu2(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
__device__ __forceinline__ uint32_t addmaxu2(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_u16x2(a, b, c); }
__device__ __forceinline__ uint32_t max3u2(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }
__device__ __forceinline__ uint32_t min3_2(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_s16x2(a, b, c); }
__device__ __forceinline__ uint32_t max3_2(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
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
KSW_EMU uint32_t maxu2(uint32_t a, uint32_t b)
{
	const uint32_t al = a & 0xffffu, bl = b & 0xffffu, ah = a >> 16, bh = b >> 16;
	return (al > bl ? al : bl) | ((ah > bh ? ah : bh) << 16);
}
KSW_EMU uint32_t addmaxu2(uint32_t a, uint32_t b, uint32_t c)   // max(a + b, c) per unsigned half, the sum wraps mod 2^16
{
	const uint32_t sl = (a + b) & 0xffffu, sh = ((a >> 16) + (b >> 16)) & 0xffffu, cl = c & 0xffffu, ch = c >> 16;
	return (sl > cl ? sl : cl) | ((sh > ch ? sh : ch) << 16);
}
KSW_EMU uint32_t max3u2(uint32_t a, uint32_t b, uint32_t c) { return maxu2(maxu2(a, b), c); }
KSW_EMU uint32_t min3_2(uint32_t a, uint32_t b, uint32_t c)
{
	return pk16(mn(mn(lo16(a), lo16(b)), lo16(c)), mn(mn(hi16(a), hi16(b)), hi16(c)));
}
KSW_EMU uint32_t max3_2(uint32_t a, uint32_t b, uint32_t c)
{
	return pk16(mx(mx(lo16(a), lo16(b)), lo16(c)), mx(mx(hi16(a), hi16(b)), hi16(c)));
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
#define KSW_KEY_RA 0x00020000u      /* columns of pair A relative to its quad: lo 0, hi 2 */
#define KSW_KEY_RB 0x00030001u      /* pair B: lo 1, hi 3 */
#define KSW_KEY_DEC 0xfffcfffcu     /* -4 in both halves: the running key maximum moves back one quad */
#define KSW_KEY_STEP 0x00040004u    /* the same step as a 32-bit subtrahend */
#define KSW_KEY_INIT 0x00800080u    /* 4 * 32 quads: never wraps below 0 and never beats a real key (>= 128 * B, B >= 1) */

struct KswFastConst {
	uint32_t neg_ei, neg_oei;                    // both halves: -e_ins, -(o_ins+e_ins)
	uint32_t neg_2ei;                            // both halves: -2*e_ins
	uint32_t neg_ed, neg_oed;                    // both halves: -e_del, -(o_del+e_del)
	uint32_t ed32, oed32;                        // both halves: e_del, o_del+e_del (subtracted with 32-bit IADDs)
	uint32_t Bpk;                                // both halves: the bias B = o_del+e_del carried by every H/E/F value
	int32_t B;
	int32_t o_del, e_del, e_ins, oe_ins, zdrop;
	// KEYED arg-max: columns of pair A / pair B relative to their quad (lo 0, hi 2 / lo 1, hi 3).  Held in registers
	// the compiler cannot fold (built from a kernel-parameter byte that is always 0), so that key = h*128 + column
	// stays one multiply-add on the FMA pipe instead of a shift-add with an immediate on the ALU pipe.
	uint32_t keyRA, keyRB;
	uint32_t one;                                // 1, unfoldable like keyRA / keyRB
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

static KSW_HD uint32_t ksw_pk2(int v) { return ((uint32_t)v & 0xffffu) | ((uint32_t)v << 16); }

static KSW_HD void ksw_fast_make_const(const KswParams &P, KswFastConst &K)
{
	K.neg_ei = ksw_pk2(-P.e_ins); K.neg_oei = ksw_pk2(-(P.o_ins + P.e_ins)); K.neg_2ei = ksw_pk2(-2 * P.e_ins);
	K.ed32 = ksw_pk2(P.e_del); K.oed32 = ksw_pk2(P.o_del + P.e_del);
	K.neg_ed = ksw_pk2(-P.e_del); K.neg_oed = ksw_pk2(-(P.o_del + P.e_del));
	K.B = P.o_del + P.e_del; K.Bpk = ksw_pk2(K.B);
	K.o_del = P.o_del; K.e_del = P.e_del; K.e_ins = P.e_ins; K.oe_ins = P.o_ins + P.e_ins; K.zdrop = P.zdrop;
	K.keyRA = KSW_KEY_RA + (uint32_t)(uint8_t)P.pad[0]; K.keyRB = KSW_KEY_RB + (uint32_t)(uint8_t)P.pad[0];
	K.one = 1u + (uint32_t)(uint8_t)P.pad[0];
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

// halfword index of column c's H inside its quad's uint4 {HA, EA, HB, EB} (A: lo c0, hi c2; B: lo c1, hi c3); E is at +2
static KSW_HD int ksw_fast_hslot(int c) { const int k = c & 3; return (k >> 1) + ((k & 1) << 2); }

// Edge look-up table (one copy per CTA in shared memory), indexed by r = a column position inside a quad, 0..4:
//   ge / lt   : halfword masks of the quad's columns >= r / < r in the (A: lo c0, hi c2; B: lo c1, hi c3) layout
//   only      : halfword mask of column r alone (all-zero for r == 4)
//   sel_left  : PRMT selector that extracts H of column r-1 from (hA, hB) into the low half, zero-extended
struct KswFastEdge {
	uint32_t geA, geB, onlyA, onlyB;
	uint32_t ltA, ltB, sel_left, pad;
};

static KSW_HD void ksw_fast_edge_entry(int r, KswFastEdge &e)
{
	// column k of a quad: k=0 -> A.lo, 1 -> B.lo, 2 -> A.hi, 3 -> B.hi
	const uint32_t colA[4] = {0x0000ffffu, 0u, 0xffff0000u, 0u};
	const uint32_t colB[4] = {0u, 0x0000ffffu, 0u, 0xffff0000u};
	e.geA = e.geB = e.ltA = e.ltB = 0u;
	for (int k = 0; k < 4; ++k) {
		if (k >= r) { e.geA |= colA[k]; e.geB |= colB[k]; }
		else        { e.ltA |= colA[k]; e.ltB |= colB[k]; }
	}
	e.onlyA = r < 4 ? colA[r] : 0u;
	e.onlyB = r < 4 ? colB[r] : 0u;
	// bytes of (hA, hB): 0,1 = A.lo (c0)  2,3 = A.hi (c2)  4,5 = B.lo (c1)  6,7 = B.hi (c3); the upper result bytes
	// replicate the sign (always 0: H >= 0), i.e. selector nibbles {lo, hi, 8|hi, 8|hi}
	const uint32_t sel[4] = {0x9910u, 0xdd54u, 0xbb32u, 0xff76u};
	e.sel_left = r >= 1 ? sel[r - 1] : sel[0];
	e.pad = 0u;
}

// Shared-memory view of one lane: quad q lives at hq[q*T]; the PRMT selectors of its two pairs in one word
// at sq[q*T] (low 16 bits: pair A, high 16 bits: pair B; PRMT ignores selector bits above 15).
template <int T>
struct KswFastMem {
	ksw_u4 *hq;
	uint32_t *sq;
	const KswFastEdge *edge;     // 5 entries, shared by the CTA
	KSW_HD uint16_t *h16(int c) const { return reinterpret_cast<uint16_t *>(&hq[(c >> 2) * T]) + ksw_fast_hslot(c); }
};

// ----------------------------------------------------------------------------- job setup
// Lane registers of a freshly fetched job (ksw.c:408-410) and the first target words.
static KSW_HD void ksw_fast_init_lane(KswFastLane &L, const DevJob &jb, const uint32_t *pool, const uint32_t *npool)
{
	const uint32_t *q2 = pool + (size_t)jb.seq_off * 4;
	L.t2 = q2 + ksw_words2(jb.qlen);
	L.tn = (jb.flags & KSW_FLAG_TN) ? npool + jb.nmask_off + ((jb.flags & KSW_FLAG_QN) ? ksw_words1(jb.qlen) : 0) : nullptr;
	L.qlen = jb.qlen; L.tlen = jb.tlen; L.h0 = jb.h0; L.w = jb.w; L.idx = jb.idx;
	L.i = 0; L.lo = 0; L.hi = jb.qlen;
	L.best = jb.h0; L.best_i = -1; L.best_j = -1; L.end_i = -1; L.end_sc = -1; L.off = 0;
	L.cells = 0;
	L.tw = jb.tlen > 0 ? L.t2[0] : 0u;
	L.tw_next = jb.tlen > 16 ? L.t2[1] : 0u;
}

// Shared-memory state of one job: row -1 (ksw.c:394-396), E = 0, and the PRMT selectors built from the 2-bit
// query.  COOPERATIVE: the quads of the job are dealt round-robin to the `n_helpers` lanes that call this with
// helper = 0..n_helpers-1 (on the GPU all 32 lanes of the warp build the state of one lane's job, so a lane that
// fetches a job alone does not make the other 31 idle through the whole setup); `owner` is the lane the job
// belongs to, hq/sq are the CTA's arrays (not offset by lane).
template <int T>
static KSW_HD void ksw_fast_setup_quads(ksw_u4 *hq, uint32_t *sq, const int owner, const int helper, const int n_helpers,
                                        const KswFastConst &K, const uint32_t seq_off, const int qlen, const int h0,
                                        const uint32_t flags, const uint32_t nmask_off,
                                        const uint32_t *pool, const uint32_t *npool)
{
	const uint32_t *q2 = pool + (size_t)seq_off * 4;
	const uint32_t *qn = (flags & KSW_FLAG_QN) ? npool + nmask_off : nullptr;
	const int nq = KSW_FAST_QUADS(qlen);
	for (int q = helper; q < nq; q += n_helpers) {
		const int c0 = q << 2;
		const uint32_t qw = (c0 < qlen) ? q2[c0 >> 4] : 0u;                  // 16 bases = 4 quads per word
		const uint32_t nw = (qn && c0 < qlen) ? qn[c0 >> 5] : 0u;
		int hv[4];
		uint32_t sb[4];
#ifdef __CUDACC__
#pragma unroll
#endif
		for (int k = 0; k < 4; ++k) {
			const int c = c0 + k;
			int v = c == 0 ? h0 : h0 - K.oe_ins - (c - 1) * K.e_ins;         // closed form of ksw.c:394-396
			hv[k] = (v > 0 ? v : 0) + K.B;
			uint32_t code = (qw >> (((c & 15)) << 1)) & 3u;
			if ((nw >> (c & 31)) & 1u) code = 4u;
			if (c >= qlen) code = 0u;
			sb[k] = code | ((8u | code) << 4);                                // byte lookup + sign replicate
		}
		ksw_u4 v4;
		v4.x = (uint32_t)hv[0] | ((uint32_t)hv[2] << 16);   // pair A: lo c0, hi c2
		v4.y = K.Bpk;                                       // E = 0 (+ bias)
		v4.z = (uint32_t)hv[1] | ((uint32_t)hv[3] << 16);   // pair B: lo c1, hi c3
		v4.w = K.Bpk;
		hq[q * T + owner] = v4;
		sq[q * T + owner] = (sb[0] | (sb[2] << 8)) | ((sb[1] | (sb[3] << 8)) << 16);
	}
}

// ----------------------------------------------------------------------------- one row
struct KswFastRowRegs {       // registers carried along a row
	uint32_t X;               // lo half = Fs entering the quad's first column, hi half = 0
	uint32_t Hc;              // hi half = H(i, c0-1): carry for the shifted H store (the previous quad's hB)
	uint32_t m, zmin;         // running max (keys relative to the current quad if KEYED) / running min of the row's interior quads, per half
	uint32_t zminE;           // running min of the (at most two) edge quads' in-band cells
	int mjl, mjh;             // !KEYED: last column where the lo / hi half reached its running max
	uint32_t hA, hB;          // H of the last processed quad
};

#if defined(__CUDA_ARCH__)
static __device__ __forceinline__ uint32_t ksw_hi16_of(uint32_t w) { return __umulhi(w, 0x10000u); }   // FMA pipe, not ALU
// (a.lo << 16) + b: moves a's low half into the high half on the FMA pipe; b's high half must be 0
static __device__ __forceinline__ uint32_t ksw_lo_to_hi_add(uint32_t a, uint32_t b)
{
	uint32_t d;
	asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(d) : "r"(a), "r"(b));
	return d;
}
#else
static inline uint32_t ksw_hi16_of(uint32_t w) { return w >> 16; }
static inline uint32_t ksw_lo_to_hi_add(uint32_t a, uint32_t b) { return (a << 16) + b; }
#endif

// One quad (4 cells).  EDGE: 0 = interior quad, 1 = a quad at the band edge: its out-of-band ("phantom")
// columns are masked with keepA/keepB, the reference's edge writes are folded into its store.
template <int T, bool KEYED, bool EDGE>
static KSW_HD void ksw_fast_quad(KswFastRowRegs &R, ksw_u4 *dst, const KswFastConst &K, const ksw_u2 mr,
                                 const int q, ksw_u4 v, const uint32_t sw, const uint32_t keepA, const uint32_t keepB,
                                 const uint32_t firstA, const uint32_t firstB, const uint32_t left0pk,
                                 const uint32_t endA, const uint32_t endB)
{
	using namespace kswdpx;
	// (v, sw) = hq[q], sq[q], loaded by the caller one quad ahead so that the LDS latency hides behind the previous quad
	if (EDGE) {
		// phantom columns get H = -8192, E = 0 (+ bias)
		v.x = (v.x & keepA) | (KSW_NEGPK & ~keepA); v.y = (v.y & keepA) | (K.Bpk & ~keepA);
		v.z = (v.z & keepB) | (KSW_NEGPK & ~keepB); v.w = (v.w & keepB) | (K.Bpk & ~keepB);
	}
	const uint32_t scA = prmt(mr.x, mr.y, sw), scB = prmt(mr.x, mr.y, ksw_hi16_of(sw));
	const uint32_t hpA = addmax2(v.x, scA, v.y), hpB = addmax2(v.z, scB, v.w);   // (h'(c0), h'(c2)), (h'(c1), h'(c3))
	// F chain on Fs = F + oe_ins (no floor, see the header): lo half c0 -> c1 -> c2, hi half c2 -> c3 -> c0'
#ifdef KSW_V_CHAIN2
	// two columns per chain step: G = (max(h'(c0)-e, h'(c1)), max(h'(c2)-e, h'(c3))) is off the chain, and
	// Fs(c+2) = max(Fs(c) - 2e, G) — the serial part of a quad is two DPX instructions and the two half moves
	const uint32_t G = addmax2(hpA, K.neg_ei, hpB);
	const uint32_t u2 = addmax2(R.X, K.neg_2ei, G);        // lo = Fs(c2)
	const uint32_t FA = ksw_lo_to_hi_add(u2, R.X);         // (Fs(c0), Fs(c2))
	const uint32_t u4 = addmax2(FA, K.neg_2ei, G);         // hi = Fs(c0 of the next quad)
	const uint32_t FB = addmax2(FA, K.neg_ei, hpA);        // (Fs(c1), Fs(c3))
	R.X = ksw_hi16_of(u4);
#else
	const uint32_t u1 = addmax2(R.X, K.neg_ei, hpA);       // lo = Fs(c1)
	const uint32_t u2 = addmax2(u1, K.neg_ei, hpB);        // lo = Fs(c2)
	const uint32_t FA = ksw_lo_to_hi_add(u2, R.X);         // (Fs(c0), Fs(c2))
	const uint32_t FB = addmax2(FA, K.neg_ei, hpA);        // (Fs(c1), Fs(c3))
	const uint32_t u4 = addmax2(FB, K.neg_ei, hpB);        // hi = Fs(c0 of the next quad)
	R.X = ksw_hi16_of(u4);
#endif
	const uint32_t hA = addmax2(FA, K.neg_oei, hpA), hB = addmax2(FB, K.neg_oei, hpB);   // max(F, h')
	// E(i+1,j) = max(E - e_del, H - oe_del, 0); every half is >= B >= oe_del, so the 32-bit subtractions cannot borrow
#ifdef KSW_V_E2
	// two fused add-max per pair instead of two subtractions and a three-way maximum: one instruction fewer per pair, all on the ALU pipe
	uint32_t eA = addmax2(v.y, K.neg_ed, addmax2(hA, K.neg_oed, K.Bpk));
	uint32_t eB = addmax2(v.w, K.neg_ed, addmax2(hB, K.neg_oed, K.Bpk));
#else
	uint32_t eA = max3_2(v.y - K.ed32, hA - K.oed32, K.Bpk);
	uint32_t eB = max3_2(v.w - K.ed32, hB - K.oed32, K.Bpk);
#endif
	// row maximum, ties to the last column (ksw.c:434)
	if (KEYED) {
#ifdef KSW_V_KEYMAX3
		// one three-way maximum instead of two two-way ones; the running key moves back by four columns with a 32-bit
		// subtraction (borrow-free: a half never falls below 4, see KSW_KEY_INIT), written as a multiply-add by a
		// register the compiler cannot fold (always 1) so that it issues on the FMA pipe
		R.m = max3u2(R.m * K.one - KSW_KEY_STEP, hA * 128u + K.keyRA, hB * 128u + K.keyRB);
#else
		R.m = addmaxu2(R.m, KSW_KEY_DEC, hA * 128u + K.keyRA);
		R.m = maxu2(R.m, hB * 128u + K.keyRB);
#endif
	} else {
		// per half (each half sees its columns in rising order)
		const int c0 = q << 2;
		bool ph, pl;
		R.m = bmax2(hA, R.m, ph, pl);
		if (pl) R.mjl = c0;
		if (ph) R.mjh = c0 + 2;
		R.m = bmax2(hB, R.m, ph, pl);
		if (pl) R.mjl = c0 + 1;
		if (ph) R.mjh = c0 + 3;
	}
	// zero detector over the in-band cells
	if (EDGE) R.zminE = min3_2(R.zminE, hA | (~keepA & 0x7fff7fffu), hB | (~keepB & 0x7fff7fffu));
	else R.zmin = min3_2(R.zmin, hA, hB);
	// store: eh[j].h = H(i, j-1), eh[j].e = E(i+1, j)
	ksw_u4 o;
#ifdef KSW_V_STORE_IMAD
	o.x = ksw_lo_to_hi_add(hB, ksw_hi16_of(R.Hc));         // (H(c0-1), H(c1)) for columns c0, c2: two FMA-pipe moves
#else
	o.x = prmt(R.Hc, hB, 0x5432u);                         // (H(c0-1), H(c1)) for columns c0, c2
#endif
	o.z = hA;                                              // (H(c0), H(c2)) for columns c1, c3
	if (EDGE) {
		// eh[beg].h = first-column value (ksw.c:429) where this quad holds column lo; eh[end].e = 0 (ksw.c:446)
		// where it holds column hi (eh[end].h = h1 is what the shifted store writes there anyway)
		o.x = (o.x & ~firstA) | (left0pk & firstA);
		o.z = (o.z & ~firstB) | (left0pk & firstB);
		eA = (eA & ~endA) | (K.Bpk & endA); eB = (eB & ~endB) | (K.Bpk & endB);
	}
	o.y = eA; o.w = eB;
	*dst = o;
	R.Hc = hB;                                             // hi half = H(c3)
	R.hA = hA; R.hB = hB;
}

// whether a word of two int16 >= B holds a half equal to B (a zero score)
static KSW_HD bool ksw_has_zero16(uint32_t x, uint32_t Bpk)
{
	x -= Bpk;
	return ((x - 0x00010001u) & ~x & 0x80008000u) != 0u;
}

// the reference's trim scans (ksw.c:463-466) over the stored eh[].h, skipping whole zero-free quads.  The stored quads
// cq_lo..cq_hi (eh[] indices 4cq_lo .. 4cq_hi+3) are known to hold no zero (the row's interior quads had none) and are
// jumped over without a look; cq_lo > cq_hi: nothing is known.
template <int T>
static KSW_HD void ksw_fast_trim_scan(const KswFastMem<T> &M, const KswFastConst &K, int rarg, int lo, int hi, int cq_lo, int cq_hi,
                                      int &new_lo, int &new_hi)
{
	int j = rarg;
	while (j >= lo) {
		const int qj = j >> 2;
		if (qj >= cq_lo && qj <= cq_hi) { j = (cq_lo << 2) - 1; continue; }
		if ((j & 3) == 3 && j - 3 >= lo) {
			const ksw_u4 v = M.hq[qj * T];
			if (!ksw_has_zero16(v.x, K.Bpk) && !ksw_has_zero16(v.z, K.Bpk)) { j -= 4; continue; }
		}
		if (*M.h16(j) == (uint16_t)K.B) break;
		--j;
	}
	new_lo = j + 1;
	j = rarg + 2;
	while (j <= hi) {
		const int qj = j >> 2;
		if (qj >= cq_lo && qj <= cq_hi) { j = (cq_hi + 1) << 2; continue; }
		if ((j & 3) == 0 && j + 3 <= hi) {
			const ksw_u4 v = M.hq[qj * T];
			if (!ksw_has_zero16(v.x, K.Bpk) && !ksw_has_zero16(v.z, K.Bpk)) { j += 4; continue; }
		}
		if (*M.h16(j) == (uint16_t)K.B) break;
		++j;
	}
	new_hi = j;
}

// Processes row L.i.  Returns true when the job is finished (results are then in L).
template <int T, bool KEYED>
static KSW_HD bool ksw_fast_row(KswFastLane &L, const KswFastMem<T> &M, const KswFastConst &K, const ksw_u2 *mrow)
{
	using namespace kswdpx;
	const int i = L.i;
	if (i >= L.tlen) return true;
	// target base of this row
	if ((i & 15) == 0 && i) {
		L.tw = L.tw_next;
		const int nx = (i >> 4) + 1;
		if ((nx << 4) < L.tlen) L.tw_next = L.t2[nx];
	}
	int t = (int)((L.tw >> ((i & 15) << 1)) & 3u);
	if (!KEYED && L.tn && ((L.tn[i >> 5] >> (i & 31)) & 1u)) t = 4;      // keyed jobs are N-free (ksw_pack.cpp)
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
	const int hi_rel = hi - (q1 << 2);                         // 1..4: where column hi sits relative to the last quad
	const KswFastEdge eL = M.edge[lo & 3], eR = M.edge[hi_rel];
	const uint32_t left0pk = (uint32_t)(left0 + K.B) * 0x10001u;
	KswFastRowRegs R;
	R.X = (uint32_t)(K.B + K.oe_ins);                          // F = 0 entering the band (ksw.c:417), as Fs, high half clear
	R.Hc = 0; R.m = KEYED ? KSW_KEY_INIT : 0u; R.zmin = R.zminE = 0x7fff7fffu; R.mjl = -1; R.mjh = -1; R.hA = 0; R.hB = 0;
	ksw_u4 *ph = M.hq + q0 * T;                                // running pointers: the loop needs no index arithmetic
	const uint32_t *ps = M.sq + q0 * T;
	ksw_u4 v = *ph;
	uint32_t sw = *ps;
	if (q1 == q0) {
		ksw_fast_quad<T, KEYED, true>(R, ph, K, mr, q0, v, sw, eL.geA & eR.ltA, eL.geB & eR.ltB, eL.onlyA, eL.onlyB, left0pk, eR.onlyA, eR.onlyB);
	} else {
		ksw_u4 vn = ph[T];                                     // software prefetch, one quad ahead
		uint32_t swn = ps[T];
		ksw_fast_quad<T, KEYED, true>(R, ph, K, mr, q0, v, sw, eL.geA, eL.geB, eL.onlyA, eL.onlyB, left0pk, 0u, 0u);
		ksw_u4 *const pl = M.hq + q1 * T;
#ifdef KSW_V_PTRIMAD
		// The interior quads q0+1 .. q1-1 by quad index: the index advances with a multiply-add by K.one (a register the
		// compiler cannot fold), so it and the two addresses derived from it (index * stride + base) are computed on the
		// FMA pipe instead of three pointer increments on the ALU pipe, which is the one the cells saturate.  Two quads
		// per turn (an odd one first); (vn, swn) always hold the next quad.
		uint32_t qi = (uint32_t)q0 + 1u;
		const uint32_t qe = (uint32_t)q1;
		if ((qe - qi) & 1u) {
			ksw_u4 *p = M.hq + qi * T;
			const uint32_t *s = M.sq + qi * T;
			v = vn; sw = swn;
			vn = p[T]; swn = s[T];
			ksw_fast_quad<T, KEYED, false>(R, p, K, mr, (int)qi, v, sw, 0u, 0u, 0u, 0u, 0u, 0u, 0u);
			qi = qi * K.one + 1u;
		}
		while (qi != qe) {
			ksw_u4 *p = M.hq + qi * T;
			const uint32_t *s = M.sq + qi * T;
			v = vn; sw = swn;
			const ksw_u4 v1 = p[T]; const uint32_t sw1 = s[T];
			vn = p[2 * T]; swn = s[2 * T];
			ksw_fast_quad<T, KEYED, false>(R, p, K, mr, (int)qi, v, sw, 0u, 0u, 0u, 0u, 0u, 0u, 0u);
			ksw_fast_quad<T, KEYED, false>(R, p + T, K, mr, (int)qi + 1, v1, sw1, 0u, 0u, 0u, 0u, 0u, 0u, 0u);
			qi = qi * K.one + 2u;
		}
#else
		int q = q0 + 1;
		ph += T; ps += T;
#ifdef __CUDACC__
#if defined(KSW_V_UNROLL4)
#pragma unroll 4
#elif defined(KSW_V_UNROLL1)
#pragma unroll 1
#else
#pragma unroll 2
#endif
#endif
		for (; ph != pl; ph += T, ps += T, ++q) {
			v = vn; sw = swn;
			vn = ph[T]; swn = ps[T];
			ksw_fast_quad<T, KEYED, false>(R, ph, K, mr, q, v, sw, 0u, 0u, 0u, 0u, 0u, 0u, 0u);
		}
#endif
		ksw_fast_quad<T, KEYED, true>(R, pl, K, mr, q1, vn, swn, eR.ltA, eR.ltB, 0u, 0u, 0u, eR.onlyA, eR.onlyB);
	}
	// H(i, hi-1): the reference's h1 after the loop
	const int left_b = (int)prmt(R.hA, R.hB, eR.sel_left);     // still biased
	const int left = left_b - K.B;
	if (hi_rel == 4) {
		// column hi opens the next quad: eh[end].h = h1, eh[end].e = 0 (ksw.c:446); the other half-words of that
		// 64-bit slot belong to column hi+2, which is rewritten before it is read again (the band end grows by at most
		// one column per row, and every row writes eh[end])
		ksw_u2 *p = reinterpret_cast<ksw_u2 *>(&M.hq[(q1 + 1) * T]);
		ksw_u2 w2; w2.x = (uint32_t)left_b; w2.y = K.Bpk;
		*p = w2;
	}
	if (hi == L.qlen) {                                        // ksw.c:447-450, ties -> last row
		if (left >= L.end_sc) L.end_i = i;
		L.end_sc = L.end_sc > left ? L.end_sc : left;
	}
	int rmax, rarg;
	if (KEYED) {
		const uint32_t k_lo = R.m & 0xffffu, k_hi = R.m >> 16;
		const uint32_t kmax = (k_lo > k_hi ? k_lo : k_hi) + (uint32_t)(q1 << 2);   // keys are relative to the last quad
		rmax = (int)(kmax >> 7) - K.B; rarg = (int)(kmax & 127u);
	} else {
		const int m_lo = (int)(int16_t)(R.m & 0xffffu), m_hi = (int)(int16_t)(R.m >> 16);
		rmax = (m_lo > m_hi ? m_lo : m_hi) - K.B;
		rarg = m_lo > m_hi ? R.mjl : (m_hi > m_lo ? R.mjh : (R.mjl > R.mjh ? R.mjl : R.mjh));
	}
	if (rmax == 0) return true;                                // ksw.c:451
	{   // ksw.c:452-461, written without divergent paths: new best, or the z-drop test against the old best
		const bool better = rmax > L.best;
		const int di = i - L.best_i, dj = rarg - L.best_j;
		const int gap = di > dj ? (di - dj) * K.e_del : (dj - di) * K.e_ins;
		const bool drop = !better && K.zdrop > 0 && (L.best - rmax - gap > K.zdrop);
		const int d = rarg > i ? rarg - i : i - rarg;
		const int off2 = L.off > d ? L.off : d;
		L.best_i = better ? i : L.best_i;
		L.best_j = better ? rarg : L.best_j;
		L.off = better ? off2 : L.off;
		L.best = better ? rmax : L.best;
		if (drop) return true;
	}
	// band trim (ksw.c:463-466)
	const bool zero_in = ((R.zmin & 0xffffu) == (uint32_t)K.B) | ((R.zmin >> 16) == (uint32_t)K.B);
	const bool zero_edge = ((R.zminE & 0xffffu) == (uint32_t)K.B) | ((R.zminE >> 16) == (uint32_t)K.B);
	if (!(zero_in | zero_edge)) {
		// every eh[j].h for j in (lo, hi] is non-zero, so only the first-column slot eh[lo].h can stop the
		// downward scan, and the upward scan runs off the end
		L.lo = left0 ? lo : lo + 1;
		L.hi = hi + 1;
	} else {
		// the columns of the interior quads q0+1 .. q1-1 are stored one slot to the right (eh[j].h = H(i, j-1)): if none of
		// them is zero, the stored quads q0+2 .. q1-1 hold no zero
		int nl, nh;
		ksw_fast_trim_scan<T>(M, K, rarg, lo, hi, zero_in ? 1 : q0 + 2, zero_in ? 0 : q1 - 1, nl, nh);
		L.lo = nl; L.hi = nh;
	}
	L.i = i + 1;
	return L.i >= L.tlen;
}

static KSW_HD void ksw_fast_result(const KswFastLane &L, DevRes &r)
{
	r.score = L.best; r.qle = L.best_j + 1; r.tle = L.best_i + 1;
	r.gtle = L.end_i + 1; r.gscore = L.end_sc; r.max_off = L.off;
}
