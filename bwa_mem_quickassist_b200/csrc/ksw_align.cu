// ksw_align.cu — local alignment with start positions and second-best score: the reference's ksw_align2
// (bwa-0.7.8/ksw.c:329-354) with its two striped SSE2 kernels ksw_u8 (ksw.c:110-236) and ksw_i16 (ksw.c:238-320), which
// mem_matesw calls for mate rescue (bwamem_pair.c:150).  SURVEY §8(f) rank 4.
//
// The reference's results depend on its vector layout (E is computed before the lazy-F loop corrects H across lane
// boundaries; the lazy-F loop stops when no lane can improve; the padding columns take part in the row maximum), so this
// kernel keeps the layout: ONE SUB-WARP OF P LANES IS ONE 128-BIT VECTOR — P = 16 lanes for the byte kernel, 8 for the
// 16-bit kernel; lane l owns the query segment l*slen .. l*slen+slen-1 exactly like SSE lane l (ksw.c:87-89), walks it
// column by column with F in a register, and the byte shifts between lanes (_mm_slli_si128) are __shfl_up_sync.  A warp
// runs 32/P jobs of similar size (the host sorts them) with warp-uniform control flow; every arithmetic step saturates the
// way the SSE2 instruction it replaces does.  Per sub-warp, shared memory holds (H | E<<16) per column, the row copied at
// the best score (Hmax, ksw.c:208-209) and the query profile (ksw.c:87-106); the second-best list (ksw.c:196-205) goes to
// a per-warp scratch slab in HBM, one (score, row) pair per qualifying row at most.
// The lazy-F loop (ksw.c:182-192) is where the SSE2 code spends its time on a real alignment: up to 16 x slen steps per row,
// each a load, a store and a vote.  When o_ins >= 1 its result has a closed form: the loop stops early only where every
// lane's own main-loop F already dominates what is left of the carried one (at a stop f - e <= H' - oe, and H' was not raised
// by f itself, else f - e <= f - oe would need o <= 0), so H' = max(H, carry) with the carry propagated all the way:
// carry into lane s = max over s' < s of (final f of lane s') - e_ins * slen * (s - s' - 1), a max-plus prefix scan over the
// lanes (log2 P shuffles), decaying by e_ins per column inside the lane.  The kernel does that scan once per row and applies
// the carry when the row is READ (next row's diagonal, the Hmax copy) — no extra pass over the row.  With o_ins == 0 (or gap
// costs that wrap in the byte kernel) the early stop is observable and the kernel runs the loop literally (LIT = true).
// Both passes of ksw_align2 run back to back in the same sub-warp: the forward pass, then — if KSW_XSTART asks for it and the
// score passes the KSW_XSUBO threshold — the pass over the reversed prefixes with KSW_XSTOP (ksw.c:342-350), which, like the
// reference, still walks all tlen rows (reversed prefix first) unless it stops at the score.
#include <cuda_runtime.h>
#include <stdlib.h>
#include "ksw_dev.cuh"
#include "ksw_launch.h"

namespace {

constexpr int XSTOP = 0x20000, XSUBO = 0x40000, XSTART = 0x80000;      // ksw.h:6-9
constexpr unsigned FULL = 0xffffffffu;

struct APass { int score, te, qe, score2, te2; };

template <int P>
__device__ __forceinline__ int sub_max(int v)
{
#pragma unroll
	for (int o = P / 2; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(FULL, v, o, P));
	return v;
}

// one call of ksw_u8 (P = 16) / ksw_i16 (P = 8) on the sub-warp's job; live = this sub-warp has a job in this pass.
// rev: the sequences are the reversed prefixes query[qe1..0], target[te1..0] followed by target[te1+1..tlen) (ksw.c:343-345)
template <int P, bool LIT>
__device__ APass align_pass(const bool live, const uint8_t *__restrict__ query, const int qlen, const uint8_t *__restrict__ target,
                            const int tlen, const int xtra, const bool rev, const int qe1, const int te1, uint32_t *HE, uint16_t *HM,
                            int8_t *PR, const int cap, uint2 *bs, const KswAlignParams &A, const int sub, const int sl)
{
	const int slen = live ? (qlen + P - 1) / P : 0;                       // ksw.c:68
	const unsigned submask = (P == 16 ? 0xffffu : 0xffu) << (sub * P);
	const int vmask = P == 16 ? 0xff : 0xffff;
	const int oe_del = (A.o_del + A.e_del) & vmask, e_del = A.e_del & vmask;     // _mm_set1_epi8/16 truncate (ksw.c:131-134)
	const int oe_ins = (A.o_ins + A.e_ins) & vmask, e_ins = A.e_ins & vmask;
	const int shift = A.shift;
	// query profile (ksw.c:87-106) and the zeroed H / E / Hmax rows (ksw.c:138-142)
	for (int j = 0; j < slen; ++j) {
		const int k = j + sl * slen;
		int code = 4;
		if (k < qlen) { code = rev ? query[qe1 - k] : query[k]; code = code > 4 ? 4 : code; }
#pragma unroll
		for (int a = 0; a < 5; ++a) PR[a * cap + j * P + sl] = k < qlen ? A.mat[a * 5 + code] : (int8_t)0;
		HE[j * P + sl] = 0u;
		HM[j * P + sl] = 0;
	}
	__syncwarp();
	int slen_w = slen, rows_w = live ? tlen : 0;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		slen_w = max(slen_w, __shfl_xor_sync(FULL, slen_w, o));
		rows_w = max(rows_w, __shfl_xor_sync(FULL, rows_w, o));
	}
	const int minsc = (xtra & XSUBO) ? (xtra & 0xffff) : 0x10000;         // ksw.c:127-128
	const int endsc = (xtra & XSTOP) ? (xtra & 0xffff) : 0x10000;
	int gmax = 0, te = -1, n_b = 0, last_sc = 0, last_row = -2;
	bool done = !live;
	int t_next = 0;
	if (live && tlen > 0) t_next = rev && 0 <= te1 ? target[te1] : target[0];
	const int decay = e_ins * slen;                                       // what a carried F loses across one whole lane
	int G = 0;                                                            // carry into this lane's segment from the previous row's lazy-F loop
	for (int i = 0; i < rows_w; ++i) {
		const bool act = !done && i < tlen;
		if (!__any_sync(FULL, act)) break;
		int t = t_next;
		t = t > 4 ? 4 : t;
		if (act && i + 1 < tlen) t_next = rev && i + 1 <= te1 ? target[te1 - i - 1] : target[i + 1];
		const int8_t *S = PR + t * cap;
		// h = H(i-1, -1): the last vector shifted up by one lane (ksw.c:147-148)
		const int hl = act ? max((int)(HE[(slen - 1) * P + sl] & 0xffffu), G - e_ins * (slen - 1)) : 0;
		int h = __shfl_up_sync(FULL, hl, 1, P);
		if (sl == 0) h = 0;
		int f = 0, mx = 0, c = G;
		for (int j = 0; j < slen_w; ++j) {
			if (act && j < slen) {
				const uint32_t w = HE[j * P + sl];
				int e = (int)(w >> 16);
				const int sc = S[j * P + sl];
				int hv;
				if (P == 16) {                                            // ksw.c:156-157
					hv = min(255, h + ((sc + shift) & 0xff));
					hv = max(0, hv - shift);
				} else {                                                  // ksw.c:268 (signed saturating add)
					hv = min(32767, max(-32768, h + sc));
				}
				hv = max(hv, max(e, f));                                  // ksw.c:159-160
				mx = max(mx, hv);
				e = max(max(0, e - e_del), max(0, hv - oe_del));          // ksw.c:164-167
				f = max(max(0, f - e_ins), max(0, hv - oe_ins));          // ksw.c:169-171
				HE[j * P + sl] = (uint32_t)hv | ((uint32_t)e << 16);
				h = max((int)(w & 0xffffu), c);                           // H'(i-1, j): the stored H with the previous row's carry
				c -= e_ins;
			}
		}
		if (!LIT) {
			// closed form of the lazy-F loop: carry into lane s = max over s' < s of f_end(s') - decay * (s - s' - 1)
			int v = __shfl_up_sync(FULL, f, 1, P);
			v = sl == 0 ? 0 : v;
#pragma unroll
			for (int d = 1; d < P; d <<= 1) {
				const int o = __shfl_up_sync(FULL, v, d, P);
				if (sl >= d) v = max(v, o - decay * d);
			}
			G = act ? max(v, 0) : 0;
		} else {
			// the lazy-F loop as it is (ksw.c:182-192 / 282-291): at most 16 rounds of slen steps; a round starts by shifting
			// f up one lane; it ends for good at the first step after which no lane of the vector has f > H - oe_ins.  If
			// every shifted f is zero the first step changes nothing and ends it, so it is skipped.
			int fs = __shfl_up_sync(FULL, f, 1, P);
			fs = sl == 0 ? 0 : fs;
			const unsigned pos = __ballot_sync(FULL, act && fs > 0);
			bool lz = act && (pos & submask) != 0u;
			int jl = 0, kl = 0;
			f = fs;
			while (__any_sync(FULL, lz)) {
				bool gt = false;
				if (lz) {
					const uint32_t w = HE[jl * P + sl];
					int hv = max((int)(w & 0xffffu), f);
					HE[jl * P + sl] = (w & 0xffff0000u) | (uint32_t)hv;
					hv = max(0, hv - oe_ins);
					f = max(0, f - e_ins);
					gt = f > hv;
				}
				const unsigned more = __ballot_sync(FULL, gt);
				if (lz) {
					if ((more & submask) == 0u) lz = false;
					else if (++jl == slen) { jl = 0; if (++kl == 16) lz = false; }
				}
				// the next round's shift (only sub-warps at a round start take it)
				fs = __shfl_up_sync(FULL, f, 1, P);
				if (lz && jl == 0) f = sl == 0 ? 0 : fs;
			}
		}
		const int imax = sub_max<P>(mx);
		if (act && imax >= minsc) {                                       // the second-best list, ksw.c:196-205
			if (n_b == 0 || last_row + 1 != i) {
				if (sl == 0) bs[n_b] = make_uint2((uint32_t)imax, (uint32_t)i);
				++n_b; last_sc = imax; last_row = i;
			} else if (last_sc < imax) {
				if (sl == 0) bs[n_b - 1] = make_uint2((uint32_t)imax, (uint32_t)i);
				last_sc = imax; last_row = i;
			}
		}
		if (act && imax > gmax) {                                         // ksw.c:206-211 (the kept row is H after the lazy-F loop)
			gmax = imax; te = i;
			int cc = G;
			for (int j = 0; j < slen; ++j, cc -= e_ins) HM[j * P + sl] = (uint16_t)max((int)(HE[j * P + sl] & 0xffffu), cc);
			if ((P == 16 && gmax + shift >= 255) || gmax >= endsc) done = true;
		}
		__syncwarp();
	}
	APass r;
	r.score = P == 16 ? (gmax + shift < 255 ? gmax : 255) : gmax;          // ksw.c:214 / 309
	r.te = te; r.qe = -1; r.score2 = -1; r.te2 = -1;
	__syncwarp();
	{
		// qe: the smallest column among those holding the maximum of the kept row (ksw.c:218-221)
		int best = -1, col = 0x7fffffff;
		for (int j = 0; j < slen; ++j) {
			const int v = HM[j * P + sl];
			if (v > best) { best = v; col = j + sl * slen; }
		}
#pragma unroll
		for (int o = P / 2; o > 0; o >>= 1) {
			const int ob = __shfl_xor_sync(FULL, best, o, P), oc = __shfl_xor_sync(FULL, col, o, P);
			if (ob > best || (ob == best && oc < col)) { best = ob; col = oc; }
		}
		if (P == 8 || r.score != 255) r.qe = col;
		// second best: the first list entry with the highest score outside [te - d, te + d] (ksw.c:223-231)
		int s2 = -1, idx2 = 0x7fffffff, row2 = -1;
		if (live && n_b > 0 && (P == 8 || r.score != 255)) {
			const int d = (r.score + A.qmax - 1) / A.qmax, low = te - d, high = te + d;
			for (int x = sl; x < n_b; x += P) {
				const uint2 en = bs[x];
				const int e = (int)en.y;
				if ((e < low || e > high) && (int)en.x > s2) { s2 = (int)en.x; idx2 = x; row2 = e; }
			}
		}
#pragma unroll
		for (int o = P / 2; o > 0; o >>= 1) {
			const int os = __shfl_xor_sync(FULL, s2, o, P), oi = __shfl_xor_sync(FULL, idx2, o, P), orow = __shfl_xor_sync(FULL, row2, o, P);
			if (os > s2 || (os == s2 && oi < idx2)) { s2 = os; idx2 = oi; row2 = orow; }
		}
		if (s2 >= 0) { r.score2 = s2; r.te2 = row2; }
	}
	__syncwarp();
	return r;
}

template <int P, bool LIT>
__global__ void __launch_bounds__(32)
ksw_align_kernel(const DevAJob *__restrict__ jobs, const uint32_t *__restrict__ order, const int n_jobs, const uint8_t *__restrict__ seq,
                 const KswAlignParams A, const int cap, uint2 *__restrict__ bscr, const int tcap, unsigned *__restrict__ counter,
                 DevARes *__restrict__ res)
{
	constexpr int G = 32 / P;
	extern __shared__ uint32_t asm_[];
	const int lane = threadIdx.x, sub = lane / P, sl = lane % P;
	uint32_t *HE = asm_ + sub * cap;
	uint16_t *HM = reinterpret_cast<uint16_t *>(asm_ + G * cap) + sub * cap;
	int8_t *PR = reinterpret_cast<int8_t *>(asm_ + G * cap + G * cap / 2) + sub * 5 * cap;
	uint2 *bs = bscr + ((size_t)blockIdx.x * G + sub) * (size_t)tcap;
	for (;;) {
		unsigned g = 0;
		if (lane == 0) g = atomicAdd(counter, 1u);
		g = __shfl_sync(FULL, g, 0);
		if ((long long)g * G >= n_jobs) break;
		const int jidx = (int)g * G + sub;
		const bool have = jidx < n_jobs;
		DevAJob jb;
		jb.seq_off = 0; jb.qlen = 1; jb.tlen = 0; jb.xtra = 0; jb.idx = 0;
		if (have) jb = jobs[order[jidx]];
		const uint8_t *query = seq + jb.seq_off, *target = query + jb.qlen;
		const APass r = align_pass<P, LIT>(have, query, jb.qlen, target, jb.tlen, jb.xtra, false, 0, -1, HE, HM, PR, cap, bs, A, sub, sl);
		// ksw.c:341: the start positions are wanted and the score passes the threshold; a saturated byte score (255) is outside the
		// reference's defined behaviour (it goes on with qe = -1): the job ends here with score 255
		const bool second = have && (jb.xtra & XSTART) && !((jb.xtra & XSUBO) && r.score < (jb.xtra & 0xffff)) && r.qe >= 0;
		const APass rr = align_pass<P, LIT>(second, query, r.qe + 1, target, jb.tlen, XSTOP | r.score, true, r.qe, r.te, HE, HM, PR, cap, bs, A, sub, sl);
		if (have && sl == 0) {
			DevARes o;
			o.score = r.score; o.te = r.te; o.qe = r.qe; o.score2 = r.score2; o.te2 = r.te2; o.tb = -1; o.qb = -1; o.pad = 0;
			if (second && r.score == rr.score) { o.tb = r.te - rr.te; o.qb = r.qe - rr.qe; }      // ksw.c:348-349
			res[jb.idx] = o;
		}
	}
}

// ------------------------------------------------------------------------------------------------------------------
// The same algorithm with TWO SSE lanes per thread as the halves of s16x2 registers (closed-form lazy-F only): thread t of a
// job's T = P/2 threads holds lanes 2t (low half) and 2t+1 (high half).  Every per-cell operation of the reference is
// element-wise, so each becomes one DPX instruction on both lanes: the saturating byte add is VIADDMNMX(h, S, 255) followed by
// VIADDMNMX(., -shift, 0), the unsigned saturating subtractions are VIADDMNMX.RELU, H = VIMNMX3(h, E, F).  The profile is not
// stored: a column keeps a 16-bit PRMT selector (two query codes) and the row's scores sit in two registers as bytes, like
// the extension kernel's look-up.  One 16-byte shared-memory word per column pair: {H, E, selector, Hmax}.  Jobs per warp:
// 4 (byte kernel) / 8 (16-bit kernel).  Used when o_ins >= 1 and, for the 16-bit kernel, no score can reach the int16
// saturation (qlen * max score <= 32000); everything else takes the int32 kernel above.
__device__ __forceinline__ uint32_t apk2(int v) { return ((uint32_t)v & 0xffffu) | ((uint32_t)v << 16); }
__device__ __forceinline__ uint32_t aprmt(uint32_t a, uint32_t b, uint32_t sel)
{
	uint32_t d;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
	return d;
}

template <int P>
__device__ APass align_pass2(const bool live, const uint8_t *__restrict__ query, const int qlen, const uint8_t *__restrict__ target,
                             const int tlen, const int xtra, const bool rev, const int qe1, const int te1, uint4 *C, uint2 *bs,
                             const uint2 *mrow, const KswAlignParams &A, const int tl)
{
	constexpr int T = P / 2;
	const int slen = live ? (qlen + P - 1) / P : 0;                       // ksw.c:68
	const int vmask = P == 16 ? 0xff : 0xffff;
	const int e_ins = A.e_ins & vmask;
	const uint32_t n_ed = apk2(-(A.e_del & vmask)), n_oed = apk2(-((A.o_del + A.e_del) & vmask));
	const uint32_t n_ei = apk2(-e_ins), n_oei = apk2(-((A.o_ins + A.e_ins) & vmask));
	const uint32_t n_shift = apk2(-A.shift), k255 = apk2(255), floor1 = apk2(-1), low16 = apk2(-32768);
	// per column pair: the PRMT selector of the two query codes (padding columns: code 5 = the table's score-0 entry), zeroed rows
	for (int j = 0; j < slen; ++j) {
		uint32_t sel = 0;
#pragma unroll
		for (int hlf = 0; hlf < 2; ++hlf) {
			const int k = j + (2 * tl + hlf) * slen;
			uint32_t code = 5;
			if (k < qlen) { code = rev ? query[qe1 - k] : query[k]; code = code > 4 ? 4 : code; }
			// byte kernel: value byte, then a zero byte (7); 16-bit kernel: value byte, then its sign
			const uint32_t nib = P == 16 ? (code | 0x70u) : (code | ((8u | code) << 4));
			sel |= nib << (8 * hlf);
		}
		C[j * T + tl] = make_uint4(0u, 0u, sel, 0u);
	}
	__syncwarp();
	int slen_w = slen, rows_w = live ? tlen : 0;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		slen_w = max(slen_w, __shfl_xor_sync(FULL, slen_w, o));
		rows_w = max(rows_w, __shfl_xor_sync(FULL, rows_w, o));
	}
	const int minsc = (xtra & XSUBO) ? (xtra & 0xffff) : 0x10000;
	const int endsc = (xtra & XSTOP) ? (xtra & 0xffff) : 0x10000;
	int gmax = 0, te = -1, n_b = 0, last_sc = 0, last_row = -2;
	bool done = !live;
	int t_next = 0;
	if (live && tlen > 0) t_next = rev && 0 <= te1 ? target[te1] : target[0];
	const int decay = min(e_ins * slen, 40000);                           // what a carried F loses across one whole lane
	const uint32_t n_last = apk2(-min(e_ins * (slen - 1), 32767));        // ... across a lane's columns but the last
	uint32_t G = 0;                                                       // carries into this thread's two lanes (previous row)
	for (int i = 0; i < rows_w; ++i) {
		const bool act = !done && i < tlen;
		if (!__any_sync(FULL, act)) break;
		int t = t_next;
		t = t > 4 ? 4 : t;
		if (act && i + 1 < tlen) t_next = rev && i + 1 <= te1 ? target[te1 - i - 1] : target[i + 1];
		const uint2 mr = mrow[t];
		// h = H(i-1, -1): the last vector shifted up by one lane (ksw.c:147-148): low half <- previous thread's high half,
		// high half <- own low half
		const uint32_t hl = act ? __vmaxs2(C[(slen - 1) * T + tl].x, __viaddmax_s16x2(G, n_last, floor1)) : 0u;
		uint32_t prev = __shfl_up_sync(FULL, hl, 1, T);
		if (tl == 0) prev = 0u;
		uint32_t h = aprmt(prev, hl, 0x5432u);
		uint32_t f = 0u, mx = 0u, c = G;
		for (int j = 0; j < slen_w; ++j) {
			if (act && j < slen) {
				uint4 *cell = &C[j * T + tl];
				const uint4 v = *cell;
				const uint32_t S = aprmt(mr.x, mr.y, v.z);
				uint32_t hv;
				if (P == 16) hv = __viaddmax_s16x2(__viaddmin_s16x2(h, S, k255), n_shift, 0u);      // ksw.c:156-157
				else hv = __viaddmax_s16x2(h, S, low16);                                            // ksw.c:268 (cannot saturate here)
				hv = __vimax3_s16x2(hv, v.y, f);                                                    // ksw.c:159-160
				mx = __vmaxs2(mx, hv);
				const uint32_t e = __viaddmax_s16x2(v.y, n_ed, __viaddmax_s16x2_relu(hv, n_oed, 0u));   // ksw.c:164-167
				f = __viaddmax_s16x2(f, n_ei, __viaddmax_s16x2_relu(hv, n_oei, 0u));                   // ksw.c:169-171
				*reinterpret_cast<uint2 *>(cell) = make_uint2(hv, e);
				h = __vmaxs2(v.x, c);                                     // H'(i-1, j): the stored H with the previous row's carry
				c = __viaddmax_s16x2(c, n_ei, floor1);
			}
		}
		{
			// closed form of the lazy-F loop (see above), two lanes per thread: x(t) = carry into lane 2t
			const int flo = (int)(f & 0xffffu), fhi = (int)(f >> 16);
			int v = max(fhi, flo - decay);
			v = __shfl_up_sync(FULL, v, 1, T);
			if (tl == 0) v = 0;
#pragma unroll
			for (int d = 1; d < T; d <<= 1) {
				const int o = __shfl_up_sync(FULL, v, d, T);
				if (tl >= d) v = max(v, o - 2 * decay * d);
			}
			const int glo = max(v, 0), ghi = max(max(flo, v - decay), 0);
			G = act ? ((uint32_t)glo | ((uint32_t)ghi << 16)) : 0u;
		}
		int imax = max((int)(mx & 0xffffu), (int)(mx >> 16));
#pragma unroll
		for (int o = T / 2; o > 0; o >>= 1) imax = max(imax, __shfl_xor_sync(FULL, imax, o, T));
		if (act && imax >= minsc) {                                       // the second-best list, ksw.c:196-205
			if (n_b == 0 || last_row + 1 != i) {
				if (tl == 0) bs[n_b] = make_uint2((uint32_t)imax, (uint32_t)i);
				++n_b; last_sc = imax; last_row = i;
			} else if (last_sc < imax) {
				if (tl == 0) bs[n_b - 1] = make_uint2((uint32_t)imax, (uint32_t)i);
				last_sc = imax; last_row = i;
			}
		}
		if (act && imax > gmax) {                                         // ksw.c:206-211
			gmax = imax; te = i;
			uint32_t cc = G;
			for (int j = 0; j < slen; ++j) {
				C[j * T + tl].w = __vmaxs2(C[j * T + tl].x, cc);
				cc = __viaddmax_s16x2(cc, n_ei, floor1);
			}
			if ((P == 16 && gmax + A.shift >= 255) || gmax >= endsc) done = true;
		}
		__syncwarp();
	}
	APass r;
	r.score = P == 16 ? (gmax + A.shift < 255 ? gmax : 255) : gmax;
	r.te = te; r.qe = -1; r.score2 = -1; r.te2 = -1;
	__syncwarp();
	{
		int best = -1, col = 0x7fffffff;
		for (int j = 0; j < slen; ++j) {
			const uint32_t w = C[j * T + tl].w;
			const int vlo = (int)(w & 0xffffu), vhi = (int)(w >> 16), clo = j + 2 * tl * slen, chi = clo + slen;
			if (vlo > best || (vlo == best && clo < col)) { best = vlo; col = clo; }
			if (vhi > best || (vhi == best && chi < col)) { best = vhi; col = chi; }
		}
#pragma unroll
		for (int o = T / 2; o > 0; o >>= 1) {
			const int ob = __shfl_xor_sync(FULL, best, o, T), oc = __shfl_xor_sync(FULL, col, o, T);
			if (ob > best || (ob == best && oc < col)) { best = ob; col = oc; }
		}
		if (P == 8 || r.score != 255) r.qe = col;
		int s2 = -1, idx2 = 0x7fffffff, row2 = -1;
		if (live && n_b > 0 && (P == 8 || r.score != 255)) {
			const int d = (r.score + A.qmax - 1) / A.qmax, low = te - d, high = te + d;
			for (int x = tl; x < n_b; x += T) {
				const uint2 en = bs[x];
				const int e = (int)en.y;
				if ((e < low || e > high) && (int)en.x > s2) { s2 = (int)en.x; idx2 = x; row2 = e; }
			}
		}
#pragma unroll
		for (int o = T / 2; o > 0; o >>= 1) {
			const int os = __shfl_xor_sync(FULL, s2, o, T), oi = __shfl_xor_sync(FULL, idx2, o, T), orow = __shfl_xor_sync(FULL, row2, o, T);
			if (os > s2 || (os == s2 && oi < idx2)) { s2 = os; idx2 = oi; row2 = orow; }
		}
		if (s2 >= 0) { r.score2 = s2; r.te2 = row2; }
	}
	__syncwarp();
	return r;
}

template <int P>
__global__ void __launch_bounds__(32)
ksw_align2x_kernel(const DevAJob *__restrict__ jobs, const uint32_t *__restrict__ order, const int n_jobs, const uint8_t *__restrict__ seq,
                   const KswAlignParams A, const int cap, uint2 *__restrict__ bscr, const int tcap, unsigned *__restrict__ counter,
                   DevARes *__restrict__ res)
{
	constexpr int T = P / 2, G = 32 / T;
	extern __shared__ uint4 asm2_[];
	const int lane = threadIdx.x, sub = lane / T, tl = lane % T;
	uint2 *mrow = reinterpret_cast<uint2 *>(asm2_);                        // 5 rows of score bytes (+ the padding entry), then the jobs
	uint4 *C = asm2_ + 4 + sub * cap;
	if (lane < 5) {
		// bytes 0..4: the scores of target base `lane` against query codes 0..4 (byte kernel: plus shift); byte 5: a padding column
		uint32_t b[8];
		for (int c = 0; c < 5; ++c) b[c] = P == 16 ? (uint32_t)((A.mat[lane * 5 + c] + A.shift) & 0xff) : (uint32_t)(uint8_t)A.mat[lane * 5 + c];
		b[5] = P == 16 ? (uint32_t)(A.shift & 0xff) : 0u;
		b[6] = b[7] = 0u;
		mrow[lane] = make_uint2(b[0] | b[1] << 8 | b[2] << 16 | b[3] << 24, b[4] | b[5] << 8);
	}
	__syncwarp();
	uint2 *bs = bscr + ((size_t)blockIdx.x * G + sub) * (size_t)tcap;
	for (;;) {
		unsigned g = 0;
		if (lane == 0) g = atomicAdd(counter, 1u);
		g = __shfl_sync(FULL, g, 0);
		if ((long long)g * G >= n_jobs) break;
		const int jidx = (int)g * G + sub;
		const bool have = jidx < n_jobs;
		DevAJob jb;
		jb.seq_off = 0; jb.qlen = 1; jb.tlen = 0; jb.xtra = 0; jb.idx = 0;
		if (have) jb = jobs[order[jidx]];
		const uint8_t *query = seq + jb.seq_off, *target = query + jb.qlen;
		const APass r = align_pass2<P>(have, query, jb.qlen, target, jb.tlen, jb.xtra, false, 0, -1, C, bs, mrow, A, tl);
		const bool second = have && (jb.xtra & XSTART) && !((jb.xtra & XSUBO) && r.score < (jb.xtra & 0xffff)) && r.qe >= 0;
		const APass rr = align_pass2<P>(second, query, r.qe + 1, target, jb.tlen, XSTOP | r.score, true, r.qe, r.te, C, bs, mrow, A, tl);
		if (have && tl == 0) {
			DevARes o;
			o.score = r.score; o.te = r.te; o.qe = r.qe; o.score2 = r.score2; o.te2 = r.te2; o.tb = -1; o.qb = -1; o.pad = 0;
			if (second && r.score == rr.score) { o.tb = r.te - rr.te; o.qb = r.qe - rr.qe; }
			res[jb.idx] = o;
		}
	}
}

template <int P>
cudaError_t launch_2x(const DevAJob *jobs, const uint32_t *order, int n_jobs, const uint8_t *seq, const KswAlignParams &A, int qmax,
                      int tmax, int sm_count, void **bscr, size_t *bscr_cap, unsigned *counter, DevARes *res, cudaStream_t st)
{
	constexpr int T = P / 2, G = 32 / T;
	if (n_jobs <= 0) return cudaSuccess;
	const int cap = ((qmax + P - 1) / P) * T;                             // uint4 words per job
	const size_t smem = 64 + (size_t)G * cap * sizeof(uint4);
	int dev = 0, optin = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
	if (e != cudaSuccess) return e;
	if (smem > (size_t)optin) return cudaErrorInvalidValue;
	e = cudaFuncSetAttribute(ksw_align2x_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
	if (e != cudaSuccess) return e;
	int per_sm = 0;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ksw_align2x_kernel<P>, 32, smem);
	if (e != cudaSuccess) return e;
	if (per_sm < 1) return cudaErrorLaunchOutOfResources;
	int blocks = sm_count * per_sm;
	const int groups = (n_jobs + G - 1) / G;
	if (blocks > groups) blocks = groups;
	const int tcap = tmax > 0 ? tmax : 1;
	const size_t need = (size_t)blocks * G * (size_t)tcap * sizeof(uint2);
	if (need > *bscr_cap) {
		if (*bscr) { e = cudaFree(*bscr); *bscr = nullptr; *bscr_cap = 0; if (e != cudaSuccess) return e; }
		e = cudaMalloc(bscr, need);
		if (e != cudaSuccess) return e;
		*bscr_cap = need;
	}
	e = cudaMemsetAsync(counter, 0, sizeof(unsigned), st);
	if (e != cudaSuccess) return e;
	ksw_align2x_kernel<P><<<blocks, 32, smem, st>>>(jobs, order, n_jobs, seq, A, cap, (uint2 *)*bscr, tcap, counter, res);
	return cudaGetLastError();
}

template <int P, bool LIT>
cudaError_t launch_one(const DevAJob *jobs, const uint32_t *order, int n_jobs, const uint8_t *seq, const KswAlignParams &A, int qmax,
                       int tmax, int sm_count, void **bscr, size_t *bscr_cap, unsigned *counter, DevARes *res, cudaStream_t st)
{
	constexpr int G = 32 / P;
	if (n_jobs <= 0) return cudaSuccess;
	const int cap = ((qmax + 15) / 16) * 16;                              // >= slen * P for every job, and keeps the regions aligned
	const size_t smem = (size_t)G * cap * 11;
	int dev = 0, optin = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
	if (e != cudaSuccess) return e;
	if (smem > (size_t)optin) return cudaErrorInvalidValue;
	e = cudaFuncSetAttribute(ksw_align_kernel<P, LIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
	if (e != cudaSuccess) return e;
	int per_sm = 0;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ksw_align_kernel<P, LIT>, 32, smem);
	if (e != cudaSuccess) return e;
	if (per_sm < 1) return cudaErrorLaunchOutOfResources;
	int blocks = sm_count * per_sm;
	const int groups = (n_jobs + G - 1) / G;
	if (blocks > groups) blocks = groups;
	const int tcap = tmax > 0 ? tmax : 1;
	const size_t need = (size_t)blocks * G * (size_t)tcap * sizeof(uint2);
	if (need > *bscr_cap) {
		if (*bscr) { e = cudaFree(*bscr); *bscr = nullptr; *bscr_cap = 0; if (e != cudaSuccess) return e; }
		e = cudaMalloc(bscr, need);
		if (e != cudaSuccess) return e;
		*bscr_cap = need;
	}
	e = cudaMemsetAsync(counter, 0, sizeof(unsigned), st);
	if (e != cudaSuccess) return e;
	ksw_align_kernel<P, LIT><<<blocks, 32, smem, st>>>(jobs, order, n_jobs, seq, A, cap, (uint2 *)*bscr, tcap, counter, res);
	return cudaGetLastError();
}

} // namespace

cudaError_t ksw_launch_align(int bytes_per_score, const DevAJob *jobs, const uint32_t *order, int n_jobs, const uint8_t *seq,
                             const KswAlignParams &A, int qmax, int tmax, int sm_count, void **bscr, size_t *bscr_cap,
                             unsigned *counter, DevARes *res, cudaStream_t st)
{
	// the closed form of the lazy-F loop needs o_ins >= 1 after the truncation the vector constants go through (ksw.c:131-134)
	const int vmask = bytes_per_score == 1 ? 0xff : 0xffff;
	const bool lit = !((A.e_ins & vmask) < ((A.o_ins + A.e_ins) & vmask)) || A.e_ins < 0 || A.o_ins < 0 || getenv("KSW_B200_ALIGN_LITERAL") != nullptr;
	// two SSE lanes per thread in s16x2 registers: every value must stay clear of the int16 saturation, the selectors need |score| <= 127
	const char *e32 = getenv("KSW_B200_ALIGN_INT32");
	const bool packed = !lit && !(e32 && e32[0] == '1') && A.e_del >= 0 && A.o_del >= 0 && (A.o_del + A.e_del) <= 30000 && (A.o_ins + A.e_ins) <= 30000 &&
	                    (bytes_per_score == 1 || (long long)qmax * A.qmax <= 32000) &&
	                    // its shared memory: 16 bytes per column pair, 4 (byte kernel) or 8 jobs per warp; longer queries take the int32 kernel
	                    (bytes_per_score == 1 ? 64 + 4 * (size_t)((qmax + 15) / 16) * 8 * 16 : 64 + 8 * (size_t)((qmax + 7) / 8) * 4 * 16) <= (size_t)160 * 1024;
	if (bytes_per_score == 1) {
		if (packed) return launch_2x<16>(jobs, order, n_jobs, seq, A, qmax, tmax, sm_count, bscr, bscr_cap, counter, res, st);
		return lit ? launch_one<16, true>(jobs, order, n_jobs, seq, A, qmax, tmax, sm_count, bscr, bscr_cap, counter, res, st)
		           : launch_one<16, false>(jobs, order, n_jobs, seq, A, qmax, tmax, sm_count, bscr, bscr_cap, counter, res, st);
	}
	if (packed) return launch_2x<8>(jobs, order, n_jobs, seq, A, qmax, tmax, sm_count, bscr, bscr_cap, counter, res, st);
	return lit ? launch_one<8, true>(jobs, order, n_jobs, seq, A, qmax, tmax, sm_count, bscr, bscr_cap, counter, res, st)
	           : launch_one<8, false>(jobs, order, n_jobs, seq, A, qmax, tmax, sm_count, bscr, bscr_cap, counter, res, st);
}
