// pb_pitch_frames.cuh — K1 + K2: frames -> normalised autocorrelation -> pitch candidates, one kernel.
//
// Follows Sound_into_PitchFrame (Praat fon/Sound_to_Pitch.cpp, AC_HANNING) per frame; see pb_pitch.cuh for the layout.
// Code-size discipline matters here: the fully unrolled radix-32 butterfly network is ~600 instructions, so the kernel
// keeps exactly ONE copy of it and runs it from a 4-step loop (2 FFTs x 2 passes); likewise one copy of the candidate
// search serves both frames of a pair.  (The first version inlined four copies: ~210 KB of SASS, 55 % of the stall
// samples were instruction-fetch misses — profiles/r01_frames_v1_*.)
#pragma once
#include "pb_pitch.cuh"

// ------------------------------------------------------------------------------------------------ sinc interpolation
// Praat NUM_interpolate_sinc (melder/NUMinterpol.cpp) on y[1..2B+1] = r[-B..B] at lag x, by 8 cooperating lanes
// (sl = lane within the group; every lane of the group passes the same x).  With phi = frac(x), il = floor(x),
// D = min(depth, B - il) the usable depth:
//   y(x) = sin(pi phi)/(2 pi) * sum_{m<D} (-1)^m [ r[il-m]   (1 + cos(pi (phi+m)   / (phi+D)))   / (phi+m)
//                                                + r[il+1+m] (1 + cos(pi (1-phi+m) / (1-phi+D))) / (1-phi+m) ]
// Lane sl takes one side (sl & 1) and every 4th m starting at sl >> 1, so side, sign and the window scale are
// loop-invariant; the loop body is one shared load, two MUFU (cos, rcp) and a handful of FP32 ops.
// `nl` (8, 16 or 32, warp-uniform) lanes cooperate on one evaluation; sl = lane within that group.
__device__ __forceinline__ float pb_sinc8(const float* __restrict__ r, int B, float x, int depth, int sl, int nl = 8) {
    const float fl = floorf(x);
    const float phi = x - fl;
    const int il = (int)fl;
    int D = B - il; if (depth < D) D = depth;
    float acc = 0.0f;
    if (phi == 0.0f) {                                        // on a sample: Praat returns y[x] (no early return:
        if (sl == 0 && depth > 0) acc = r[abs(il)];           // the other groups of the warp still shuffle below)
    } else if (D > 0) {
        const int side = sl & 1, j0 = sl >> 1, hs = nl >> 1;  // hs (4, 8, 16) is even: the sign of a lane's terms is fixed
        const float e = side ? 1.0f - phi : phi;
        const float k = __fdividef(PB_PI_F, e + (float)D);
        const float fhs = (float)hs;
        float d = e + (float)j0;
        int idx = side ? il + 1 + j0 : il - j0;
        const int step = side ? hs : -hs;
#ifndef PB_SIMT_EMU
#pragma unroll 2
#endif
        for (int m = j0; m < D; m += hs) {
            const float yv = r[abs(idx)];
            acc += __fdividef(yv * (1.0f + __cosf(d * k)), d);
            d += fhs; idx += step;
        }
        if (j0 & 1) acc = -acc;
        acc *= sinpif(phi) * (0.5f / PB_PI_F);
    }
    for (int o = 1; o < nl; o <<= 1) acc += __shfl_xor_sync(PB_FULL_MASK, acc, o);
    return acc;
}
// vertex of the parabola through (xa,fa),(xb,fb),(xc,fc), xa < xb < xc; xb if not concave
__device__ __forceinline__ float pb_parabola(float xa, float fa, float xb, float fb, float xc, float fc) {
    const float a = xb - xa, b = xb - xc;
    const float num = a * a * (fb - fc) - b * b * (fb - fa);
    const float den = a * (fb - fc) - b * (fb - fa);
    return den > 0.0f ? xb - 0.5f * __fdividef(num, den) : xb;
}

// ------------------------------------------------------------------------------------------------ rare path: too many maxima
// More maxima than candidate slots (tonal high-frequency content): Praat's sequential insertion — each maximum gets
// its first-pass frequency (parabola) and strength (sinc, depth 30) and replaces the weakest stored candidate, ranked
// by strength - octaveCost*log2(minPitch/f), if it beats it.  One warp; returns the candidate count (= max_cand).
// Kept out of line so the hot path stays compact in the instruction cache.
__device__ PB_NOINLINE int pb_candidates_overflow(const float* __restrict__ r, float* scratch, const PbPitchGeomDev& gm, int lane) {
    float* cf = scratch;
    float* cs = scratch + PB_MAXC;
    int* imax = (int*)(scratch + 2 * PB_MAXC);
    const int B = gm.brent_ixmax, lim = gm.scan_lim, maxc = gm.max_cand;
    const int sub = lane >> 3, sl = lane & 7;
    int ncf = 1;
    for (int base = 2; base < lim; base += 32) {
        const int i = base + lane;
        const bool pk = i < lim && r[i] > gm.half_voicing && r[i] > r[i - 1] && r[i] >= r[i + 1];
        unsigned mask = __ballot_sync(PB_FULL_MASK, pk);
        while (mask) {                                   // four maxima at a time, 8 lanes each
            unsigned m = mask;
            for (int q = 0; q < sub; q++) m &= m - 1;
            const bool have = m != 0;
            const int ip = have ? base + __ffs((int)m) - 1 : 2;
            const float dr = 0.5f * (r[ip + 1] - r[ip - 1]), d2r = 2.0f * r[ip] - r[ip - 1] - r[ip + 1];
            const float x0 = (float)ip + ((have && d2r > 0.0f) ? __fdividef(dr, d2r) : 0.0f);
            float st = pb_sinc8(r, B, x0, have ? 30 : 0, sl);
            if (st > 1.0f) st = __fdividef(1.0f, st);
            const float fq0 = __fdividef(gm.sr, x0);
            for (int q = 0; q < 4; q++) {
                const int hv = __shfl_sync(PB_FULL_MASK, (int)have, q * 8);
                if (!hv) break;
                const float fq = __shfl_sync(PB_FULL_MASK, fq0, q * 8), sq = __shfl_sync(PB_FULL_MASK, st, q * 8);
                const int iq = __shfl_sync(PB_FULL_MASK, ip, q * 8);
                int place = 0;
                if (ncf < maxc) place = ncf++;
                else {
                    // weakest of slots 1..maxc-1; the first minimum wins (Praat scans upward with a strict '<')
                    float ls = 3.0e38f; int li = lane;
                    if (lane >= 1 && lane < maxc) ls = cs[lane] - gm.octave_cost * log2f(gm.min_pitch / cf[lane]);
                    PB_UNROLL for (int o = 16; o > 0; o >>= 1) {
                        const float os = __shfl_xor_sync(PB_FULL_MASK, ls, o);
                        const int oi = __shfl_xor_sync(PB_FULL_MASK, li, o);
                        if (os < ls || (os == ls && oi < li)) { ls = os; li = oi; }
                    }
                    if (sq - gm.octave_cost * log2f(gm.min_pitch / fq) > ls) place = li;
                }
                if (place && lane == 0) { cf[place] = fq; cs[place] = sq; imax[place] = iq; }
                __syncwarp();
            }
            for (int q = 0; q < 4 && mask; q++) mask &= mask - 1;
        }
    }
    return ncf;
}

// ------------------------------------------------------------------------------------------------ rare path: a peak that is not parabola-shaped
// When neither parabola vertex beats the best grid point, the interpolated curve has a plateau or two bumps inside
// [i-1, i+1] and the four evaluations may sit on another bump than the one Praat's Brent search ends on (measured on a
// 1-hour recording: 4 of 250 000 voiced frames, F0 off by up to 0.6 %).  Those candidates — about 0.5 % — get the real
// thing: golden-section / parabolic minimisation of -y(x) over [i-1, i+1] (Brent 1973, the routine Praat calls), to a
// lag tolerance of 1e-3 samples (4e-5 relative at the shortest refined lag), one candidate at a time with all 32 lanes on
// each sinc evaluation.
__device__ PB_NOINLINE void pb_brent_refine(const float* __restrict__ r, int B, float fi, int depth, int lane, float* bx, float* by) {
    const float golden = 0.38196601125f, tol = 1.0e-3f;
    float a = fi - 1.0f, b = fi + 1.0f;
    float t = a + golden * (b - a);
    float x = t, v = t, w = t, fx = 0.0f, fv = 0.0f, fw = 0.0f;
    for (int iter = 0; iter < 32; iter++) {                   // every lane carries the same state: the loop is warp-uniform
        const float ft = -pb_sinc8(r, B, t, depth, lane, 32);
        if (iter == 0) { fx = fv = fw = ft; }
        else if (ft <= fx) { if (t < x) b = x; else a = x; v = w; w = x; x = t; fv = fw; fw = fx; fx = ft; }
        else {
            if (t < x) a = t; else b = t;
            if (ft <= fw || w == x) { v = w; w = t; fv = fw; fw = ft; }
            else if (ft <= fv || v == x || v == w) { v = t; fv = ft; }
        }
        const float range = b - a, mid = 0.5f * (a + b);
        if (fabsf(x - mid) + 0.5f * range <= 2.0f * tol) break;
        float step = golden * (x < mid ? b - x : a - x);
        if (fabsf(x - w) >= tol) {
            const float tt = (x - w) * (fx - fv);
            float q = (x - v) * (fx - fw);
            float pq = (x - v) * q - (x - w) * tt;
            q = 2.0f * (q - tt);
            if (q > 0.0f) pq = -pq; else q = -q;
            if (fabsf(pq) < fabsf(step * q) && pq > q * (a - x + 2.0f * tol) && pq < q * (b - x - 2.0f * tol)) step = pq / q;
        }
        if (fabsf(step) < tol) step = step > 0.0f ? tol : -tol;
        t = x + step;
    }
    *bx = x; *by = -fx;
}

// ------------------------------------------------------------------------------------------------ candidates of one frame
// One warp. r: normalised autocorrelation for lags 0..B (shared memory). scratch: 3*PB_MAXC words of shared memory.
// Maxima of r above voicingThreshold/2 between lag 2 and scan_lim-1 become candidates (the weakest is replaced when
// there are more than max_cand-1), each is then refined on the sinc-interpolated curve (depth 70, or 700 above 0.3/dx).
// Praat refines with Brent (tol 1e-10, <= 60 iterations); here the maximiser is found with FOUR evaluations:
// the two half-sample points next to the maximum, the vertex of the parabola through the best three of the
// five-point grid, and the vertex of the parabola through that point and its grid neighbours.  On speech this
// reproduces Brent's optimum to ~1e-4 relative in lag and ~1e-6 in strength (DESIGN.md "candidate refinement").
// Maxima below min_refine_lag stay above the pitch ceiling wherever in [i-1, i+1] their refinement lands: the path
// finder treats them as voiceless whatever their strength, so they keep their first-pass values.
__device__ __forceinline__ void pb_frame_candidates(const float* __restrict__ r, float* scratch, const PbPitchGeomDev& gm,
                                                    int lane, float* __restrict__ out_f, float* __restrict__ out_s,
                                                    uint8_t* __restrict__ out_n, const float* __restrict__ half_tab) {
    int* imax = (int*)(scratch + 2 * PB_MAXC); // lag of the maximum
    const int B = gm.brent_ixmax, lim = gm.scan_lim, maxc = gm.max_cand;
    // ---- the maxima, in lag order; slot arrays hold PB_MAXC-1 of them, anything beyond max_cand-1 is the rare path
    int total = 0;
    for (int base = 2; base < lim; base += 32) {
        const int i = base + lane;
        const bool pk = i < lim && r[i] > gm.half_voicing && r[i] > r[i - 1] && r[i] >= r[i + 1];
        const unsigned mask = __ballot_sync(PB_FULL_MASK, pk);
        const int slot = 1 + total + __popc(mask & ((1u << lane) - 1u));
        if (pk && slot < maxc) imax[slot] = i;
        total += __popc(mask);
    }
    const bool overflow = total > maxc - 1;
    int ncf = 1 + total;
    if (overflow) ncf = pb_candidates_overflow(r, scratch, gm, lane);      // out of line: tonal high-frequency content only
    __syncwarp();
    if (lane == 0) { out_f[0] = 0.0f; out_s[0] = 0.0f; *out_n = (uint8_t)ncf; }
    // ---- candidates that can never be voiced keep first-pass values (sorted by lag unless the overflow path ran)
    int c_first = 1;
    if (!overflow) {
        int nskip = 0;
        for (int c = 1 + lane; c < ncf; c += 32) {
            const int i = imax[c];
            const bool skip = i < gm.min_refine_lag;
            if (skip) {
                const float r0 = r[i], rm = r[i - 1], rp = r[i + 1];
                out_f[c] = __fdividef(gm.sr, (float)i + 0.5f * __fdividef(rp - rm, 2.0f * r0 - rm - rp));
                out_s[c] = r0;
            }
            nskip += skip;
        }
        c_first = 1 + pb_warp_sum_i(nskip);
    }
    // ---- refine the others on the sinc curve; the lanes are split evenly over the candidates of a round:
    //      32 lanes for a single candidate, 16 each for two, otherwise 8 each and 4 candidates per round
    int* flagged = (int*)scratch;                       // candidates whose peak is not parabola-shaped (see pb_brent_refine)
    int n_flagged = 0;
    const int nref = ncf - c_first;
    const int nl = nref <= 1 ? 32 : nref == 2 ? 16 : 8;
    const int per_round = 32 / nl;
    const int sub = lane / nl, sl = lane & (nl - 1);
    for (int c0 = c_first; c0 < ncf; c0 += per_round) {
        const int c = c0 + sub;
        const bool have = c < ncf;
        const int i = have ? imax[c] : 2;
        const float fi = (float)i;
        const float r0 = r[i], rm = r[i - 1], rp = r[i + 1];
        const float den0 = 2.0f * r0 - rm - rp;
        const float x_first = fi + ((have && den0 > 0.0f) ? 0.5f * __fdividef(rp - rm, den0) : 0.0f);   // Praat's first guess
        const int depth = !have ? 0 : (x_first < (1.0f / 0.3f)) ? 700 : 70;                    // f > 0.3/dx; idle groups do no work
        // The two half-sample points have phi = 1/2: their windowed-sinc coefficients do not depend on the candidate
        // (depth 70, away from the end of r), so both are dot products against one 70-entry table — no MUFU.
        const bool tab = have && depth == 70 && (B - i) >= 70;
        float ta = 0.0f, tb = 0.0f;
        if (tab) {
            for (int m = sl; m < 70; m += nl) {
                const float cm = half_tab[m];
                ta = fmaf(cm, r[abs(i - 1 - m)] + r[i + m], ta);
                tb = fmaf(cm, r[abs(i - m)] + r[i + 1 + m], tb);
            }
        }
        for (int o = 1; o < nl; o <<= 1) { ta += __shfl_xor_sync(PB_FULL_MASK, ta, o); tb += __shfl_xor_sync(PB_FULL_MASK, tb, o); }
        // four evaluations through ONE call site (code size): y(i-.5), y(i+.5), y(x1), y(x2)
        float xe = fi - 0.5f, ya = 0.0f, xc = fi, yc = r0, yl = 0.0f, yr = 0.0f, x1 = fi, y1 = 0.0f, y2 = 0.0f;
#ifndef PB_SIMT_EMU
#pragma unroll 1
#endif
        for (int e = 0; e < 4; e++) {
            float y = pb_sinc8(r, B, xe, (tab && e < 2) ? 0 : depth, sl, nl);
            if (tab && e < 2) y = e ? tb : ta;
            if (e == 0) { ya = y; xe = fi + 0.5f; }
            else if (e == 1) {
                // half-sample grid r[i-1], y(i-.5), r[i], y(i+.5), r[i+1]: best of the middle three and its neighbours
                const float yb = y;
                yl = ya; yr = yb;
                if (ya > yc && ya >= yb) { xc = fi - 0.5f; yc = ya; yl = rm; yr = r0; }
                else if (yb > yc) { xc = fi + 0.5f; yc = yb; yl = r0; yr = rp; }
                x1 = pb_parabola(xc - 0.5f, yl, xc, yc, xc + 0.5f, yr);
                x1 = fminf(fmaxf(x1, xc - 0.5f), xc + 0.5f);
                xe = x1;
            } else if (e == 2) {
                // second parabola: x1 with the grid points that bracket it
                y1 = y;
                float x2 = x1;
                if (x1 > xc) x2 = pb_parabola(xc, yc, x1, y1, xc + 0.5f, yr);
                else if (x1 < xc) x2 = pb_parabola(xc - 0.5f, yl, x1, y1, xc, yc);
                x2 = fminf(fmaxf(x2, fminf(x1, xc) - 0.25f), fmaxf(x1, xc) + 0.25f);
                x2 = fminf(fmaxf(x2, fi - 1.0f), fi + 1.0f);
                xe = x2;
            } else y2 = y;
        }
        float bx = xe, by = y2;
        if (y1 > by) { bx = x1; by = y1; }
        const bool flat = have && depth > 0 && !(by > yc);      // no vertex beat the best grid point: not a parabola-shaped peak
        if (yc > by) { bx = xc; by = yc; }
        // remember the flagged candidates (scratch words [0, PB_MAXC) are free here); they are redone after the loop, where
        // nothing of this iteration is live across the out-of-line call
        const unsigned fm = __ballot_sync(PB_FULL_MASK, flat && sl == 0);
        if (fm) {
            if (flat && sl == 0) flagged[n_flagged + __popc(fm & ((1u << lane) - 1u))] = c | (depth == 700 ? 0x100 : 0);
            n_flagged += __popc(fm);
        }
        if (by > 1.0f) by = __fdividef(1.0f, by);
        if (have && sl == 0) { out_f[c] = __fdividef(gm.sr, bx); out_s[c] = by; }
    }
    if (n_flagged) {
        __syncwarp();
        for (int k = 0; k < n_flagged; k++) {
            const int code = flagged[k], c = code & 0xff;
            float bx, by;
            pb_brent_refine(r, B, (float)imax[c], (code & 0x100) ? 700 : 70, lane, &bx, &by);
            if (by > 1.0f) by = __fdividef(1.0f, by);
            if (lane == 0) { out_f[c] = __fdividef(gm.sr, bx); out_s[c] = by; }
        }
    }
}

// ------------------------------------------------------------------------------------------------ sample prefetch (cp.async)
__device__ __forceinline__ void pb_cp_async16(void* smem_dst, const void* gmem_src) {
#ifdef PB_SIMT_EMU
    memcpy(smem_dst, gmem_src, 16);
#else
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
#endif
}
__device__ __forceinline__ void pb_cp_async_commit() {
#ifndef PB_SIMT_EMU
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int PENDING> __device__ __forceinline__ void pb_cp_async_wait() {
#ifndef PB_SIMT_EMU
    asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
#endif
}

// Where a frame pair's samples sit: part index of frame A's sample 0, distance to frame B, and the offset of frame A's
// sample `span_lo` inside the staged (16-byte aligned) range.
struct PbPairPos { long long start0; int hop; int shift; };

// Frame position, Praat's Sampled_indexToX / Sampled_xToLowIndex in float64 with explicit rounding (no FMA contraction):
// returns the 1-based part index of frame sample n = 0.
__device__ __forceinline__ long long pb_frame_start(double t1, double x1, int frame, const PbPitchGeomDev& gm) {
    const double t = __dadd_rn(t1, __dmul_rn((double)frame, gm.dt));
    const long long left = (long long)floor(__ddiv_rn(__dsub_rn(t, x1), gm.dx)) + 1;
    return left + 1 - gm.half_nw;
}

// Stage the samples the pair (frames fA, fA+1 of unit u) will read into `dst` with 16-byte cp.async copies issued by the
// GT threads of the group.  Only chunks entirely inside the pcm buffer are copied asynchronously; the (at most two) chunks
// straddling its ends are filled sample by sample.  Values outside the unit's part / file are masked at read time.
// The float64 frame positions of every pair, computed once by a small kernel ahead of the frames kernel: {start0, hop}.
// (They used to be computed inside the frames kernel: two float64 divisions per pair and ~250 instructions of its loop body.)
__global__ void __launch_bounds__(256)
pb_pair_pos_kernel(const PbUnitDev* __restrict__ units, const int32_t* __restrict__ pair_off, PbPitchGeomDev gm, int2* __restrict__ out) {
    for (int item = blockIdx.x * blockDim.x + threadIdx.x; item < gm.n_pairs; item += gridDim.x * blockDim.x) {
        const int u = pb_upper_unit(pair_off, gm.n_units, item);
        const PbUnitDev* ud = units + u;
        const int fA = 2 * (item - ud->pair_off);
        const long long s0 = pb_frame_start(ud->t1, ud->x1, fA, gm);
        const int hop = fA + 1 < ud->n_frames ? (int)(pb_frame_start(ud->t1, ud->x1, fA + 1, gm) - s0) : 0;
        out[item] = make_int2((int)s0, hop);
    }
}

template <int GT>
__device__ __forceinline__ PbPairPos pb_prefetch_pair(const int16_t* __restrict__ pcm, const PbUnitDev* __restrict__ units, int u, int2 sp,
                                                      const PbPitchGeomDev& gm, int span_lo, int span_len, int16_t* dst, int g) {
    const PbUnitDev* ud = units + u;
    PbPairPos pos;
    pos.start0 = sp.x;
    pos.hop = sp.y;
    // frame A sample n is part sample start0+n = pcm sample pcm_off + ix1 + start0 + n - 2
    const long long gs = ud->pcm_off + ud->ix1 + pos.start0 - 2 + span_lo;
    const long long byte0 = (long long)(size_t)pcm + 2 * gs;            // may lie outside the buffer: never dereferenced there
    const long long a0 = byte0 & ~15LL;
    pos.shift = (int)((byte0 - a0) >> 1);
    const int chunks = (pos.shift + span_len + pos.hop + 7) >> 3;
    const long long buf_lo = (long long)(size_t)pcm, buf_hi = buf_lo + 2 * gm.pcm_len;
    for (int j = g; j < chunks; j += GT) {
        const long long src = a0 + 16LL * j;
        if (src >= buf_lo && src + 16 <= buf_hi) pb_cp_async16(dst + 8 * j, (const void*)(size_t)src);
        else {
            PB_UNROLL for (int e = 0; e < 8; e++) {
                const long long b = src + 2 * e;
                dst[8 * j + e] = (b >= buf_lo && b + 2 <= buf_hi) ? *(const int16_t*)(size_t)b : (int16_t)0;
            }
        }
    }
    pb_cp_async_commit();
    return pos;
}

// ------------------------------------------------------------------------------------------------ K1 + K2
template <int LOG2N>
__global__ void __launch_bounds__(PbFftCfg<LOG2N>::WARPS_PER_CTA * 32, PbFftCfg<LOG2N>::MIN_CTAS)
pb_pitch_frames_kernel(const int16_t* __restrict__ pcm, const PbUnitDev* __restrict__ units, const int32_t* __restrict__ pair_off,
                       const int2* __restrict__ pairpos, PbPitchGeomDev gm, float* __restrict__ cand_f, float* __restrict__ cand_s,
                       uint8_t* __restrict__ ncand, float* __restrict__ intensity) {
    typedef PbFftCfg<LOG2N> C;
    constexpr int R = C::R, LR = C::LR, N = C::N, G = C::G, GT = C::GT;
    PB_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = warp / G, wg = warp % G;          // group in CTA, warp in group
    const int g = wg * 32 + lane;                       // thread in group = butterfly index
    const int bar_id = 1 + (G == 1 ? (group & 7) : group);   // one-warp groups may share an id: every arrival completes the barrier
    // per-group shared memory: FFT buffer, a small reduction scratch, the sample staging buffer
    const size_t group_bytes = (size_t)(C::BUF + 8 * G) * sizeof(float2) + (size_t)gm.pre_cap * sizeof(int16_t);
    unsigned char* gbase = smem_raw + (size_t)group * group_bytes;
    float2* buf = (float2*)gbase;
    float* red = (float*)(buf + C::BUF);                // [G][4] floats
    int16_t* pre = (int16_t*)(buf + C::BUF + 8 * G);    // [pre_cap] staged samples of one pair
    // CTA-wide: the half-sample sinc coefficients (72 floats after the last group's region)
    float* half_tab = (float*)(smem_raw + (size_t)C::GROUPS_PER_CTA * group_bytes);
    if (threadIdx.x < 72) half_tab[threadIdx.x] = threadIdx.x < 70 ? __ldg(&gm.half_tab[threadIdx.x]) : 0.0f;
    __syncthreads();
    const int B = gm.brent_ixmax;
    const int rstride = (B + 4) & ~1;
    float* rbase = (float*)buf;                         // r of frame A, r of frame B, then candidate scratch
    const int pk_lo = max(0, gm.half_nw - gm.half_period), pk_hi = min(gm.nw, gm.half_nw + gm.half_period);   // [lo, hi)
    const int mean_n0 = gm.half_nw - gm.nsamp_period;   // local mean spans frame samples [mean_n0, mean_n0 + 2 P)
    const int span_lo = min(0, mean_n0);                // staged range of frame samples: [span_lo, span_lo + span_len)
    const int span_len = max(gm.nw, mean_n0 + 2 * gm.nsamp_period) - span_lo;

    // Blocked distribution: a group walks a contiguous range of frame pairs, so consecutive iterations stay in the same
    // unit and the next pair's samples can be staged (cp.async, double buffered) while the current pair is transformed.
    const int n_groups = gridDim.x * C::GROUPS_PER_CTA;
    const int per_group = (gm.n_pairs + n_groups - 1) / n_groups;
    const int item_begin = (blockIdx.x * C::GROUPS_PER_CTA + group) * per_group;
    const int item_end = min(gm.n_pairs, item_begin + per_group);
    const bool has_items = item_begin < item_end;
    if (!has_items && !gm.phase_sync) return;
    int u_next = 0, u_next_end = 0;
    PbPairPos pos_next = {0, 0, 0};
    if (has_items) {
        u_next = pb_upper_unit(pair_off, gm.n_units, item_begin);
        u_next_end = pair_off[u_next + 1];
        pos_next = pb_prefetch_pair<GT>(pcm, units, u_next, pairpos[item_begin], gm, span_lo, span_len, pre, g);
    }
    int u = -1;
    PbUnitDev ud;
    for (int it = 0; it < per_group; it++) {
        // The loop body is ~60 KB of code, twice the SM's 32 KB instruction cache.  Starting every iteration together keeps
        // the CTA's warps within a few hundred instructions of each other, so one warp's instruction fetches serve all.
        if (gm.phase_sync && it % gm.phase_sync == 0) __syncthreads();     // phase_sync = every how many items
        const int item = item_begin + it;
        if (item >= item_end) {
            if (!gm.phase_sync) break;
            continue;
        }
        const PbPairPos pos = pos_next;
        if (u != u_next) { u = u_next; ud = units[u]; }
        const int16_t* sm = pre;
        pb_cp_async_wait<0>();                          // this pair's samples (requested during the previous pair) have landed
        pb_group_sync<G>(bar_id);                       // ... and are visible to every lane of the group
        const int fA = 2 * (item - ud.pair_off);
        const bool hasB = fA + 1 < ud.n_frames;
        const bool global_silent = ud.global_peak == 0.0;
        float2 v[R];
        float pkA = 0.0f, pkB = 0.0f, sA = 1.0f, sB = 1.0f;
        bool active = false;

        {
            {
                // ---- which frame samples exist: sample n of frame f is part sample start+n = file sample ix1+start+n-2;
                //      Praat zero-fills outside the part and outside the file
                int nlo[2], nhi[2], sb[2]; float lmean[2];
                PB_UNROLL for (int f = 0; f < 2; f++) {
                    const long long start = pos.start0 + (f ? pos.hop : 0);
                    long long lo = 1 - start, hi = ud.nx - start + 1;
                    const long long lo2 = 2 - ud.ix1 - start, hi2 = (long long)ud.file_nx - ud.ix1 - start + 2;
                    if (lo2 > lo) lo = lo2; if (hi2 < hi) hi = hi2;
                    const long long BIG = 1 << 30;
                    nlo[f] = (int)(lo < -BIG ? -BIG : (lo > BIG ? BIG : lo));
                    nhi[f] = (int)(hi < -BIG ? -BIG : (hi > BIG ? BIG : hi));
                    sb[f] = pos.shift - span_lo + (f ? pos.hop : 0);       // staged index of frame sample n is sb[f] + n
                    // local mean: one longest period to both sides of the frame centre (exact in integers)
                    int s = 0;
                    if (nlo[f] <= mean_n0 && nhi[f] >= mean_n0 + 2 * gm.nsamp_period) {   // interior frame: no range checks
                        const int16_t* ms = sm + sb[f] + mean_n0;
                        for (int q = lane; q < 2 * gm.nsamp_period; q += 32) s += (int)ms[q];
                    } else {
                        for (int q = lane; q < 2 * gm.nsamp_period; q += 32) {
                            const int n = mean_n0 + q;
                            s += (n >= nlo[f] && n < nhi[f]) ? (int)sm[sb[f] + n] : 0;
                        }
                    }
                    s = pb_warp_sum_i(s);
                    lmean[f] = (float)(((double)s / 32768.0) / (double)(2 * gm.nsamp_period));
                }
                if (!hasB) { nlo[1] = 0; nhi[1] = 0; sb[1] = sb[0]; }
                // ---- window both frames into the FFT buffer, z = a + i b (natural order, zero padded); a compact
                //      rolled loop: the range checks live here once instead of in 32 unrolled copies
                float mxA = 0.0f, mxB = 0.0f;
                const float2 nmean = make_float2(-lmean[0], -lmean[1]), q15 = make_float2(1.0f / 32768.0f, 1.0f / 32768.0f);
                const float hb = hasB ? 1.0f : 0.0f;
                if (nlo[0] <= 0 && nhi[0] >= gm.nw && nlo[1] <= 0 && nhi[1] >= gm.nw) {
                    // interior pair (all but the first / last frames of a unit): no range checks, packed-pair arithmetic
                    const int16_t* pa = sm + sb[0]; const int16_t* pb = sm + sb[1];
                    for (int n = g; n < gm.nw; n += GT) {
                        const float w = __ldg(&gm.window[n]);
                        const float2 ab = __fmul2_rn(__ffma2_rn(make_float2((float)pa[n], (float)pb[n]), q15, nmean), make_float2(w, w * hb));
                        const float aa = fabsf(ab.x), bb = fabsf(ab.y);
                        mxA = fmaxf(mxA, aa); mxB = fmaxf(mxB, bb);
                        if (n >= pk_lo && n < pk_hi) { pkA = fmaxf(pkA, aa); pkB = fmaxf(pkB, bb); }
                        buf[pb_pad5(n)] = ab;
                    }
                } else {
                    for (int n = g; n < gm.nw; n += GT) {
                        const float w = __ldg(&gm.window[n]);
                        const int sa = (n >= nlo[0] && n < nhi[0]) ? (int)sm[sb[0] + n] : 0;
                        const int sbv = (n >= nlo[1] && n < nhi[1]) ? (int)sm[sb[1] + n] : 0;
                        const float2 ab = __fmul2_rn(__ffma2_rn(make_float2((float)sa, (float)sbv), q15, nmean), make_float2(w, w * hb));
                        const float aa = fabsf(ab.x), bb = fabsf(ab.y);
                        mxA = fmaxf(mxA, aa); mxB = fmaxf(mxB, bb);
                        if (n >= pk_lo && n < pk_hi) { pkA = fmaxf(pkA, aa); pkB = fmaxf(pkB, bb); }
                        buf[pb_pad5(n)] = ab;
                    }
                }
                for (int n = g + ((gm.nw - g + GT - 1) / GT) * GT; n < N; n += GT) buf[pb_pad5(n)] = make_float2(0.0f, 0.0f);   // zero padding
                mxA = pb_warp_max(mxA); mxB = pb_warp_max(mxB); pkA = pb_warp_max(pkA); pkB = pb_warp_max(pkB);
                if (G > 1) {
                    if (lane == 0) { red[wg * 4 + 0] = mxA; red[wg * 4 + 1] = mxB; red[wg * 4 + 2] = pkA; red[wg * 4 + 3] = pkB; }
                    pb_group_sync<G>(bar_id);
                    for (int k = 0; k < G; k++) {
                        mxA = fmaxf(mxA, red[k * 4 + 0]); mxB = fmaxf(mxB, red[k * 4 + 1]);
                        pkA = fmaxf(pkA, red[k * 4 + 2]); pkB = fmaxf(pkB, red[k * 4 + 3]);
                    }
                    pb_group_sync<G>(bar_id);
                }
                active = !global_silent && (pkA > 0.0f || pkB > 0.0f);
                // bring both frames to comparable magnitude (power-of-two scales are exact and cancel in r = ac/ac[0]):
                // keeps the weaker frame of a pair out of the stronger one's rounding noise
                // (exponent arithmetic on the float bits: scale = 2^-(exponent of the maximum))
                sA = mxA > 0.0f ? __int_as_float((254 - ((__float_as_int(mxA) >> 23) & 0xff)) << 23) : 1.0f;
                sB = mxB > 0.0f ? __int_as_float((254 - ((__float_as_int(mxB) >> 23) & 0xff)) << 23) : 1.0f;
                pb_group_sync<G>(bar_id);               // the windowed frames are in the buffer
            }
        }
        // the staged samples are consumed: request the next pair's now, they land while this pair is transformed
        if (item + 1 < item_end) {
            if (item + 1 >= u_next_end) { do { u_next++; u_next_end = pair_off[u_next + 1]; } while (item + 1 >= u_next_end); }
            pos_next = pb_prefetch_pair<GT>(pcm, units, u_next, pairpos[item + 1], gm, span_lo, span_len, pre, g);
        }
        // ---- two FFTs x two passes through ONE copy of the butterfly code: FFT (step >> 1), pass (step & 1).  The step
        //      index is made opaque so the optimiser neither peels nor unswitches the loop (either duplicates ~450
        //      instructions of butterflies, and instruction fetch is what this kernel stalls on).
        int n_steps = active ? 4 : 0;
        asm volatile("" : "+r"(n_steps));
#ifndef PB_SIMT_EMU
#pragma unroll 1
#endif
        for (int it = 0; it < n_steps; it++) {
            int step = it;
            asm volatile("" : "+r"(step));
            int pass = step & 1;
            asm volatile("" : "+r"(pass));              // ... nor thread the step == 2 path (where pass is known) through a private copy
            if (step == 2) {
                // ---- power spectra of both frames: P_a = |Z_k + conj Z_-k|^2 / 4, P_b = |Z_k - conj Z_-k|^2 / 4
                for (int k = g; k <= N / 2; k += GT) {
                    const int k2 = (N - k) & (N - 1);
                    const float2 za = buf[pb_pad5(k)], zb = buf[pb_pad5(k2)];
                    const float S = za.x * za.x + za.y * za.y + zb.x * zb.x + zb.y * zb.y;
                    const float Cc = 2.0f * (za.x * zb.x - za.y * zb.y);
                    const float2 w = make_float2(S + Cc, S - Cc);
                    buf[pb_pad5(k)] = w; buf[pb_pad5(k2)] = w;
                }
                pb_group_sync<G>(bar_id);
            }
            {
                // pass 1 reads natural order (pad5): the windowed frames (step 0) or the spectra (step 2);
                // pass 2 reads the pass-1 layout and applies the inter-pass twiddles
                if (R == 32) {
                    // i = g + 32 G t  ->  i + (i >> 5) = g + (g >> 5) + 33 G t: one base, compile-time offsets
                    const float2* src = buf + (g + (g >> 5));
                    PB_UNROLL for (int t = 0; t < R; t++) v[t] = src[t * (33 * G)];
                } else {
                    const int sh = pass ? LR : 5;
                    PB_UNROLL for (int t = 0; t < R; t++) { const int i = g + t * (N / R); v[t] = buf[i + (i >> sh)]; }
                }
                if (step == 0) { const float2 sc = make_float2(sA, sB); PB_UNROLL for (int t = 0; t < R; t++) v[t] = __fmul2_rn(v[t], sc); }
                if (pass) {
                    const int k = g & (R - 1);
                    PB_UNROLL for (int t = 1; t < R; t++) {
                        const float2 w = __ldg(&gm.tw_a[t * R + k]);
                        const float2 x = v[t];
                        v[t] = __ffma2_rn(make_float2(x.y, x.y), make_float2(-w.y, w.x), __fmul2_rn(make_float2(x.x, x.x), w));   // x * w
                    }
                }
                pb_group_sync<G>(bar_id);               // every load of this step is done before any store
            }
            pb_dft<R>(v);
            {
                // pass 1 (Ns = 1): out[g*R + t], skew (index >> LR);  pass 2 (Ns = R): out[(g/R) R^2 + g%R + t R], skew 5
                if (R == 32) {
                    // pass 1: 32 g + t -> + (index >> 5) = 33 g + t;  pass 2: (g>>5) 1024 + (g&31) + 32 t -> (g>>5) 1056 + (g&31) + 33 t
                    float2* dst = buf + (pass ? ((g >> 5) * 1056 + (g & 31)) : 33 * g);
                    const int ds = pass ? 33 : 1;
                    PB_UNROLL for (int t = 0; t < R; t++) dst[t * ds] = v[pb_bitrev(t, LR)];
                } else {
                    const int ob = pass ? ((g >> LR) * (R * R) + (g & (R - 1))) : g * R;
                    const int os = pass ? R : 1;
                    const int osh = pass ? 5 : LR;
                    PB_UNROLL for (int t = 0; t < R; t++) { const int o = ob + t * os; buf[o + (o >> osh)] = v[pb_bitrev(t, LR)]; }
                }
            }
            pb_group_sync<G>(bar_id);
            if (pass && C::F > 1) {
                // ---- final pass (radix F, Ns = R*R): butterflies are in place
                constexpr int F = C::F > 1 ? C::F : 2, LF = pb_ilog2(F);
                PB_UNROLL for (int b = 0; b < C::FB; b++) {
                    const int j = g + b * GT;             // 0 .. N/F-1 = R*R-1
                    float2 a[F];
                    PB_UNROLL for (int t = 0; t < F; t++) a[t] = buf[pb_pad5(j + t * (R * R))];
                    PB_UNROLL for (int t = 1; t < F; t++) {
                        const float2 w = __ldg(&gm.tw_b[t * (R * R) + j]);
                        const float2 x = a[t];
                        a[t] = make_float2(x.x * w.x - x.y * w.y, x.x * w.y + x.y * w.x);
                    }
                    pb_dft<F>(a);
                    PB_UNROLL for (int t = 0; t < F; t++) buf[pb_pad5(j + t * (R * R))] = a[pb_bitrev(t, LF)];
                }
                pb_group_sync<G>(bar_id);
            }
        }

        if (active) {
            // ---- r[lag] = ac[lag] / (ac[0] * windowR[lag]) for lags 0..B+1, into shared memory
            float ra[C::RPL], rb[C::RPL];
            const float2 ac0 = buf[0];
            const float iA = ac0.x > 0.0f ? 1.0f / ac0.x : 0.0f, iB = ac0.y > 0.0f ? 1.0f / ac0.y : 0.0f;
            PB_UNROLL for (int q = 0; q < C::RPL; q++) {
                const int lag = g + q * GT;
                ra[q] = 0.0f; rb[q] = 0.0f;
                if (lag <= B + 1) {
                    const float2 a = buf[pb_pad5(lag)];
                    const float iw = lag <= B ? __ldg(&gm.inv_wr[lag]) : 0.0f;
                    ra[q] = a.x * iA * iw; rb[q] = a.y * iB * iw;
                }
            }
            pb_group_sync<G>(bar_id);
            PB_UNROLL for (int q = 0; q < C::RPL; q++) {
                const int lag = g + q * GT;
                if (lag <= B + 1) { rbase[lag] = lag == 0 ? 1.0f : ra[q]; rbase[rstride + lag] = lag == 0 ? 1.0f : rb[q]; }
            }
            pb_group_sync<G>(bar_id);
        }
        // ---- candidates: warp 0 of the group takes frame A, warp 1 (or warp 0 again) frame B
        {
            const float gpk = (float)ud.global_peak;
#ifndef PB_SIMT_EMU
#pragma unroll 1
#endif
            for (int f = 0; f < 2; f++) {
                const int owner = (G > 1) ? f : 0;
                if (wg != owner || (f == 1 && !hasB)) continue;
                const int64_t fr = ud.frame_off + fA + f;
                const float pk = f ? pkB : pkA;
                float* of = cand_f + fr * gm.max_cand; float* os = cand_s + fr * gm.max_cand;
                if (!active || pk == 0.0f) {
                    if (lane == 0) { of[0] = 0.0f; os[0] = 0.0f; ncand[fr] = 1; intensity[fr] = 0.0f; }
                } else {
                    if (lane == 0) { const float it = pk / gpk; intensity[fr] = it > 1.0f ? 1.0f : it; }
                    pb_frame_candidates(rbase + f * rstride, rbase + 2 * rstride + f * (3 * PB_MAXC), gm, lane, of, os, ncand + fr, half_tab);
                }
            }
        }
        pb_group_sync<G>(bar_id);     // buf and the staging buffer just consumed are reused by the next iterations
    }
}
