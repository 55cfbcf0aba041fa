// pb_pitch_cand.cuh — K2: normalised autocorrelation -> pitch candidates, one warp per frame.
//
// Follows the second half of Sound_into_PitchFrame (Praat fon/Sound_to_Pitch.cpp): maxima of r above voicingThreshold / 2,
// Praat's candidate insertion (incl. the replace-the-weakest path) and the refinement of every candidate on the
// depth-70 / 700 windowed-sinc curve (NUM_interpolate_sinc + NUMimproveMaximum).  r of every active frame was written to
// a global scratch by K1 (pb_pitch_acf.cuh); here each warp pulls its next frame's r into shared memory with one
// cp.async.bulk (TMA) while it works on the current one (two buffers per warp, one mbarrier each).  Frames are handed
// out in chunks of 32 slots through an atomic counter: silent stretches cost nothing and the warps stay balanced.
// The kernel needs ~48 registers, so 10 CTAs x 4 warps are resident per SM and the shared-memory / special-function
// latencies of the sinc loops are hidden by other warps (in the fused round-1 kernel they ran at 4 warps per scheduler).
#pragma once
#include "pb_async.cuh"
#include "pb_pitch.cuh"

// ------------------------------------------------------------------------------------------------ sinc interpolation
// r in shared memory is preceded by its own mirror image, r[-j] = r[j] for j < PB_MIR, so the depth-70 sums (and the
// peak scan's left neighbour of lag 0) index it directly; only the depth-700 evaluations (candidates above 0.3 / dx, i.e. lags
// below 3.33: refined when the ceiling is close to the Nyquist frequency, or when a frame has more maxima than candidate slots)
// reach further down: those take |index| and the direct cosine.
#define PB_MIR 72

// Praat NUM_interpolate_sinc (melder/NUMinterpol.cpp) on y[1..2B+1] = r[-B..B] at lag x, by `nl` cooperating lanes
// (8, 16 or 32, warp-uniform; sl = lane within the group; every lane of the group passes the same x).  With
// phi = frac(x), il = floor(x), D = min(depth, B - il) the usable depth:
//   y(x) = sin(pi phi)/(2 pi) * sum_{m<D} (-1)^m [ r[il-m]   (1 + cos(pi (phi+m)   / (phi+D)))   / (phi+m)
//                                                + r[il+1+m] (1 + cos(pi (1-phi+m) / (1-phi+D))) / (1-phi+m) ]
// Lane sl takes one side (sl & 1) and every (nl/2)-th m starting at sl >> 1, so side, sign and the window scale are
// loop-invariant.  The window's cosine advances by a fixed angle per term: it is carried by the three-term recurrence
// c[m+1] = 2 cos(delta) c[m] - c[m-1] (one FFMA) instead of a multiply and a MUFU per term; over the <= 18 terms a lane
// sums its error stays below 1e-5 of a factor that multiplies the smallest terms.  Loop body: one shared load, one
// MUFU (rcp) and five FP32 operations.  Depth 700 keeps the direct cosine (up to 350 terms per lane) and |index|.
// 1 / x in one MUFU (x is a sample distance >= 2^-24 here: no range fix-up as in __fdividef)
__device__ __forceinline__ float pb_rcp(float x) {
#ifdef PB_SIMT_EMU
    return 1.0f / x;
#else
    float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#endif
}

__device__ __forceinline__ float pb_sinc8(const float* __restrict__ r, int B, float x, int depth, int sl, int nl = 8) {
    const float fl = floorf(x);
    const float phi = x - fl;
    const int il = (int)fl;
    int D = B - il; if (depth < D) D = depth;
    float acc = 0.0f;
    if (phi == 0.0f) {                                        // on a sample: Praat returns y[x] (no early return:
        if (sl == 0 && depth > 0) acc = r[il];                // the other groups of the warp still shuffle below)
    } else if (D > 0) {
        const int side = sl & 1, j0 = sl >> 1, hs = nl >> 1;  // hs (4, 8, 16) is even: the sign of a lane's terms is fixed
        const float e = side ? 1.0f - phi : phi;
        const float k = __fdividef(PB_PI_F, e + (float)D);
        const float fhs = (float)hs;
        float d = e + (float)j0;
        int idx = side ? il + 1 + j0 : il - j0;
        const int step = side ? hs : -hs;
        if (depth > 70) {                                     // rare: candidates above 0.3 / dx (lags below 3.33)
            for (int m = j0; m < D; m += hs) {
                const float yv = r[abs(idx)];
                acc += __fdividef(yv * (1.0f + __cosf(d * k)), d);
                d += fhs; idx += step;
            }
        } else {
            const float dl = fhs * k;
            const float tw = 2.0f * __cosf(dl);
            float c = __cosf(d * k), cp = __cosf(d * k - dl);
            const float* p = r + idx;
            float a2 = 0.0f;
            const int n_it = D > j0 ? (D - j0 + hs - 1) >> (__ffs(hs) - 1) : 0;      // hs is 4, 8 or 16
#ifndef PB_SIMT_EMU
#pragma unroll 2
#endif
            for (int it = 0; it < n_it; it++) {
                const float t = *p * pb_rcp(d);
                acc += t; a2 = fmaf(t, c, a2);
                const float cn = fmaf(tw, c, -cp); cp = c; c = cn;
                d += fhs; p += step;
            }
            acc += a2;
        }
        if (j0 & 1) acc = -acc;
        acc *= sinpif(phi) * (0.5f / PB_PI_F);
    }
    // sum over the nl >= 8 lanes of the group
    acc += __shfl_xor_sync(PB_FULL_MASK, acc, 1); acc += __shfl_xor_sync(PB_FULL_MASK, acc, 2); acc += __shfl_xor_sync(PB_FULL_MASK, acc, 4);
    if (nl > 8) acc += __shfl_xor_sync(PB_FULL_MASK, acc, 8);
    if (nl > 16) acc += __shfl_xor_sync(PB_FULL_MASK, acc, 16);
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
struct PbOverflowArgs { int B, lim, maxc; float half_voicing, octave_cost, min_pitch, sr; };      // by value: a reference to the kernel's geometry struct would force a local-memory copy of it
__device__ PB_NOINLINE int pb_candidates_overflow(const float* __restrict__ r, float* scratch, const PbOverflowArgs gm, int lane) {
    float* cf = scratch;
    float* cs = scratch + PB_MAXC;
    int* imax = (int*)(scratch + 2 * PB_MAXC);
    const int B = gm.B, lim = gm.lim, maxc = gm.maxc;
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
                __syncwarp();                               // every lane has read its slot before lane 0 overwrites one
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
    // four consecutive lags per lane and pass (one 16-byte load plus the two neighbours); maxima come out in lag order:
    // pass, lane, bit
    int total = 0;
    const float hv = gm.half_voicing;
    for (int base = 0; base < lim; base += 128) {
        const int i0 = base + 4 * lane;
        unsigned m4 = 0;
        if (i0 < lim) {
            const float4 q = *reinterpret_cast<const float4*>(r + i0);
            const float pv = r[i0 - 1], nx = r[i0 + 4];
            m4 = ((unsigned)((q.x > hv) & (q.x > pv) & (q.x >= q.y))) | ((unsigned)((q.y > hv) & (q.y > q.x) & (q.y >= q.z)) << 1) |
                 ((unsigned)((q.z > hv) & (q.z > q.y) & (q.z >= q.w)) << 2) | ((unsigned)((q.w > hv) & (q.w > q.z) & (q.w >= nx)) << 3);
            if (i0 == 0) m4 &= ~3u;                              // candidate lags start at 2
            const int rem = lim - i0;
            if (rem < 4) m4 &= (1u << rem) - 1u;
        }
        const int cnt = __popc(m4);
        int incl = cnt;
        PB_UNROLL for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(PB_FULL_MASK, incl, o); if (lane >= o) incl += t; }
        int slot = 1 + total + incl - cnt;
        while (m4) {
            const int j = __ffs((int)m4) - 1;
            m4 &= m4 - 1;
            if (slot < maxc) imax[slot] = i0 + j;
            slot++;
        }
        total += __shfl_sync(PB_FULL_MASK, incl, 31);
    }
    const bool overflow = total > maxc - 1;
    int ncf = 1 + total;
    if (overflow) {                                     // out of line: tonal high-frequency content only
        const PbOverflowArgs oa = {B, lim, maxc, gm.half_voicing, gm.octave_cost, gm.min_pitch, gm.sr};
        ncf = pb_candidates_overflow(r, scratch, oa, lane);
    }
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
        // each lane of the group takes a run of consecutive m: both sums walk r outwards from the maximum, so every step
        // loads one new value per side and reuses the previous one
        float ta = 0.0f, tb = 0.0f;
        if (tab) {
            const int per = (70 + nl - 1) / nl;                 // 9, 5 or 3 coefficients per lane
            const int m0 = sl * per, m1 = min(70, m0 + per);
            if (m0 < m1) {
                float lc = r[i - m0], rc = r[i + m0];
                for (int m = m0; m < m1; m++) {
                    const float cm = half_tab[m];
                    const float ln = r[i - 1 - m], rn = r[i + 1 + m];
                    ta = fmaf(cm, ln + rc, ta);
                    tb = fmaf(cm, lc + rn, tb);
                    lc = ln; rc = rn;
                }
            }
        }
        PB_UNROLL for (int o = 1; o < 8; o <<= 1) { ta += __shfl_xor_sync(PB_FULL_MASK, ta, o); tb += __shfl_xor_sync(PB_FULL_MASK, tb, o); }
        if (nl > 8) { ta += __shfl_xor_sync(PB_FULL_MASK, ta, 8); tb += __shfl_xor_sync(PB_FULL_MASK, tb, 8); }
        if (nl > 16) { ta += __shfl_xor_sync(PB_FULL_MASK, ta, 16); tb += __shfl_xor_sync(PB_FULL_MASK, tb, 16); }
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


// ------------------------------------------------------------------------------------------------ K2
#define PB_CAND_WARPS 4

// MIN_CTAS: residency target (10 -> 48 registers with a few spilled words, 8 -> 64 registers); chosen at launch (PB_CAND_CTAS)
template <int MIN_CTAS>
__global__ void __launch_bounds__(PB_CAND_WARPS * 32, MIN_CTAS)
pb_pitch_cand_kernel(const float* __restrict__ racf, const long long* __restrict__ slot_fr, int n_slots, int rstride_g, PbPitchGeomDev gm,
                     float* __restrict__ cand_f, float* __restrict__ cand_s, uint8_t* __restrict__ ncand, unsigned* __restrict__ work_counter) {
    PB_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // per warp: two r buffers, each behind PB_MIR words for its mirror image (16-byte aligned: rstride_g and PB_MIR are multiples
    // of 4), candidate scratch, two mbarriers; CTA-wide: half-sample table
    const int rrow = rstride_g + PB_MIR;
    const size_t warp_bytes = (size_t)(2 * rrow + 3 * PB_MAXC) * sizeof(float) + 2 * sizeof(pbMbar);
    unsigned char* wbase = smem_raw + (size_t)warp * warp_bytes;
    float* rbuf = (float*)wbase;
    float* scratch = rbuf + 2 * rrow;
    pbMbar* mbar = (pbMbar*)(scratch + 3 * PB_MAXC);
    float* half_tab = (float*)(smem_raw + (size_t)PB_CAND_WARPS * warp_bytes);
    if (threadIdx.x < 72) half_tab[threadIdx.x] = threadIdx.x < 70 ? __ldg(&gm.half_tab[threadIdx.x]) : 0.0f;
    if (lane == 0) { pb_mbar_init(mbar, 1); pb_mbar_init(mbar + 1, 1); }
    pb_mbar_init_fence();
    __syncthreads();
    unsigned phase0 = 0, phase1 = 0;
    const unsigned r_bytes = (unsigned)rstride_g * sizeof(float);
    const int mc = gm.max_cand;
    for (;;) {
        unsigned c = 0;
        if (lane == 0) c = atomicAdd(work_counter, 1u);
        c = __shfl_sync(PB_FULL_MASK, c, 0);
        const long long s0 = (long long)c * 32;
        if (s0 >= n_slots) break;
        const long long fr_l = s0 + lane < n_slots ? slot_fr[s0 + lane] : -1;
        unsigned mask = __ballot_sync(PB_FULL_MASK, fr_l >= 0);
        if (!mask) continue;
        int b = 0;
        // first active slot of the chunk into buffer 0
        if (lane == 0) {
            const int j = __ffs((int)mask) - 1;
            pb_mbar_expect_tx(mbar, r_bytes);
            pb_bulk_g2s(rbuf + PB_MIR, racf + (size_t)(s0 + j) * rstride_g, r_bytes, mbar);
        }
        while (mask) {
            const int j = __ffs((int)mask) - 1;
            mask &= mask - 1;
            if (mask && lane == 0) {                   // the next active slot lands in the other buffer meanwhile
                const int jn = __ffs((int)mask) - 1;
                pb_mbar_expect_tx(mbar + (b ^ 1), r_bytes);
                pb_bulk_g2s(rbuf + (b ^ 1) * rrow + PB_MIR, racf + (size_t)(s0 + jn) * rstride_g, r_bytes, mbar + (b ^ 1));
            }
            if (b) { pb_mbar_wait(mbar + 1, phase1); phase1 ^= 1u; } else { pb_mbar_wait(mbar, phase0); phase0 ^= 1u; }
#ifdef PB_SIMT_EMU
            __syncwarp();
#endif
            const long long fr = __shfl_sync(PB_FULL_MASK, fr_l, j);
            float* r = rbuf + b * rrow + PB_MIR;
            for (int q = 1 + lane; q < PB_MIR && q < rstride_g; q += 32) r[-q] = r[q];     // the mirror image
            __syncwarp();
            pb_frame_candidates(r, scratch, gm, lane, cand_f + fr * mc, cand_s + fr * mc, ncand + fr, half_tab);
            __syncwarp();                              // every lane is done with this buffer before it is refilled
            b ^= 1;
        }
    }
}
