// pb_pitch.cuh — device code of the F0 path (K0 unit stats, K1+K2 frames -> candidates, K3 path finder).
//
// Computes what parselmouth's Sound.to_pitch(pitch_floor, pitch_ceiling) (Praat Sound_to_Pitch_ac, AC_HANNING)
// computes for the reference's get_median_pitch (/root/reference/Code/audioPipeline.py:326-335), one
// independent analysis per unit (file, t0, t1).  B200-first layout:
//   * a GROUP of G warps (G = 1 for FFT sizes <= 1024) owns a PAIR of consecutive frames end to end: PCM load,
//     local mean, Hanning window, one complex FFT carrying both real frames, power spectra, second FFT back to
//     lags, normalisation, peak picking and sinc refinement.  The autocorrelation never leaves shared memory.
//   * FFTs are register-blocked: every lane runs one radix-R butterfly (R = 32 for N >= 1024) per pass on 2R
//     registers; passes exchange data through a skew-padded shared buffer (bank-conflict free), in place.
//   * no tensor cores (nothing here is a dense contraction), no block-wide barriers when G == 1.
#pragma once
#include "pb_rt.h"
#include <math.h>

#define PB_MAXC 32            // hard cap on candidates per frame (Praat default 15)
#define PB_PI_F 3.14159265358979323846f

struct PbUnitDev {
    int64_t pcm_off;      // first sample of the unit's file in the pcm buffer
    int64_t ix1;          // 1-based file index of the first sample of the extracted part (Sound_extractPart)
    int64_t nx;           // samples in the extracted part (zero-filled beyond the file)
    int64_t frame_off;    // first frame of this unit in the per-frame arrays
    double x1;            // time of the first sample of the part
    double t1;            // time of the first analysis frame (Sampled_shortTermAnalysis)
    double mean;          // K0: mean of the part
    double global_peak;   // K0: max |x - mean| over the part
    int32_t file_nx;
    int32_t n_frames;
    int32_t pair_off;     // first frame pair of this unit
    int32_t out_index;    // index of the unit in the caller's arrays
};

struct PbPitchGeomDev {
    int nw;               // nsamp_window
    int half_nw;          // halfnsamp_window
    int nsamp_period;
    int half_period;      // halfnsamp_period
    int scan_lim;         // min(maximumLag, brent_ixmax): candidate lags are 2 .. scan_lim-1
    int brent_ixmax;      // B: r is known for lags 0..B
    int max_cand;         // maxnCandidates (incl. the voiceless one)
    int n_units;
    int n_pairs;
    int pad0;
    float sr;             // 1/dx
    float half_voicing;   // 0.5 * voicingThreshold
    float octave_cost;
    float min_pitch;
    double dx, dt;
    double ceiling, silence_threshold, voicing_threshold, octave_cost_d, octave_jump_cost, voiced_unvoiced_cost;
    const float* window;  // [nw]   Hanning window, window[n] = Praat window[n+1]
    const float* inv_wr;  // [B+2]  1 / (normalised autocorrelation of the window)
    const float2* tw_a;   // [R][R]     pass-2 twiddles  exp(-2 pi i t k / R^2)
    const float2* tw_b;   // [F][R*R]   final-pass twiddles exp(-2 pi i t k / N)
};

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ int pb_sample(const int16_t* __restrict__ pcm, const PbUnitDev& u, long long i) {
    // i: 1-based index in the extracted part; Praat zero-fills outside the file
    if (i < 1 || i > u.nx) return 0;
    long long fi = u.ix1 - 1 + (i - 1);
    if (fi < 0 || fi >= (long long)u.file_nx) return 0;
    return (int)pcm[u.pcm_off + fi];
}

__device__ __forceinline__ float pb_warp_max(float v) {
    PB_UNROLL for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(PB_FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ int pb_warp_sum_i(int v) {
    PB_UNROLL for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PB_FULL_MASK, v, o);
    return v;
}

// find u with off[u] <= x < off[u+1]   (off has n+1 entries, non-decreasing)
__device__ __forceinline__ int pb_upper_unit(const int32_t* __restrict__ off, int n, int x) {
    int lo = 0, hi = n;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (off[mid] <= x) lo = mid; else hi = mid; }
    return lo;
}

// ------------------------------------------------------------------------------------------------ K0
// Per-unit mean and global peak (Praat: globalPeak = max |x - mean| over the whole analysed sound).
// Integer sum / min / max of the int16 samples are exact, so mean and peak equal the float64 reference's.
__global__ void __launch_bounds__(256) pb_unit_stats_kernel(const int16_t* __restrict__ pcm, PbUnitDev* __restrict__ units, int n_units) {
    __shared__ long long s_sum[8];
    __shared__ int s_min[8], s_max[8];
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        PbUnitDev ud = units[u];
        // the part of [ix1-1, ix1-1+nx) that lies inside the file
        long long a = ud.ix1 - 1, b = a + ud.nx;
        long long lo = a < 0 ? 0 : a, hi = b > ud.file_nx ? ud.file_nx : b;
        bool padded = (lo > a) || (hi < b) || (hi <= lo);
        long long sum = 0; int mn = 32767, mx = -32768;
        const int16_t* p = pcm + ud.pcm_off;
        for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            int v = p[i]; sum += v; mn = min(mn, v); mx = max(mx, v);
        }
        PB_UNROLL for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(PB_FULL_MASK, sum, o);
            mn = min(mn, __shfl_xor_sync(PB_FULL_MASK, mn, o));
            mx = max(mx, __shfl_xor_sync(PB_FULL_MASK, mx, o));
        }
        int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        if (l == 0) { s_sum[w] = sum; s_min[w] = mn; s_max[w] = mx; }
        __syncthreads();
        if (threadIdx.x == 0) {
            int nwarp = (blockDim.x + 31) >> 5;
            for (int k = 1; k < nwarp; k++) { sum += s_sum[k]; mn = min(mn, s_min[k]); mx = max(mx, s_max[k]); }
            if (padded) { mn = min(mn, 0); mx = max(mx, 0); }
            if (hi <= lo) { mn = 0; mx = 0; }
            double mean = ((double)sum / 32768.0) / (double)ud.nx;
            double p1 = fabs((double)mx / 32768.0 - mean), p2 = fabs((double)mn / 32768.0 - mean);
            units[u].mean = mean;
            units[u].global_peak = p1 > p2 ? p1 : p2;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ radix-R DFT in registers
__host__ __device__ constexpr float pb_c32(int m) {   // cos(2 pi m / 32), 0 <= m <= 16
    return m == 0 ? 1.0f : m == 1 ? 0.98078528040323043f : m == 2 ? 0.92387953251128674f : m == 3 ? 0.83146961230254524f
         : m == 4 ? 0.70710678118654752f : m == 5 ? 0.55557023301960218f : m == 6 ? 0.38268343236508977f
         : m == 7 ? 0.19509032201612825f : m == 8 ? 0.0f : m == 9 ? -0.19509032201612825f : m == 10 ? -0.38268343236508977f
         : m == 11 ? -0.55557023301960218f : m == 12 ? -0.70710678118654752f : m == 13 ? -0.83146961230254524f
         : m == 14 ? -0.92387953251128674f : m == 15 ? -0.98078528040323043f : -1.0f;
}
__host__ __device__ constexpr float pb_s32(int m) { return pb_c32(m <= 8 ? 8 - m : m - 8); }  // sin(2 pi m / 32)
__host__ __device__ constexpr int pb_bitrev(int i, int bits) {
    int r = 0;
    for (int b = 0; b < bits; b++) if (i & (1 << b)) r |= 1 << (bits - 1 - b);
    return r;
}
__host__ __device__ constexpr int pb_ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

// In-place decimation-in-frequency radix-2 network; on return v[i] = X[bitrev(i)].
template <int R> __device__ __forceinline__ void pb_dft(float2 (&v)[R]) {
    constexpr int LR = pb_ilog2(R);
    PB_UNROLL for (int st = 0; st < LR; st++) {
        const int s = (R / 2) >> st;
        PB_UNROLL for (int i = 0; i < R / 2; i++) {
            const int blk = i / s, k = i % s;
            const int a = blk * 2 * s + k, b = a + s;
            const int m = k * (16 / s);           // w_{2s}^k = w_32^{16k/s}
            const float2 p = v[a], q = v[b];
            v[a] = make_float2(p.x + q.x, p.y + q.y);
            const float dxr = p.x - q.x, dyi = p.y - q.y;
            if (m == 0) v[b] = make_float2(dxr, dyi);
            else if (m == 8) v[b] = make_float2(dyi, -dxr);
            else {
                const float c = pb_c32(m), sn = pb_s32(m);
                v[b] = make_float2(dxr * c + dyi * sn, dyi * c - dxr * sn);   // (d) * (c - i sn)
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ FFT of N points by one group
template <int LOG2N> struct PbFftCfg {
    static constexpr int N = 1 << LOG2N;
    static constexpr int R = LOG2N >= 10 ? 32 : (LOG2N == 9 ? 16 : 8);
    static constexpr int LR = pb_ilog2(R);
    static constexpr int F = N / (R * R);                 // radix of the final pass (1 = none)
    static constexpr int G = (N / R) / 32;                // warps per group
    static constexpr int GT = 32 * G;                     // threads per group
    static constexpr int BUF = N + (N >> 3);              // float2 slots per group (worst-case skew padding)
    static constexpr int WARPS_PER_CTA = G >= 4 ? G : 4;
    static constexpr int GROUPS_PER_CTA = WARPS_PER_CTA / G;
    static constexpr int FB = F > 1 ? (N / F) / GT : 0;   // final-pass butterflies per thread
    static constexpr int RPL = (N / 3 + 2 + GT - 1) / GT + 1;   // lags per thread when extracting r
};

template <int G> __device__ __forceinline__ void pb_group_sync(int bar_id) {
    if (G == 1) __syncwarp(); else PB_GROUP_SYNC(bar_id, 32 * G);
}
__device__ __forceinline__ int pb_pad5(int i) { return i + (i >> 5); }

// v holds the pass-1 inputs z[g + t*N/R] of thread g; on return buf (pad5 layout) holds the spectrum in natural order.
template <int LOG2N>
__device__ __forceinline__ void pb_fft_group(float2 (&v)[PbFftCfg<LOG2N>::R], float2* buf, int g, int bar_id,
                                             const float2* __restrict__ tw_a, const float2* __restrict__ tw_b) {
    typedef PbFftCfg<LOG2N> C;
    constexpr int R = C::R, LR = C::LR, N = C::N;
    // ---- pass 1 (Ns = 1): no twiddles; out[g*R + t], skew-padded by (index >> LR)
    pb_dft<R>(v);
    PB_UNROLL for (int t = 0; t < R; t++) { int o = g * R + t; buf[o + (o >> LR)] = v[pb_bitrev(t, LR)]; }
    pb_group_sync<C::G>(bar_id);
    // ---- pass 2 (Ns = R): in[g + t*N/R] * w^(t*k), k = g mod R
    const int k = g & (R - 1);
    PB_UNROLL for (int t = 0; t < R; t++) { int i = g + t * (N / R); v[t] = buf[i + (i >> LR)]; }
    PB_UNROLL for (int t = 1; t < R; t++) {
        const float2 w = __ldg(&tw_a[t * R + k]);
        const float2 x = v[t];
        v[t] = make_float2(x.x * w.x - x.y * w.y, x.x * w.y + x.y * w.x);
    }
    pb_dft<R>(v);
    pb_group_sync<C::G>(bar_id);                       // every load of this pass is done before any store
    {
        const int base = (g >> LR) * (R * R) + k;
        PB_UNROLL for (int t = 0; t < R; t++) buf[pb_pad5(base + t * R)] = v[pb_bitrev(t, LR)];
    }
    pb_group_sync<C::G>(bar_id);
    // ---- final pass (radix F, Ns = R*R): butterflies are in place
    if (C::F > 1) {
        constexpr int F = C::F > 1 ? C::F : 2, LF = pb_ilog2(F);
        PB_UNROLL for (int b = 0; b < C::FB; b++) {
            const int j = g + b * C::GT;              // 0 .. N/F-1 = R*R-1
            float2 a[F];
            PB_UNROLL for (int t = 0; t < F; t++) a[t] = buf[pb_pad5(j + t * (R * R))];
            PB_UNROLL for (int t = 1; t < F; t++) {
                const float2 w = __ldg(&tw_b[t * (R * R) + j]);
                const float2 x = a[t];
                a[t] = make_float2(x.x * w.x - x.y * w.y, x.x * w.y + x.y * w.x);
            }
            pb_dft<F>(a);
            PB_UNROLL for (int t = 0; t < F; t++) buf[pb_pad5(j + t * (R * R))] = a[pb_bitrev(t, LF)];
        }
        pb_group_sync<C::G>(bar_id);
    }
}

// ------------------------------------------------------------------------------------------------ sinc interpolation
// Praat NUM_interpolate_sinc on y[1..2B+1] = r[-B..B] at lag x, evaluated by `nl` cooperating lanes
// (this lane is `sl`); returns this lane's partial sum (caller reduces).  With phi = frac(x), D the usable depth:
//   result = sin(pi phi)/(2 pi) * sum_m (-1)^m [ r[il-m] (1+cos(pi (phi+m)/(phi+D)))/(phi+m)
//                                             + r[il+1+m] (1+cos(pi (1-phi+m)/(1-phi+D)))/(1-phi+m) ]
__device__ __forceinline__ float pb_sinc_partial(const float* __restrict__ r, int B, float x, int depth, int sl, int nl) {
    const float fl = floorf(x);
    const float phi = x - fl;
    const int il = (int)fl;
    if (phi == 0.0f) return sl == 0 ? r[abs(il)] : 0.0f;
    int D = B - il; if (depth < D) D = depth;
    if (D <= 0) return 0.0f;
    const float phi1 = 1.0f - phi;
    const float kl = PB_PI_F / (phi + (float)D), kr = PB_PI_F / (phi1 + (float)D);
    float acc = 0.0f;
    for (int q = sl; q < 2 * D; q += nl) {
        const int m = q >> 1, side = q & 1;
        const float d = (side ? phi1 : phi) + (float)m;
        const int idx = side ? il + 1 + m : il - m;
        const float yv = r[abs(idx)];
        const float wgt = __fdividef(1.0f + __cosf(d * (side ? kr : kl)), d);
        const float term = yv * wgt;
        acc += (m & 1) ? -term : term;
    }
    return acc * (sinpif(phi) * (0.5f / PB_PI_F));
}
__device__ __forceinline__ float pb_reduce8(float v) {
    v += __shfl_xor_sync(PB_FULL_MASK, v, 1);
    v += __shfl_xor_sync(PB_FULL_MASK, v, 2);
    v += __shfl_xor_sync(PB_FULL_MASK, v, 4);
    return v;
}
// vertex of the parabola through (xa,fa),(xb,fb),(xc,fc), xa < xb < xc; xb if not concave
__device__ __forceinline__ float pb_parabola(float xa, float fa, float xb, float fb, float xc, float fc) {
    const float a = xb - xa, b = xb - xc;
    const float num = a * a * (fb - fc) - b * b * (fb - fa);
    const float den = a * (fb - fc) - b * (fb - fa);
    return den > 0.0f ? xb - 0.5f * num / den : xb;
}

// ------------------------------------------------------------------------------------------------ candidates of one frame
// One warp. r: normalised autocorrelation for lags 0..B (shared memory). scratch: 4*PB_MAXC floats of shared memory.
// Follows Sound_into_PitchFrame (Praat fon/Sound_to_Pitch.cpp): maxima of r above voicingThreshold/2 between lag 2
// and scan_lim-1 become candidates (weakest replaced when more than max_cand-1), each then refined on the
// sinc-interpolated curve (depth 70, or 700 above 0.3/dx).  Praat refines with Brent (tol 1e-10, <= 60 its);
// here the maximiser is found with a fixed number of evaluations (half-sample grid, two parabolic steps), which
// reproduces Brent's optimum far inside the F0 / strength tolerances.
__device__ __forceinline__ void pb_frame_candidates(const float* __restrict__ r, float* scratch, const PbPitchGeomDev& gm,
                                                    int lane, float* __restrict__ out_f, float* __restrict__ out_s,
                                                    uint8_t* __restrict__ out_n) {
    float* cf = scratch;                       // first-pass frequency
    float* cs = scratch + PB_MAXC;             // first-pass strength
    int* imax = (int*)(scratch + 2 * PB_MAXC); // lag of the maximum
    const int B = gm.brent_ixmax, lim = gm.scan_lim, maxc = gm.max_cand;
    const int sub = lane >> 3, sl = lane & 7;
    // ---- count the maxima
    int total = 0;
    for (int base = 2; base < lim; base += 32) {
        const int i = base + lane;
        const bool pk = i < lim && r[i] > gm.half_voicing && r[i] > r[i - 1] && r[i] >= r[i + 1];
        total += __popc(__ballot_sync(PB_FULL_MASK, pk));
    }
    int ncf = 1;
    const bool overflow = total > maxc - 1;
    for (int base = 2; base < lim; base += 32) {
        const int i = base + lane;
        const bool pk = i < lim && r[i] > gm.half_voicing && r[i] > r[i - 1] && r[i] >= r[i + 1];
        unsigned mask = __ballot_sync(PB_FULL_MASK, pk);
        if (!overflow) {
            // common case: every maximum gets its own slot, in lag order
            if (pk) imax[ncf + __popc(mask & ((1u << lane) - 1u))] = i;
            ncf += __popc(mask);
        } else {
            // rare (tonal high-frequency content): Praat's sequential insert / replace-the-weakest
            while (mask) {
                unsigned m = mask;
                for (int q = 0; q < sub; q++) m &= m - 1;
                const bool have = m != 0;
                const int ip = have ? base + __ffs((int)m) - 1 : 2;
                const float dr = 0.5f * (r[ip + 1] - r[ip - 1]), d2r = 2.0f * r[ip] - r[ip - 1] - r[ip + 1];
                const float x0 = (float)ip + (have ? dr / d2r : 0.0f);
                float st = pb_reduce8(pb_sinc_partial(r, B, x0, 30, sl, 8));
                if (st > 1.0f) st = 1.0f / st;
                const float fq0 = gm.sr / x0;
                for (int q = 0; q < 4; q++) {
                    const int hv = __shfl_sync(PB_FULL_MASK, (int)have, q * 8);
                    if (!hv) break;
                    const float fq = __shfl_sync(PB_FULL_MASK, fq0, q * 8), sq = __shfl_sync(PB_FULL_MASK, st, q * 8);
                    const int iq = __shfl_sync(PB_FULL_MASK, ip, q * 8);
                    int place = 0;
                    if (ncf < maxc) place = ncf++;
                    else {
                        // weakest of slots 1..maxc-1 by strength - octaveCost*log2(minPitch/f); first minimum wins
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
    }
    __syncwarp();
    // ---- refine every candidate on the sinc curve: 4 candidates at a time, 8 lanes each
    if (lane == 0) { out_f[0] = 0.0f; out_s[0] = 0.0f; }
    for (int c0 = 1; c0 < ncf; c0 += 4) {
        const int c = c0 + sub;
        const bool have = c < ncf;
        const int i = have ? imax[c] : 2;
        const float fi = (float)i;
        const float r0 = r[i], rm = r[i - 1], rp = r[i + 1];
        const float x_first = fi + 0.5f * (rp - rm) / (2.0f * r0 - rm - rp);      // Praat's parabolic first guess
        const int depth = (gm.sr / x_first > 0.3f * gm.sr) ? 700 : 70;
        // half-sample grid: r[i-1], f(i-.5), r[i], f(i+.5), r[i+1]
        const float fa = pb_reduce8(pb_sinc_partial(r, B, fi - 0.5f, depth, sl, 8));
        const float fb = pb_reduce8(pb_sinc_partial(r, B, fi + 0.5f, depth, sl, 8));
        float xb = fi, yb = r0, xl = fi - 0.5f, yl = fa, xr = fi + 0.5f, yr = fb;     // best point and its neighbours
        if (fa > yb && fa >= fb) { xb = fi - 0.5f; yb = fa; xl = fi - 1.0f; yl = rm; xr = fi; yr = r0; }
        else if (fb > yb) { xb = fi + 0.5f; yb = fb; xl = fi; yl = r0; xr = fi + 1.0f; yr = rp; }
        float x1 = pb_parabola(xl, yl, xb, yb, xr, yr);
        x1 = fminf(fmaxf(x1, xb - 0.5f), xb + 0.5f);
        const float h2 = 0.125f;
        const float xa2 = fmaxf(x1 - h2, fi - 1.0f), xc2 = fminf(x1 + h2, fi + 1.0f);
        const float ya2 = pb_reduce8(pb_sinc_partial(r, B, xa2, depth, sl, 8));
        const float y1 = pb_reduce8(pb_sinc_partial(r, B, x1, depth, sl, 8));
        const float yc2 = pb_reduce8(pb_sinc_partial(r, B, xc2, depth, sl, 8));
        float x2 = x1;
        if (xa2 < x1 && x1 < xc2) x2 = pb_parabola(xa2, ya2, x1, y1, xc2, yc2);
        x2 = fminf(fmaxf(x2, xa2 - h2), xc2 + h2);
        x2 = fminf(fmaxf(x2, fi - 1.0f), fi + 1.0f);
        const float y2 = pb_reduce8(pb_sinc_partial(r, B, x2, depth, sl, 8));
        float bx = x2, by = y2;
        if (y1 > by) { bx = x1; by = y1; }
        if (ya2 > by) { bx = xa2; by = ya2; }
        if (yc2 > by) { bx = xc2; by = yc2; }
        if (yb > by) { bx = xb; by = yb; }
        if (by > 1.0f) by = 1.0f / by;
        if (have && sl == 0) { out_f[c] = gm.sr / bx; out_s[c] = by; }
    }
    if (lane == 0) *out_n = (uint8_t)ncf;
}

// ------------------------------------------------------------------------------------------------ K1 + K2
template <int LOG2N>
__global__ void __launch_bounds__(PbFftCfg<LOG2N>::WARPS_PER_CTA * 32)
pb_pitch_frames_kernel(const int16_t* __restrict__ pcm, const PbUnitDev* __restrict__ units, const int32_t* __restrict__ pair_off,
                       PbPitchGeomDev gm, float* __restrict__ cand_f, float* __restrict__ cand_s,
                       uint8_t* __restrict__ ncand, float* __restrict__ intensity) {
    typedef PbFftCfg<LOG2N> C;
    constexpr int R = C::R, N = C::N, G = C::G, GT = C::GT;
    PB_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = warp / G, wg = warp % G;          // group in CTA, warp in group
    const int g = wg * 32 + lane;                       // thread in group = butterfly index
    const int bar_id = 1 + group;
    // per-group shared memory: FFT buffer, then a small reduction scratch
    float2* buf = (float2*)smem_raw + (size_t)group * (C::BUF + 8 * G);
    float* red = (float*)(buf + C::BUF);                // [G][4] floats
    const int B = gm.brent_ixmax;
    const int rstride = (B + 4) & ~1;

    for (int item = blockIdx.x * C::GROUPS_PER_CTA + group; item < gm.n_pairs; item += gridDim.x * C::GROUPS_PER_CTA) {
        const int u = pb_upper_unit(pair_off, gm.n_units, item);
        const PbUnitDev ud = units[u];
        const int fA = 2 * (item - ud.pair_off), fB = fA + 1;
        const bool hasB = fB < ud.n_frames;
        const float mean_scale = 1.0f / 32768.0f;
        // ---- frame positions (float64, Praat's Sampled_indexToX / Sampled_xToLowIndex)
        long long start[2]; float lmean[2];
        PB_UNROLL for (int f = 0; f < 2; f++) {
            const int fi = f ? fB : fA;
            const double t = __dadd_rn(ud.t1, __dmul_rn((double)fi, gm.dt));
            const long long left = (long long)floor(__ddiv_rn(__dsub_rn(t, ud.x1), gm.dx)) + 1, right = left + 1;
            start[f] = right - gm.half_nw;              // part index (1-based) of frame sample n = 0
            // local mean over [right - nsamp_period, left + nsamp_period]
            int s = 0;
            const long long m0 = right - gm.nsamp_period;
            for (int q = lane; q < 2 * gm.nsamp_period; q += 32) s += pb_sample(pcm, ud, m0 + q);
            s = pb_warp_sum_i(s);
            lmean[f] = (float)(((double)s / 32768.0) / (double)(2 * gm.nsamp_period));
        }
        // ---- load + window both frames straight into the pass-1 registers
        float2 v[R];
        float mxA = 0.0f, mxB = 0.0f, pkA = 0.0f, pkB = 0.0f;
        const int pk_lo = max(0, gm.half_nw - gm.half_period), pk_hi = min(gm.nw, gm.half_nw + gm.half_period);  // [lo, hi)
        PB_UNROLL for (int t = 0; t < R; t++) {
            const int n = g + t * (N / R);
            float a = 0.0f, b = 0.0f;
            if (n < gm.nw) {
                const float w = __ldg(&gm.window[n]);
                a = ((float)pb_sample(pcm, ud, start[0] + n) * mean_scale - lmean[0]) * w;
                if (hasB) b = ((float)pb_sample(pcm, ud, start[1] + n) * mean_scale - lmean[1]) * w;
                const float aa = fabsf(a), ab = fabsf(b);
                mxA = fmaxf(mxA, aa); mxB = fmaxf(mxB, ab);
                if (n >= pk_lo && n < pk_hi) { pkA = fmaxf(pkA, aa); pkB = fmaxf(pkB, ab); }
            }
            v[t] = make_float2(a, b);
        }
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
        // bring both frames to comparable magnitude (power-of-two scales are exact and cancel in r = ac/ac[0]);
        // keeps the weaker frame of a pair out of the stronger one's rounding noise
        {
            int eA = 0, eB = 0;
            if (mxA > 0.0f) frexpf(mxA, &eA);
            if (mxB > 0.0f) frexpf(mxB, &eB);
            const float sA = ldexpf(1.0f, -eA), sB = ldexpf(1.0f, -eB);
            PB_UNROLL for (int t = 0; t < R; t++) { v[t].x *= sA; v[t].y *= sB; }
        }
        const bool global_silent = ud.global_peak == 0.0;
        if (!global_silent && (pkA > 0.0f || pkB > 0.0f)) {
            // ---- FFT 1: z = a + i b
            pb_fft_group<LOG2N>(v, buf, g, bar_id, gm.tw_a, gm.tw_b);
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
            // ---- FFT 2 (the spectra are real and even, so a forward transform returns both autocorrelations)
            PB_UNROLL for (int t = 0; t < R; t++) v[t] = buf[pb_pad5(g + t * (N / R))];
            pb_group_sync<G>(bar_id);
            pb_fft_group<LOG2N>(v, buf, g, bar_id, gm.tw_a, gm.tw_b);
            // ---- r[lag] = ac[lag] / (ac[0] * windowR[lag]) for lags 0..B+1, into shared memory
            float ra[C::RPL], rb[C::RPL];
            const float2 ac0 = buf[0];
            const float iA = ac0.x > 0.0f ? 1.0f / ac0.x : 0.0f, iB = ac0.y > 0.0f ? 1.0f / ac0.y : 0.0f;
            PB_UNROLL for (int q = 0; q < C::RPL; q++) {
                const int lag = g + q * GT;
                ra[q] = 0.0f; rb[q] = 0.0f;
                if (lag <= B + 1 && lag < N) {
                    const float2 a = buf[pb_pad5(lag)];
                    const float iw = lag <= B ? __ldg(&gm.inv_wr[lag]) : 0.0f;
                    ra[q] = a.x * iA * iw; rb[q] = a.y * iB * iw;
                }
            }
            pb_group_sync<G>(bar_id);
            float* rA = (float*)buf; float* rB = rA + rstride;
            PB_UNROLL for (int q = 0; q < C::RPL; q++) {
                const int lag = g + q * GT;
                if (lag <= B + 1) { rA[lag] = lag == 0 ? 1.0f : ra[q]; rB[lag] = lag == 0 ? 1.0f : rb[q]; }
            }
            pb_group_sync<G>(bar_id);
        }
        // ---- candidates: warp 0 of the group takes frame A, warp 1 (or warp 0 again) frame B
        {
            float* rA = (float*)buf; float* rB = rA + rstride;
            float* scr = rB + rstride;
            const float gpk = (float)ud.global_peak;
            PB_UNROLL for (int f = 0; f < 2; f++) {
                const int owner = (G > 1) ? f : 0;
                if (wg != owner) continue;
                if (f == 1 && !hasB) continue;
                const int64_t fr = ud.frame_off + (f ? fB : fA);
                const float pk = f ? pkB : pkA;
                float* of = cand_f + fr * gm.max_cand; float* os = cand_s + fr * gm.max_cand;
                if (global_silent || pk == 0.0f) {
                    if (lane == 0) { of[0] = 0.0f; os[0] = 0.0f; ncand[fr] = 1; intensity[fr] = global_silent ? 0.0f : 0.0f; }
                } else {
                    if (lane == 0) { const float it = pk / gpk; intensity[fr] = it > 1.0f ? 1.0f : it; }
                    pb_frame_candidates(f ? rB : rA, scr + f * (4 * PB_MAXC), gm, lane, of, os, ncand + fr);
                }
            }
        }
        pb_group_sync<G>(bar_id);     // buf is reused by the next item
    }
}

// ------------------------------------------------------------------------------------------------ K3: path finder + median
// One warp per unit; lane c2 owns candidate c2 of the current frame.  Praat Pitch_pathFinder (fon/Pitch.cpp):
// Viterbi over the candidate lattice in float64, strict '>' so the lowest-index predecessor wins ties.
// Back-pointers go to global memory (one byte per candidate); the backtrack then writes selected_array
// (frequency, strength of the chosen candidate per frame).  Finally np.median of the frequencies > 0
// (mean of the two middle values) by bisection on the float bit patterns.
__device__ __forceinline__ double pb_shfl_d(double v, int src) { return __shfl_sync(PB_FULL_MASK, v, src); }

__global__ void __launch_bounds__(128)
pb_pitch_path_kernel(const PbUnitDev* __restrict__ units, PbPitchGeomDev gm, const float* __restrict__ cand_f,
                     const float* __restrict__ cand_s, const uint8_t* __restrict__ ncand, const float* __restrict__ intensity,
                     uint8_t* __restrict__ psi, float* __restrict__ sel_f, float* __restrict__ sel_s,
                     double* __restrict__ median_out, int32_t* __restrict__ nvoiced_out) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const int maxc = gm.max_cand;
    const double tcorr = 0.01 / gm.dt;
    const double ojc = gm.octave_jump_cost * tcorr, vuc = gm.voiced_unvoiced_cost * tcorr;
    for (int u = blockIdx.x * wpb + (threadIdx.x >> 5); u < gm.n_units; u += gridDim.x * wpb) {
        const PbUnitDev ud = units[u];
        const int nF = ud.n_frames;
        const int64_t f0 = ud.frame_off;
        if (ud.global_peak == 0.0) {
            // Praat returns before the path finder: every frame voiceless
            for (int f = lane; f < nF; f += 32) { sel_f[f0 + f] = 0.0f; sel_s[f0 + f] = 0.0f; }
            if (lane == 0) { median_out[ud.out_index] = 0.0; nvoiced_out[ud.out_index] = 0; }
            continue;
        }
        double delta_prev = 0.0, l2_prev = 0.0;     // of candidate `lane` in the previous frame
        int voiced_prev = 0, nc_prev = 0;
        // prefetch frame 0
        int nc_n = ncand[f0];
        float cf_n = lane < nc_n ? cand_f[f0 * maxc + lane] : 0.0f, cs_n = lane < nc_n ? cand_s[f0 * maxc + lane] : 0.0f;
        float in_n = intensity[f0];
        for (int f = 0; f < nF; f++) {
            const int nc = nc_n; const float cf = cf_n, cs = cs_n, inten = in_n;
            if (f + 1 < nF) {
                const int64_t fr = f0 + f + 1;
                nc_n = ncand[fr];
                cf_n = lane < nc_n ? cand_f[fr * maxc + lane] : 0.0f; cs_n = lane < nc_n ? cand_s[fr * maxc + lane] : 0.0f;
                in_n = intensity[fr];
            }
            const double fr_d = (double)cf;
            const int voiced = fr_d > 0.0 && fr_d < gm.ceiling;
            double us = gm.silence_threshold <= 0.0 ? 0.0 : 2.0 - (double)inten / (gm.silence_threshold / (1.0 + gm.voicing_threshold));
            us = gm.voicing_threshold + (us > 0.0 ? us : 0.0);
            const double l2 = voiced ? log2(fr_d) : 0.0;
            const double local = voiced ? (double)cs - gm.octave_cost_d * (log2(gm.ceiling) - l2) : us;
            double best = local; int place = 0;
            if (f > 0) {
                best = -1.0e30; place = -1;
                for (int c1 = 0; c1 < nc_prev; c1++) {
                    const double dp = pb_shfl_d(delta_prev, c1), lp = pb_shfl_d(l2_prev, c1);
                    const int vp = __shfl_sync(PB_FULL_MASK, voiced_prev, c1);
                    double cost;
                    if (!voiced) cost = vp ? vuc : 0.0;
                    else cost = vp ? ojc * fabs(lp - l2) : vuc;
                    const double value = __dadd_rn(__dsub_rn(dp, cost), local);
                    if (value > best) { best = value; place = c1; }
                }
                if (lane < nc) psi[(f0 + f) * maxc + lane] = (uint8_t)place;
            }
            delta_prev = best; l2_prev = l2; voiced_prev = voiced; nc_prev = nc;
        }
        // terminal candidate: first maximum
        double bv = lane < nc_prev ? delta_prev : -1.0e300; int bi = lane;
        PB_UNROLL for (int o = 16; o > 0; o >>= 1) {
            const double ov = pb_shfl_d(bv, lane ^ o); const int oi = __shfl_xor_sync(PB_FULL_MASK, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        __syncwarp();
        // backtrack (lane 0), writing selected_array
        if (lane == 0) {
            int place = bi;
            for (int f = nF - 1; f >= 0; f--) {
                const int64_t fr = f0 + f;
                sel_f[fr] = cand_f[fr * maxc + place]; sel_s[fr] = cand_s[fr * maxc + place];
                if (f > 0) place = psi[fr * maxc + place];
            }
        }
        __syncwarp();
        // ---- np.median(freqs[freqs > 0]) : positive floats order like their bit patterns
        int nv = 0;
        for (int f = lane; f < nF; f += 32) nv += sel_f[f0 + f] > 0.0f;
        nv = pb_warp_sum_i(nv);
        double med = 0.0;
        if (nv > 0) {
            const int k = (nv - 1) >> 1;                       // 0-based rank of the lower middle
            unsigned lo = 0u, hi = 0x7f800000u;                // smallest pattern with count(<= pattern) >= k+1
            while (lo < hi) {
                const unsigned mid = lo + ((hi - lo) >> 1);
                int c = 0;
                for (int f = lane; f < nF; f += 32) { const float v = sel_f[f0 + f]; c += (v > 0.0f && __float_as_uint(v) <= mid); }
                c = pb_warp_sum_i(c);
                if (c >= k + 1) hi = mid; else lo = mid + 1;
            }
            const float lower = __uint_as_float(lo);
            float upper = lower;
            if ((nv & 1) == 0) {
                // the next order statistic: lower again if enough duplicates, else the smallest value above it
                int c = 0; float nxt = 3.0e38f;
                for (int f = lane; f < nF; f += 32) {
                    const float v = sel_f[f0 + f];
                    if (v > 0.0f) { if (v <= lower) c++; else nxt = fminf(nxt, v); }
                }
                c = pb_warp_sum_i(c);
                PB_UNROLL for (int o = 16; o > 0; o >>= 1) nxt = fminf(nxt, __shfl_xor_sync(PB_FULL_MASK, nxt, o));
                upper = (c >= k + 2) ? lower : nxt;
            }
            med = ((double)lower + (double)upper) / 2.0;
        }
        if (lane == 0) { median_out[ud.out_index] = med; nvoiced_out[ud.out_index] = nv; }
    }
}
