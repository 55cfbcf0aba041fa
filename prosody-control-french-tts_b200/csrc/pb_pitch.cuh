// pb_pitch.cuh — shared device code of the F0 path: descriptors, K0 (unit stats), the register-blocked DFT.
//
// The F0 path computes what parselmouth's Sound.to_pitch(pitch_floor, pitch_ceiling) (Praat Sound_to_Pitch_ac,
// AC_HANNING) computes for the reference's get_median_pitch (/root/reference/Code/audioPipeline.py:326-335), one
// independent analysis per unit (file, t0, t1).  B200-first layout:
//   * K0  pb_unit_stats_kernel     per-unit mean / global peak (exact integer reductions)          [HBM-bound]
//   * K1+K2 pb_pitch_frames_kernel (pb_pitch_frames.cuh): a GROUP of G warps (G = 1 for FFT sizes <= 1024) owns a
//     PAIR of consecutive frames end to end: PCM load, local mean, Hanning window, one complex FFT carrying both
//     real frames, power spectra, second FFT back to lags, normalisation, peak picking and sinc refinement.
//     The autocorrelation never leaves shared memory.                                               [FP32-pipe-bound]
//   * K3  pb_pitch_path_kernel (pb_pitch_path.cuh): Viterbi over the candidate lattice + median of the voiced.
//   FFTs are register-blocked: every lane runs one radix-R butterfly (R = 32 for N >= 1024) per pass on 2R
//   registers; passes exchange data through a skew-padded shared buffer (bank-conflict free), in place.
//   No tensor cores (nothing here is a dense contraction).
#pragma once
#include "pb_rt.h"
#include "pb_stream.cuh"
#include <math.h>

#ifndef PB_WPC
#define PB_WPC 4            // warps per CTA of the frames kernel (when one group needs fewer)
#endif
#ifndef PB_BIG_WARPS
#define PB_BIG_WARPS 16      // resident warps per SM the multi-warp-group FFT sizes (N >= 2048) are compiled for: 128 registers, no spills;
                             // measured at 8 / 12 / 16 warps: 5.07 / 4.22 / 3.94 ms (N = 2048), 10.46 / 8.54 / 7.98 ms (N = 4096) for 599 k frames
#endif
#define PB_MAXC 32           // hard cap on candidates per frame (Praat default 15)
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
    int min_refine_lag;   // maxima at smaller lags stay above the ceiling whatever the refinement: never voiced
    int pre_cap;          // samples per staging buffer of the cp.async prefetch (multiple of 8)
    int phase_sync;          // 1: the CTA's warps start every work item together (instruction-cache locality)
    int path_long;        // K3: units with more frames than this go through the blocked path finder (pb_path_block_*_kernel)
    int path_block;       // K3: frames per block of the blocked path finder (<= PB_PATHL_BLOCK_MAX)
    long long pcm_len;    // samples in the pcm buffer (prefetch copies stay inside it)
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
    const float* half_tab;// [70]  windowed-sinc coefficients at phi = 1/2, depth 70 (incl. the 1/(2 pi) factor)
};

// ------------------------------------------------------------------------------------------------ helpers
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
// One warp per unit, 16-byte loads (pb_stream.cuh): the kernel is a pure HBM stream of 2 B/sample.
__global__ void __launch_bounds__(256) pb_unit_stats_kernel(const int16_t* __restrict__ pcm, PbUnitDev* __restrict__ units, int n_units, long long long_nx) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int u = blockIdx.x * wpb + (threadIdx.x >> 5); u < n_units; u += gridDim.x * wpb) {
        const PbUnitDev ud = units[u];
        if (ud.nx > long_nx) continue;                 // long units: pb_unit_stats_long_kernel
        // the part of [ix1-1, ix1-1+nx) that lies inside the file
        const long long a = ud.ix1 - 1, b = a + ud.nx;
        const long long lo = a < 0 ? 0 : a, hi = b > ud.file_nx ? ud.file_nx : b;
        const bool padded = (lo > a) || (hi < b) || (hi <= lo);
        long long sum = 0; int mn = 32767, mx = -32768;
        pb_warp_foreach_s16(pcm + ud.pcm_off, lo, hi, lane, [&](int v) { sum += v; mn = min(mn, v); mx = max(mx, v); });
        PB_UNROLL for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(PB_FULL_MASK, sum, o);
            mn = min(mn, __shfl_xor_sync(PB_FULL_MASK, mn, o));
            mx = max(mx, __shfl_xor_sync(PB_FULL_MASK, mx, o));
        }
        if (lane == 0) {
            if (padded) { mn = min(mn, 0); mx = max(mx, 0); }
            if (hi <= lo) { mn = 0; mx = 0; }
            const double mean = ((double)sum / 32768.0) / (double)ud.nx;
            const double p1 = fabs((double)mx / 32768.0 - mean), p2 = fabs((double)mn / 32768.0 - mean);
            units[u].mean = mean;
            units[u].global_peak = p1 > p2 ? p1 : p2;
        }
    }
}

// K0 for long units (a one-hour recording analysed as one sound is 79 M samples: one warp would stream it for tens of
// milliseconds): the units above long_nx samples are listed, every warp of the grid takes 64 Ki-sample pieces of each in turn and
// merges its exact integer sum / minimum / maximum with atomics, and a last small kernel turns them into mean and global peak.
struct PbStatsLong { int unit; int mn; int mx; int pad; unsigned long long sum; };
#define PB_STATS_PIECE 65536
__global__ void pb_unit_stats_long_index_kernel(const PbUnitDev* __restrict__ units, int n_units, long long long_nx, PbStatsLong* __restrict__ longs, int* __restrict__ counter) {
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += gridDim.x * blockDim.x) {
        if (units[u].nx <= long_nx) continue;
        PbStatsLong e; e.unit = u; e.mn = 32767; e.mx = -32768; e.pad = 0; e.sum = 0ull;
        longs[atomicAdd(counter, 1)] = e;
    }
}
__global__ void __launch_bounds__(256) pb_unit_stats_long_kernel(const int16_t* __restrict__ pcm, const PbUnitDev* __restrict__ units, PbStatsLong* __restrict__ longs, const int* __restrict__ counter) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int li = 0; li < *counter; li++) {
        const PbUnitDev ud = units[longs[li].unit];
        const long long a = ud.ix1 - 1, b = a + ud.nx;
        const long long lo = a < 0 ? 0 : a, hi = b > ud.file_nx ? ud.file_nx : b;
        const long long pieces = hi > lo ? (hi - lo + PB_STATS_PIECE - 1) / PB_STATS_PIECE : 0;
        long long sum = 0; int mn = 32767, mx = -32768;
        for (long long pc = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); pc < pieces; pc += (long long)gridDim.x * wpb) {
            const long long p0 = lo + pc * PB_STATS_PIECE, p1 = p0 + PB_STATS_PIECE < hi ? p0 + PB_STATS_PIECE : hi;
            pb_warp_foreach_s16(pcm + ud.pcm_off, p0, p1, lane, [&](int v) { sum += v; mn = min(mn, v); mx = max(mx, v); });
        }
        PB_UNROLL for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(PB_FULL_MASK, sum, o);
            mn = min(mn, __shfl_xor_sync(PB_FULL_MASK, mn, o));
            mx = max(mx, __shfl_xor_sync(PB_FULL_MASK, mx, o));
        }
        if (lane == 0 && mn <= mx) {
            atomicAdd(&longs[li].sum, (unsigned long long)sum);          // two's complement: wraps to the exact signed total
            atomicMin(&longs[li].mn, mn); atomicMax(&longs[li].mx, mx);
        }
    }
}
__global__ void pb_unit_stats_long_fin_kernel(PbUnitDev* __restrict__ units, const PbStatsLong* __restrict__ longs, const int* __restrict__ counter) {
    for (int li = blockIdx.x * blockDim.x + threadIdx.x; li < *counter; li += gridDim.x * blockDim.x) {
        const PbStatsLong e = longs[li];
        const PbUnitDev ud = units[e.unit];
        const long long a = ud.ix1 - 1, b = a + ud.nx;
        const long long lo = a < 0 ? 0 : a, hi = b > ud.file_nx ? ud.file_nx : b;
        const bool padded = (lo > a) || (hi < b) || (hi <= lo);
        int mn = e.mn, mx = e.mx;
        if (padded) { mn = min(mn, 0); mx = max(mx, 0); }
        if (hi <= lo) { mn = 0; mx = 0; }
        const double mean = ((double)(long long)e.sum / 32768.0) / (double)ud.nx;
        const double p1 = fabs((double)mx / 32768.0 - mean), p2 = fabs((double)mn / 32768.0 - mean);
        units[e.unit].mean = mean;
        units[e.unit].global_peak = p1 > p2 ? p1 : p2;
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
            // sm_100 packed-pair arithmetic (FADD2 / FMUL2 / FFMA2): one instruction per complex add / scale, which
            // cuts the network from 442 to 270 instructions — the kernel is issue- and fetch-bound, not FP32-pipe-bound
            const float2 p = v[a], q = v[b];
            v[a] = __fadd2_rn(p, q);
            const float2 d = __fadd2_rn(p, make_float2(-q.x, -q.y));
            if (m == 0) v[b] = d;
            else if (m == 8) v[b] = make_float2(d.y, -d.x);
            else {
                const float c = pb_c32(m), sn = pb_s32(m);
                v[b] = __ffma2_rn(make_float2(d.y, d.x), make_float2(sn, -sn), __fmul2_rn(d, make_float2(c, c)));   // d * (c - i sn)
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ FFT geometry
// N = R * R * F points handled by G warps: passes of radix R, R and (if F > 1) F.
template <int LOG2N> struct PbFftCfg {
    static constexpr int N = 1 << LOG2N;
    static constexpr int R = LOG2N >= 10 ? 32 : (LOG2N == 9 ? 16 : 8);
    static constexpr int LR = pb_ilog2(R);
    static constexpr int F = N / (R * R);                 // radix of the final pass (1 = none)
    static constexpr int G = (N / R) / 32;                // warps per group
    static constexpr int GT = 32 * G;                     // threads per group
    static constexpr int BUF = N + (N >> LR) + 8;         // float2 slots per group: N plus the skew padding (index >> LR)
    static constexpr int WARPS_PER_CTA = G >= PB_WPC ? G : PB_WPC;
    static constexpr int GROUPS_PER_CTA = WARPS_PER_CTA / G;
    static constexpr int FB = F > 1 ? (N / F) / GT : 0;   // final-pass butterflies per thread
    static constexpr int RPL = (N / 3 + 2 + GT - 1) / GT + 1;   // lags per thread when extracting r
    static constexpr int MIN_CTAS = LOG2N <= 10 ? 16 / WARPS_PER_CTA : (PB_BIG_WARPS / WARPS_PER_CTA > 0 ? PB_BIG_WARPS / WARPS_PER_CTA : 1);   // occupancy target: 128 registers, no spills (a fifth CTA at 96 registers measured no faster)
};

template <int G> __device__ __forceinline__ void pb_group_sync(int bar_id) {
    // A named barrier even for a single warp: ptxas clones code across __syncwarp() (it kept a second copy of the
    // unrolled butterfly network for the first loop iteration) but not across bar.sync.
    PB_GROUP_SYNC(bar_id, 32 * G);
}
__device__ __forceinline__ int pb_pad5(int i) { return i + (i >> 5); }
