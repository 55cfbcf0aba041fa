// pb_plan.h — host-side planning in float64: which samples and frames each reference call touches.
//
// These are index computations, not signal processing: they decide *where* the kernels read, and must agree with
// the reference's third-party libraries to the sample (SURVEY.md §7 "index parity lives on the host in float64").
//   Praat  fon/Sound.cpp Sound_extractPart, fon/Sound_to_Pitch.cpp Sound_to_Pitch_any, fon/Sampled.cpp
//          Sampled_shortTermAnalysis          <- parselmouth calls at Code/audioPipeline.py:327-333
//   pydub  audio_segment.py __getitem__/_parse_position/duration_seconds
//                                              <- Code/audioPipeline.py:319-323, 340-348, 361
//   pyloudnorm util.valid_audio (>= 0.4 s) and the reference's fallbacks at Code/audioPipeline.py:345-358
// Compiled with FP contraction off so every product/sum rounds like CPython's float arithmetic.
#pragma once
#include <cmath>
#include <cstdint>
#include "../../include/prosody_b200.h"

struct PbGeomHost {
    int64_t nsamp_period, half_period, nw, half_nw, min_lag, max_lag, brent_ixmax, nfft_praat;
    int32_t max_cand, log2n;      // log2n: FFT size used by the kernels (any N >= nw + brent_ixmax is exact)
    double dx, dt, dt_window, ceiling;
};

struct PbUnitPlan {
    int32_t status, n_frames;
    int64_t ix1, nx;
    double x1, t1;
};

static inline int64_t pb_ifloor(double x) { return (int64_t)std::floor(x); }

// The part of Sound_to_Pitch_any that depends only on the sampling period and the parameters.
static inline int pb_geom_for_rate(double sr, const PbPitchParams& p, PbGeomHost& g) {
    const double dx = 1.0 / sr;
    double ceiling = p.pitch_ceiling, dt = p.time_step;
    int64_t maxc = p.max_candidates;
    if ((double)maxc < ceiling / p.pitch_floor) maxc = pb_ifloor(ceiling / p.pitch_floor);
    if (dt <= 0.0) dt = p.periods_per_window / p.pitch_floor / 4.0;
    g.dx = dx; g.dt = dt;
    g.nsamp_period = pb_ifloor(1.0 / dx / p.pitch_floor);
    g.half_period = g.nsamp_period / 2 + 1;
    if (ceiling > 0.5 / dx) ceiling = 0.5 / dx;
    g.ceiling = ceiling;
    g.dt_window = p.periods_per_window / p.pitch_floor;
    g.nw = pb_ifloor(g.dt_window / dx);
    g.half_nw = g.nw / 2 - 1;
    if (g.half_nw < 2) return PB_UNIT_WINDOW;
    g.nw = g.half_nw * 2;
    g.min_lag = pb_ifloor(1.0 / dx / ceiling);
    if (g.min_lag < 2) g.min_lag = 2;
    g.max_lag = pb_ifloor((double)g.nw / p.periods_per_window) + 2;
    if (g.max_lag > g.nw) g.max_lag = g.nw;
    g.nfft_praat = 1;
    while ((double)g.nfft_praat < (double)g.nw * 1.5) g.nfft_praat *= 2;   // AC_HANNING: interpolation depth 0.5
    g.brent_ixmax = pb_ifloor((double)g.nw * 0.5);
    g.max_cand = (int32_t)maxc;
    int l2 = 8;
    while (((int64_t)1 << l2) < g.nw + g.brent_ixmax + 1) l2++;
    g.log2n = l2;
    return PB_UNIT_OK;
}

// Sound_extractPart (rectangular, relative width 1, preserve_times) on a file-loaded Sound (x1 = dx/2, xmin = 0).
static inline int pb_extract_part(int64_t file_nx, double sr, int has_t1, double t0, double t1, int64_t* ix1, int64_t* nx, double* x1_part) {
    const double dx = 1.0 / sr, x1 = 0.5 / sr;
    if (!has_t1) { *ix1 = 1; *nx = file_nx; *x1_part = x1; return PB_UNIT_OK; }
    if (t0 == t1) { t0 = 0.0; t1 = (double)file_nx * dx; }
    const int64_t i1 = 1 + (int64_t)std::ceil((t0 - x1) / dx);
    const int64_t i2 = 1 + (int64_t)std::floor((t1 - x1) / dx);
    if (i2 < i1) return PB_UNIT_NO_SAMPLES;
    *ix1 = i1; *nx = i2 - i1 + 1; *x1_part = x1 + (double)(i1 - 1) * dx;
    if (has_t1 == 2) *x1_part -= t0;      // preserve_times = false: Praat shifts the part's time axis to start at 0
    return PB_UNIT_OK;
}

// The per-sound part of Sound_to_Pitch_any + Sampled_shortTermAnalysis.
static inline void pb_plan_pitch_unit(int64_t file_nx, double sr, int has_t1, double t0, double t1,
                                      const PbPitchParams& p, const PbGeomHost& g, int geom_status, PbUnitPlan& u) {
    u.n_frames = 0; u.t1 = 0.0;
    u.status = pb_extract_part(file_nx, sr, has_t1, t0, t1, &u.ix1, &u.nx, &u.x1);
    if (u.status != PB_UNIT_OK) return;
    volatile double duration = g.dx * (double)u.nx;
    if (p.pitch_floor < p.periods_per_window / duration) { u.status = PB_UNIT_TOO_SHORT; return; }
    if (geom_status != PB_UNIT_OK) { u.status = geom_status; return; }
    if (g.dt_window > duration) { u.status = PB_UNIT_TOO_SHORT; return; }
    u.n_frames = (int32_t)(pb_ifloor((duration - g.dt_window) / g.dt) + 1);
    const double mid = u.x1 - 0.5 * g.dx + 0.5 * duration;
    const double span = (double)u.n_frames * g.dt;
    u.t1 = mid - 0.5 * span + 0.5 * g.dt;
}

// ---- pydub
static inline int64_t pb_pydub_len_ms(int64_t n_frames, double rate) { return (int64_t)std::nearbyint(1000.0 * ((double)n_frames / rate)); }

// audio[s_ms:e_ms] -> real samples [a, b) followed by npad zeros, with the millisecond positions as pydub receives them (ints on the step's path, floats on the legacy one)
static inline int pb_pydub_slice_ms(int64_t n_frames, double rate, double s_ms, double e_ms, int64_t* a, int64_t* b, int64_t* npad) {
    const double L = (double)pb_pydub_len_ms(n_frames, rate);
    *a = *b = *npad = 0;
    if (!(s_ms >= 0.0) || !(e_ms >= 0.0)) return PB_UNIT_SLICE_ERROR;   // outside the reference's usage
    if (s_ms > L) s_ms = L; if (e_ms > L) e_ms = L;
    const double per_ms = rate / 1000.0;
    const int64_t sf = (int64_t)(s_ms * per_ms), ef = (int64_t)(e_ms * per_ms);
    const int64_t aa = sf < n_frames ? sf : n_frames;
    int64_t bb = ef < n_frames ? ef : n_frames; if (bb < aa) bb = aa;
    const int64_t expected = ef > sf ? ef - sf : 0;
    const int64_t missing = expected - (bb - aa);
    *a = aa; *b = bb;
    if (missing) {
        if ((double)missing > 2.0 * per_ms) return PB_UNIT_SLICE_ERROR;   // TooManyMissingFrames
        *npad = (bb - aa) > 0 ? missing : 0;                        // silence is cloned from the first frame: none if empty
    }
    return PB_UNIT_OK;
}

// audio[int(t0*1000):int(t1*1000)] -> real samples [a, b) followed by npad zeros. Returns 0 or PB_UNIT_SLICE_ERROR.
static inline int pb_pydub_slice(int64_t n_frames, double rate, double t0, double t1, int64_t* a, int64_t* b, int64_t* npad) {
    return pb_pydub_slice_ms(n_frames, rate, (double)(int64_t)(t0 * 1000.0), (double)(int64_t)(t1 * 1000.0), a, b, npad);
}

static inline double pb_part_duration(int64_t n_frames, double rate, int has_t1, double t0, double t1, int* status) {
    *status = PB_UNIT_OK;
    double d;
    if (has_t1) {
        int64_t a, b, npad;
        *status = pb_pydub_slice(n_frames, rate, t0, t1, &a, &b, &npad);
        d = (double)(b - a + npad) / rate;
    } else d = (double)n_frames / rate;
    return d != 0.0 ? d : 1e-4;                                      // "or 1e-4"
}

// get_lufs control flow -> the sample range whose loudness is returned. status carries the fallback / error flags.
static inline int pb_lufs_resolve(int64_t n_frames, double rate, double meter_rate, int has_t1, double t0, double t1,
                                  int64_t* a, int64_t* b, int64_t* npad) {
    int st = PB_UNIT_OK;
    if (has_t1) {
        st = pb_pydub_slice(n_frames, rate, t0, t1, a, b, npad);
        if (st != PB_UNIT_OK) return st;
    } else { *a = 0; *b = n_frames; *npad = 0; }
    if ((*b - *a + *npad) == 0) { *a = 0; *b = n_frames; *npad = 0; st |= PB_UNIT_LUFS_FALLBACK; }
    if ((double)(*b - *a + *npad) < 0.4 * meter_rate) {
        *a = 0; *b = n_frames; *npad = 0; st |= PB_UNIT_LUFS_FALLBACK;
        if ((double)n_frames < 0.4 * meter_rate) st |= PB_UNIT_LUFS_ERROR;
    }
    return st;
}

// pyloudnorm IIRfilter coefficients (RBJ cookbook forms used by iirfilter.py), normalised by a0.
static inline void pb_kweight_coeffs(double rate, double* b1, double* a1, double* b2, double* a2) {
    const double PI = 3.14159265358979323846;
    {
        const double G = 4.0, Q = 1.0 / std::sqrt(2.0), fc = 1500.0;
        const double A = std::pow(10.0, G / 40.0), w0 = 2.0 * PI * (fc / rate), al = std::sin(w0) / (2.0 * Q), cw = std::cos(w0), sA = std::sqrt(A);
        const double a0 = (A + 1) - (A - 1) * cw + 2 * sA * al;
        b1[0] = A * ((A + 1) + (A - 1) * cw + 2 * sA * al) / a0;
        b1[1] = -2 * A * ((A - 1) + (A + 1) * cw) / a0;
        b1[2] = A * ((A + 1) + (A - 1) * cw - 2 * sA * al) / a0;
        a1[0] = 1.0; a1[1] = 2 * ((A - 1) - (A + 1) * cw) / a0; a1[2] = ((A + 1) - (A - 1) * cw - 2 * sA * al) / a0;
    }
    {
        const double Q = 0.5, fc = 38.0;
        const double w0 = 2.0 * PI * (fc / rate), al = std::sin(w0) / (2.0 * Q), cw = std::cos(w0);
        const double a0 = 1 + al;
        b2[0] = (1 + cw) / 2 / a0; b2[1] = -(1 + cw) / a0; b2[2] = (1 + cw) / 2 / a0;
        a2[0] = 1.0; a2[1] = -2 * cw / a0; a2[2] = (1 - al) / a0;
    }
}

static inline int64_t pb_lufs_num_blocks(int64_t n, double rate) {
    const double T = (double)n / rate;
    const int64_t nb = (int64_t)(std::nearbyint((T - 0.4) / (0.4 * 0.25)) + 1.0);
    return nb < 0 ? 0 : nb;
}
