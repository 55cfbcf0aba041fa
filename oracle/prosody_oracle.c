/*
 * prosody_oracle.c — CPU ORACLE. TEST INFRASTRUCTURE ONLY. **PARITY UNPINNED**.
 *
 * Float64 restatement of the third-party numerics behind the reference's hot path
 *   /root/reference/Code/audioPipeline.py:326-335  get_median_pitch  -> parselmouth 0.4.5 (Praat 6.1.38)
 *   /root/reference/Code/audioPipeline.py:338-358  get_lufs          -> pyloudnorm 0.1.x
 * None of these packages (nor any golden vector) is shipped with the reference or installable here,
 * so this file restates the PUBLISHED algorithms (Praat fon/Sound_to_Pitch.cpp, fon/Pitch.cpp,
 * melder/NUMinterpol.cpp, dwsys/NUM2.cpp, fon/Sampled.cpp, fon/Sound.cpp; pyloudnorm meter.py,
 * iirfilter.py) and is pinned only by first-principles known-answer tests (tests/test_oracle_kat.py), by torchaudio's
 * independent BS.1770 implementation for the loudness part (tests/test_oracle_vs_torchaudio.py) and by a second, independently
 * written numpy / scipy restatement of the published pitch algorithm (tests/test_oracle_independent.py) — none of which is Praat.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library. The product (prosody-control-french-tts_b200/) never does.
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define PO_OK 0
#define PO_ERR_TOO_SHORT 1   /* Praat: "minimum pitch must not be less than ..." / shorter than window */
#define PO_ERR_NO_SAMPLES 2  /* Praat: "Extracted Sound would contain no samples." */
#define PO_ERR_WINDOW 3      /* Praat: "Analysis window too short." */
#define PO_ERR_LUFS_SHORT 4  /* pyloudnorm: ValueError "Audio must have length greater than the block size." */
#define PO_ERR_ALLOC 9

#define PO_PI 3.1415926535897932384626433832795028841972

typedef struct {
    double dt;                 /* <=0 -> auto (periodsPerWindow / minimumPitch / 4) */
    double minimumPitch;
    double periodsPerWindow;   /* 3.0 (AC_HANNING) */
    int32_t maxnCandidates;    /* 15 */
    double silenceThreshold;   /* 0.03 */
    double voicingThreshold;   /* 0.45 */
    double octaveCost;         /* 0.01 */
    double octaveJumpCost;     /* 0.35 */
    double voicedUnvoicedCost; /* 0.14 */
    double ceiling;            /* pitch ceiling */
} PoPitchParams;

typedef struct {
    int64_t nsamp_period, halfnsamp_period, nsamp_window, halfnsamp_window;
    int64_t minimumLag, maximumLag, nsampFFT, brent_ixmax, nFrames, maxnCandidates;
    double t1, dt, ceiling, dt_window;
} PoPitchGeom;

/* ------------------------------------------------------------------ counters (work model) */
typedef struct {
    int64_t frames, candidates, sinc_evals, sinc_terms, brent_iters;
} PoCounters;
static PoCounters g_cnt;
#ifdef _OPENMP
#pragma omp threadprivate(g_cnt)
#endif
void po_counters_reset(void) { memset(&g_cnt, 0, sizeof g_cnt); }
void po_counters_get(PoCounters *out) { *out = g_cnt; }

/* ------------------------------------------------------------------ FFT (radix-2, double) */
/* Praat uses an FFTPACK-derived real FFT (NUMfft_forward/backward). Any exact DFT is equivalent up to
 * float64 rounding; we use an in-place iterative radix-2 complex FFT with a real-input packing. */
typedef struct {
    int64_t n;       /* real length (power of two) */
    int64_t nc;      /* complex length n/2 */
    double *cosT, *sinT;   /* twiddles for length nc complex FFT: k=0..nc/2-1 */
    double *cosR, *sinR;   /* real-unpack twiddles exp(-2 pi i k / n): k=0..nc */
    int32_t *rev;
} PoFFT;

static int po_fft_init(PoFFT *t, int64_t n) {
    t->n = n; t->nc = n / 2;
    int64_t nc = t->nc;
    t->cosT = (double *)malloc(sizeof(double) * (nc / 2 + 1));
    t->sinT = (double *)malloc(sizeof(double) * (nc / 2 + 1));
    t->cosR = (double *)malloc(sizeof(double) * (nc + 1));
    t->sinR = (double *)malloc(sizeof(double) * (nc + 1));
    t->rev = (int32_t *)malloc(sizeof(int32_t) * nc);
    if (!t->cosT || !t->sinT || !t->cosR || !t->sinR || !t->rev) return PO_ERR_ALLOC;
    for (int64_t k = 0; k < nc / 2 + 1; k++) {
        t->cosT[k] = cos(2.0 * PO_PI * (double)k / (double)nc);
        t->sinT[k] = sin(2.0 * PO_PI * (double)k / (double)nc);
    }
    for (int64_t k = 0; k <= nc; k++) {
        t->cosR[k] = cos(2.0 * PO_PI * (double)k / (double)n);
        t->sinR[k] = sin(2.0 * PO_PI * (double)k / (double)n);
    }
    int bits = 0; while (((int64_t)1 << bits) < nc) bits++;
    for (int64_t i = 0; i < nc; i++) {
        int64_t r = 0;
        for (int b = 0; b < bits; b++) if (i & ((int64_t)1 << b)) r |= (int64_t)1 << (bits - 1 - b);
        t->rev[i] = (int32_t)r;
    }
    return PO_OK;
}
static void po_fft_free(PoFFT *t) {
    free(t->cosT); free(t->sinT); free(t->cosR); free(t->sinR); free(t->rev);
}
/* in-place complex FFT of length nc; sign=-1 forward (exp(-i..)), +1 backward (unnormalised) */
static void po_cfft(const PoFFT *t, double *re, double *im, int sign) {
    int64_t nc = t->nc;
    for (int64_t i = 0; i < nc; i++) {
        int64_t j = t->rev[i];
        if (j > i) { double a = re[i]; re[i] = re[j]; re[j] = a; a = im[i]; im[i] = im[j]; im[j] = a; }
    }
    for (int64_t len = 2; len <= nc; len <<= 1) {
        int64_t half = len >> 1, step = nc / len;
        for (int64_t i = 0; i < nc; i += len) {
            for (int64_t k = 0; k < half; k++) {
                double wr = t->cosT[k * step], wi = sign * t->sinT[k * step];
                double xr = re[i + k + half], xi = im[i + k + half];
                double tr = wr * xr - wi * xi, ti = wr * xi + wi * xr;
                re[i + k + half] = re[i + k] - tr; im[i + k + half] = im[i + k] - ti;
                re[i + k] += tr; im[i + k] += ti;
            }
        }
    }
}
/* Autocorrelation via power spectrum: in x[0..n-1] (zero padded), out ac[0..n-1] (unnormalised, real).
 * work: 4*nc doubles. */
static void po_autocorr(const PoFFT *t, const double *x, double *ac, double *work) {
    int64_t nc = t->nc, n = t->n;
    double *zr = work, *zi = work + nc, *pr = work + 2 * nc, *pi_ = work + 3 * nc;
    for (int64_t i = 0; i < nc; i++) { zr[i] = x[2 * i]; zi[i] = x[2 * i + 1]; }
    po_cfft(t, zr, zi, -1);
    /* unpack to X_k, k=0..nc ; power P_k */
    /* X_k = 0.5*(Z_k + conj(Z_{nc-k})) - 0.5 i * W^k * (Z_k - conj(Z_{nc-k})),  W = exp(-2 pi i / n) */
    /* then pack the (real, even) power spectrum for the inverse: Y_k = (P_k + P_{nc-k}) + i W^{-k} (P_k - P_{nc-k}) */
    double *P = ac; /* reuse as temp for P_0..P_nc (needs nc+1 <= n) */
    for (int64_t k = 0; k <= nc; k++) {
        int64_t k1 = (k == nc) ? 0 : k, k2 = (k == 0) ? 0 : nc - k;
        double ar = zr[k1], ai = zi[k1], br = zr[k2], bi = -zi[k2]; /* b = conj(Z_{nc-k}) */
        double sr = 0.5 * (ar + br), si = 0.5 * (ai + bi);
        double dr = 0.5 * (ar - br), di = 0.5 * (ai - bi);
        /* -i * W^k * d, W^k = cos - i sin */
        double wr = t->cosR[k], wi = -t->sinR[k];
        double er = wr * dr - wi * di, ei = wr * di + wi * dr; /* W^k d */
        double xr = sr + ei, xi = si - er;                      /* s - i*(W^k d) */
        P[k] = xr * xr + xi * xi;
    }
    for (int64_t k = 0; k < nc; k++) {
        double a = P[k], b = P[nc - k];
        double s = a + b, d = a - b;
        /* i * W^{-k} * d ; W^{-k} = cos + i sin */
        pr[k] = s - t->sinR[k] * d;
        pi_[k] = t->cosR[k] * d;
    }
    po_cfft(t, pr, pi_, +1);
    for (int64_t i = 0; i < nc; i++) { ac[2 * i] = pr[i]; ac[2 * i + 1] = pi_[i]; }
    (void)n;
}

/* ------------------------------------------------------------------ NUM_interpolate_sinc */
/* y is 1-based: y[1..n] (pass pointer to element 0 = unused). Praat melder/NUMinterpol.cpp. */
static double po_interpolate_sinc(const double *y, int64_t n, double x, int64_t maxDepth) {
    int64_t ix, midleft = (int64_t)floor(x), midright = midleft + 1, left, right;
    double result = 0.0, a, halfsina, aa, daa;
    if (n < 1) return NAN;
    if (x > (double)n) return y[n];
    if (x < 1.0) return y[1];
    if (x == (double)midleft) return y[midleft];
    if (maxDepth > midright - 1) maxDepth = midright - 1;
    if (maxDepth > n - midleft) maxDepth = n - midleft;
    if (maxDepth <= 0) return y[(int64_t)floor(x + 0.5)];
    if (maxDepth == 1) return y[midleft] + (x - midleft) * (y[midright] - y[midleft]);
    if (maxDepth == 2) {
        double yl = y[midleft], yr = y[midright];
        double dyl = 0.5 * (yr - y[midleft - 1]), dyr = 0.5 * (y[midright + 1] - yl);
        double fil = x - midleft, fir = midright - x;
        return yl * fir + yr * fil - fil * fir * (0.5 * (dyr - dyl) + (fil - 0.5) * (dyl + dyr - 2 * (yr - yl)));
    }
    left = midright - maxDepth; right = midleft + maxDepth;
    a = PO_PI * (x - midleft);
    halfsina = 0.5 * sin(a);
    aa = a / (x - left + 1);
    daa = PO_PI / (x - left + 1);
    for (ix = midleft; ix >= left; ix--) {
        double d = halfsina / a * (1.0 + cos(aa));
        result += y[ix] * d;
        a += PO_PI; aa += daa; halfsina = -halfsina;
    }
    a = PO_PI * (midright - x);
    halfsina = 0.5 * sin(a);
    aa = a / (right - x + 1);
    daa = PO_PI / (right - x + 1);
    for (ix = midright; ix <= right; ix++) {
        double d = halfsina / a * (1.0 + cos(aa));
        result += y[ix] * d;
        a += PO_PI; aa += daa; halfsina = -halfsina;
    }
    g_cnt.sinc_evals++; g_cnt.sinc_terms += 2 * maxDepth;
    return result;
}

/* ------------------------------------------------------------------ NUMimproveMaximum via Brent */
typedef struct { const double *y; int64_t n; int64_t depth; } PoImprove;
static double po_improve_eval(double x, const PoImprove *p) { return -po_interpolate_sinc(p->y, p->n, x, p->depth); }

/* Praat dwsys/NUM2.cpp NUMminimize_brent */
static double po_minimize_brent(const PoImprove *p, double a, double b, double tol, double *fx) {
    double x, v, fv, w, fw;
    const double golden = 1.0 - 0.6180339887498948482045868343656381177203;
    const double sqrt_epsilon = sqrt(DBL_EPSILON);
    int itermax = 60;
    v = a + golden * (b - a);
    fv = po_improve_eval(v, p);
    x = v; w = v; *fx = fv; fw = fv;
    for (int iter = 1; iter <= itermax; iter++) {
        double range = b - a;
        double middle_range = (a + b) / 2.0;
        double tol_act = sqrt_epsilon * fabs(x) + tol / 3.0;
        double new_step;
        g_cnt.brent_iters++;
        if (fabs(x - middle_range) + range / 2.0 <= 2.0 * tol_act) return x;
        new_step = golden * (x < middle_range ? b - x : a - x);
        if (fabs(x - w) >= tol_act) {
            double pp, q, t;
            t = (x - w) * (*fx - fv);
            q = (x - v) * (*fx - fw);
            pp = (x - v) * q - (x - w) * t;
            q = 2.0 * (q - t);
            if (q > 0.0) pp = -pp; else q = -q;
            if (fabs(pp) < fabs(new_step * q) && pp > q * (a - x + 2.0 * tol_act) && pp < q * (b - x - 2.0 * tol_act))
                new_step = pp / q;
        }
        if (fabs(new_step) < tol_act) new_step = new_step > 0.0 ? tol_act : -tol_act;
        {
            double t = x + new_step;
            double ft = po_improve_eval(t, p);
            if (ft <= *fx) {
                if (t < x) b = x; else a = x;
                v = w; w = x; x = t;
                fv = fw; fw = *fx; *fx = ft;
            } else {
                if (t < x) a = t; else b = t;
                if (ft <= fw || w == x) { v = w; w = t; fv = fw; fw = ft; }
                else if (ft <= fv || v == x || v == w) { v = t; fv = ft; }
            }
        }
    }
    return x;
}

/* NUMimproveExtremum (isMaximum = true), interpolation depth 70 or 700 */
static double po_improve_maximum(const double *y, int64_t n, int64_t ixmid, int64_t depth, double *ixmid_real) {
    if (ixmid <= 1) { *ixmid_real = 1.0; return y[1]; }
    if (ixmid >= n) { *ixmid_real = (double)n; return y[n]; }
    PoImprove p = { y, n, depth };
    double result;
    *ixmid_real = po_minimize_brent(&p, (double)(ixmid - 1), (double)(ixmid + 1), 1e-10, &result);
    return -result;
}

/* ------------------------------------------------------------------ geometry: Sound_to_Pitch_any preamble */
static int64_t po_ifloor(double x) { return (int64_t)floor(x); }

int po_pitch_geometry(int64_t nx, double dx, double x1, const PoPitchParams *p, PoPitchGeom *g) {
    double minimumPitch = p->minimumPitch, periodsPerWindow = p->periodsPerWindow, ceiling = p->ceiling, dt = p->dt;
    int64_t maxnCandidates = p->maxnCandidates;
    if ((double)maxnCandidates < ceiling / minimumPitch) maxnCandidates = po_ifloor(ceiling / minimumPitch);
    if (dt <= 0.0) dt = periodsPerWindow / minimumPitch / 4.0;
    double interpolation_depth = 0.5; /* AC_HANNING */
    volatile double duration = dx * (double)nx;
    if (minimumPitch < periodsPerWindow / duration) return PO_ERR_TOO_SHORT;
    g->nsamp_period = po_ifloor(1.0 / dx / minimumPitch);
    g->halfnsamp_period = g->nsamp_period / 2 + 1;
    if (ceiling > 0.5 / dx) ceiling = 0.5 / dx;
    double dt_window = periodsPerWindow / minimumPitch;
    g->nsamp_window = po_ifloor(dt_window / dx);
    g->halfnsamp_window = g->nsamp_window / 2 - 1;
    if (g->halfnsamp_window < 2) return PO_ERR_WINDOW;
    g->nsamp_window = g->halfnsamp_window * 2;
    g->minimumLag = po_ifloor(1.0 / dx / ceiling);
    if (g->minimumLag < 2) g->minimumLag = 2;
    g->maximumLag = po_ifloor((double)g->nsamp_window / periodsPerWindow) + 2;
    if (g->maximumLag > g->nsamp_window) g->maximumLag = g->nsamp_window;
    /* Sampled_shortTermAnalysis (fon/Sampled.cpp) */
    {
        volatile double myDuration = dx * (double)nx;
        if (dt_window > myDuration) return PO_ERR_TOO_SHORT;
        g->nFrames = po_ifloor((myDuration - dt_window) / dt) + 1;
        double ourMidTime = x1 - 0.5 * dx + 0.5 * myDuration;
        double thyDuration = (double)g->nFrames * dt;
        g->t1 = ourMidTime - 0.5 * thyDuration + 0.5 * dt;
    }
    g->nsampFFT = 1;
    while ((double)g->nsampFFT < (double)g->nsamp_window * (1.0 + interpolation_depth)) g->nsampFFT *= 2;
    g->brent_ixmax = po_ifloor((double)g->nsamp_window * interpolation_depth);
    g->maxnCandidates = maxnCandidates;
    g->dt = dt; g->ceiling = ceiling; g->dt_window = dt_window;
    return PO_OK;
}

/* ------------------------------------------------------------------ Sound_to_Pitch_any + Pitch_pathFinder */
/*
 * x[0..nx-1]: the (already extracted) sound, float64. dx, x1: Sampled geometry.
 * Outputs (caller-allocated, g->nFrames rows, g->maxnCandidates columns, row-major):
 *   cand_f, cand_s : candidates AFTER the path finder's swap (slot 0 = selected)   [may be NULL]
 *   pre_f, pre_s   : candidates BEFORE the path finder (slot 0 = voiceless)        [may be NULL]
 *   ncand          : candidates per frame
 *   intensity      : per frame
 *   sel_f, sel_s   : selected_array['frequency'/'strength'] (raw slot 0)
 */
int po_pitch_ac(const double *x, int64_t nx, double dx, double x1, const PoPitchParams *p, const PoPitchGeom *g,
                double *sel_f, double *sel_s, int32_t *ncand_out, double *intensity_out,
                double *cand_f_out, double *cand_s_out, double *pre_f_out, double *pre_s_out)
{
    const int64_t nF = g->nFrames, maxC = g->maxnCandidates, nw = g->nsamp_window, nfft = g->nsampFFT;
    const int64_t B = g->brent_ixmax;
    const double minimumPitch = p->minimumPitch, ceiling = g->ceiling;
    int status = PO_OK;

    double *cf = (double *)calloc((size_t)(nF * maxC), sizeof(double));
    double *cs = (double *)calloc((size_t)(nF * maxC), sizeof(double));
    int32_t *nc = (int32_t *)calloc((size_t)nF, sizeof(int32_t));
    double *inten = (double *)calloc((size_t)nF, sizeof(double));
    double *window = (double *)calloc((size_t)(nw + 1), sizeof(double));
    double *windowR = (double *)calloc((size_t)(nfft + 1), sizeof(double));
    double *frame = (double *)calloc((size_t)nfft, sizeof(double));
    double *ac = (double *)calloc((size_t)nfft, sizeof(double));
    double *work = (double *)calloc((size_t)(2 * nfft), sizeof(double));
    double *rbuf = (double *)calloc((size_t)(2 * B + 3), sizeof(double));
    int64_t *imax = (int64_t *)calloc((size_t)(maxC + 1), sizeof(int64_t));
    PoFFT fft; int fft_ok = 0;
    if (!cf || !cs || !nc || !inten || !window || !windowR || !frame || !ac || !work || !rbuf || !imax) { status = PO_ERR_ALLOC; goto done; }

    /* global peak: max |x - mean| over the whole sound */
    double globalPeak = 0.0;
    {
        double sum = 0.0;
        for (int64_t i = 0; i < nx; i++) sum += x[i];
        double mean = sum / (double)nx;
        for (int64_t i = 0; i < nx; i++) { double v = fabs(x[i] - mean); if (v > globalPeak) globalPeak = v; }
    }
    /* voiceless candidate is always present */
    for (int64_t f = 0; f < nF; f++) { nc[f] = 1; }
    if (globalPeak == 0.0) {
        /* Praat returns the Pitch before Sound_into_PitchFrame AND before the path finder:
           frames have nCandidates=1, frequency 0, intensity 0. */
        goto emit;
    }
    if (po_fft_init(&fft, nfft) != PO_OK) { status = PO_ERR_ALLOC; goto done; }
    fft_ok = 1;

    for (int64_t i = 1; i <= nw; i++) window[i] = 0.5 - 0.5 * cos((double)i * 2.0 * PO_PI / (double)(nw + 1));
    /* normalised autocorrelation of the window */
    memset(frame, 0, sizeof(double) * (size_t)nfft);
    for (int64_t i = 1; i <= nw; i++) frame[i - 1] = window[i];
    po_autocorr(&fft, frame, ac, work);
    for (int64_t i = 1; i < nw; i++) windowR[i] = ac[i] / ac[0];   /* windowR[i] ~ Praat windowR[i+1] */
    windowR[0] = 1.0;

    double *r = rbuf + B + 1;  /* r[-B-1 .. B+1] addressable; y[k] = r[k - B - 1] for k=1..2B+1 */
    const double *y = r - B - 1; /* y[1] = r[-B] */
    const int64_t ny = 2 * B + 1;

    for (int64_t iframe = 1; iframe <= nF; iframe++) {
        double t = g->t1 + (double)(iframe - 1) * g->dt;
        int64_t leftSample = po_ifloor((t - x1) / dx) + 1, rightSample = leftSample + 1;
        int64_t startSample, endSample;
        double *Cf = cf + (iframe - 1) * maxC, *Cs = cs + (iframe - 1) * maxC;
        g_cnt.frames++;
        /* local mean: one longest period to both sides */
        startSample = rightSample - g->nsamp_period;
        endSample = leftSample + g->nsamp_period;
        double localMean = 0.0;
        for (int64_t i = startSample; i <= endSample; i++) localMean += (i >= 1 && i <= nx) ? x[i - 1] : 0.0;
        localMean /= (double)(2 * g->nsamp_period);
        /* window the frame */
        startSample = rightSample - g->halfnsamp_window;
        for (int64_t j = 1, i = startSample; j <= nw; j++, i++)
            frame[j - 1] = (((i >= 1 && i <= nx) ? x[i - 1] : 0.0) - localMean) * window[j];
        for (int64_t j = nw; j < nfft; j++) frame[j] = 0.0;
        /* local peak: half a longest period to both sides */
        double localPeak = 0.0;
        if ((startSample = g->halfnsamp_window + 1 - g->halfnsamp_period) < 1) startSample = 1;
        if ((endSample = g->halfnsamp_window + g->halfnsamp_period) > nw) endSample = nw;
        for (int64_t j = startSample; j <= endSample; j++) { double v = fabs(frame[j - 1]); if (v > localPeak) localPeak = v; }
        inten[iframe - 1] = localPeak > globalPeak ? 1.0 : localPeak / globalPeak;
        /* autocorrelation, normalised by lag 0 and by the window's autocorrelation */
        po_autocorr(&fft, frame, ac, work);
        r[0] = 1.0;
        for (int64_t i = 1; i <= B; i++) r[-i] = r[i] = ac[i] / (ac[0] * windowR[i]);
        Cf[0] = 0.0; Cs[0] = 0.0; nc[iframe - 1] = 1;
        if (localPeak == 0.0) continue;
        /* first pass: maxima of r */
        imax[0] = 0;
        int64_t ncf = 1;
        for (int64_t i = 2; i < g->maximumLag && i < B; i++) {
            if (r[i] > 0.5 * p->voicingThreshold && r[i] > r[i - 1] && r[i] >= r[i + 1]) {
                int64_t place = 0;
                double dr = 0.5 * (r[i + 1] - r[i - 1]), d2r = 2.0 * r[i] - r[i - 1] - r[i + 1];
                double frequencyOfMaximum = 1.0 / dx / ((double)i + dr / d2r);
                int64_t offset = -B - 1;
                double strengthOfMaximum = po_interpolate_sinc(y, ny, 1.0 / dx / frequencyOfMaximum - (double)offset, 30);
                if (strengthOfMaximum > 1.0) strengthOfMaximum = 1.0 / strengthOfMaximum;
                if (ncf < maxC) {
                    place = ncf++;           /* 0-based slot */
                } else {
                    double weakest = 2.0;
                    for (int64_t iweak = 1; iweak < maxC; iweak++) {
                        double localStrength = Cs[iweak] - p->octaveCost * log2(minimumPitch / Cf[iweak]);
                        if (localStrength < weakest) { weakest = localStrength; place = iweak; }
                    }
                    if (strengthOfMaximum - p->octaveCost * log2(minimumPitch / frequencyOfMaximum) <= weakest) place = 0;
                }
                if (place) { Cf[place] = frequencyOfMaximum; Cs[place] = strengthOfMaximum; imax[place] = i; }
            }
        }
        nc[iframe - 1] = (int32_t)ncf;
        g_cnt.candidates += ncf - 1;
        /* second pass: sinc70/700 + Brent */
        for (int64_t i = 1; i < ncf; i++) {
            double xmid, ymid;
            int64_t offset = -B - 1;
            ymid = po_improve_maximum(y, ny, imax[i] - offset, Cf[i] > 0.3 / dx ? 700 : 70, &xmid);
            xmid += (double)offset;
            Cf[i] = 1.0 / dx / xmid;
            if (ymid > 1.0) ymid = 1.0 / ymid;
            Cs[i] = ymid;
        }
    }

    if (pre_f_out) memcpy(pre_f_out, cf, sizeof(double) * (size_t)(nF * maxC));
    if (pre_s_out) memcpy(pre_s_out, cs, sizeof(double) * (size_t)(nF * maxC));

    /* ---- Pitch_pathFinder (fon/Pitch.cpp) */
    {
        double timeStepCorrection = 0.01 / g->dt;
        double octaveJumpCost = p->octaveJumpCost * timeStepCorrection;
        double voicedUnvoicedCost = p->voicedUnvoicedCost * timeStepCorrection;
        double *delta = (double *)calloc((size_t)(nF * maxC), sizeof(double));
        int32_t *psi = (int32_t *)calloc((size_t)(nF * maxC), sizeof(int32_t));
        if (!delta || !psi) { free(delta); free(psi); status = PO_ERR_ALLOC; goto done; }
        for (int64_t f = 0; f < nF; f++) {
            double unvoicedStrength = p->silenceThreshold <= 0 ? 0.0 :
                2.0 - inten[f] / (p->silenceThreshold / (1.0 + p->voicingThreshold));
            unvoicedStrength = p->voicingThreshold + (unvoicedStrength > 0.0 ? unvoicedStrength : 0.0);
            for (int64_t c = 0; c < nc[f]; c++) {
                double fr = cf[f * maxC + c];
                int voiceless = !(fr > 0.0 && fr < ceiling);
                delta[f * maxC + c] = voiceless ? unvoicedStrength : cs[f * maxC + c] - p->octaveCost * log2(ceiling / fr);
            }
        }
        for (int64_t f = 1; f < nF; f++) {
            double *prevDelta = delta + (f - 1) * maxC, *curDelta = delta + f * maxC;
            int32_t *curPsi = psi + f * maxC;
            for (int64_t c2 = 0; c2 < nc[f]; c2++) {
                double f2 = cf[f * maxC + c2];
                volatile double maximum = -1e30, value;
                int32_t place = -1;
                for (int64_t c1 = 0; c1 < nc[f - 1]; c1++) {
                    double f1 = cf[(f - 1) * maxC + c1];
                    double transitionCost;
                    int previousVoiceless = !(f1 > 0.0 && f1 < ceiling);
                    int currentVoiceless = !(f2 > 0.0 && f2 < ceiling);
                    if (currentVoiceless) transitionCost = previousVoiceless ? 0.0 : voicedUnvoicedCost;
                    else transitionCost = previousVoiceless ? voicedUnvoicedCost : octaveJumpCost * fabs(log2(f1 / f2));
                    value = prevDelta[c1] - transitionCost + curDelta[c2];
                    if (value > maximum) { maximum = value; place = (int32_t)c1; }
                }
                curDelta[c2] = maximum;
                curPsi[c2] = place;
            }
        }
        int32_t place = 0;
        double maximum = delta[(nF - 1) * maxC + 0];
        for (int64_t c = 1; c < nc[nF - 1]; c++)
            if (delta[(nF - 1) * maxC + c] > maximum) { place = (int32_t)c; maximum = delta[(nF - 1) * maxC + c]; }
        for (int64_t f = nF - 1; f >= 0; f--) {
            double hf = cf[f * maxC], hs = cs[f * maxC];
            cf[f * maxC] = cf[f * maxC + place]; cs[f * maxC] = cs[f * maxC + place];
            cf[f * maxC + place] = hf; cs[f * maxC + place] = hs;
            place = psi[f * maxC + place];
        }
        free(delta); free(psi);
    }

emit:
    for (int64_t f = 0; f < nF; f++) {
        if (sel_f) sel_f[f] = cf[f * maxC];
        if (sel_s) sel_s[f] = cs[f * maxC];
        if (ncand_out) ncand_out[f] = nc[f];
        if (intensity_out) intensity_out[f] = inten[f];
    }
    if (cand_f_out) memcpy(cand_f_out, cf, sizeof(double) * (size_t)(nF * maxC));
    if (cand_s_out) memcpy(cand_s_out, cs, sizeof(double) * (size_t)(nF * maxC));
    if (globalPeak == 0.0) {
        if (pre_f_out) memset(pre_f_out, 0, sizeof(double) * (size_t)(nF * maxC));
        if (pre_s_out) memset(pre_s_out, 0, sizeof(double) * (size_t)(nF * maxC));
    }
done:
    if (fft_ok) po_fft_free(&fft);
    free(cf); free(cs); free(nc); free(inten); free(window); free(windowR); free(frame); free(ac); free(work); free(rbuf); free(imax);
    return status;
}

/* ------------------------------------------------------------------ Sound_extractPart + to_pitch + median */
/*
 * ≙ get_median_pitch (audioPipeline.py:326-335) on an in-memory mono s16 file.
 *   pcm[0..file_nx-1], sample rate sr. has_t1==0 -> whole file (to_pitch on the Sound itself);
 *   1 -> extract_part(from_time=t0, to_time=t1, rectangular, relwidth 1, preserve_times=True);
 *   2 -> the same with preserve_times=False (parselmouth's default, used by the legacy compute_pitch_adjustments.py).
 * Praat WAV decode: s16 / 32768. Sound: dx=1/sr, x1=0.5/sr, xmin=0, xmax=nx*dx.
 * Returns status; *median = np.median(freqs[freqs>0]) or 0.0; counts out.
 * If sel_f != NULL it must hold the frames (use po_pitch_unit_nframes first).
 */
static int po_extract_geometry(int64_t file_nx, double sr, int has_t1, double t0, double t1,
                               int64_t *ix1, int64_t *nx, double *x1_part, double *dx_out)
{
    double dx = 1.0 / sr, x1 = 0.5 / sr;
    *dx_out = dx;
    if (!has_t1) { *ix1 = 1; *nx = file_nx; *x1_part = x1; return PO_OK; }
    if (t0 == t1) { t0 = 0.0; t1 = (double)file_nx * dx; } /* Praat: t1==t2 -> whole domain [xmin,xmax] */
    int64_t i1 = 1 + (int64_t)ceil((t0 - x1) / dx);
    int64_t i2 = 1 + (int64_t)floor((t1 - x1) / dx);
    if (i2 < i1) return PO_ERR_NO_SAMPLES;
    *ix1 = i1; *nx = i2 - i1 + 1;
    *x1_part = x1 + (double)(i1 - 1) * dx;
    if (has_t1 == 2) *x1_part -= t0;   /* preserve_times = false (Praat: thy xmin = 0; thy xmax -= t1; thy x1 -= t1) */
    return PO_OK;
}

int po_pitch_unit_geometry(int64_t file_nx, double sr, int has_t1, double t0, double t1,
                           const PoPitchParams *p, PoPitchGeom *g, int64_t *ix1_out, int64_t *nx_out, double *x1_out)
{
    int64_t ix1, nx; double x1p, dx;
    int st = po_extract_geometry(file_nx, sr, has_t1, t0, t1, &ix1, &nx, &x1p, &dx);
    if (st != PO_OK) return st;
    if (ix1_out) *ix1_out = ix1;
    if (nx_out) *nx_out = nx;
    if (x1_out) *x1_out = x1p;
    return po_pitch_geometry(nx, dx, x1p, p, g);
}

static int po_cmp_double(const void *a, const void *b) {
    double x = *(const double *)a, y = *(const double *)b;
    return (x > y) - (x < y);
}

int po_median_pitch_i16(const int16_t *pcm, int64_t file_nx, double sr, int has_t1, double t0, double t1,
                        const PoPitchParams *p, double *median, int32_t *n_voiced, int32_t *n_frames,
                        double *sel_f, double *sel_s, int32_t *ncand, double *intensity,
                        double *pre_f, double *pre_s)
{
    PoPitchGeom g; int64_t ix1, nx; double x1p;
    double dx = 1.0 / sr;
    *median = 0.0; if (n_voiced) *n_voiced = 0; if (n_frames) *n_frames = 0;
    int st = po_pitch_unit_geometry(file_nx, sr, has_t1, t0, t1, p, &g, &ix1, &nx, &x1p);
    if (st != PO_OK) return st;
    double *x = (double *)calloc((size_t)nx, sizeof(double));
    double *f = sel_f ? sel_f : (double *)calloc((size_t)g.nFrames, sizeof(double));
    if (!x || !f) { free(x); if (!sel_f) free(f); return PO_ERR_ALLOC; }
    for (int64_t i = 0; i < nx; i++) {
        int64_t src = ix1 - 1 + i; /* 0-based file index */
        x[i] = (src >= 0 && src < file_nx) ? (double)pcm[src] / 32768.0 : 0.0;
    }
    st = po_pitch_ac(x, nx, dx, x1p, p, &g, f, sel_s, ncand, intensity, NULL, NULL, pre_f, pre_s);
    if (st == PO_OK) {
        int64_t nv = 0;
        double *v = (double *)malloc(sizeof(double) * (size_t)(g.nFrames > 0 ? g.nFrames : 1));
        if (!v) st = PO_ERR_ALLOC;
        else {
            for (int64_t i = 0; i < g.nFrames; i++) if (f[i] > 0.0) v[nv++] = f[i];
            if (nv > 0) {
                qsort(v, (size_t)nv, sizeof(double), po_cmp_double);
                *median = (nv & 1) ? v[nv / 2] : (v[nv / 2 - 1] + v[nv / 2]) / 2.0; /* np.median: mean of the two middles */
            }
            if (n_voiced) *n_voiced = (int32_t)nv;
            if (n_frames) *n_frames = (int32_t)g.nFrames;
            free(v);
        }
    }
    free(x); if (!sel_f) free(f);
    return st;
}

/* ------------------------------------------------------------------ pyloudnorm: K-weighting + gated loudness */
/* RBJ-form coefficients exactly as pyloudnorm iirfilter.py generate_coefficients (float64). */
void po_kweight_coeffs(double rate, double *b_shelf, double *a_shelf, double *b_hp, double *a_hp) {
    {   /* high_shelf: G=4 dB, Q=1/sqrt(2), fc=1500 */
        double G = 4.0, Q = 1.0 / sqrt(2.0), fc = 1500.0;
        double A = pow(10.0, G / 40.0);
        double w0 = 2.0 * PO_PI * (fc / rate);
        double alpha = sin(w0) / (2.0 * Q);
        double b0 = A * ((A + 1) + (A - 1) * cos(w0) + 2 * sqrt(A) * alpha);
        double b1 = -2 * A * ((A - 1) + (A + 1) * cos(w0));
        double b2 = A * ((A + 1) + (A - 1) * cos(w0) - 2 * sqrt(A) * alpha);
        double a0 = (A + 1) - (A - 1) * cos(w0) + 2 * sqrt(A) * alpha;
        double a1 = 2 * ((A - 1) - (A + 1) * cos(w0));
        double a2 = (A + 1) - (A - 1) * cos(w0) - 2 * sqrt(A) * alpha;
        b_shelf[0] = b0 / a0; b_shelf[1] = b1 / a0; b_shelf[2] = b2 / a0;
        a_shelf[0] = a0 / a0; a_shelf[1] = a1 / a0; a_shelf[2] = a2 / a0;
    }
    {   /* high_pass: G=0, Q=0.5, fc=38 */
        double Q = 0.5, fc = 38.0;
        double w0 = 2.0 * PO_PI * (fc / rate);
        double alpha = sin(w0) / (2.0 * Q);
        double b0 = (1 + cos(w0)) / 2;
        double b1 = -(1 + cos(w0));
        double b2 = (1 + cos(w0)) / 2;
        double a0 = 1 + alpha;
        double a1 = -2 * cos(w0);
        double a2 = 1 - alpha;
        b_hp[0] = b0 / a0; b_hp[1] = b1 / a0; b_hp[2] = b2 / a0;
        a_hp[0] = a0 / a0; a_hp[1] = a1 / a0; a_hp[2] = a2 / a0;
    }
}

/* scipy.signal.lfilter, order 2, zero initial state (direct form II transposed), in place */
static void po_lfilter2(const double *b, const double *a, double *x, int64_t n) {
    double z0 = 0.0, z1 = 0.0;
    for (int64_t i = 0; i < n; i++) {
        double xi = x[i];
        double yi = b[0] * xi + z0;
        z0 = b[1] * xi - a[1] * yi + z1;
        z1 = b[2] * xi - a[2] * yi;
        x[i] = yi;
    }
}

/* Meter(rate).integrated_loudness(data) for mono float64 data (pyloudnorm meter.py). data is NOT modified. */
int po_integrated_loudness(const double *data, int64_t n, double rate, double *lufs) {
    const double T_g = 0.4, Gamma_a = -70.0, overlap = 0.75, step = 1.0 - overlap;
    if ((double)n < T_g * rate) return PO_ERR_LUFS_SHORT;   /* util.valid_audio */
    double bs[3], as[3], bh[3], ah[3];
    po_kweight_coeffs(rate, bs, as, bh, ah);
    double *y = (double *)malloc(sizeof(double) * (size_t)n);
    if (!y) return PO_ERR_ALLOC;
    memcpy(y, data, sizeof(double) * (size_t)n);
    po_lfilter2(bs, as, y, n);
    po_lfilter2(bh, ah, y, n);
    double T = (double)n / rate;
    int64_t numBlocks = (int64_t)(nearbyint((T - T_g) / (T_g * step)) + 1.0); /* np.round = half-to-even */
    if (numBlocks < 0) numBlocks = 0;
    double *z = (double *)calloc((size_t)(numBlocks > 0 ? numBlocks : 1), sizeof(double));
    double *l = (double *)calloc((size_t)(numBlocks > 0 ? numBlocks : 1), sizeof(double));
    if (!z || !l) { free(y); free(z); free(l); return PO_ERR_ALLOC; }
    for (int64_t j = 0; j < numBlocks; j++) {
        int64_t lo = (int64_t)(T_g * ((double)j * step) * rate);
        int64_t up = (int64_t)(T_g * ((double)j * step + 1.0) * rate);
        if (lo > n) lo = n; if (up > n) up = n;      /* numpy slice clipping */
        double s = 0.0;
        for (int64_t i = lo; i < up; i++) s += y[i] * y[i];
        z[j] = (1.0 / (T_g * rate)) * s;
        l[j] = -0.691 + 10.0 * log10(z[j]);          /* log10(0) = -inf, as numpy (with a warning) */
    }
    /* absolute gate (>=), mean, relative gate */
    double sum = 0.0; int64_t cnt = 0;
    for (int64_t j = 0; j < numBlocks; j++) if (l[j] >= Gamma_a) { sum += z[j]; cnt++; }
    double z_avg = cnt > 0 ? sum / (double)cnt : NAN;     /* np.mean([]) = nan */
    double Gamma_r = -0.691 + 10.0 * log10(z_avg) - 10.0;
    sum = 0.0; cnt = 0;
    for (int64_t j = 0; j < numBlocks; j++) if (l[j] > Gamma_r && l[j] > Gamma_a) { sum += z[j]; cnt++; }
    z_avg = cnt > 0 ? sum / (double)cnt : 0.0;            /* nan_to_num(nan) = 0 */
    if (isnan(z_avg)) z_avg = 0.0;
    *lufs = -0.691 + 10.0 * log10(z_avg);                 /* may be -inf */
    free(y); free(z); free(l);
    return PO_OK;
}

/* ≙ the numeric part of get_lufs (audioPipeline.py:343-358) on a resolved sample range:
 * samples = pcm[a:b] followed by npad zeros (pydub's missing-frame padding); peak-normalise; meter at meter_rate. */
int po_lufs_i16(const int16_t *pcm, int64_t a, int64_t b, int64_t npad, double meter_rate, double *lufs) {
    int64_t n = (b - a) + npad;
    if (n < 0) n = 0;
    double *x = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
    if (!x) return PO_ERR_ALLOC;
    double peak = 0.0;
    for (int64_t i = 0; i < b - a; i++) { x[i] = (double)pcm[a + i]; double v = fabs(x[i]); if (v > peak) peak = v; }
    if (peak == 0.0) peak = 1.0;                  /* "or 1.0" */
    for (int64_t i = 0; i < n; i++) x[i] = x[i] / peak;
    int st = po_integrated_loudness(x, n, meter_rate, lufs);
    free(x);
    return st;
}

/* ------------------------------------------------------------------ batch drivers (OpenMP) for the CPU baseline */
/* unit arrays are SoA; pcm is one concatenated buffer, file_off[] gives each unit's file start. */
int po_batch_median_pitch(const int16_t *pcm, int64_t n_units, const int64_t *file_off, const int64_t *file_nx,
                          const double *sr, const int32_t *has_t1, const double *t0, const double *t1,
                          const PoPitchParams *p, double *median, int32_t *n_voiced, int32_t *n_frames, int32_t *status,
                          int n_threads)
{
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t u = 0; u < n_units; u++) {
        status[u] = po_median_pitch_i16(pcm + file_off[u], file_nx[u], sr[u], has_t1[u], t0[u], t1[u], p,
                                        &median[u], &n_voiced[u], &n_frames[u], NULL, NULL, NULL, NULL, NULL, NULL);
    }
    return PO_OK;
}

int po_batch_lufs(const int16_t *pcm, int64_t n_units, const int64_t *file_off, const int64_t *a, const int64_t *b,
                  const int64_t *npad, const double *meter_rate, double *lufs, int32_t *status, int n_threads)
{
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t u = 0; u < n_units; u++)
        status[u] = po_lufs_i16(pcm + file_off[u], a[u], b[u], npad[u], meter_rate[u], &lufs[u]);
    return PO_OK;
}

int po_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
