"""A second, independently written restatement of the pitch tracker, used only to cross-check ``oracle/prosody_oracle.c``.

TEST INFRASTRUCTURE.  The C oracle follows Praat's source routine by routine (loops, 1-based indices, in-place FFT).  This file
restates the same published algorithm — P. Boersma, "Accurate short-term analysis of the fundamental frequency and the
harmonics-to-noise ratio of a sampled sound", IFA Proceedings 17 (1993), and the Praat manual pages "Sound: To Pitch (ac)..." /
"Pitch: path finder" — from the formulas, with numpy's FFT, vectorised windowed-sinc sums, scipy's bounded scalar minimiser instead
of a hand-written Brent, and a dynamic programme written from the manual's cost function.  Two implementations that share no code and
no structure agreeing to optimiser tolerance on real speech is what this adds over the first-principles known-answer tests: it sees
the window and its autocorrelation, the local-mean span, the lag range, the sinc depth, the candidate rule, the transition costs and
the tie-breaking.  It is still not Praat itself: a constant remembered wrongly in BOTH restatements would pass.
"""
import math
import wave
from pathlib import Path

import numpy as np
import pytest
from scipy.optimize import minimize_scalar

from conftest import speechlike

CLIPS = Path(__file__).resolve().parent / "golden" / "clips"

# Praat's defaults for Sound: To Pitch (ac)... as parselmouth's to_pitch() leaves them
MAX_CANDIDATES = 15
SILENCE_THRESHOLD = 0.03
VOICING_THRESHOLD = 0.45
OCTAVE_COST = 0.01
OCTAVE_JUMP_COST = 0.35
VOICED_UNVOICED_COST = 0.14
PERIODS_PER_WINDOW = 3.0


def sinc_interp(y, x, depth):
    """Windowed-sinc interpolation of y (0-based array standing for Praat's y[1..n]) at the 1-based position x: the sum over `depth`
    samples on either side of x of y * sinc(distance) * raised-cosine window of half-width (depth + fractional part)."""
    n = len(y)
    left = math.floor(x)
    if x > n:
        return y[n - 1]
    if x < 1:
        return y[0]
    if x == left:
        return y[left - 1]
    depth = min(depth, left, n - left)          # midright - 1 = left, n - midleft
    assert depth > 2
    out = 0.0
    for centre_dist, idx in ((x - left, left - np.arange(depth)), ((left + 1) - x, left + 1 + np.arange(depth))):
        d = centre_dist + np.arange(depth)                      # distance of every sample on this side from x
        width = centre_dist + depth                             # where the raised cosine reaches zero
        out += np.sum(y[idx - 1] * np.sin(np.pi * d) / (np.pi * d) * 0.5 * (1.0 + np.cos(np.pi * d / width)))
    return out


class Tracker:
    def __init__(self, sr, floor, ceiling):
        self.dx = 1.0 / sr
        self.floor, self.ceiling = floor, min(ceiling, 0.5 * sr)
        self.dt = PERIODS_PER_WINDOW / floor / 4.0               # time step 0: four frames per window
        dt_window = PERIODS_PER_WINDOW / floor
        nw = int(math.floor(dt_window / self.dx))
        self.half_nw = nw // 2 - 1
        self.nw = 2 * self.half_nw
        self.min_lag = max(2, int(math.floor(1.0 / self.dx / self.ceiling)))
        self.max_lag = min(int(math.floor(self.nw / PERIODS_PER_WINDOW)) + 2, self.nw)
        self.dt_window = dt_window
        self.nfft = 1
        while self.nfft < self.nw * 1.5:
            self.nfft *= 2
        self.B = int(math.floor(self.nw * 0.5))                   # lags the interpolation may touch
        self.period = int(math.floor(1.0 / self.dx / floor))      # samples in one period of the floor
        self.half_period = self.period // 2 + 1
        i = np.arange(1, self.nw + 1)
        self.window = 0.5 - 0.5 * np.cos(i * 2.0 * np.pi / (self.nw + 1))
        W = np.fft.rfft(self.window, self.nfft)
        wr = np.fft.irfft(np.abs(W) ** 2, self.nfft)
        self.window_r = wr / wr[0]

    def frames(self, nx):
        duration = nx * self.dx
        n_frames = int(math.floor((duration - self.dt_window) / self.dt)) + 1
        mid = 0.5 * duration                                      # x1 - dx/2 + duration/2 with x1 = dx/2
        t1 = mid - 0.5 * n_frames * self.dt + 0.5 * self.dt
        return n_frames, t1

    def candidates(self, x, mean, global_peak, t):
        """(frequencies, strengths, intensity) of one frame centred at time t; candidate 0 is the voiceless one."""
        dx = self.dx
        left = int(math.floor((t - 0.5 * dx) / dx)) + 1           # 1-based sample left of t (x1 = dx / 2)
        right = left + 1
        sample = lambda k: x[k - 1] if 1 <= k <= len(x) else 0.0
        local_mean = np.mean([sample(k) for k in range(right - self.period, left + self.period + 1)])
        start = right - self.half_nw
        frame = np.array([sample(start + j) - local_mean for j in range(self.nw)]) * self.window
        lo = max(1, self.half_nw + 1 - self.half_period); hi = min(self.nw, self.half_nw + self.half_period)
        local_peak = np.max(np.abs(frame[lo - 1:hi]))
        intensity = min(1.0, local_peak / global_peak)
        freqs, strengths = [0.0], [0.0]
        if local_peak == 0.0:
            return np.array(freqs), np.array(strengths), intensity
        F = np.fft.rfft(frame, self.nfft)
        ac = np.fft.irfft(np.abs(F) ** 2, self.nfft)
        r_pos = ac[:self.B + 1] / (ac[0] * self.window_r[:self.B + 1])
        r_pos[0] = 1.0
        r = np.concatenate([r_pos[:0:-1], r_pos])                  # lags -B .. B; lag l sits at index l + B (1-based position l + B + 1)
        at = lambda l: r[l + self.B]
        imax = [0]
        for i in range(2, min(self.max_lag, self.B)):
            if at(i) > 0.5 * VOICING_THRESHOLD and at(i) > at(i - 1) and at(i) >= at(i + 1):
                dr, d2r = 0.5 * (at(i + 1) - at(i - 1)), 2.0 * at(i) - at(i - 1) - at(i + 1)
                f_max = 1.0 / dx / (i + dr / d2r)
                s_max = sinc_interp(r, 1.0 / dx / f_max + self.B + 1, 30)
                if s_max > 1.0:
                    s_max = 1.0 / s_max
                if len(freqs) < MAX_CANDIDATES:
                    freqs.append(f_max); strengths.append(s_max); imax.append(i)
                else:                                              # replace the weakest, octave-cost corrected
                    weakest, place = 2.0, 0
                    for z in range(1, MAX_CANDIDATES):
                        loc = strengths[z] - OCTAVE_COST * math.log2(self.floor / freqs[z])
                        if loc < weakest:
                            weakest, place = loc, z
                    if s_max - OCTAVE_COST * math.log2(self.floor / f_max) > weakest:
                        freqs[place], strengths[place], imax[place] = f_max, s_max, i
        for c in range(1, len(freqs)):                             # second pass: the maximum of the interpolated curve near the sample maximum
            depth = 700 if freqs[c] > 0.3 / dx else 70
            pos = imax[c] + self.B + 1
            res = minimize_scalar(lambda p: -sinc_interp(r, p, depth), bounds=(pos - 1, pos + 1), method="bounded", options=dict(xatol=1e-11))
            xmid, ymid = res.x - self.B - 1, -res.fun
            freqs[c] = 1.0 / dx / xmid
            strengths[c] = 1.0 / ymid if ymid > 1.0 else ymid
        return np.array(freqs), np.array(strengths), intensity

    def best_path(self, cand_f, cand_s, intensity):
        """Global optimum of  sum(local score) - sum(transition cost)  by dynamic programming; ties go to the lower index."""
        correction = 0.01 / self.dt
        oj, vu = OCTAVE_JUMP_COST * correction, VOICED_UNVOICED_COST * correction
        voiced = lambda f: 0.0 < f < self.ceiling
        n = len(cand_f)
        local = []
        for f, s, inten in zip(cand_f, cand_s, intensity):
            unvoiced = VOICING_THRESHOLD + max(0.0, 2.0 - inten / (SILENCE_THRESHOLD / (1.0 + VOICING_THRESHOLD)))
            local.append([s_ - OCTAVE_COST * math.log2(self.ceiling / f_) if voiced(f_) else unvoiced for f_, s_ in zip(f, s)])
        delta, back = [local[0]], [None]
        for t in range(1, n):
            d_t, b_t = [], []
            for j, fj in enumerate(cand_f[t]):
                best, arg = -1e30, 0
                for i_, fi in enumerate(cand_f[t - 1]):
                    vi, vj = voiced(fi), voiced(fj)
                    cost = 0.0 if not vi and not vj else (vu if vi != vj else oj * abs(math.log2(fi / fj)))
                    v = delta[t - 1][i_] - cost + local[t][j]
                    if v > best:
                        best, arg = v, i_
                d_t.append(best); b_t.append(arg)
            delta.append(d_t); back.append(b_t)
        j = int(np.argmax(delta[-1]))                              # first maximum
        path = [j]
        for t in range(n - 1, 0, -1):
            j = back[t][j]
            path.append(j)
        return path[::-1]


def _read_clip(name, seconds, start=0.0):
    with wave.open(str(CLIPS / name), "rb") as w:
        sr = w.getframerate()
        assert w.getnchannels() == 1 and w.getsampwidth() == 2
        w.setpos(int(start * sr))
        return np.frombuffer(w.readframes(int(seconds * sr)), np.int16).copy(), sr


# the last case has more autocorrelation maxima than candidate slots (replace-the-weakest rule) and candidates above 0.3 / dx
# (interpolation depth 700)
CASES = [("clip", "segment_ph2.wav", 150.0), ("clip", "segment_ph7.wav", 75.0), ("synth", 16000, 75.0), ("synth", 24000, 150.0),
         ("tonal", 16000, 75.0)]


@pytest.mark.parametrize("kind,src,floor", CASES)
def test_oracle_candidates_and_path_match_independent_restatement(oracle, kind, src, floor):
    if kind == "clip":
        if not (CLIPS / src).exists():
            pytest.skip("clip fixture not present")
        pcm, sr = _read_clip(src, 0.9, start=1.0)
    elif kind == "tonal":
        from test_emu_parity import _many_maxima
        sr = src
        pcm = _many_maxima(sr, 0.3, 5)
    else:
        sr = src
        pcm = speechlike(1, 1.0, sr, seed=21)[0]
    o = oracle.pitch_track(pcm, sr, params=oracle.pitch_params(floor, 600.0), want_candidates=True)
    tr = Tracker(sr, floor, 600.0)
    g = o["geom"]
    assert (tr.nw, tr.max_lag, tr.nfft, tr.B) == (g.nsamp_window, g.maximumLag, g.nsampFFT, g.brent_ixmax)
    n_frames, t1 = tr.frames(len(pcm))
    assert n_frames == o["n_frames"] and abs(t1 - o["t1"]) < 1e-12 and abs(tr.dt - o["dt"]) < 1e-15
    x = pcm.astype(np.float64) / 32768.0
    mean = x.mean()
    global_peak = np.max(np.abs(x - mean))
    cand_f, cand_s, inten = [], [], []
    step = max(1, n_frames // 60)                                   # every frame feeds the path check; candidates are compared on a subset
    worst_f = worst_s = 0.0
    n_voiced_cands = 0
    for k in range(n_frames):
        nc = int(o["ncand"][k])
        if k % step == 0 or kind != "clip":
            f, s, it = tr.candidates(x, mean, global_peak, t1 + k * tr.dt)
            assert len(f) == nc, (k, len(f), nc)
            assert abs(it - o["intensity"][k]) < 1e-12
            of, os_ = o["pre_f"][k, :nc], o["pre_s"][k, :nc]
            assert of[0] == 0.0
            if nc > 1:
                worst_f = max(worst_f, float(np.max(np.abs(f[1:] - of[1:]) / of[1:])))
                worst_s = max(worst_s, float(np.max(np.abs(s[1:] - os_[1:]))))
                n_voiced_cands += nc - 1
        cand_f.append(o["pre_f"][k, :nc]); cand_s.append(o["pre_s"][k, :nc]); inten.append(o["intensity"][k])
    assert n_voiced_cands > 15
    if kind == "tonal":
        assert int(o["ncand"].max()) == MAX_CANDIDATES                # the replacement rule ran
    # two optimisers (Brent to 1e-10 in the oracle, scipy's bounded minimiser here) on the same smooth maximum: the heights agree to
    # rounding, the positions to what a float64 maximum can be located to (~sqrt(eps) of the peak's width)
    assert worst_f < 5e-6 and worst_s < 1e-10, (worst_f, worst_s)
    path = tr.best_path(cand_f, cand_s, inten)
    sel = np.array([cand_f[k][j] for k, j in enumerate(path)])
    sel[~((sel > 0) & (sel < tr.ceiling))] = 0.0                    # parselmouth's selected_array reports voiceless frames as 0
    assert np.array_equal(sel, o["frequency"]), int(np.sum(sel != o["frequency"]))
    voiced = sel[sel > 0]
    assert (np.median(voiced) if len(voiced) else 0.0) == o["median"]


def test_oracle_intensity_window_and_calibration(oracle):
    """Praat's Sound_to_Intensity: the Bessel function behind its Kaiser window against scipy's exact I0 (Abramowitz & Stegun 9.8.1 /
    9.8.2 are good to 2e-7), and the manual's calibration: a stationary sine of amplitude A reads 10 log10(A^2 / 2 / 4e-10) dB."""
    from scipy.special import i0
    xs = np.concatenate([np.linspace(0.0, 3.75, 200), np.linspace(3.75, 21.0, 400)])
    approx = np.array([oracle.bessel_i0_f(float(v)) for v in xs])
    assert np.max(np.abs(approx / i0(xs) - 1.0)) < 2.5e-7
    sr, amp = 16000, 0.5
    t = np.arange(int(0.6 * sr)) / sr
    pcm = np.round(amp * 32768.0 * np.sin(2 * np.pi * 440.0 * t)).astype(np.int16)
    db = oracle.intensity(pcm, sr)
    assert len(db) == int(math.floor((0.6 - 0.064) / 0.008)) + 1
    assert np.max(np.abs(db - 10.0 * math.log10(amp * amp / 2.0 / 4e-10))) < 0.01
    # the same frames from an independent evaluation with the exact Bessel function
    half = int(math.floor(0.032 * sr))
    k = np.arange(-half, half + 1)
    w = i0((2 * np.pi ** 2 + 0.5) * np.sqrt(np.clip(1.0 - (k / sr / 0.032) ** 2, 0.0, None)))
    x = pcm / 32768.0
    n_frames = len(db)
    t_first = 0.5 * 0.6 - 0.5 * n_frames * 0.008 + 0.5 * 0.008
    for f in (0, n_frames // 2, n_frames - 1):
        mid = int(math.floor((t_first + f * 0.008) * sr + 1.0))      # nearest 1-based sample to the frame centre (samples sit at (i - 1/2) / sr)
        lo, hi = max(1, mid - half), min(len(x), mid + half)        # the window is clipped to the sound
        seg = x[lo - 1:hi]
        ww = w[lo - mid + half:hi - mid + half + 1]
        seg = seg - seg.mean()
        assert abs(10.0 * math.log10(np.sum(seg * seg * ww) / np.sum(ww) / 4e-10) - db[f]) < 1e-5
