"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  **PARITY UNPINNED.**

ctypes front-end of ``oracle/prosody_oracle.c`` plus float64/integer restatements of the pydub slicing
rules the reference's closures rely on.  Mirrors, on in-memory mono s16 PCM, the four closures of
``/root/reference/Code/audioPipeline.py``:

    get_part_duration  :314-323      get_median_pitch  :326-335
    get_lufs           :338-358      get_duration      :360-361

The arithmetic behind them lives in packages that are neither vendored in the reference nor installable
here (praat-parselmouth 0.4.5 / Praat 6.1.38, pyloudnorm 0.1.x, pydub 0.25.1) and the reference ships no
golden vectors, so this oracle restates their published algorithms and is pinned only by first-principles
known-answer tests, by torchaudio's independent BS.1770 implementation (loudness) and by a second, independently written
numpy / scipy restatement of the published pitch algorithm (tests/test_oracle_independent.py) — none of which is Praat.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module; the product never does.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
import wave
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libprosody_oracle.so"

PO_OK, PO_ERR_TOO_SHORT, PO_ERR_NO_SAMPLES, PO_ERR_WINDOW, PO_ERR_LUFS_SHORT = 0, 1, 2, 3, 4


class PraatError(RuntimeError):
    """Raised where Praat/parselmouth would throw (slice shorter than 3/floor s, empty extraction)."""


class PoPitchParams(C.Structure):
    _fields_ = [("dt", C.c_double), ("minimumPitch", C.c_double), ("periodsPerWindow", C.c_double),
                ("maxnCandidates", C.c_int32), ("silenceThreshold", C.c_double), ("voicingThreshold", C.c_double),
                ("octaveCost", C.c_double), ("octaveJumpCost", C.c_double), ("voicedUnvoicedCost", C.c_double),
                ("ceiling", C.c_double)]


class PoPitchGeom(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("nsamp_period", "halfnsamp_period", "nsamp_window", "halfnsamp_window",
                                          "minimumLag", "maximumLag", "nsampFFT", "brent_ixmax", "nFrames",
                                          "maxnCandidates")] + \
               [(n, C.c_double) for n in ("t1", "dt", "ceiling", "dt_window")]


class PoCounters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("frames", "candidates", "sinc_evals", "sinc_terms", "brent_iters")]


def build(force: bool = False) -> Path:
    """Compile the C restatement (gcc, seconds). Building the checker is not using it."""
    src = _HERE / "prosody_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        dp, ip, i64p, i16p = (C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int16))
        L.po_pitch_unit_geometry.argtypes = [C.c_int64, C.c_double, C.c_int, C.c_double, C.c_double,
                                             C.POINTER(PoPitchParams), C.POINTER(PoPitchGeom), i64p, i64p, dp]
        L.po_pitch_unit_geometry.restype = C.c_int
        L.po_median_pitch_i16.argtypes = [i16p, C.c_int64, C.c_double, C.c_int, C.c_double, C.c_double,
                                          C.POINTER(PoPitchParams), dp, ip, ip, dp, dp, ip, dp, dp, dp]
        L.po_median_pitch_i16.restype = C.c_int
        L.po_integrated_loudness.argtypes = [dp, C.c_int64, C.c_double, dp]
        L.po_integrated_loudness.restype = C.c_int
        L.po_lufs_i16.argtypes = [i16p, C.c_int64, C.c_int64, C.c_int64, C.c_double, dp]
        L.po_lufs_i16.restype = C.c_int
        L.po_kweight_coeffs.argtypes = [C.c_double, dp, dp, dp, dp]
        L.po_batch_median_pitch.argtypes = [i16p, C.c_int64, i64p, i64p, dp, ip, dp, dp, C.POINTER(PoPitchParams),
                                            dp, ip, ip, ip, C.c_int]
        L.po_batch_lufs.argtypes = [i16p, C.c_int64, i64p, i64p, i64p, i64p, dp, dp, ip, C.c_int]
        L.po_counters_get.argtypes = [C.POINTER(PoCounters)]
        L.po_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def pitch_params(pitch_floor=150.0, pitch_ceiling=600.0, time_step=0.0) -> PoPitchParams:
    """parselmouth ``Sound.to_pitch(time_step, pitch_floor, pitch_ceiling)`` == Praat Sound_to_Pitch ->
    Sound_to_Pitch_ac(dt, floor, 3.0, 15, accurate=false, 0.03, 0.45, 0.01, 0.35, 0.14, ceiling)."""
    return PoPitchParams(time_step, pitch_floor, 3.0, 15, 0.03, 0.45, 0.01, 0.35, 0.14, pitch_ceiling)


# ------------------------------------------------------------------------------------------ WAV decode
def read_wav(path) -> tuple[np.ndarray, int]:
    """Mono 16-bit PCM WAV -> (int16 samples, rate). Praat divides by 32768; pydub keeps the integers."""
    with wave.open(str(path), "rb") as w:
        if w.getsampwidth() != 2 or w.getnchannels() != 1:
            raise ValueError("oracle supports mono 16-bit PCM only (what the reference pipeline produces)")
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.int16)
        return pcm, w.getframerate()


# ------------------------------------------------------------------------------------------ Praat pitch
def _has_t1(t1, preserve_times):
    return 0 if t1 is None else (1 if preserve_times else 2)


def pitch_geometry(file_nx: int, sr: float, t0=0.0, t1=None, params: PoPitchParams | None = None, preserve_times=True):
    params = params or pitch_params()
    g = PoPitchGeom()
    ix1, nx, x1 = C.c_int64(), C.c_int64(), C.c_double()
    st = lib().po_pitch_unit_geometry(file_nx, float(sr), _has_t1(t1, preserve_times), float(t0), float(t1 or 0.0),
                                      C.byref(params), C.byref(g), C.byref(ix1), C.byref(nx), C.byref(x1))
    return st, g, ix1.value, nx.value, x1.value


def pitch_track(pcm: np.ndarray, sr: float, t0=0.0, t1=None, params: PoPitchParams | None = None,
                want_candidates=False, preserve_times=True) -> dict:
    """Full ``to_pitch`` on (a slice of) an in-memory file. Raises PraatError where Praat throws."""
    params = params or pitch_params()
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    st, g, ix1, nx, x1 = pitch_geometry(len(pcm), sr, t0, t1, params, preserve_times)
    if st != PO_OK:
        raise PraatError(f"Praat would throw (status {st}) for slice t0={t0} t1={t1}")
    nF, maxC = g.nFrames, g.maxnCandidates
    sel_f = np.zeros(nF); sel_s = np.zeros(nF); ncand = np.zeros(nF, np.int32); inten = np.zeros(nF)
    pre_f = np.zeros((nF, maxC)) if want_candidates else None
    pre_s = np.zeros((nF, maxC)) if want_candidates else None
    med, nv, nfr = C.c_double(), C.c_int32(), C.c_int32()
    st = lib().po_median_pitch_i16(_p(pcm, C.c_int16), len(pcm), float(sr), _has_t1(t1, preserve_times), float(t0),
                                   float(t1 or 0.0), C.byref(params), C.byref(med), C.byref(nv), C.byref(nfr),
                                   _p(sel_f, C.c_double), _p(sel_s, C.c_double), _p(ncand, C.c_int32),
                                   _p(inten, C.c_double),
                                   _p(pre_f, C.c_double) if want_candidates else None,
                                   _p(pre_s, C.c_double) if want_candidates else None)
    if st != PO_OK:
        raise PraatError(f"Praat would throw (status {st})")
    return dict(frequency=sel_f, strength=sel_s, ncand=ncand, intensity=inten, median=med.value,
                n_voiced=nv.value, n_frames=nfr.value, geom=g, ix1=ix1, nx=nx, x1=x1,
                t1=g.t1, dt=g.dt, pre_f=pre_f, pre_s=pre_s)


def median_pitch(pcm, sr, t0=0.0, t1=None, pitch_floor=150.0, pitch_ceiling=600.0) -> float:
    """≙ get_median_pitch (audioPipeline.py:326-335)."""
    return pitch_track(pcm, sr, t0, t1, pitch_params(pitch_floor, pitch_ceiling))["median"]


# ------------------------------------------------------------------------------------------ pydub slicing
def pydub_len_ms(n_frames: int, rate: int) -> int:
    """len(AudioSegment) = round(1000 * frames / rate) (python round: half-to-even)."""
    return round(1000 * (n_frames / rate))


def pydub_slice(n_frames: int, rate: int, t0: float, t1: float):
    """``audio[int(t0*1000):int(t1*1000)]`` -> (a, b, npad): real samples [a,b) then npad zeros.
    Raises OverflowError-like ValueError where pydub raises TooManyMissingFrames."""
    L = pydub_len_ms(n_frames, rate)
    s_ms, e_ms = int(t0 * 1000), int(t1 * 1000)
    if s_ms < 0 or e_ms < 0:
        raise ValueError("negative slice positions are outside the reference's usage")
    s_ms, e_ms = min(s_ms, L), min(e_ms, L)
    per_ms = rate / 1000.0
    sf, ef = int(s_ms * per_ms), int(e_ms * per_ms)
    a = min(sf, n_frames)
    b = max(a, min(ef, n_frames))
    expected = max(ef - sf, 0)
    missing = expected - (b - a)
    npad = 0
    if missing:
        if missing > 2 * per_ms:
            raise ValueError("pydub TooManyMissingFrames")
        npad = missing if (b - a) > 0 else 0   # silence is built from data[:frame_width]; empty data -> no padding
    return a, b, npad


def part_duration(n_frames: int, rate: int, t0=0.0, t1=None) -> float:
    """≙ get_part_duration (audioPipeline.py:314-323)."""
    if t1 is not None:
        a, b, npad = pydub_slice(n_frames, rate, t0, t1)
        return ((b - a + npad) / rate) or 1e-4
    return (n_frames / rate) or 1e-4


def duration(n_frames: int, rate: int) -> float:
    """≙ get_duration (audioPipeline.py:360-361)."""
    return (n_frames / rate) or 1e-4


# ------------------------------------------------------------------------------------------ loudness
def integrated_loudness(data: np.ndarray, rate: float) -> float:
    """pyloudnorm ``Meter(rate).integrated_loudness(data)`` for mono float64. Raises ValueError if too short."""
    data = np.ascontiguousarray(data, dtype=np.float64)
    out = C.c_double()
    st = lib().po_integrated_loudness(_p(data, C.c_double), len(data), float(rate), C.byref(out))
    if st == PO_ERR_LUFS_SHORT:
        raise ValueError("Audio must have length greater than the block size.")
    if st != PO_OK:
        raise RuntimeError(f"oracle loudness failed: {st}")
    return out.value


def lufs_resolve(n_frames: int, rate: int, meter_rate: float, t0=0.0, t1=None):
    """The control flow of get_lufs (audioPipeline.py:338-358) as a resolved sample range.
    Returns (a, b, npad, used_fallback) or raises ValueError when even the whole file is < 0.4 s."""
    if t1 is not None:
        a, b, npad = pydub_slice(n_frames, rate, t0, t1)
    else:
        a, b, npad = 0, n_frames, 0
    fallback = False
    if (b - a + npad) == 0:                       # empty slice -> whole file (:345-348)
        a, b, npad, fallback = 0, n_frames, 0, True
    if (b - a + npad) < 0.4 * meter_rate:         # ValueError -> whole-file loudness (:353-358)
        a, b, npad, fallback = 0, n_frames, 0, True
        if n_frames < 0.4 * meter_rate:
            raise ValueError("Audio must have length greater than the block size.")
    return a, b, npad, fallback


def lufs(pcm: np.ndarray, rate: int, meter_rate: float, t0=0.0, t1=None) -> float:
    """≙ get_lufs (audioPipeline.py:338-358) with ``meter = pyln.Meter(meter_rate)``."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    a, b, npad, _ = lufs_resolve(len(pcm), rate, meter_rate, t0, t1)
    out = C.c_double()
    st = lib().po_lufs_i16(_p(pcm, C.c_int16), a, b, npad, float(meter_rate), C.byref(out))
    if st != PO_OK:
        raise RuntimeError(f"oracle lufs failed: {st}")
    return out.value


def kweight_coeffs(rate: float):
    bs, as_, bh, ah = (np.zeros(3) for _ in range(4))
    lib().po_kweight_coeffs(float(rate), _p(bs, C.c_double), _p(as_, C.c_double), _p(bh, C.c_double), _p(ah, C.c_double))
    return bs, as_, bh, ah


# ------------------------------------------------------------------------------------------ batch (CPU baseline)
def batch_median_pitch(pcm_cat, file_off, file_nx, sr, has_t1, t0, t1, params=None, n_threads=0):
    params = params or pitch_params()
    n = len(file_off)
    med = np.zeros(n); nv = np.zeros(n, np.int32); nf = np.zeros(n, np.int32); st = np.zeros(n, np.int32)
    args = [np.ascontiguousarray(file_off, np.int64), np.ascontiguousarray(file_nx, np.int64),
            np.ascontiguousarray(sr, np.float64), np.ascontiguousarray(has_t1, np.int32),
            np.ascontiguousarray(t0, np.float64), np.ascontiguousarray(t1, np.float64)]
    pcm_cat = np.ascontiguousarray(pcm_cat, np.int16)
    lib().po_batch_median_pitch(_p(pcm_cat, C.c_int16), n, _p(args[0], C.c_int64), _p(args[1], C.c_int64),
                                _p(args[2], C.c_double), _p(args[3], C.c_int32), _p(args[4], C.c_double),
                                _p(args[5], C.c_double), C.byref(params), _p(med, C.c_double), _p(nv, C.c_int32),
                                _p(nf, C.c_int32), _p(st, C.c_int32), int(n_threads))
    return med, nv, nf, st


def batch_lufs(pcm_cat, file_off, a, b, npad, meter_rate, n_threads=0):
    n = len(file_off)
    out = np.zeros(n); st = np.zeros(n, np.int32)
    args = [np.ascontiguousarray(x, np.int64) for x in (file_off, a, b, npad)]
    mr = np.ascontiguousarray(meter_rate, np.float64)
    pcm_cat = np.ascontiguousarray(pcm_cat, np.int16)
    lib().po_batch_lufs(_p(pcm_cat, C.c_int16), n, *[_p(x, C.c_int64) for x in args], _p(mr, C.c_double),
                        _p(out, C.c_double), _p(st, C.c_int32), int(n_threads))
    return out, st


def counters() -> dict:
    c = PoCounters()
    lib().po_counters_get(C.byref(c))
    return {n: getattr(c, n) for n, _ in PoCounters._fields_}


def max_threads() -> int:
    return lib().po_max_threads()


# ------------------------------------------------------------------------------------------ "next" rows
def bessel_i0_f(x: float) -> float:
    """Praat NUMbessel_i0_f (melder/NUMspecfunc.cpp): Abramowitz & Stegun 9.8.1 / 9.8.2."""
    x = abs(x)
    if x < 3.75:
        t = (x / 3.75) ** 2
        return 1.0 + t * (3.5156229 + t * (3.0899424 + t * (1.2067492 + t * (0.2659732 + t * (0.0360768 + t * 0.0045813)))))
    t = 3.75 / x
    return math.exp(x) / math.sqrt(x) * (0.39894228 + t * (0.01328592 + t * (0.00225319 + t * (-0.00157565 + t * (
        0.00916281 + t * (-0.02057706 + t * (0.02635537 + t * (-0.01647633 + t * 0.00392377))))))))


def intensity_geometry(nx: int, sr: float, minimum_pitch=100.0, time_step=0.0):
    """Praat Sound_to_Intensity geometry -> (status, n_frames, t_first, dt, half_window_samples)."""
    dx, x1 = 1.0 / sr, 0.5 / sr
    dt = time_step if time_step > 0 else 0.8 / minimum_pitch
    window = 6.4 / minimum_pitch
    half = int(math.floor(0.5 * window / dx))
    duration = dx * nx
    if window > duration:
        return PO_ERR_TOO_SHORT, 0, 0.0, dt, half
    n_frames = int(math.floor((duration - window) / dt)) + 1
    mid = x1 - 0.5 * dx + 0.5 * duration
    t_first = mid - 0.5 * (n_frames * dt) + 0.5 * dt
    return PO_OK, n_frames, t_first, dt, half


def intensity(pcm: np.ndarray, sr: float, minimum_pitch=100.0, time_step=0.0, subtract_mean=True) -> np.ndarray:
    """≙ parselmouth Sound.to_intensity().values[0] (Praat fon/Sound_to_Intensity.cpp) for a mono s16 file, dB re 4e-10."""
    x = np.asarray(pcm, np.float64) / 32768.0
    nx = len(x)
    st, n_frames, t_first, dt, half = intensity_geometry(nx, sr, minimum_pitch, time_step)
    if st != PO_OK:
        raise PraatError("Praat: sound shorter than the intensity window")
    dx, x1 = 1.0 / sr, 0.5 / sr
    half_dur = 0.5 * 6.4 / minimum_pitch
    i = np.arange(-half, half + 1)
    xx = i * dx / half_dur
    root = np.sqrt(np.clip(1.0 - xx * xx, 0.0, None))
    window = np.array([bessel_i0_f((2.0 * math.pi * math.pi + 0.5) * r) for r in root])
    out = np.empty(n_frames)
    for f in range(n_frames):
        t = t_first + f * dt
        mid = int(math.floor((t - x1) / dx + 1.0 + 0.5))
        left, right = max(1, mid - half), min(nx, mid + half)
        a = x[left - 1:right].copy()
        w = window[left - mid + half:right - mid + half + 1]
        if subtract_mean:
            a -= a.sum() / (right - left + 1)
        inten = float((a * a * w).sum() / w.sum()) / 4e-10
        out[f] = -300.0 if inten < 1e-30 else 10.0 * math.log10(inten)
    return out


def legacy_loudness(pcm: np.ndarray, rate: int, start: float, end: float) -> float:
    """≙ _calculate_loudness (Code/Pipeline/compute_loudness_adjustments.py:8-25): RMS dB of the slice with the
    reference's int16 wrap-around in ``samples ** 2``. pydub slice positions are ``start*1000`` / ``end*1000`` (float ms)."""
    n = len(pcm)
    L = pydub_len_ms(n, rate)
    s_ms, e_ms = min(start * 1000, L), min(end * 1000, L)
    per_ms = rate / 1000.0
    sf, ef = int(s_ms * per_ms), int(e_ms * per_ms)
    a = min(sf, n); b = max(a, min(ef, n))
    seg = np.asarray(pcm[a:b], np.int16)
    missing = max(ef - sf, 0) - (b - a)               # pydub pads an overshoot of < 2 ms with zeros, else raises
    if missing:
        if missing > 2 * per_ms:
            return float("nan")
        if len(seg):
            seg = np.concatenate([seg, np.zeros(missing, np.int16)])
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        S = seg ** 2                                   # int16: wraps modulo 2^16 exactly like the reference
        rms = np.sqrt(np.abs(np.mean(S))) if len(S) else float("nan")
        return float(20 * np.log10(rms))


def legacy_pitch_segment(pcm: np.ndarray, sr: float, start: float, end: float) -> float:
    """≙ calculate_pitch_segment (Code/Pipeline/compute_pitch_adjustments.py:167-208): extract_part (times not preserved),
    first pitch floor of [75, 100, 150, 200] that yields a voiced frame, geometric mean of the voiced frequencies."""
    import statistics
    total = len(pcm) / sr
    if start >= end or start < 0 or end > total:
        return 0
    for floor in (75, 100, 150, 200):
        try:
            tr = pitch_track(pcm, sr, start, end, pitch_params(float(floor), 600.0), preserve_times=False)
        except PraatError:
            continue
        v = tr["frequency"][tr["frequency"] != 0]
        if len(v):
            return statistics.geometric_mean(v)
    return 0


# ------------------------------------------------------------------ pydub.silence (pydub 0.25.1 silence.py), mono s16
def _audioop_rms(seg: np.ndarray) -> int:
    """audioop.rms: ``(unsigned int) sqrt(sum_squares / n)`` with the sum accumulated in a C double (exact here)."""
    if len(seg) == 0:
        return 0
    s = int(np.sum(seg.astype(np.int64) ** 2))
    return int(math.sqrt(float(s) / float(len(seg))))


def _pydub_slice_samples(pcm: np.ndarray, rate: int, s_ms, e_ms) -> np.ndarray:
    """AudioSegment.__getitem__(slice) on mono s16 data: clip to len(), int(ms * rate/1000) frames, pad < 2 ms of zeros."""
    n = len(pcm)
    L = pydub_len_ms(n, rate)
    s_ms, e_ms = min(s_ms, L), min(e_ms, L)
    per_ms = rate / 1000.0
    sf, ef = int(s_ms * per_ms), int(e_ms * per_ms)
    a = min(sf, n); b = max(a, min(ef, n))
    seg = np.asarray(pcm[a:b], np.int16)
    missing = max(ef - sf, 0) - (b - a)
    if missing:
        if missing > 2 * per_ms:
            raise ValueError("TooManyMissingFrames")
        if len(seg):
            seg = np.concatenate([seg, np.zeros(missing, np.int16)])
    return seg


def detect_silence(pcm: np.ndarray, rate: int, min_silence_len=1000, silence_thresh=-16, seek_step=1, naive=False):
    """≙ pydub.silence.detect_silence.  naive=True slices and squares every window like pydub does (small inputs only);
    otherwise window sums come from an exact int64 prefix sum (same integers, so the same decisions)."""
    n = len(pcm)
    seg_len = pydub_len_ms(n, rate)
    if seg_len < min_silence_len:
        return []
    thresh = (10 ** (silence_thresh / 20.0)) * 32768          # db_to_float(silence_thresh) * max_possible_amplitude
    last_slice_start = seg_len - min_silence_len
    starts = list(range(0, last_slice_start + 1, seek_step))
    if last_slice_start % seek_step:
        starts.append(last_slice_start)
    per_ms = rate / 1000.0
    if not naive:
        P = np.concatenate([[0], np.cumsum(np.asarray(pcm, np.int64) ** 2)])
    silence_starts = []
    for i in starts:
        if naive:
            rms = _audioop_rms(_pydub_slice_samples(pcm, rate, i, i + min_silence_len))
        else:
            sf, ef = int(i * per_ms), int(min(i + min_silence_len, seg_len) * per_ms)
            a = min(sf, n); b = max(a, min(ef, n))
            cnt = (ef - sf) if b > a else 0
            rms = int(math.sqrt(float(int(P[b] - P[a])) / float(cnt))) if cnt else 0
        if rms <= thresh:
            silence_starts.append(i)
    if not silence_starts:
        return []
    silent_ranges = []
    prev_i = silence_starts.pop(0)
    current_range_start = prev_i
    for silence_start_i in silence_starts:
        continuous = (silence_start_i == prev_i + seek_step)
        silence_has_gap = silence_start_i > (prev_i + min_silence_len)
        if not continuous and silence_has_gap:
            silent_ranges.append([current_range_start, prev_i + min_silence_len])
            current_range_start = silence_start_i
        prev_i = silence_start_i
    silent_ranges.append([current_range_start, prev_i + min_silence_len])
    return silent_ranges


def detect_nonsilent(pcm, rate, min_silence_len=1000, silence_thresh=-16, seek_step=1, naive=False):
    """≙ pydub.silence.detect_nonsilent."""
    silent_ranges = detect_silence(pcm, rate, min_silence_len, silence_thresh, seek_step, naive)
    len_seg = pydub_len_ms(len(pcm), rate)
    if not silent_ranges:
        return [[0, len_seg]]
    if silent_ranges[0][0] == 0 and silent_ranges[0][1] == len_seg:
        return []
    prev_end_i = 0
    nonsilent_ranges = []
    for start_i, end_i in silent_ranges:
        nonsilent_ranges.append([prev_end_i, start_i])
        prev_end_i = end_i
    if end_i != len_seg:
        nonsilent_ranges.append([prev_end_i, len_seg])
    if nonsilent_ranges[0] == [0, 0]:
        nonsilent_ranges.pop(0)
    return nonsilent_ranges


def split_on_silence(pcm, rate, min_silence_len=1000, silence_thresh=-16, keep_silence=100, seek_step=1, naive=False):
    """≙ pydub.silence.split_on_silence (Code/Preprocessing/preprocess_audio.py:41-46 calls it with 1000 ms / -50 dBFS / 300 ms).
    Returns the [start_ms, end_ms) ranges that the final ``audio_segment[max(start,0):min(end,len)]`` slices use."""
    len_seg = pydub_len_ms(len(pcm), rate)
    if isinstance(keep_silence, bool):
        keep_silence = len_seg if keep_silence else 0
    output_ranges = [[s - keep_silence, e + keep_silence]
                     for s, e in detect_nonsilent(pcm, rate, min_silence_len, silence_thresh, seek_step, naive)]
    for k in range(len(output_ranges) - 1):
        range_i, range_ii = output_ranges[k], output_ranges[k + 1]
        last_end, next_start = range_i[1], range_ii[0]
        if next_start < last_end:
            range_i[1] = (last_end + next_start) // 2
            range_ii[0] = range_i[1]
    return [(max(s, 0), min(e, len_seg)) for s, e in output_ranges]
