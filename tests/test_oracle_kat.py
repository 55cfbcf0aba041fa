"""The oracle is unpinned against the real parselmouth / pyloudnorm / pydub (none installable here, the reference ships
no golden vectors): these known-answer tests pin it from first principles instead."""
import math

import numpy as np
import pytest


def tone(f, sr, dur, amp=0.5):
    t = np.arange(int(sr * dur)) / sr
    return (amp * 32767 * np.sin(2 * np.pi * f * t)).astype(np.int16)


@pytest.mark.parametrize("sr", [16000, 22050, 44100])
@pytest.mark.parametrize("f", [110.0, 220.0, 440.0])
def test_pure_tone_f0(oracle, sr, f):
    o = oracle.pitch_track(tone(f, sr, 0.6), sr, params=oracle.pitch_params(75.0, 600.0))
    assert o["n_voiced"] == o["n_frames"]
    assert abs(o["median"] - f) / f < 2e-4
    assert np.all(o["strength"] > 0.99)


def test_silence_is_unvoiced(oracle):
    o = oracle.pitch_track(np.zeros(16000, np.int16), 16000)
    assert o["n_voiced"] == 0 and o["median"] == 0.0 and np.all(o["frequency"] == 0)


def test_white_noise_mostly_unvoiced(oracle):
    x = (np.random.default_rng(0).standard_normal(32000) * 3000).astype(np.int16)
    o = oracle.pitch_track(x, 16000)
    assert o["n_voiced"] < 0.1 * o["n_frames"]


def test_frame_grid_formulas(oracle):
    """Praat's geometry at the BASELINE configs (computed by hand from the formulas)."""
    st, g, ix1, nx, x1 = oracle.pitch_geometry(80000, 16000.0, params=oracle.pitch_params(75.0, 600.0))
    assert (st, g.nsamp_window, g.maximumLag, g.nsampFFT, g.brent_ixmax, g.nFrames, g.maxnCandidates) == (0, 638, 214, 1024, 319, 497, 15)
    assert abs(g.dt - 0.01) < 1e-15
    st, g, *_ = oracle.pitch_geometry(int(24000 * 20), 24000.0, params=oracle.pitch_params(75.0, 600.0))
    assert (g.nsamp_window, g.maximumLag, g.nsampFFT, g.nFrames) == (958, 321, 2048, 1997)
    st, g, *_ = oracle.pitch_geometry(938543, 44100.0, params=oracle.pitch_params(150.0, 600.0))
    assert (g.nsamp_window, g.maximumLag, g.nsampFFT, g.nFrames) == (880, 295, 2048, 4253)


def test_too_short_slice_raises_like_praat(oracle):
    x = tone(200.0, 16000, 1.0)
    with pytest.raises(oracle.PraatError):
        oracle.pitch_track(x, 16000, 0.5, 0.515)           # 15 ms < 3/150 s
    assert oracle.pitch_track(x, 16000, 0.5, 0.53)["n_frames"] >= 1


def test_extract_part_zero_fills_beyond_file(oracle):
    x = tone(200.0, 16000, 1.0)
    o = oracle.pitch_track(x, 16000, 0.9, 1.3)
    assert o["nx"] == 6400 and o["n_frames"] > 0
    assert o["frequency"][-1] == 0.0                        # the tail is digital silence


def test_lufs_bs1770_calibration(oracle):
    """BS.1770: a 997 Hz sine at 0 dBFS reads -3.01 LUFS (K-weighting is ~0 dB there up to the shelf's +0.7 dB skirt)."""
    sr = 48000
    t = np.arange(sr * 3) / sr
    x = np.sin(2 * np.pi * 997.0 * t)
    l = oracle.integrated_loudness(x, sr)
    assert abs(l - (-3.01)) < 0.05
    # 20 dB down reads 20 LU lower (linearity)
    assert abs(oracle.integrated_loudness(0.1 * x, sr) - (l - 20.0)) < 1e-9


def test_lufs_kweighting_matches_scipy(oracle):
    sig = pytest.importorskip("scipy.signal")
    bs, as_, bh, ah = oracle.kweight_coeffs(44100.0)
    x = np.random.default_rng(1).standard_normal(44100)
    y = sig.lfilter(bh, ah, sig.lfilter(bs, as_, x))
    # block 0 energy by hand -> single-block loudness
    z = np.sum(y[:int(0.4 * 44100)] ** 2) / (0.4 * 44100)
    l = oracle.integrated_loudness(x[:int(0.4 * 44100)], 44100.0)
    assert abs(l - (-0.691 + 10 * math.log10(z))) < 1e-9


def test_lufs_short_raises_and_reference_fallbacks(oracle):
    x = tone(300.0, 16000, 1.0)
    with pytest.raises(ValueError):
        oracle.integrated_loudness(x[:6000].astype(float), 16000)
    whole = oracle.lufs(x, 16000, 16000.0)
    assert oracle.lufs(x, 16000, 16000.0, 0.1, 0.3) == whole        # < 0.4 s -> whole file
    assert oracle.lufs(x, 16000, 16000.0, 5.0, 6.0) == whole        # empty slice -> whole file
    assert oracle.lufs(np.zeros(16000, np.int16), 16000, 16000.0) == -math.inf


def test_pydub_ms_slicing_identities(oracle):
    # int(t*1000) after ms/1000 loses a millisecond for some values
    lost = [ms for ms in range(0, 5000) if int((ms / 1000) * 1000) != ms]
    assert 1001 in lost and len(lost) > 0
    a, b, npad = oracle.pydub_slice(16000, 16000, 0.25, 0.75)
    assert (a, b, npad) == (4000, 12000, 0)
    a, b, npad = oracle.pydub_slice(16000, 16000, 0.9, 1.5)          # clipped to len(audio)
    assert (a, b, npad) == (14400, 16000, 0)
    assert oracle.part_duration(16000, 16000, 2.0, 3.0) == 1e-4      # "or 1e-4"
    assert oracle.pydub_len_ms(44100 * 3 + 22, 44100) == 3000
