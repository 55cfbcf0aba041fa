"""Third-party pin for the loudness oracle: torchaudio.functional.loudness is an independent BS.1770-4
implementation shipped in this image (K-weighting = RBJ high shelf 4 dB / 1500 Hz / Q 1/sqrt 2 then RBJ high-pass 38 Hz /
Q 0.5, 400 ms blocks at 75 % overlap, -70 LUFS absolute and -10 LU relative gates — the same forms pyloudnorm's meter
uses, which is why the reference's numbers come out of either).  The real pyloudnorm is not installable here
(profiles/r02_pip_real_packages.log); this is the closest independent implementation the oracle can be held against,
and it is not ours.  Praat's pitch tracker has no such stand-in: that part of the oracle stays pinned by
first-principles known-answer tests only (tests/test_oracle_kat.py)."""
import math

import numpy as np
import pytest

torch = pytest.importorskip("torch")
taF = pytest.importorskip("torchaudio.functional")


def _signal(sr, n, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    x = 0.3 * np.sin(2 * np.pi * 220 * t) * (0.5 + 0.5 * np.sin(2 * np.pi * 0.7 * t)) + 0.02 * rng.standard_normal(n)
    x[: n // 5] *= 1e-4                       # a stretch the absolute gate removes
    x[n // 2: n // 2 + n // 7] *= 0.05        # and one the relative gate removes
    return x


@pytest.mark.parametrize("sr", [16000, 22050, 24000, 44100, 48000])
def test_integrated_loudness_matches_torchaudio(oracle, sr):
    gate = int(round(0.4 * sr)); step = int(round(gate * 0.25))
    for k, seed in ((3, 1), (37, 2), (120, 3)):
        n = gate + step * k                   # both block rules (pyloudnorm's rounded count, torchaudio's unfold) agree here
        x = _signal(sr, n, seed)
        ours = oracle.integrated_loudness(x, float(sr))
        ref = float(taF.loudness(torch.from_numpy(x)[None, :].double(), sr))
        assert abs(ours - ref) < 1e-9, (sr, k, ours, ref)


@pytest.mark.parametrize("rate", [16000.0, 24000.0, 44100.0])
def test_kweighting_filters_match_torchaudio_biquads(oracle, rate):
    """Impulse responses of the two K-weighting stages, coefficient set by coefficient set."""
    b1, a1, b2, a2 = oracle.kweight_coeffs(rate)
    imp = torch.zeros(1, 4096, dtype=torch.float64); imp[0, 0] = 1.0
    from scipy.signal import lfilter
    h1 = lfilter(b1, a1, imp[0].numpy()); h2 = lfilter(b2, a2, imp[0].numpy())
    # torchaudio clamps filter outputs to [-1, 1] by default: scale the impulse down to stay clear of it
    t1 = taF.treble_biquad(imp * 0.25, int(rate), 4.0, 1500.0, 1 / math.sqrt(2))[0].numpy() * 4.0
    t2 = taF.highpass_biquad(imp * 0.25, int(rate), 38.0, 0.5)[0].numpy() * 4.0
    assert np.max(np.abs(h1 - t1)) < 1e-12
    assert np.max(np.abs(h2 - t2)) < 1e-12


def test_measured_clip_loudness_through_the_reference_closure(oracle):
    """get_lufs control flow (pydub slice -> peak normalisation -> meter) with the meter swapped for torchaudio's:
    the oracle's closure-level value is reproduced on a slice whose length both block rules agree on."""
    sr = 16000
    x = (_signal(sr, sr * 3, 9) * 20000).astype(np.int16)
    t0, t1 = 0.5, 2.5                                     # 2.0 s = 6400 + 16 * 1600 samples
    ours = oracle.lufs(x, sr, float(sr), t0, t1)
    a, b = int(t0 * 1000) * sr // 1000, int(t1 * 1000) * sr // 1000
    seg = x[a:b].astype(np.float64)
    seg = seg / (np.max(np.abs(seg)) or 1.0)              # audioPipeline.py:349-350
    # torchaudio clamps its biquad outputs to [-1, 1]; the shelf lifts a peak-normalised signal above 1, so measure it
    # 12 dB down and add the 20 log10(4) back (loudness is exactly linear in gain)
    ref = float(taF.loudness(torch.from_numpy(seg * 0.25)[None, :], sr)) + 20.0 * math.log10(4.0)
    assert abs(ours - ref) < 1e-9
