"""Pins the oracle against the REAL third-party packages whenever they are importable (mitigation 1).

praat-parselmouth, pyloudnorm and pydub are not installable in the build container (no network, not in the wheelhouse),
so these tests skip there and the oracle stays "parity unpinned"; on a machine that has them they turn the oracle's
restatement of Praat / pyloudnorm / pydub into a checked one, on the same seeded inputs the GPU parity tests use."""
import wave

import numpy as np
import pytest

from conftest import speechlike

# attempted in BOTH suites: the CPU one here, and (gpu-marked twin) on the GPU box, whichever machine happens to have the packages;
# the install attempt on the GPU box is logged in profiles/r02_pip_real_packages.log (no network, not in /opt/wheelhouse)
BOTH = pytest.mark.parametrize("where", ["build-box", pytest.param("gpu-box", marks=pytest.mark.gpu)])


def _wav(path, pcm, sr):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(sr)
        w.writeframes(np.ascontiguousarray(pcm, "<i2").tobytes())
    return str(path)


@BOTH
def test_pitch_against_parselmouth(tmp_path, oracle, where):
    parselmouth = pytest.importorskip("parselmouth")
    for sr, floor in ((16000, 75.0), (44100, 150.0), (24000, 150.0)):
        x = speechlike(1, 2.0, sr, seed=sr // 100)[0]
        snd = parselmouth.Sound(_wav(tmp_path / f"a{sr}.wav", x, sr))
        for t0, t1 in ((0.0, None), (0.3, 1.7)):
            part = snd if t1 is None else snd.extract_part(from_time=t0, to_time=t1, preserve_times=True)
            ref = part.to_pitch(pitch_floor=floor, pitch_ceiling=600.0)
            f_ref = ref.selected_array["frequency"]; s_ref = ref.selected_array["strength"]
            o = oracle.pitch_track(x, sr, t0, t1, params=oracle.pitch_params(floor, 600.0))
            assert o["n_frames"] == len(f_ref)
            assert abs(o["t1"] - ref.xs()[0]) < 1e-9
            assert np.array_equal(o["frequency"] > 0, f_ref > 0)
            v = f_ref > 0
            assert np.max(np.abs(o["frequency"][v] - f_ref[v]) / f_ref[v]) < 1e-6
            assert np.max(np.abs(o["strength"] - s_ref)) < 1e-6
        inten = snd.to_intensity().values[0]
        assert np.max(np.abs(oracle.intensity(x, sr) - inten)) < 1e-6


@BOTH
def test_loudness_against_pyloudnorm(oracle, where):
    pyln = pytest.importorskip("pyloudnorm")
    for sr in (16000, 24000, 44100):
        x = speechlike(1, 3.0, sr, seed=sr // 50)[0]
        data = x.astype(np.float64)
        data = data / np.max(np.abs(data))
        for mr in (sr, 44100 if sr != 44100 else 16000):
            assert abs(oracle.lufs(x, sr, float(mr)) - pyln.Meter(mr).integrated_loudness(data)) < 1e-9


@BOTH
def test_slicing_and_silence_against_pydub(tmp_path, oracle, where):
    pydub = pytest.importorskip("pydub")
    from pydub.silence import split_on_silence
    from test_emu_parity import _gappy
    for sr in (16000, 22050):
        x = _gappy(sr, 6.0, ((0.4, 1.7), (2.5, 3.9), (5.2, 6.1)), sr)
        seg = pydub.AudioSegment.from_file(_wav(tmp_path / f"s{sr}.wav", x, sr))
        assert len(seg) == oracle.pydub_len_ms(len(x), sr)
        for a, b in ((0, 1000), (1234, 5999), (5990, 7000)):
            got = np.array(seg[a:b].get_array_of_samples(), np.int16)
            assert np.array_equal(got, oracle._pydub_slice_samples(x, sr, a, b))
        parts = split_on_silence(seg, min_silence_len=1000, silence_thresh=-50, keep_silence=300)
        ref = oracle.split_on_silence(x, sr, 1000, -50, 300)
        assert [len(p) for p in parts] == [e - s for s, e in ref]
        for p, (s, e) in zip(parts, ref):
            assert np.array_equal(np.array(p.get_array_of_samples(), np.int16), oracle._pydub_slice_samples(x, sr, s, e))
