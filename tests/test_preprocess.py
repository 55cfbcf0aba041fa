"""Silence-segmentation drop-ins (preprocess.py) end to end on files, through the SIMT-emulated kernels (CPU) — the GPU
parity of the same call is in test_gpu_parity.py."""
import wave

import numpy as np

from test_emu_parity import _gappy


def _write_wav(path, pcm, sr):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(sr)
        w.writeframes(np.ascontiguousarray(pcm, "<i2").tobytes())


def test_main_writes_the_reference_segments(tmp_path, emu_lib, oracle):
    import prosody_b200 as pb
    from prosody_b200 import preprocess
    ex = pb.Extractor(0, lib=emu_lib)
    sr = 16000
    x = _gappy(sr, 3.0, ((0.5, 0.9), (1.6, 2.1), (2.8, 3.1)), 21)
    _write_wav(tmp_path / "brute.wav", x, sr)
    stats = preprocess.main(tmp_path / "brute.wav", tmp_path / "audio", min_silence_len=300, silence_thresh=-50, keep_silence=100, extractor=ex)
    ref = oracle.split_on_silence(x, sr, 300, -50, 100)
    assert stats["nombre_segments"] == len(ref) >= 2
    assert abs(stats["duree_totale"] - sum(e - s for s, e in ref) / 1000) < 1e-9
    for i, (s, e) in enumerate(ref):
        got, rate = oracle.read_wav(tmp_path / "audio" / f"segment_ph{i+1}.wav")
        assert rate == sr and np.array_equal(got, oracle._pydub_slice_samples(x, sr, s, e))
    many = preprocess.segment_audio_files([tmp_path / "brute.wav", tmp_path / "audio" / "segment_ph1.wav"], 300, -50, 100, extractor=ex)
    assert len(many) == 2 and len(many[0]) == len(ref)
    ex.close()
