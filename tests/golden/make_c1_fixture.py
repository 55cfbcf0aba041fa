"""BASELINE config 1 as a fixture: the reference's ten example clips
(Data/voice/records/audio/segment_ph{2..11}.wav: 44.1 kHz mono s16, 161.85 s — copied byte for byte into
tests/golden/clips/, they are data, not source) plus what the repo does NOT ship for them and the step needs:

  * one TextGrid per clip — a deterministic word grid (seed 0 + clip number: words U(0.12, 0.55) s, a pause after a word
    with p = 0.18 lasting U(0.05, 0.9) s, a sentence end every 8-20 words), written in the long format the Whisper aligner
    step produces (Code/Aligners/use_whisper_timestamped.py:330-395);
  * the "raw synthesis" twin of every clip — the same clip resampled to 16 kHz and time-scaled x0.93
    (scipy.signal.resample_poly(x, 248, 735): 44100 * 248 / 735 = 14880 = 0.93 * 16000 samples per original second).

build_voice(root) lays them out the way AudioPipeline expects (Code/audioPipeline.py:88-107):
    root/voice/records/audio/segment_ph*.wav, root/voice/records_raw/audio/segment_ph*.wav,
    root/voice/records/WhisperTS_textgrid_files/segment_ph*.TextGrid
Run as a script it rebuilds the layout under the given directory and prints a manifest (sha1 of every generated file)."""
from __future__ import annotations

import hashlib
import re
import shutil
import sys
import wave
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
CLIPS = HERE / "clips"


def read_wav(path):
    with wave.open(str(path), "rb") as w:
        assert w.getnchannels() == 1 and w.getsampwidth() == 2
        return np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.int16), w.getframerate()


def write_wav(path, pcm, sr):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(sr); w.writeframes(np.asarray(pcm, "<i2").tobytes())


def clip_paths():
    return sorted(CLIPS.glob("segment_ph*.wav"), key=lambda p: int(re.search(r"segment_ph(\d+)", p.stem).group(1)))


def raw_twin(pcm):
    from scipy.signal import resample_poly
    y = resample_poly(pcm.astype(np.float64), 248, 735)
    return np.clip(np.rint(y), -32768, 32767).astype(np.int16)


def word_grid(name, dur_s):
    from prosody_b200 import synth
    k = int(re.search(r"(\d+)", name).group(1))
    return synth.make_word_grid(1, dur_s, seed=k)[0]          # seed 0 + clip number


def build_voice(root, voice="records"):
    """-> dict(voice_dir, raw_audio_dir, textgrid_dir, segments=[(name, nat_pcm, nat_sr, syn_pcm, syn_sr, grid)])"""
    from prosody_b200 import textgrid as TG
    root = Path(root)
    vdir = root / "voice" / voice
    raw = root / "voice" / f"{voice}_raw" / "audio"
    tg = vdir / "WhisperTS_textgrid_files"
    for d in (vdir / "audio", raw, tg):
        d.mkdir(parents=True, exist_ok=True)
    segs = []
    for p in clip_paths():
        pcm, sr = read_wav(p)
        shutil.copyfile(p, vdir / "audio" / p.name)
        syn = raw_twin(pcm)
        write_wav(raw / p.name, syn, 16000)
        grid = word_grid(p.stem, len(pcm) / sr)
        TG.write(tg / f"{p.stem}.TextGrid", {"words": grid})
        segs.append((p.stem, pcm, sr, syn, 16000, grid))
    return dict(voice_dir=vdir, raw_audio_dir=raw, textgrid_dir=tg, segments=segs)


if __name__ == "__main__":
    out = Path(sys.argv[1] if len(sys.argv) > 1 else "/tmp/c1_voice")
    v = build_voice(out)
    for f in sorted(out.rglob("*")):
        if f.is_file():
            print(hashlib.sha1(f.read_bytes()).hexdigest(), f.relative_to(out))
    print(f"{len(v['segments'])} segments, {sum(len(s[1]) / s[2] for s in v['segments']):.2f} s of natural audio")
