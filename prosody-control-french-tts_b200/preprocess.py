"""Silence segmentation drop-ins: same names and argument meaning as
/root/reference/Code/Preprocessing/preprocess_audio.py (segment_audio_file :20-50, save_segments :52-73,
analyze_segment_lengths :75-98, main :100-118), with pydub.split_on_silence replaced by the batched GPU call.

Only what the pipeline itself produces is accepted: mono 16-bit PCM WAV.  `segment_audio_files` is the batched form
(many files, one GPU call); the single-file functions go through it.
"""
from __future__ import annotations

import logging
import os
import wave
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from .batch import Extractor, Units
from .pipeline import default_extractor, read_wav

logger = logging.getLogger(__name__)


@dataclass
class AudioSlice:
    """The part of pydub's AudioSegment the reference uses on a segment: len() in ms and export()."""
    samples: np.ndarray     # int16, zero padding pydub would append already included
    frame_rate: int

    def __len__(self) -> int:        # pydub: round(1000 * frame_count / frame_rate)
        return int(round(1000.0 * (len(self.samples) / self.frame_rate)))

    def export(self, out_f, format: str = "wav"):
        if format != "wav":
            raise ValueError("only wav export is supported (the reference's save_segments default)")
        with wave.open(str(out_f), "wb") as w:
            w.setnchannels(1); w.setsampwidth(2); w.setframerate(int(self.frame_rate))
            w.writeframes(np.ascontiguousarray(self.samples, "<i2").tobytes())
        return out_f


def segment_audio_files(input_files, min_silence_len, silence_thresh, keep_silence, extractor: Extractor | None = None):
    """Batched segment_audio_file: one list of AudioSlice per input file, one GPU call for all of them."""
    ex = extractor or default_extractor()
    pcs, rates = [], []
    for f in input_files:
        if not os.path.exists(f):
            raise FileNotFoundError(f"Le fichier {f} n'existe pas")
        pcm, sr = read_wav(f)
        pcs.append(pcm); rates.append(sr)
    off = np.cumsum([0] + [len(p) for p in pcs])
    units = Units(off[:-1], [len(p) for p in pcs], rates, np.zeros(len(pcs), np.int32), np.zeros(len(pcs)), np.zeros(len(pcs)))
    r = ex.split_on_silence(np.concatenate(pcs) if pcs else np.zeros(0, np.int16), units, min_silence_len, silence_thresh, keep_silence)
    out = []
    for i, (pcm, sr) in enumerate(zip(pcs, rates)):
        segs = []
        for k in range(int(r["seg_off"][i]), int(r["seg_off"][i + 1])):
            a, n, pad = int(r["first_sample"][k]), int(r["n_samples"][k]), int(r["n_pad"][k])
            s = pcm[a:a + n]
            segs.append(AudioSlice(np.concatenate([s, np.zeros(pad, np.int16)]) if pad else s, sr))
        out.append(segs)
    return out


def segment_audio_file(input_file, min_silence_len, silence_thresh, keep_silence, extractor: Extractor | None = None):
    """Segmente un fichier audio en utilisant les silences comme points de découpe (reference :20-50)."""
    logger.info(f"Chargement du fichier audio: {input_file}")
    segments = segment_audio_files([input_file], min_silence_len, silence_thresh, keep_silence, extractor)[0]
    logger.info(f"Segmentation terminée. {len(segments)} segments créés.")
    return segments


def save_segments(segments, output_dir, format="wav"):
    """Writes segment_ph{i+1}.{format} into output_dir (reference :52-73)."""
    Path(output_dir).mkdir(parents=True, exist_ok=True)
    for i, segment in enumerate(segments):
        segment.export(os.path.join(output_dir, f"segment_ph{i+1}.{format}"), format=format)


def analyze_segment_lengths(segments):
    """Same statistics dict as the reference (:75-98); like it, raises on an empty list (min of empty)."""
    lengths = [len(segment) for segment in segments]
    return {
        "nombre_segments": len(segments),
        "duree_moyenne": np.mean(lengths) / 1000,
        "duree_min": min(lengths) / 1000,
        "duree_max": max(lengths) / 1000,
        "duree_totale": sum(lengths) / 1000,
    }


def main(input_file, output_dir, min_silence_len=1000, silence_thresh=-50, keep_silence=300, extractor: Extractor | None = None):
    segments = segment_audio_file(input_file, min_silence_len, silence_thresh, keep_silence, extractor)
    stats = analyze_segment_lengths(segments)
    for key, value in stats.items():
        logger.info(f"{key}: {value}")
    save_segments(segments, output_dir, format="wav")
    return stats
