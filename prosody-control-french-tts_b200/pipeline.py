"""File-level drop-ins with the reference's own names and signatures.

    get_part_duration(wav_path, t0=0.0, t1=None)   Code/audioPipeline.py:314-323
    get_median_pitch(wav_path, t0=0.0, t1=None)    Code/audioPipeline.py:326-335
    get_lufs(wav_path, meter, t0=0.0, t1=None)     Code/audioPipeline.py:338-358
    get_duration(wav_path)                         Code/audioPipeline.py:360-361
    measure_prosody_and_build_ssml(self)           Code/audioPipeline.py:261-711  (bind it onto AudioPipeline)

The scalar closures exist for API parity and tests; each call is one GPU unit, so real work should go through
measure_prosody_and_build_ssml (one GPU call for the whole voice) or prosody_b200.step directly.
"""
from __future__ import annotations

import logging
import re
import wave
from pathlib import Path

import numpy as np

from . import intervals as IV
from . import ssml as SSML
from . import step as S
from . import textgrid as TG
from .batch import Extractor, Units, part_durations, pitch_params


class CouldntDecodeError(Exception):
    """Stands in for pydub.exceptions.CouldntDecodeError (the reference falls back to natural audio on it)."""


class UnsupportedAudioError(Exception):
    """A WAV file the reference WOULD decode but this path does not measure: anything but mono 16-bit PCM.

    The reference hands any container to pydub / Praat: for interleaved stereo its loudness closure treats the interleaved
    samples as a mono signal of twice the length (np.array(AudioSegment.get_array_of_samples()), Code/audioPipeline.py:343-348)
    while Praat analyses the channels jointly, and 8 / 24 / 32-bit files are rescaled by each library in its own way.  None of
    that is reproduced here (the pipeline's own steps only ever write mono s16).  It is deliberately NOT a CouldntDecodeError:
    that one makes the step fall back to natural-audio metrics for a raw file (:385-388, :506-509), which would silently change
    results; this one stops the step with the file name."""


class Meter:
    """Shim for ``pyln.Meter(rate)``: the hot path only needs the rate the meter was built with."""

    def __init__(self, rate):
        self.rate = float(rate)


def read_wav(path):
    """Mono 16-bit PCM WAV -> (int16 samples, rate). Anything else is outside what the pipeline writes."""
    try:
        with wave.open(str(path), "rb") as w:
            if w.getsampwidth() != 2 or w.getnchannels() != 1 or w.getcomptype() != "NONE":
                raise UnsupportedAudioError(f"{path}: {w.getnchannels()} channel(s), {8 * w.getsampwidth()}-bit, {w.getcomptype()}: only mono 16-bit "
                                            f"PCM WAV (what the pipeline itself writes) is measured; convert the file first")
            return np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").astype(np.int16), w.getframerate()
    except (wave.Error, EOFError, FileNotFoundError) as e:
        raise CouldntDecodeError(str(e)) from e


_default_ex: Extractor | None = None


def default_extractor() -> Extractor:
    global _default_ex
    if _default_ex is None:
        _default_ex = Extractor(0)
    return _default_ex


def _unit(pcm, sr, t0, t1, meter_rate=None):
    return Units.from_list([(0, len(pcm), sr, float(t0), None if t1 is None else float(t1), float(meter_rate or sr))])


def get_median_pitch(wav_path, t0=0.0, t1=None, extractor: Extractor | None = None) -> float:
    pcm, sr = read_wav(wav_path)
    r = (extractor or default_extractor()).median_pitch(pcm, _unit(pcm, sr, t0, t1), pitch_params(**S.REFERENCE_PITCH))
    if r["status"][0] != 0:
        raise S.PraatError(f"Praat refuses this sound (status {int(r['status'][0])}): shorter than 3 / pitch_floor s or empty")
    return float(r["median_f0"][0])


def get_lufs(wav_path, meter, t0=0.0, t1=None, extractor: Extractor | None = None) -> float:
    pcm, sr = read_wav(wav_path)
    out, st = (extractor or default_extractor()).lufs(pcm, _unit(pcm, sr, t0, t1, meter.rate))
    if st[0] & 96:
        raise ValueError("Audio must have length greater than the block size.")
    return float(out[0])


def get_part_duration(wav_path, t0=0.0, t1=None) -> float:
    pcm, sr = read_wav(wav_path)
    d, _ = part_durations(_unit(pcm, sr, t0, t1))
    return float(d[0])


def get_duration(wav_path) -> float:
    return get_part_duration(wav_path)


def load_voice(audio_dir, raw_audio_dir, textgrid_dir):
    """Reads segment_ph*.wav (sorted by number, :364-367), the raw-synth twins and the TextGrids into one PCM buffer."""
    seg_files = sorted(Path(audio_dir).glob("*.wav"), key=lambda p: int(re.search(r"segment_ph(\d+)", p.stem).group(1)))
    bufs, segs, off = [], [], 0
    # all alignments in one native call (host threads); a file it refuses goes through the Python reader for its error
    tg_paths = [Path(textgrid_dir) / f"{wav.stem}.TextGrid" for wav in seg_files]
    tg_status, tg_words = TG.read_tier_batch(tg_paths, tier=0) if seg_files else ([], [])
    for k, wav in enumerate(seg_files):
        pcm, sr = read_wav(wav)
        words = tg_words[k] if tg_status[k] == TG.STATUS_OK else TG.word_intervals(tg_paths[k])
        seg = S.Segment(wav.stem, off, len(pcm), sr, words)
        bufs.append(pcm); off += len(pcm)
        try:
            spcm, ssr = read_wav(Path(raw_audio_dir) / f"{wav.stem}.wav")
            seg.syn_off, seg.syn_nx, seg.syn_sr = off, len(spcm), ssr
            bufs.append(spcm); off += len(spcm)
        except CouldntDecodeError:
            logging.warning(f"Couldn’t decode raw audio {wav.stem}.wav; falling back to natural metrics")
        segs.append(seg)
    return (np.concatenate(bufs) if bufs else np.zeros(0, np.int16)), segs


def measure_prosody_and_build_ssml(self, extractor: Extractor | None = None, pos_of: IV.PosFn | None = None):
    """Drop-in for AudioPipeline.measure_prosody_and_build_ssml: same inputs on disk, same three CSVs out."""
    logging.info(">>> Measure Prosody & Build SSML")
    pcm, segs = load_voice(self.voice_dir / "audio", self.raw_audio_dir, self.textgrid_dir)
    if not segs:
        logging.error("No audio segments found!")
        return
    prosody = dict(pitch_semitones=self.p_st, pitch_lower_clip_factor=self.pitch_lower_clip_factor, volume_pct=self.v_pct,
                   rate_percent=self.r_pct_clamp, smoothing_alpha=self.alpha, max_jump_percent=self.max_jump,
                   end_punctuation_pause_ms=self.end_pause_ms, baseline_window=self.baseline_window,
                   inter_syntagme_pause_factor=self.inter_syntagme_pause_factor,
                   threshold_duration_before_slowing_down=self.threshold_duration_before_slowing_down,
                   slow_floor_per_sec=self.slow_floor_per_sec)
    if pos_of is None:
        # Like the reference (which fails at import without fr_core_news_sm, audioPipeline.py:26), a missing tagger is an
        # error: silently dropping the comma / pause filter would change which slices are measured.  Callers that WANT
        # no filter pass pos_of=intervals.NO_POS explicitly.
        try:
            pos_of = IV.spacy_pos()
        except Exception as e:
            raise RuntimeError("spaCy model fr_core_news_sm is required for the comma / pause POS filter "
                               "(pass pos_of=prosody_b200.intervals.NO_POS to run without it)") from e
    ex = extractor or default_extractor()
    pl = S.plan(segs, prosody, pos_of)
    out = S.measure(ex, pcm, pl, prosody)
    names = [segs[i].name for i in pl.syn_seg]
    # the three tables, formatted natively into the bytes pandas' to_csv(index=False) would write for the Python emitters' rows
    # (ssml.build / ssml.write_csvs: same output, tests/test_ssml_pipeline.py), which matters once a voice has 10^5 rows
    tables = SSML.build_csv_bytes(SSML.TextPools(names, pl.syn_words), pl.syn_pause_ms, out["sm_pitch"], out["sm_rate"], out["raw_volume"],
                                  self.azure_voice, self.inter_syntagme_pause_factor, lib=ex._lib)
    SSML.write_csv_bytes(tables, self.bdd_ssml_csv, self.bdd_syntagme_ssml_csv, self.bdd_syntagme_synth_csv)
    return out


def export_training_json(self):
    """Drop-in for AudioPipeline.export_training_json (/root/reference/Code/audioPipeline.py:840-854): the syntagme-level
    CSV of this voice -> training_data_<name>.json, then bdd.json over every voice folder under <out_dir>/results."""
    from . import training_export as TE
    TE.create_training_data(str(self.bdd_syntagme_ssml_csv), str(self.results_dir / f"training_data_{self.name}.json"))
    TE.combine_training_jsons(str(self.out_dir / "results"), str(self.out_dir / "results" / "bdd.json"))
