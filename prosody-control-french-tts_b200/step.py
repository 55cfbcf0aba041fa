"""The "Measure & Build SSML" step, batched: every measurement of pass 1 (per segment) and pass 2 (per syntagme) of
AudioPipeline.measure_prosody_and_build_ssml (/root/reference/Code/audioPipeline.py:261-711) goes to the GPU in ONE
call; baselines, deltas, smoothing and SSML strings follow on the host in float64.

What stays identical to the reference, on purpose: pitch is re-analysed per slice with its own
frame grid; floor 150 / ceiling 600; the natural timeline's (t0, t1) is applied to the synthetic file; meters are
built from the natural file's rate (first file for pass 1, the segment's own for pass 2) and used on synthetic audio;
peak normalisation before loudness; < 0.4 s and empty slices fall back to the whole file; pause rows are measured too.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _native as N
from . import intervals as IV
from .batch import Extractor, Units, pitch_params

DEFAULT_PROSODY = dict(                      # code defaults, Code/audioPipeline.py:127-139
    pitch_semitones=2.0, pitch_lower_clip_factor=0.7, volume_pct=7.0, rate_percent=15.0, smoothing_alpha=0.4,
    max_jump_percent=5.0, end_punctuation_pause_ms=150, baseline_window=None, inter_syntagme_pause_factor=1,
    threshold_duration_before_slowing_down=1.0, slow_floor_per_sec=2.0)
REFERENCE_PITCH = dict(pitch_floor=150.0, pitch_ceiling=600.0)     # hard-coded at Code/audioPipeline.py:329,332


class PraatError(RuntimeError):
    """The reference step would have died here: parselmouth raises on slices shorter than 3 / pitch_floor seconds."""


@dataclass
class Segment:
    """One segment_ph*.wav with its synthetic twin and tier-0 alignment, all PCM living in one shared int16 buffer."""
    name: str
    nat_off: int
    nat_nx: int
    nat_sr: int
    intervals: list                          # [(tmin, tmax, mark)]
    syn_off: Optional[int] = None            # None <-> CouldntDecodeError: natural audio stands in (:385-388, :506-509)
    syn_nx: Optional[int] = None
    syn_sr: Optional[int] = None


@dataclass
class Plan:
    """Everything the host knows before measuring: sequences, syntagmes and the flat unit list."""
    segments: Sequence[Segment]
    word_counts: np.ndarray                  # pass-1 wc per segment
    syn_seg: np.ndarray                      # segment index of every syntagme row
    syn_words: list
    syn_start_ms: np.ndarray
    syn_end_ms: np.ndarray
    syn_pause_ms: np.ndarray
    syn_wc: np.ndarray
    units: Units
    want_pitch: np.ndarray
    want_lufs: np.ndarray
    n_seg: int = 0
    n_syn: int = 0


def plan(segments: Sequence[Segment], prosody: dict | None = None, pos_of: IV.PosFn = IV.NO_POS) -> Plan:
    """Host-only. Unit layout: [nat whole]*S, [syn whole]*S, then per syntagme (nat slice, syn slice)."""
    prm = dict(DEFAULT_PROSODY); prm.update(prosody or {})
    S = len(segments)
    if S == 0:
        raise ValueError("No audio segments found!")
    wc = np.zeros(S, np.int32)
    seg_of, words, s_ms, e_ms, p_ms, swc = [], [], [], [], [], []
    for i, seg in enumerate(segments):
        wc[i] = IV.word_count(IV.words_and_pauses(seg.intervals))
        for w, a, b, p in IV.syntagmes(IV.segment_sequence(seg.intervals, pos_of, prm["end_punctuation_pause_ms"])):
            seg_of.append(i); words.append(w); s_ms.append(a); e_ms.append(b); p_ms.append(p); swc.append(len(w.split()))
    seg_of = np.asarray(seg_of, np.int64)
    K = len(seg_of)
    nat_off = np.array([s.nat_off for s in segments], np.int64); nat_nx = np.array([s.nat_nx for s in segments], np.int64)
    nat_sr = np.array([s.nat_sr for s in segments], np.float64)
    has_syn = np.array([s.syn_off is not None for s in segments])
    syn_off = np.where(has_syn, [s.syn_off if s.syn_off is not None else 0 for s in segments], nat_off).astype(np.int64)
    syn_nx = np.where(has_syn, [s.syn_nx if s.syn_nx is not None else 0 for s in segments], nat_nx).astype(np.int64)
    syn_sr = np.where(has_syn, [s.syn_sr if s.syn_sr is not None else 0 for s in segments], nat_sr).astype(np.float64)
    meter0 = float(segments[0].nat_sr)                              # :373 one meter, from the FIRST natural file
    t0 = np.asarray(s_ms, np.float64) / 1000.0                       # :496-497  ms / 1000
    t1 = np.asarray(e_ms, np.float64) / 1000.0
    n = 2 * S + 2 * K
    file_off = np.empty(n, np.int64); file_nx = np.empty(n, np.int64); rate = np.empty(n); has_t1 = np.zeros(n, np.int32)
    ut0 = np.zeros(n); ut1 = np.zeros(n); meter = np.empty(n)
    want_p = np.zeros(n, np.uint8); want_l = np.zeros(n, np.uint8)
    file_off[:S], file_nx[:S], rate[:S], meter[:S] = nat_off, nat_nx, nat_sr, meter0
    want_p[:S] = 1; want_l[:S] = 1
    file_off[S:2 * S], file_nx[S:2 * S], rate[S:2 * S], meter[S:2 * S] = syn_off, syn_nx, syn_sr, meter0
    want_l[S:2 * S] = 1
    a = slice(2 * S, n, 2); b = slice(2 * S + 1, n, 2)
    file_off[a], file_nx[a], rate[a], meter[a] = nat_off[seg_of], nat_nx[seg_of], nat_sr[seg_of], nat_sr[seg_of]
    file_off[b], file_nx[b], rate[b], meter[b] = syn_off[seg_of], syn_nx[seg_of], syn_sr[seg_of], nat_sr[seg_of]   # :493 meter_seg
    has_t1[2 * S:] = 1
    ut0[a] = t0; ut1[a] = t1; ut0[b] = t0; ut1[b] = t1
    want_p[a] = 1; want_l[b] = 1
    units = Units(file_off, file_nx, rate, has_t1, ut0, ut1, meter)
    return Plan(segments, wc, seg_of, words, np.asarray(s_ms, np.int64), np.asarray(e_ms, np.int64),
                np.asarray(p_ms, np.int32), np.asarray(swc, np.int32), units, want_p, want_l, S, K)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def baselines(p_nat, l_nat, rate_ratio, window: int | None, lib=None):
    """Global or sliding-window medians per segment (:401-424), np.median semantics, in C (pb_segment_baselines).
    -> (f0, loud, rate) arrays."""
    lib = lib if lib is not None else N.load()
    S = len(p_nat)
    p = np.ascontiguousarray(p_nat, np.float64); l = np.ascontiguousarray(l_nat, np.float64); r = np.ascontiguousarray(rate_ratio, np.float64)
    f0 = np.empty(S); loud = np.empty(S); rate = np.empty(S)
    N.check(lib, None, lib.pb_segment_baselines(S, _dp(p), _dp(l), _dp(r), -1 if window is None else int(window), _dp(f0), _dp(loud), _dp(rate)),
            "pb_segment_baselines")
    return f0, loud, rate


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def measure(ex: Extractor, pcm, pl: Plan, prosody: dict | None = None, pitch: dict | None = None, strict: bool = True) -> dict:
    """GPU measurements for every unit of the plan, then baselines, raw deltas and smoothing. No strings yet."""
    prm = dict(DEFAULT_PROSODY); prm.update(prosody or {})
    pp = pitch_params(**(pitch or REFERENCE_PITCH))
    S, K = pl.n_seg, pl.n_syn
    r = ex.extract(pcm, pl.units, pp, want_pitch=pl.want_pitch, want_lufs=pl.want_lufs)
    st = r["status"]
    _raise_like_the_reference(st, pl, strict)
    out = finish(pl, r["median_f0"], r["lufs"], r["duration_s"], prm, ex._lib)
    out["status"] = st
    out["timings"] = ex.timings()
    return out


def submit(ex: Extractor, pcm, pl: Plan, pitch: dict | None = None) -> None:
    """Asynchronous first half of measure(): everything up to the GPU results is enqueued on `ex`; collect() finishes the step.
    With two Extractors a caller keeps two batches in flight: batch k+1 is planned and enqueued, batch k-1 post-processed, while the
    GPU runs batch k."""
    ex.submit(pcm, pl.units, pitch_params(**(pitch or REFERENCE_PITCH)), want_pitch=pl.want_pitch, want_lufs=pl.want_lufs)


def collect(ex: Extractor, pl: Plan, prosody: dict | None = None, strict: bool = True) -> dict:
    """Second half of measure() for a batch submitted with submit()."""
    prm = dict(DEFAULT_PROSODY); prm.update(prosody or {})
    r = ex.wait()
    _raise_like_the_reference(r["status"], pl, strict)
    out = finish(pl, r["median_f0"], r["lufs"], r["duration_s"], prm, ex._lib)
    out["status"] = r["status"]
    out["timings"] = ex.timings()
    return out


def _raise_like_the_reference(st, pl: Plan, strict: bool) -> None:
    if not strict:
        return
    bad = np.nonzero(((st & N.PB_UNIT_PITCH_MASK) != 0) & (pl.want_pitch != 0))[0]
    if len(bad):
        u = int(bad[0])
        raise PraatError(f"unit {u} (t0={pl.units.t0[u]}, t1={pl.units.t1[u]}): Praat refuses this slice "
                         f"(status {int(st[u]) & N.PB_UNIT_PITCH_MASK}); the reference step aborts here")
    bad = np.nonzero((st & (N.PB_UNIT_LUFS_ERROR | N.PB_UNIT_SLICE_ERROR)) != 0)[0]
    if len(bad):
        raise ValueError(f"unit {int(bad[0])}: Audio must have length greater than the block size.")


def finish(pl: Plan, med, lufs, dur, prm: dict, lib=None) -> dict:
    """The host half of the step on per-unit measurements laid out like the plan's units: pass-1 statistics and
    baselines (:375-424), per-syntagme raw deltas (:515-577), EMA + jump clamp (:593-602).  float64 throughout."""
    lib = lib if lib is not None else N.load()
    S, K = pl.n_seg, pl.n_syn
    med = np.asarray(med, np.float64); lufs = np.asarray(lufs, np.float64); dur = np.asarray(dur, np.float64)
    # ---- pass 1 (:375-400)
    p_nat, l_nat, d_nat = med[:S], lufs[:S], dur[:S]
    l_syn, d_syn = lufs[S:2 * S], dur[S:2 * S]
    wc = pl.word_counts.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        rate_ratio = np.where((pl.word_counts > 0) & (d_syn > 0), (wc / d_nat) / (wc / d_syn), 1.0)
    b_f0, b_loud, b_rate = baselines(p_nat, l_nat, rate_ratio, prm["baseline_window"], lib)
    # ---- pass 2 (:495-589)
    a = slice(2 * S, 2 * S + 2 * K, 2); b = slice(2 * S + 1, 2 * S + 2 * K, 2)
    sp_nat = np.ascontiguousarray(med[a]); sl_syn = np.ascontiguousarray(lufs[b])
    nat_total = np.ascontiguousarray(dur[a]); syn_total = np.ascontiguousarray(dur[b])
    raw_p = np.empty(K); raw_v = np.empty(K); raw_r = np.empty(K)
    dprm = N.PbDeltaParams(prm["pitch_semitones"], prm["pitch_lower_clip_factor"], prm["volume_pct"], prm["rate_percent"],
                           prm["threshold_duration_before_slowing_down"], prm["slow_floor_per_sec"])
    bf = np.ascontiguousarray(b_f0[pl.syn_seg]); bl = np.ascontiguousarray(b_loud[pl.syn_seg])
    N.check(lib, None, lib.pb_syntagme_deltas(K, _dp(sp_nat), _dp(bf), _dp(bl), _dp(sl_syn), _ip(pl.syn_wc), _dp(nat_total),
                                              _dp(syn_total), _ip(pl.syn_pause_ms), C.byref(dprm), _dp(raw_p), _dp(raw_v), _dp(raw_r)),
            "pb_syntagme_deltas")
    if K == 0:
        raise KeyError(0)                     # df.loc[0, ...] on an empty frame (:593)
    sm_p = np.empty(K); sm_r = np.empty(K)
    N.check(lib, None, lib.pb_ema_clamp(_dp(raw_p), K, prm["smoothing_alpha"], prm["max_jump_percent"], _dp(sm_p)), "pb_ema_clamp")
    N.check(lib, None, lib.pb_ema_clamp(_dp(raw_r), K, prm["smoothing_alpha"], prm["max_jump_percent"], _dp(sm_r)), "pb_ema_clamp")
    return dict(seg_stats=dict(p_nat=p_nat, l_nat=l_nat, l_syn=l_syn, d_nat=d_nat, d_syn=d_syn, wc=pl.word_counts, rate_ratio=rate_ratio),
                baselines=dict(f0=b_f0, loud=b_loud, rate=b_rate),
                syn=dict(p_nat=sp_nat, l_syn=sl_syn, nat_total=nat_total, syn_total=syn_total),
                raw_pitch=raw_p, raw_volume=raw_v, raw_rate=raw_r, sm_pitch=sm_p, sm_rate=sm_r)
