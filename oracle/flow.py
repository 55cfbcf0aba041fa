"""CPU ORACLE (host-logic part) — TEST INFRASTRUCTURE ONLY.  **PARITY UNPINNED** for the numerics it calls.

Loop-by-loop restatement of the reference's "Measure & Build SSML" step on in-memory inputs:

    /root/reference/Code/Preprocessing/gen_break_ssml.py:12-42   extract_words_and_pauses
    /root/reference/Code/audioPipeline.py:64-81                  remove_spurious_commas (POS predicate injected)
    /root/reference/Code/audioPipeline.py:265-311                construct_syntagmes_seq
    /root/reference/Code/audioPipeline.py:364-424                pass 1: segment stats + (sliding) baselines
    /root/reference/Code/audioPipeline.py:436-589                pass 2: per-syntagme deltas
    /root/reference/Code/audioPipeline.py:592-602                EMA + jump clamp
    /root/reference/Code/audioPipeline.py:604-711                three SSML emitters

Deliberately scalar and sequential (one Python loop per reference loop) so it can be read against the
reference; the product's host side is an independent, batched implementation.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field
from typing import Callable, Optional
from xml.sax.saxutils import escape as xml_escape

import numpy as np

from . import oracle as O

FORBIDDEN_POS = {"DET", "ADP", "CCONJ", "SCONJ", "PART", "PRON"}       # audioPipeline.py:27
INITIAL_PAUSE_THRESHOLD = 150                                           # gen_break_ssml.py:9

DEFAULT_PARAMS = dict(                                                  # code defaults, audioPipeline.py:127-139
    pitch_semitones=2.0, pitch_lower_clip_factor=0.7, volume_pct=7.0, rate_percent=15.0, smoothing_alpha=0.4,
    max_jump_percent=5.0, end_punctuation_pause_ms=150, baseline_window=None, inter_syntagme_pause_factor=1,
    threshold_duration_before_slowing_down=1.0, slow_floor_per_sec=2.0)


@dataclass
class Segment:
    name: str
    nat_pcm: np.ndarray
    nat_sr: int
    syn_pcm: Optional[np.ndarray]          # None ≙ CouldntDecodeError -> natural fallback (:385-388, :506-509)
    syn_sr: Optional[int]
    intervals: list                        # tier 0: [(minTime, maxTime, mark)]


def words_and_pauses(intervals) -> list:
    """gen_break_ssml.py:12-42 on already-parsed tier-0 intervals."""
    seq, ignore_initial = [], True
    for (tmin, tmax, mark) in intervals:
        text = mark.strip()
        dur = round(tmax * 1000) - round(tmin * 1000)
        if not text or text == " ":
            if not ignore_initial or dur >= INITIAL_PAUSE_THRESHOLD:
                seq.append(("pause", None, dur))
        else:
            seq.append(("word", text, dur))
            ignore_initial = False
    return seq


_TOK = re.compile(r"\[\*\]|\w+(?:['’\-]\w+)*['’]?|[^\w\s]", re.UNICODE)


def strip_spurious_commas(text: str, pos_of: Callable[[str], str]) -> str:
    """audioPipeline.py:64-81 with the spaCy tagger replaced by an injected ``pos_of(word) -> POS``.
    Tokenisation stand-in: words vs single punctuation characters.  As in spaCy, whitespace TRAILS its token
    (``text_with_ws``), so a dropped comma takes the space after it along ("de, la" -> "dela")."""
    toks, prev_end = [], 0
    for m in _TOK.finditer(text):
        if toks:
            toks[-1][1] += text[prev_end:m.start()]
        lead = text[prev_end:m.start()] if not toks else ""
        toks.append([m.group(0), "", lead])
        prev_end = m.end()
    if not toks:
        return text
    toks[-1][1] += text[prev_end:]
    kept, last_pos = [], None
    for tok, ws, lead in toks:
        is_word = bool(re.match(r"\w", tok))
        if (tok == "," or tok == "[*]") and kept and last_pos in FORBIDDEN_POS:
            continue
        kept.append(lead + tok + ws)
        last_pos = pos_of(tok) if is_word else "PUNCT"
    return "".join(kept)


def build_syntagmes(seq) -> list:
    """audioPipeline.py:265-311."""
    synts, cursor, current, start = [], 0, [], 0
    for kind, tok, dur in seq:
        if kind == "word":
            if not current:
                start = cursor
            current.append(tok.strip())
            cursor += dur
        else:
            if current:
                synts.append(dict(words=" ".join(current), start_ms=start, end_ms=cursor, pause_ms=0))
                current = []
            synts.append(dict(words="", start_ms=cursor, end_ms=cursor + dur, pause_ms=dur))
            cursor += dur
    if current:
        synts.append(dict(words=" ".join(current), start_ms=start, end_ms=cursor, pause_ms=0))
    return synts


def segment_sequence(intervals, pos_of, end_pause_ms) -> list:
    """audioPipeline.py:441-489: comma strip, POS pause filter, punctuation clamp / inject."""
    raw = [(k, strip_spurious_commas(t, pos_of) if k == "word" else t, d) for k, t, d in words_and_pauses(intervals)]
    filtered, prev = [], None
    for item in raw:
        kind, tok, dur = item
        if kind == "pause" and prev is not None and prev[0] == "word":
            first = _TOK.search(prev[1].strip())
            pos = pos_of(first.group(0)) if first and re.match(r"\w", first.group(0)) else "PUNCT"
            if pos in FORBIDDEN_POS:
                prev = item
                continue
        filtered.append(item)
        prev = item
    out = []
    for i, (kind, tok, dur) in enumerate(filtered):
        if kind == "pause" and i > 0:
            pk, pt, _ = filtered[i - 1]
            if pk == "word" and pt.strip().endswith((".", "?", "!")):
                dur = max(dur, end_pause_ms)
        out.append((kind, tok, dur))
        if kind == "word" and tok.strip().endswith((".", "?", "!")):
            if not (i + 1 < len(filtered) and filtered[i + 1][0] == "pause"):
                out.append(("pause", "", end_pause_ms))
    return out


def _median_pitch(pcm, sr, t0=0.0, t1=None):
    return O.median_pitch(pcm, sr, t0, t1, 150.0, 600.0)          # floor/ceiling hard-coded at :329,:332


def syntagme_deltas(p_nat, base, l_syn, wc_syn, nat_total, syn_total, pause_ms, prm):
    """audioPipeline.py:515-577 for one syntagme -> (raw_pitch, raw_volume, raw_rate)."""
    pause_s = pause_ms / 1000.0
    d_nat = max(nat_total - pause_s, 1e-4)
    d_syn = max(syn_total - pause_s, 1e-4)
    P_ST = prm["pitch_semitones"]
    if p_nat > 0:
        st = 12 * np.log2(p_nat / base["f0"])
        st = np.clip(st, -P_ST * prm["pitch_lower_clip_factor"], P_ST)
        p_pct = (2 ** (st / 12) - 1) * 100
    else:
        p_pct = 0.0
    v_pct = (10 ** ((base["loud"] - l_syn) / 20) - 1.0) * 100.0
    v_pct = np.clip(v_pct, -prm["volume_pct"], +prm["volume_pct"])
    if wc_syn > 0:
        nat_r, syn_r = wc_syn / d_nat, wc_syn / d_syn
        rp = (nat_r - syn_r) / syn_r * 100
    else:
        rp = 0.0
    length_s = d_nat
    slow, fast = (1.0, 1.0) if length_s <= 1.0 else (length_s ** 1.5, np.sqrt(length_s))
    rp = rp * slow if rp < 0 else rp / fast
    rp = rp - max(0.0, length_s - prm["threshold_duration_before_slowing_down"]) * prm["slow_floor_per_sec"]
    R = prm["rate_percent"]
    lo, hi = (R * 1.5, R * 0.5) if length_s > 5.0 else (R, R)
    rp = np.clip(rp, -lo, +hi)
    return float(p_pct), float(v_pct), float(rp)


def smooth(values, alpha, max_jump):
    """audioPipeline.py:593-602: EMA over all rows, then forward jump clamp."""
    sm = [values[0]]
    for i in range(1, len(values)):
        sm.append(alpha * values[i] + (1 - alpha) * sm[-1])
    for i in range(1, len(sm)):
        if abs(sm[i] - sm[i - 1]) > max_jump:
            sm[i] = sm[i - 1] + np.sign(sm[i] - sm[i - 1]) * max_jump
    return [float(v) for v in sm]


def prosody_piece(row, p_adj, r_adj, factor, with_break=True):
    """The <prosody> element of the three emitters (:607-625, :652-667, :687-693)."""
    pros = (f'<prosody pitch="{p_adj:+.2f}%" rate="{r_adj:+.2f}%" volume="{row["raw_volume"]:+.2f}%">'
            f'{xml_escape(row["syntagme"])}')
    if with_break and row["pause"] >= 50:
        last = row["syntagme"][-1] if row["syntagme"] else None
        dur = row["pause"] if (last is not None and last in ".?!") else int(row["pause"] * factor)
        pros += f'<break time="{dur}ms"/>'
    return pros + "</prosody>"


_SPEAK_MSTTS = ('<speak xmlns="http://www.w3.org/2001/10/synthesis" xmlns:mstts="http://www.w3.org/2001/mstts" '
                'version="1.0" xml:lang="fr-FR">')
_SPEAK_PLAIN = '<speak xmlns="http://www.w3.org/2001/10/synthesis" version="1.0" xml:lang="fr-FR">'
_LEAD = '<mstts:silence type="Leading-exact" value="0"/>'
_TAIL = '<mstts:silence type="Tailing-exact" value="0"/>'


def emit_ssml(raw_rows, sm_p, sm_r, voice, factor):
    """:604-711 -> (bdd_ssml rows, bdd_syntagme_ssml rows, bdd_syntagme_for_synth rows)."""
    by_seg = {}
    syn_rows, synth_rows = [], []
    for row, p, r in zip(raw_rows, sm_p, sm_r):
        piece = prosody_piece(row, p, r, factor, True)
        by_seg.setdefault(row["segment"], []).append(piece)
        syn_rows.append(dict(segment=row["segment"], syntagme=row["syntagme"], pause=row["pause"],
                             ssml=f'{_SPEAK_PLAIN}<voice name="{voice}">{piece}</voice></speak>'))
        nb = prosody_piece(row, p, r, factor, False)
        synth_rows.append(dict(segment=row["segment"], syntagme=row["syntagme"], pause=row["pause"],
                               ssml=f'{_SPEAK_MSTTS}<voice name="{voice}">{_LEAD}{nb}{_TAIL}</voice></speak>'))
    final = [dict(segment=s, ssml=f'{_SPEAK_MSTTS}<voice name="{voice}">{_LEAD}{"".join(p)}{_TAIL}</voice></speak>')
             for s, p in by_seg.items()]
    return final, syn_rows, synth_rows


def measure_and_build(segments: list, params: dict | None = None, pos_of: Callable[[str], str] = lambda w: "X",
                      azure_voice: str = "fr-FR-HenriNeural") -> dict:
    """The whole step (audioPipeline.py:261-711) on in-memory segments (already sorted by segment number)."""
    prm = dict(DEFAULT_PARAMS); prm.update(params or {})
    meter_rate = segments[0].nat_sr                                   # :373 one meter from the FIRST file
    stats = []
    for s in segments:                                                # pass 1 (:375-400)
        seq = words_and_pauses(s.intervals)
        wc = sum(1 for k, t, m in seq if k == "word" and t.strip())
        p_nat = _median_pitch(s.nat_pcm, s.nat_sr)
        l_nat = O.lufs(s.nat_pcm, s.nat_sr, meter_rate)
        if s.syn_pcm is not None:
            l_syn = O.lufs(s.syn_pcm, s.syn_sr, meter_rate)
            d_syn = O.duration(len(s.syn_pcm), s.syn_sr)
        else:
            l_syn = l_nat
            d_syn = O.duration(len(s.nat_pcm), s.nat_sr)
        d_nat = O.duration(len(s.nat_pcm), s.nat_sr)
        rate_ratio = (wc / d_nat) / (wc / d_syn) if wc > 0 and d_syn > 0 else 1.0
        stats.append(dict(segment=s.name, p_nat=p_nat, l_nat=l_nat, l_syn=l_syn, d_nat=d_nat, d_syn=d_syn, wc=wc,
                          rate_ratio=rate_ratio))
    n_seg, win = len(stats), prm["baseline_window"]

    def _base(window):
        voiced = [w["p_nat"] for w in window if w["p_nat"] > 0]
        f0 = (float(np.median(voiced)) if voiced else float("nan")) or 1.0
        return dict(f0=f0, loud=float(np.median([w["l_nat"] for w in window])),
                    rate=float(np.median([w["rate_ratio"] for w in window])))
    if win is None or win >= n_seg:                                   # :403-408
        baselines = [_base(stats)] * n_seg
    else:                                                             # :410-424
        half = win // 2
        baselines = [_base(stats[max(0, i - half):min(n_seg, i + half + 1)]) for i in range(n_seg)]

    raw_rows, units = [], []
    for idx, s in enumerate(segments):                                # pass 2 (:437-589)
        seq = segment_sequence(s.intervals, pos_of, prm["end_punctuation_pause_ms"])
        seg_meter_rate = s.nat_sr                                     # :493 meter_seg from the natural file
        base = baselines[idx]
        for syn in build_syntagmes(seq):
            t0, t1 = syn["start_ms"] / 1000, syn["end_ms"] / 1000
            wc_syn = len(syn["words"].split())
            p_nat = _median_pitch(s.nat_pcm, s.nat_sr, t0, t1)
            if s.syn_pcm is not None:
                l_syn = O.lufs(s.syn_pcm, s.syn_sr, seg_meter_rate, t0, t1)
                syn_total = O.part_duration(len(s.syn_pcm), s.syn_sr, t0, t1)
            else:
                l_syn = O.lufs(s.nat_pcm, s.nat_sr, seg_meter_rate, t0, t1)
                syn_total = O.part_duration(len(s.nat_pcm), s.nat_sr, t0, t1)
            nat_total = O.part_duration(len(s.nat_pcm), s.nat_sr, t0, t1)
            rp, rv, rr = syntagme_deltas(p_nat, base, l_syn, wc_syn, nat_total, syn_total, syn["pause_ms"], prm)
            raw_rows.append(dict(segment=s.name, syntagme=syn["words"], pause=syn["pause_ms"],
                                 raw_pitch=rp, raw_volume=rv, raw_rate=rr))
            units.append(dict(segment=s.name, t0=t0, t1=t1, p_nat=p_nat, l_syn=l_syn, nat_total=nat_total,
                              syn_total=syn_total))
    if not raw_rows:
        raise KeyError(0)                                             # df.loc[0, ...] on an empty frame (:593)
    sm_p = smooth([r["raw_pitch"] for r in raw_rows], prm["smoothing_alpha"], prm["max_jump_percent"])
    sm_r = smooth([r["raw_rate"] for r in raw_rows], prm["smoothing_alpha"], prm["max_jump_percent"])
    final, syn_rows, synth_rows = emit_ssml(raw_rows, sm_p, sm_r, azure_voice, prm["inter_syntagme_pause_factor"])
    return dict(seg_stats=stats, baselines=baselines, units=units, raw_rows=raw_rows, sm_p=sm_p, sm_r=sm_r,
                bdd_ssml=final, bdd_syntagme_ssml=syn_rows, bdd_syntagme_synth=synth_rows)
