"""SSML emitters and the three CSVs of the step (/root/reference/Code/audioPipeline.py:604-711).

Byte-for-byte the reference's strings: ``f'{x:+.2f}%'`` number formatting, xml.sax escape of the text, the break tag
only when pause >= 50 ms (full duration after . ? !, else int(pause * factor)), and the three <speak> wrappers that
differ in the mstts namespace and the Leading/Tailing silence tags.  Consumers parse these strings back with
``float(pitch.strip('%'))`` (Code/baseline_models/bilstm.py:44-50), so the format is part of the contract.
"""
from __future__ import annotations

from pathlib import Path
from typing import Sequence
from xml.sax.saxutils import escape as xml_escape

SPEAK_MSTTS = ('<speak xmlns="http://www.w3.org/2001/10/synthesis" xmlns:mstts="http://www.w3.org/2001/mstts" '
               'version="1.0" xml:lang="fr-FR">')
SPEAK_PLAIN = '<speak xmlns="http://www.w3.org/2001/10/synthesis" version="1.0" xml:lang="fr-FR">'
LEADING = '<mstts:silence type="Leading-exact" value="0"/>'
TAILING = '<mstts:silence type="Tailing-exact" value="0"/>'


def prosody_open(pitch: float, rate: float, volume: float, text: str) -> str:
    return f'<prosody pitch="{pitch:+.2f}%" rate="{rate:+.2f}%" volume="{volume:+.2f}%">{xml_escape(text)}'


def break_tag(text: str, pause_ms: int, factor) -> str:
    if pause_ms < 50:
        return ""
    dur = pause_ms if (text and text[-1] in ".?!") else int(pause_ms * factor)
    return f'<break time="{dur}ms"/>'


def build(segment_names: Sequence[str], words: Sequence[str], pauses: Sequence[int], sm_pitch, sm_rate, raw_volume,
          voice: str, factor=1):
    """-> (bdd_ssml rows, bdd_syntagme_ssml rows, bdd_syntagme_for_synth rows) as lists of dicts (CSV column order)."""
    by_seg: dict = {}
    syn_rows, synth_rows = [], []
    head_plain = f'{SPEAK_PLAIN}<voice name="{voice}">'
    head_mstts = f'{SPEAK_MSTTS}<voice name="{voice}">{LEADING}'
    tail_mstts = f'{TAILING}</voice></speak>'
    for seg, text, pause, p, r, v in zip(segment_names, words, pauses, sm_pitch, sm_rate, raw_volume):
        pause = int(pause)
        opened = prosody_open(float(p), float(r), float(v), text)
        piece = opened + break_tag(text, pause, factor) + "</prosody>"
        by_seg.setdefault(seg, []).append(piece)
        syn_rows.append({"segment": seg, "syntagme": text, "pause": pause, "ssml": f"{head_plain}{piece}</voice></speak>"})
        synth_rows.append({"segment": seg, "syntagme": text, "pause": pause, "ssml": f"{head_mstts}{opened}</prosody>{tail_mstts}"})
    final = [{"segment": seg, "ssml": head_mstts + "".join(pieces) + tail_mstts} for seg, pieces in by_seg.items()]
    return final, syn_rows, synth_rows


def write_csvs(final, syn_rows, synth_rows, bdd_ssml_csv, bdd_syntagme_ssml_csv, bdd_syntagme_synth_csv) -> None:
    """pandas.DataFrame(rows).to_csv(path, index=False) — same writer the reference uses (:647, :682, :711)."""
    import pandas as pd
    for rows, path in ((final, bdd_ssml_csv), (syn_rows, bdd_syntagme_ssml_csv), (synth_rows, bdd_syntagme_synth_csv)):
        Path(path).parent.mkdir(parents=True, exist_ok=True)
        pd.DataFrame(rows).to_csv(path, index=False)


def _pool(strings):
    """-> (UTF-8 bytes of all strings back to back, int64 offsets [n + 1])"""
    import numpy as np
    enc = [s.encode("utf-8") for s in strings]
    off = np.zeros(len(enc) + 1, np.int64)
    np.cumsum([len(e) for e in enc], out=off[1:])
    return b"".join(enc), off


class TextPools:
    """The per-row strings of a plan (segment names, syntagme texts) packed once: they do not change from step to step."""

    def __init__(self, segment_names: Sequence[str], words: Sequence[str]):
        self.n = len(words)
        self.seg, self.seg_off = _pool(segment_names)
        self.txt, self.txt_off = _pool(words)


def build_csv_bytes(pools: TextPools, pauses, sm_pitch, sm_rate, raw_volume, voice: str, factor=1, lib=None, n_threads: int = 0):
    """The three CSV tables as bytes — what write_csvs(build(...)) puts on disk — formatted natively (pb_ssml_csv) on host threads.
    -> (BDD_ssml.csv, BDD_syntagme_ssml.csv, BDD_syntagme_for_synth.csv)"""
    import ctypes as C
    import numpy as np
    from . import _native as N
    lib = lib if lib is not None else N.load()
    n = pools.n
    pa = np.ascontiguousarray(pauses, np.int32); p = np.ascontiguousarray(sm_pitch, np.float64)
    r = np.ascontiguousarray(sm_rate, np.float64); v = np.ascontiguousarray(raw_volume, np.float64)
    if not (len(pa) == len(p) == len(r) == len(v) == n):
        raise ValueError("row arrays must have one entry per syntagme")
    outs = [C.c_void_p() for _ in range(3)]; lens = [C.c_int64() for _ in range(3)]
    i64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int64))
    dbl = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rc = lib.pb_ssml_csv(n, pools.seg, i64(pools.seg_off), pools.txt, i64(pools.txt_off), pa.ctypes.data_as(C.POINTER(C.c_int32)), dbl(p), dbl(r), dbl(v),
                         float(factor), voice.encode("utf-8"), int(n_threads), C.byref(outs[0]), C.byref(lens[0]), C.byref(outs[1]), C.byref(lens[1]),
                         C.byref(outs[2]), C.byref(lens[2]))
    N.check(lib, None, rc, "pb_ssml_csv")
    try:
        return tuple(C.string_at(o, l.value) for o, l in zip(outs, lens))
    finally:
        for o in outs:
            lib.pb_ssml_free(o)


def write_csv_bytes(tables, bdd_ssml_csv, bdd_syntagme_ssml_csv, bdd_syntagme_synth_csv) -> None:
    for data, path in zip(tables, (bdd_ssml_csv, bdd_syntagme_ssml_csv, bdd_syntagme_synth_csv)):
        Path(path).parent.mkdir(parents=True, exist_ok=True)
        Path(path).write_bytes(data)
