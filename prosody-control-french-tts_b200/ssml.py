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
