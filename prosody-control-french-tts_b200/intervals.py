"""Word / pause sequences and syntagmes from TextGrid tier-0 intervals — the integer-millisecond bookkeeping that
decides WHICH (t0, t1) slices are measured.

Mirrors, behaviour for behaviour (incl. the cursor drift):
    extract_words_and_pauses     /root/reference/Code/Preprocessing/gen_break_ssml.py:12-42
    remove_spurious_commas       /root/reference/Code/audioPipeline.py:64-81
    POS pause filter             /root/reference/Code/audioPipeline.py:451-465
    punctuation clamp / inject   /root/reference/Code/audioPipeline.py:470-489
    construct_syntagmes_seq      /root/reference/Code/audioPipeline.py:265-311

The reference asks spaCy (fr_core_news_sm) for part-of-speech tags; that model is a boundary input here: callers
inject ``pos_of(word) -> POS`` (``spacy_pos()`` builds one when spaCy is installed).  The default tags nothing, i.e.
no comma or pause is dropped.
"""
from __future__ import annotations

import re
from typing import Callable, Iterable

FORBIDDEN_POS = frozenset({"DET", "ADP", "CCONJ", "SCONJ", "PART", "PRON"})      # audioPipeline.py:27
INITIAL_PAUSE_THRESHOLD_MS = 150                                                   # gen_break_ssml.py:9
PAUSE_MARKERS = frozenset({"[*]"})                                                 # audioPipeline.py:65
SENTENCE_END = (".", "?", "!")

PosFn = Callable[[str], str]
NO_POS: PosFn = lambda word: "X"

# word-ish tokens (keeping French elisions / hyphens together), the "[*]" marker, or one punctuation character
_TOKENS = re.compile(r"\[\*\]|\w+(?:['’\-]\w+)*['’]?|[^\w\s]", re.UNICODE)


def spacy_pos(model: str = "fr_core_news_sm") -> PosFn:
    """POS predicate backed by the reference's own tagger (only if spaCy and the model are installed)."""
    import spacy
    nlp = spacy.load(model, disable=["ner"])
    return lambda word: (lambda d: d[0].pos_ if len(d) else "X")(nlp(word))


def words_and_pauses(intervals: Iterable) -> list:
    """[(tmin, tmax, mark)] -> [("word", text, ms) | ("pause", None, ms)], dropping a short leading pause."""
    seq = []
    leading = True
    for tmin, tmax, mark in intervals:
        text = mark.strip()
        ms = round(tmax * 1000) - round(tmin * 1000)
        if text:
            seq.append(("word", text, ms))
            leading = False
        elif not leading or ms >= INITIAL_PAUSE_THRESHOLD_MS:
            seq.append(("pause", None, ms))
    return seq


def _split(text: str):
    """-> [(token, trailing_whitespace)] plus the leading whitespace."""
    out, last = [], 0
    lead = ""
    for m in _TOKENS.finditer(text):
        gap = text[last:m.start()]
        if out:
            out[-1][1] += gap
        else:
            lead = gap
        out.append([m.group(0), ""])
        last = m.end()
    if out:
        out[-1][1] += text[last:]
    else:
        lead = text
    return lead, out


def _pos(token: str, pos_of: PosFn) -> str:
    return pos_of(token) if re.match(r"\w", token, re.UNICODE) else "PUNCT"


def strip_spurious_commas(text: str, pos_of: PosFn = NO_POS) -> str:
    """Drop a comma (or pause marker) that directly follows a function word; keep the original spacing."""
    lead, toks = _split(text)
    kept = []
    for tok, ws in toks:
        if (tok == "," or tok in PAUSE_MARKERS) and kept and _pos(kept[-1][0], pos_of) in FORBIDDEN_POS:
            continue
        kept.append((tok, ws))
    return lead + "".join(t + w for t, w in kept)


def first_pos(word: str, pos_of: PosFn) -> str:
    _, toks = _split(word.strip())
    return _pos(toks[0][0], pos_of) if toks else "X"


def segment_sequence(intervals: Iterable, pos_of: PosFn = NO_POS, end_pause_ms: int = 150) -> list:
    """The sequence pass 2 measures: comma strip, POS pause filter, then clamp / inject sentence-final pauses."""
    raw = [(k, strip_spurious_commas(t, pos_of) if k == "word" else t, d) for k, t, d in words_and_pauses(intervals)]
    kept, prev = [], None
    for item in raw:
        if item[0] == "pause" and prev is not None and prev[0] == "word" and first_pos(prev[1], pos_of) in FORBIDDEN_POS:
            prev = item                     # the reference moves on without keeping the pause
            continue
        kept.append(item)
        prev = item
    out = []
    for i, (kind, tok, ms) in enumerate(kept):
        if kind == "pause" and i > 0 and kept[i - 1][0] == "word" and kept[i - 1][1].strip().endswith(SENTENCE_END):
            ms = max(ms, end_pause_ms)
        out.append((kind, tok, ms))
        if kind == "word" and tok.strip().endswith(SENTENCE_END) and not (i + 1 < len(kept) and kept[i + 1][0] == "pause"):
            out.append(("pause", "", end_pause_ms))
    return out


def syntagmes(seq: Iterable) -> list:
    """Runs of words between pauses, and the pauses themselves, on a cursor that starts at 0 and advances by the
    (possibly modified) durations.  -> [(words, start_ms, end_ms, pause_ms)]"""
    out, cursor, words, start = [], 0, [], 0
    for kind, tok, ms in seq:
        if kind == "word":
            if not words:
                start = cursor
            words.append(tok.strip())
            cursor += ms
            continue
        if words:
            out.append((" ".join(words), start, cursor, 0))
            words = []
        out.append(("", cursor, cursor + ms, ms))
        cursor += ms
    if words:
        out.append((" ".join(words), start, cursor, 0))
    return out


def word_count(seq: Iterable) -> int:
    return sum(1 for k, t, _ in seq if k == "word" and t.strip())
