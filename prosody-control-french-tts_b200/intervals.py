"""Word / pause sequences and syntagmes from TextGrid tier-0 intervals — the integer-millisecond bookkeeping that
decides WHICH (t0, t1) slices are measured.

Mirrors, behaviour for behaviour (incl. the cursor drift):
    extract_words_and_pauses     /root/reference/Code/Preprocessing/gen_break_ssml.py:12-42
    remove_spurious_commas       /root/reference/Code/audioPipeline.py:64-81
    POS pause filter             /root/reference/Code/audioPipeline.py:451-465
    punctuation clamp / inject   /root/reference/Code/audioPipeline.py:470-489
    construct_syntagmes_seq      /root/reference/Code/audioPipeline.py:265-311

The reference asks spaCy (fr_core_news_sm) for part-of-speech tags; that model is a boundary input here.  Callers
inject either a TAGGER — an object with ``tag(text) -> [(token, trailing_ws, POS)]`` that tokenises and tags a whole
mark the way ``_nlp(text)`` does (``spacy_pos()`` builds one on the reference's own model: elisions such as "l'homme,"
are then split and tagged exactly as the reference sees them) — or, for tests and callers without spaCy, a plain
``pos_of(word) -> POS`` that is applied to regex-split tokens.  The default tags nothing: no comma or pause is dropped.
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


class SpacyTagger:
    """The reference's own tagger (`_nlp = spacy.load("fr_core_news_sm")`, audioPipeline.py:26): tag(text) runs the
    model on the WHOLE mark and returns spaCy's tokens, so the comma filter sees the real previous token ("homme" in
    "l'homme,") and the pause filter the first token of the previous mark, as at audioPipeline.py:70-79,459."""

    def __init__(self, nlp):
        self.nlp = nlp

    def tag(self, text: str):
        return [(t.text, t.whitespace_, t.pos_) for t in self.nlp(text)]

    def __call__(self, word: str) -> str:
        d = self.nlp(word)
        return d[0].pos_ if len(d) else "X"


def spacy_pos(model: str = "fr_core_news_sm") -> "SpacyTagger":
    """Tagger backed by the reference's own model (raises when spaCy or the model is not installed)."""
    import spacy
    return SpacyTagger(spacy.load(model))


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
    if hasattr(pos_of, "tag"):
        # the reference's loop on the tagger's own tokens of the whole mark (audioPipeline.py:70-81); like
        # `"".join(t.text_with_ws ...)`, leading whitespace the tokeniser swallowed is not restored
        kept = []
        for tok, ws, pos in pos_of.tag(text):
            if (tok == "," or tok in PAUSE_MARKERS) and kept and kept[-1][2] in FORBIDDEN_POS:
                continue
            kept.append((tok, ws, pos))
        return "".join(t + w for t, w, _ in kept)
    lead, toks = _split(text)
    kept = []
    for tok, ws in toks:
        if (tok == "," or tok in PAUSE_MARKERS) and kept and _pos(kept[-1][0], pos_of) in FORBIDDEN_POS:
            continue
        kept.append((tok, ws))
    return lead + "".join(t + w for t, w in kept)


def first_pos(word: str, pos_of: PosFn) -> str:
    """`_nlp(ptok.strip())[0].pos_` (audioPipeline.py:459): POS of the first token of the mark."""
    if hasattr(pos_of, "tag"):
        toks = pos_of.tag(word.strip())
        return toks[0][2] if toks else "X"
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
