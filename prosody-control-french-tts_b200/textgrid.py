"""Minimal Praat TextGrid reader / writer (long and short text formats).

The reference reads its alignments with the `textgrid` package (1.6.1; Code/Preprocessing/gen_break_ssml.py:19-26),
which is not installable here.  Semantics kept: times are rounded to 5 decimals on read (that package's default
`round_digits`), intervals with min >= max are dropped, `""` inside a mark is an escaped quote, and tier 0 is the
word tier.  The writer emits the long format that Code/Aligners/use_whisper_timestamped.py:330-395 produces.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from pathlib import Path

_TOKEN = re.compile(r'"(?:[^"]|"")*"|\[[^\]]*\]|<exists>|<absent>|[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?')
ROUND_DIGITS = 5


@dataclass
class Tier:
    name: str
    kind: str                       # "IntervalTier" | "TextTier"
    xmin: float
    xmax: float
    intervals: list = field(default_factory=list)   # IntervalTier: (tmin, tmax, mark); TextTier: (time, mark)


@dataclass
class TextGrid:
    xmin: float
    xmax: float
    tiers: list


def _decode(raw: bytes) -> str:
    if raw[:2] in (b"\xff\xfe", b"\xfe\xff"):
        return raw.decode("utf-16")
    if raw[:3] == b"\xef\xbb\xbf":
        return raw[3:].decode("utf-8")
    try:
        return raw.decode("utf-8")
    except UnicodeDecodeError:
        return raw.decode("latin-1")


def parse(text: str) -> TextGrid:
    toks = [t for t in _TOKEN.findall(text) if not t.startswith("[")]
    pos = 0

    def nxt():
        nonlocal pos
        if pos >= len(toks):
            raise ValueError("truncated TextGrid")
        t = toks[pos]; pos += 1
        return t

    def num():
        t = nxt()
        if t.startswith('"') or t.startswith("<"):
            raise ValueError(f"expected a number in TextGrid, got {t!r}")
        return round(float(t), ROUND_DIGITS)

    def string():
        t = nxt()
        if not t.startswith('"'):
            raise ValueError(f"expected a string in TextGrid, got {t!r}")
        return t[1:-1].replace('""', '"')

    if string() != "ooTextFile" or string() != "TextGrid":
        raise ValueError("not a TextGrid text file")
    xmin, xmax = num(), num()
    if nxt() != "<exists>":
        return TextGrid(xmin, xmax, [])
    n_tiers = int(float(nxt()))
    tiers = []
    for _ in range(n_tiers):
        kind, name = string(), string()
        tmin, tmax = num(), num()
        n = int(float(nxt()))
        tier = Tier(name, kind, tmin, tmax)
        for _ in range(n):
            if kind == "IntervalTier":
                a, b, mark = num(), num(), string()
                if a < b:                                   # textgrid 1.6.1 refuses non-positive intervals
                    tier.intervals.append((a, b, mark))
            else:
                tier.intervals.append((num(), string()))
        tiers.append(tier)
    return TextGrid(xmin, xmax, tiers)


def read(path) -> TextGrid:
    return parse(_decode(Path(path).read_bytes()))


def word_intervals(path_or_grid) -> list:
    """Tier 0 as [(minTime, maxTime, mark)] — what extract_words_and_pauses iterates over."""
    tg = path_or_grid if isinstance(path_or_grid, TextGrid) else read(path_or_grid)
    if not tg.tiers:
        raise IndexError("TextGrid has no tiers")
    return list(tg.tiers[0].intervals)


def write(path, tiers: dict, xmin: float = 0.0, xmax: float | None = None) -> None:
    """tiers: {name: [(tmin, tmax, mark), ...]} -> long-format TextGrid."""
    if xmax is None:
        xmax = max((iv[-1][1] for iv in tiers.values() if iv), default=0.0)
    esc = lambda s: s.replace('"', '""')
    out = ['File type = "ooTextFile"', 'Object class = "TextGrid"', "", f"xmin = {xmin}", f"xmax = {xmax}",
           "tiers? <exists>", f"size = {len(tiers)}", "item []:"]
    for k, (name, ivs) in enumerate(tiers.items(), 1):
        out += [f"    item [{k}]:", '        class = "IntervalTier"', f'        name = "{esc(name)}"', f"        xmin = {xmin}",
                f"        xmax = {xmax}", f"        intervals: size = {len(ivs)}"]
        for j, (a, b, mark) in enumerate(ivs, 1):
            out += [f"        intervals [{j}]:", f"            xmin = {a}", f"            xmax = {b}", f'            text = "{esc(mark)}"']
    Path(path).write_text("\n".join(out) + "\n", encoding="utf-8")


STATUS_OK, STATUS_UNREADABLE, STATUS_NOT_TEXTGRID, STATUS_NO_TIER = 0, 1, 2, 3


def read_tier_batch(paths, tier: int = 0, threads: int = 0, lib=None):
    """One tier of many TextGrid files through the native batch parser (pb_textgrid_parse_files, host threads).
    -> (status per file, [[(tmin, tmax, mark), ...] per file]); a file with status != 0 has an empty list."""
    import ctypes as C

    import numpy as np

    from . import _native as N
    lib = lib if lib is not None else N.load()
    enc = [str(p).encode() for p in paths]
    arr = (C.c_char_p * len(enc))(*enc)
    h = C.c_void_p()
    N.check(lib, None, lib.pb_textgrid_parse_files(arr, len(enc), int(tier), int(threads), C.byref(h)), "pb_textgrid_parse_files")
    try:
        nf, ni, mb = C.c_int64(), C.c_int64(), C.c_int64()
        lib.pb_textgrid_sizes(h, C.byref(nf), C.byref(ni), C.byref(mb))
        st = np.zeros(nf.value, np.int32); off = np.zeros(nf.value + 1, np.int64)
        t0 = np.zeros(ni.value); t1 = np.zeros(ni.value); mo = np.zeros(ni.value + 1, np.int64)
        pool = C.create_string_buffer(max(1, mb.value))
        dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        N.check(lib, None, lib.pb_textgrid_copy(h, st.ctypes.data_as(ip), None, None, off.ctypes.data_as(lp), t0.ctypes.data_as(dp),
                                                t1.ctypes.data_as(dp), mo.ctypes.data_as(lp), pool), "pb_textgrid_copy")
    finally:
        lib.pb_textgrid_free(h)
    raw = pool.raw
    out = []
    for f in range(nf.value):
        a, b = int(off[f]), int(off[f + 1])
        out.append([(float(t0[k]), float(t1[k]), raw[mo[k]:mo[k + 1]].decode("utf-8")) for k in range(a, b)])
    return st, out


def whisper_json_to_intervals(data: dict) -> tuple[list, float]:
    """The word tier that `json_to_textgrid` builds from a whisper-timestamped result
    (/root/reference/Code/Aligners/use_whisper_timestamped.py:330-395): a word whose start >= end gets end = start + 0.01,
    a " " interval fills every gap before a word, "[*]" inside a word becomes " ", and a result without any word yields a
    single "..." interval up to the last segment's end (1.0 s if unknown).  -> ([(tmin, tmax, mark)], maxTime)."""
    out, current, n_words = [], 0.0, 0
    for segment in data["segments"]:
        for word in segment["words"]:
            n_words += 1
            start, end = word["start"], word["end"]
            if start >= end:
                end = start + 0.01
            if start > current:
                out.append((current, start, " "))
            out.append((start, end, word["text"].replace("[*]", " ")))
            current = end
    if n_words == 0:
        xmax = 1.0
        if data["segments"] and "end" in data["segments"][-1]:
            xmax = data["segments"][-1]["end"]
        out.append((0.0, xmax, "..."))
        current = xmax
    return out, current


def write_whisper_textgrid(path, data: dict) -> None:
    """whisper-timestamped JSON -> TextGrid with the single tier `words` (what the Whisper aligner step leaves on disk)."""
    ivs, xmax = whisper_json_to_intervals(data)
    write(path, {"words": ivs}, 0.0, xmax)
