"""Host logic of the step (intervals, unit planning, baselines, deltas, smoothing) against the loop-by-loop oracle
restatement of Code/audioPipeline.py:261-711, with the kernels running in the SIMT emulator (CPU, tiny inputs)."""
import numpy as np
import pytest

from conftest import speechlike

POS = {"le": "DET", "la": "DET", "de": "ADP", "et": "CCONJ", "que": "SCONJ", "il": "PRON"}
pos_of = lambda w: POS.get(w.lower().strip(",.?!"), "NOUN")

GRID_A = [(0.0, 0.08, ""), (0.08, 0.41, "Bonjour,"), (0.41, 0.62, "le"), (0.62, 0.70, ""), (0.70, 1.02, "chat"), (1.02, 1.31, "dort."),
          (1.31, 1.50, ""), (1.50, 1.83, "de,"), (1.83, 2.20, "jour")]
GRID_B = [(0.0, 0.21, ""), (0.21, 0.55, "Il"), (0.55, 0.95, "mange."), (0.95, 1.25, "Et"), (1.25, 1.32, ""), (1.32, 1.78, "puis?"), (1.78, 2.0, "")]


def test_interval_logic_matches_oracle_flow():
    from oracle import flow as F
    from prosody_b200 import intervals as IV
    for grid in (GRID_A, GRID_B):
        assert IV.words_and_pauses(grid) == F.words_and_pauses(grid)
        for ep in (150, 400):
            seq = IV.segment_sequence(grid, pos_of, ep)
            assert seq == F.segment_sequence(grid, pos_of, ep)
            mine = IV.syntagmes(seq)
            ref = F.build_syntagmes(seq)
            assert [(d["words"], d["start_ms"], d["end_ms"], d["pause_ms"]) for d in ref] == mine
    assert IV.strip_spurious_commas("de, la, maison", pos_of) == F.strip_spurious_commas("de, la, maison", pos_of) == "delamaison"


def test_textgrid_roundtrip(tmp_path):
    from prosody_b200 import textgrid as TG
    p = tmp_path / "a.TextGrid"
    TG.write(p, {"words": [(0.0, 0.5, 'dit "oui"'), (0.5, 0.5, "x"), (0.5, 1.25, "")]}, xmax=1.25)
    ivs = TG.word_intervals(p)
    assert ivs == [(0.0, 0.5, 'dit "oui"'), (0.5, 1.25, "")]          # the empty-length interval is dropped
    short = 'File type = "ooTextFile"\nObject class = "TextGrid"\n\n0\n1.25\n<exists>\n1\n"IntervalTier"\n"words"\n0\n1.25\n2\n0\n0.5\n"a"\n0.5\n1.25\n""\n'
    assert TG.parse(short).tiers[0].intervals == [(0.0, 0.5, "a"), (0.5, 1.25, "")]


@pytest.mark.parametrize("window", [None, 2])
def test_step_matches_oracle_flow_on_emulator(emu_lib, oracle, window):
    import prosody_b200 as pb
    from prosody_b200 import step as S
    from oracle import flow as F
    nat = speechlike(3, 2.2, 16000, seed=31)
    syn = speechlike(3, 2.0, 24000, seed=32)
    bufs, segs, fsegs, off = [], [], [], 0
    grids = [GRID_A, GRID_B, GRID_A]
    for i in range(3):
        has_syn = i != 1                                   # segment 1: undecodable synth -> natural fallback
        n_off = off; bufs.append(nat[i]); off += len(nat[i])
        s_off = None
        if has_syn:
            s_off = off; bufs.append(syn[i]); off += len(syn[i])
        segs.append(S.Segment(f"segment_ph{i+1}", n_off, len(nat[i]), 16000, grids[i], s_off, len(syn[i]) if has_syn else None, 24000 if has_syn else None))
        fsegs.append(F.Segment(f"segment_ph{i+1}", nat[i], 16000, syn[i] if has_syn else None, 24000 if has_syn else None, grids[i]))
    pcm = np.concatenate(bufs)
    prm = dict(baseline_window=window, pitch_semitones=1.3, smoothing_alpha=0.2, end_punctuation_pause_ms=400)
    ref = F.measure_and_build(fsegs, prm, pos_of)
    with pb.Extractor(0, lib=emu_lib) as ex:
        pl = S.plan(segs, prm, pos_of)
        out = S.measure(ex, pcm, pl, prm)
    assert pl.n_syn == len(ref["raw_rows"])
    assert pl.syn_words == [r["syntagme"] for r in ref["raw_rows"]]
    assert list(pl.syn_pause_ms) == [r["pause"] for r in ref["raw_rows"]]
    # loudness / durations are float64 on both sides; pitch carries the FP32 kernel tolerance (0.5 %)
    np.testing.assert_allclose(out["seg_stats"]["l_nat"], [s["l_nat"] for s in ref["seg_stats"]], atol=1e-9)
    np.testing.assert_allclose(out["seg_stats"]["l_syn"], [s["l_syn"] for s in ref["seg_stats"]], atol=1e-9)
    assert list(out["seg_stats"]["d_syn"]) == [s["d_syn"] for s in ref["seg_stats"]]
    np.testing.assert_allclose(out["seg_stats"]["p_nat"], [s["p_nat"] for s in ref["seg_stats"]], rtol=5e-3)
    np.testing.assert_allclose(out["syn"]["l_syn"], [u["l_syn"] for u in ref["units"]], atol=1e-9)
    assert list(out["syn"]["nat_total"]) == [u["nat_total"] for u in ref["units"]]
    assert list(out["syn"]["syn_total"]) == [u["syn_total"] for u in ref["units"]]
    np.testing.assert_allclose(out["raw_volume"], [r["raw_volume"] for r in ref["raw_rows"]], atol=1e-9)
    np.testing.assert_allclose(out["raw_rate"], [r["raw_rate"] for r in ref["raw_rows"]], atol=1e-12)
    np.testing.assert_allclose(out["raw_pitch"], [r["raw_pitch"] for r in ref["raw_rows"]], atol=0.6)   # 0.5 % of F0 ~ 0.5 pct-points
    np.testing.assert_allclose(out["sm_rate"], ref["sm_r"], atol=1e-12)


def test_delta_and_smoothing_helpers_match_python_arithmetic(native_lib):
    """pb_syntagme_deltas / pb_ema_clamp (host C, no GPU) vs the reference's scalar Python expressions, bit for bit."""
    import ctypes as C
    from oracle import flow as F
    from prosody_b200 import _native as N
    rng = np.random.default_rng(9)
    n = 4000
    p_nat = np.where(rng.random(n) < 0.2, 0.0, rng.uniform(80, 400, n)); f0 = rng.uniform(100, 300, n)
    loud = rng.uniform(-30, -10, n); l_syn = np.where(rng.random(n) < 0.02, -np.inf, rng.uniform(-35, -8, n))
    wc = rng.integers(0, 9, n).astype(np.int32); nat = rng.uniform(0.0, 7.0, n); syn = rng.uniform(0.0, 7.0, n)
    pause = np.where(rng.random(n) < 0.3, rng.integers(0, 900, n), 0).astype(np.int32)
    prm = dict(F.DEFAULT_PARAMS)
    dp = N.PbDeltaParams(prm["pitch_semitones"], prm["pitch_lower_clip_factor"], prm["volume_pct"], prm["rate_percent"],
                         prm["threshold_duration_before_slowing_down"], prm["slow_floor_per_sec"])
    rp, rv, rr = np.empty(n), np.empty(n), np.empty(n)
    d = lambda a: a.ctypes.data_as(C.POINTER(C.c_double)); ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    assert native_lib.pb_syntagme_deltas(n, d(p_nat), d(f0), d(loud), d(l_syn), ip(wc), d(nat), d(syn), ip(pause), C.byref(dp), d(rp), d(rv), d(rr)) == 0
    for i in range(n):
        e = F.syntagme_deltas(p_nat[i], dict(f0=f0[i], loud=loud[i]), l_syn[i], int(wc[i]), nat[i], syn[i], int(pause[i]), prm)
        assert (rp[i], rv[i], rr[i]) == e, (i, (rp[i], rv[i], rr[i]), e)
    x = rng.normal(0, 6, 3000); out = np.empty(3000)
    assert native_lib.pb_ema_clamp(d(x), 3000, 0.4, 5.0, d(out)) == 0
    assert list(out) == F.smooth(list(x), 0.4, 5.0)


def test_segment_baselines_match_numpy_medians():
    """pb_segment_baselines against the reference's own expression (np.median over list comprehensions, :401-424)."""
    import warnings
    from prosody_b200 import step as S
    rng = np.random.default_rng(5)
    for n, win in ((1, None), (7, None), (7, 10), (40, 10), (41, 3), (200, 11), (64, 0), (30, 1)):
        p = rng.uniform(80, 300, n); p[rng.random(n) < 0.3] = 0.0
        if n >= 40:
            p[10:22] = 0.0                                         # a window without any voiced segment -> NaN baseline
        l = rng.uniform(-35, -12, n); r = rng.uniform(0.7, 1.3, n)
        if n > 5:
            l[3] = -np.inf
            l[5] = np.nan
        f0, loud, rate = S.baselines(p, l, r, win)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for i in range(n):
                if win is None or win >= n:
                    lo, hi = 0, n
                else:
                    lo, hi = max(0, i - win // 2), min(n, i + win // 2 + 1)
                ef = float(np.median([x for x in p[lo:hi] if x > 0])) or 1.0
                el = float(np.median(list(l[lo:hi]))); er = float(np.median(list(r[lo:hi])))
                for got, exp in ((f0[i], ef), (loud[i], el), (rate[i], er)):
                    assert (np.isnan(got) and np.isnan(exp)) or got == exp, (n, win, i, got, exp)


class _FakeSpacyTagger:
    """Tokenises like fr_core_news_sm does on elisions and inverted clitics (what the reference's `_nlp(text)` sees)."""
    TABLE = {"l'homme,": [("l'", "", "DET"), ("homme", "", "NOUN"), (",", "", "PUNCT")],
             "dit-il,": [("dit", "", "VERB"), ("-il", "", "PRON"), (",", "", "PUNCT")],
             "l'homme": [("l'", "", "DET"), ("homme", "", "NOUN")],
             "de, la": [("de", "", "ADP"), (",", " ", "PUNCT"), ("la", "", "DET")]}

    def tag(self, text):
        return self.TABLE[text]

    def __call__(self, word):
        return self.TABLE[word][0][2]


def test_comma_filter_uses_the_taggers_own_tokens():
    from prosody_b200 import intervals as IV
    """ADVICE r1: the reference tags the whole mark and looks at the real previous spaCy token (audioPipeline.py:70-79)."""
    tg = _FakeSpacyTagger()
    assert IV.strip_spurious_commas("l'homme,", tg) == "l'homme,"          # previous token is NOUN "homme": comma stays
    assert IV.strip_spurious_commas("dit-il,", tg) == "dit-il"             # previous token is PRON "-il": comma goes
    assert IV.strip_spurious_commas("de, la", tg) == "dela"                # text_with_ws of the dropped comma goes with it
    assert IV.first_pos(" l'homme ", tg) == "DET"                          # `_nlp(ptok.strip())[0].pos_`
    # the pause after "l'homme" is dropped (first token DET), as the reference's filter does
    grid = [(0.0, 0.5, "l'homme"), (0.5, 0.9, ""), (0.9, 1.4, "dit-il,")]
    seq = IV.segment_sequence(grid, tg, 150)
    assert [k for k, _, _ in seq] == ["word", "word"] and seq[1][1] == "dit-il"


def test_submit_wait_is_measure_in_two_halves(emu_lib):
    """pb_extract_submit / pb_extract_wait: same results as the blocking call, one pending batch per handle, and two handles
    keep two batches in flight (here on the emulator, where 'in flight' only means 'not collected yet')."""
    import prosody_b200 as pb
    from prosody_b200 import step as S
    nat = speechlike(2, 1.2, 16000, seed=61); syn = speechlike(2, 1.1, 16000, seed=62)
    bufs, segs, off = [], [], 0
    for i, grid in enumerate((GRID_A[:5], GRID_B[:4])):
        bufs += [nat[i], syn[i]]
        segs.append(S.Segment(f"segment_ph{i + 1}", off, len(nat[i]), 16000, grid, off + len(nat[i]), len(syn[i]), 16000))
        off += len(nat[i]) + len(syn[i])
    pcm = np.concatenate(bufs)
    pl = S.plan(segs, None, pos_of)
    with pb.Extractor(0, lib=emu_lib) as a, pb.Extractor(0, lib=emu_lib) as b:
        want = S.measure(a, pcm, pl)
        S.submit(a, pcm, pl)
        S.submit(b, pcm, pl)                                   # a second handle takes the next batch meanwhile
        with pytest.raises(Exception):
            S.submit(a, pcm, pl)                               # one pending batch per handle
        with pytest.raises(Exception):
            a.lufs(pcm, pl.units)                              # nor any other work on that handle
        got_a = S.collect(a, pl); got_b = S.collect(b, pl)
        with pytest.raises(Exception):
            a.wait()
        for got in (got_a, got_b):
            for k in ("raw_pitch", "raw_volume", "raw_rate", "sm_pitch", "sm_rate"):
                assert np.array_equal(got[k], want[k], equal_nan=True)
            assert np.array_equal(got["status"], want["status"])
        assert np.array_equal(S.measure(a, pcm, pl)["sm_pitch"], want["sm_pitch"])      # the handle is reusable afterwards
