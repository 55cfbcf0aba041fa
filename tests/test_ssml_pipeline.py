"""SSML emitters and the file-level drop-in: WAV + TextGrid files on disk -> the three CSVs of the reference step
(Code/audioPipeline.py:604-711), against the oracle's loop-by-loop restatement.  Kernels run in the SIMT emulator."""
import re
import types
import wave

import numpy as np
import pytest

from conftest import speechlike

POS = {"le": "DET", "de": "ADP", "et": "CCONJ", "il": "PRON"}
pos_of = lambda w: POS.get(w.lower().strip(",.?!"), "NOUN")
GRIDS = [
    [(0.0, 0.2, ""), (0.2, 0.55, "Il"), (0.55, 0.95, "mange."), (0.95, 1.25, "Et"), (1.25, 1.32, ""), (1.32, 1.78, "puis?"), (1.78, 2.0, "")],
    [(0.0, 0.05, ""), (0.05, 0.45, "L'homme"), (0.45, 0.9, "<&>"), (0.9, 1.4, ""), (1.4, 1.9, "rit")],
]


def _write_wav(path, pcm, sr):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(sr); w.writeframes(pcm.astype("<i2").tobytes())


def test_emitters_match_oracle_flow_byte_for_byte():
    from oracle import flow as F
    from prosody_b200 import ssml as SSML
    rng = np.random.default_rng(3)
    names, words, pauses = [], [], []
    texts = ["Bonjour le monde.", "", "il <dit> & \"rit\"", "fin!", "", "d'accord"]
    for k in range(60):
        t = texts[k % len(texts)]
        names.append(f"segment_ph{1 + k // 7}"); words.append(t); pauses.append(0 if t else int(rng.integers(0, 900)))
    pauses[0] = 300; pauses[3] = 49; pauses[9] = 50
    p = rng.normal(0, 8, 60); r = rng.normal(0, 10, 60); v = np.clip(rng.normal(0, 6, 60), -7, 7)
    p[5] = -0.0; r[6] = 0.004999; v[7] = 12.345
    rows = [dict(segment=n, syntagme=w, pause=pa, raw_pitch=0.0, raw_volume=float(vv), raw_rate=0.0) for n, w, pa, vv in zip(names, words, pauses, v)]
    for factor in (1, 0.5):
        f_ref, s_ref, y_ref = F.emit_ssml(rows, list(p), list(r), "fr-FR-HenriNeural", factor)
        f, s, y = SSML.build(names, words, pauses, p, r, v, "fr-FR-HenriNeural", factor)
        assert f == f_ref and s == s_ref and y == y_ref
    assert '<break time="300ms"/>' in f[0]["ssml"] and "&lt;dit&gt; &amp;" in f[0]["ssml"]
    assert re.search(r'pitch="[+-]\d+\.\d\d%" rate="[+-]\d+\.\d\d%" volume="[+-]\d+\.\d\d%"', s[0]["ssml"])


def test_file_level_dropin_writes_the_reference_csvs(tmp_path, emu_lib, oracle):
    import pandas as pd
    import prosody_b200 as pb
    from oracle import flow as F
    from prosody_b200 import pipeline as P
    from prosody_b200 import textgrid as TG
    voice = tmp_path / "Data" / "voice" / "v"
    (voice / "audio").mkdir(parents=True); (voice / "tg").mkdir(); raw = tmp_path / "Data" / "voice" / "v_raw" / "audio"; raw.mkdir(parents=True)
    nat = speechlike(2, 2.0, 16000, seed=51); syn = speechlike(2, 1.9, 24000, seed=52)
    fsegs = []
    for i in range(2):
        name = f"segment_ph{i + 2}"
        _write_wav(voice / "audio" / f"{name}.wav", nat[i], 16000)
        if i == 0:
            _write_wav(raw / f"{name}.wav", syn[i], 24000)          # segment_ph3 has no raw twin: natural fallback
        TG.write(voice / "tg" / f"{name}.TextGrid", {"words": GRIDS[i]})
        fsegs.append(F.Segment(name, nat[i], 16000, syn[i] if i == 0 else None, 24000 if i == 0 else None, GRIDS[i]))
    res = tmp_path / "Out"
    self = types.SimpleNamespace(
        voice_dir=voice, raw_audio_dir=raw, textgrid_dir=voice / "tg", p_st=1.3, pitch_lower_clip_factor=0.7, v_pct=7.0, r_pct_clamp=15.0,
        alpha=0.2, max_jump=5.0, end_pause_ms=400, baseline_window=None, inter_syntagme_pause_factor=1,
        threshold_duration_before_slowing_down=1.0, slow_floor_per_sec=2.0, azure_voice="fr-FR-HenriNeural",
        bdd_ssml_csv=res / "BDD_ssml.csv", bdd_syntagme_ssml_csv=res / "BDD_syntagme_ssml.csv", bdd_syntagme_synth_csv=res / "BDD_syntagme_for_synth.csv")
    with pb.Extractor(0, lib=emu_lib) as ex:
        out = P.measure_prosody_and_build_ssml(self, extractor=ex, pos_of=pos_of)
        # the scalar closures agree with the batched step
        assert P.get_duration(voice / "audio" / "segment_ph2.wav") == 2.0
        assert abs(P.get_lufs(voice / "audio" / "segment_ph2.wav", P.Meter(16000), extractor=ex) - out["seg_stats"]["l_nat"][0]) < 1e-12
        assert P.get_median_pitch(voice / "audio" / "segment_ph2.wav", extractor=ex) == out["seg_stats"]["p_nat"][0]
        with pytest.raises(pb.step.PraatError):
            P.get_median_pitch(voice / "audio" / "segment_ph2.wav", 0.5, 0.51, extractor=ex)
    prm = dict(pitch_semitones=1.3, smoothing_alpha=0.2, end_punctuation_pause_ms=400)
    ref = F.measure_and_build(fsegs, prm, pos_of, "fr-FR-HenriNeural")
    got_seg = pd.read_csv(self.bdd_ssml_csv); got_syn = pd.read_csv(self.bdd_syntagme_ssml_csv, keep_default_na=False)
    got_synth = pd.read_csv(self.bdd_syntagme_synth_csv, keep_default_na=False)
    assert list(got_seg.columns) == ["segment", "ssml"] and list(got_syn.columns) == ["segment", "syntagme", "pause", "ssml"] == list(got_synth.columns)
    assert list(got_syn["syntagme"]) == [r["syntagme"] for r in ref["bdd_syntagme_ssml"]]
    assert list(got_syn["pause"]) == [r["pause"] for r in ref["bdd_syntagme_ssml"]]
    num = re.compile(r'pitch="([+-][\d.]+)%" rate="([+-][\d.]+)%" volume="([+-][\d.]+)%"')
    strip = lambda s: num.sub("P", s)
    for got, want in ((got_syn, ref["bdd_syntagme_ssml"]), (got_synth, ref["bdd_syntagme_synth"])):
        for g, w in zip(got["ssml"], [r["ssml"] for r in want]):
            assert strip(g) == strip(w)                                  # text, escaping, break tags, wrappers: identical
            (gp, gr, gv), (wp, wr, wv) = num.search(g).groups(), num.search(w).groups()
            assert gr == wr and gv == wv                                 # float64 paths: identical strings
            assert abs(float(gp) - float(wp)) <= 0.02                    # pitch carries the FP32 F0 (documented rounding boundary)
    assert [strip(s) for s in got_seg["ssml"]] == [strip(r["ssml"]) for r in ref["bdd_ssml"]]


def test_native_csv_tables_equal_the_pandas_writer_byte_for_byte(tmp_path, native_lib):
    """pb_ssml_csv: the three tables formatted natively are the bytes pandas.DataFrame(rows).to_csv(index=False) writes for the
    Python emitters' rows (which are checked against the oracle's restatement of audioPipeline.py:604-711 above) — incl. XML
    escapes, CSV quoting of commas / quotes / line breaks, non-ASCII text, the .2f lattice (-0.00, ties), NaN / inf, break tags
    on both sides of the 50 ms threshold and after sentence-final punctuation, a fractional pause factor, and a segment name that
    comes back after another one."""
    from prosody_b200 import ssml as SSML
    rng = np.random.default_rng(11)
    texts = ["Bonjour le monde.", "", 'il <dit> & "rit", puis', "fin!", "", "d'accord", "garçon naïf à l'école?", "ligne\nbrisée", "x,y", "ça va ?"]
    names, words, pauses = [], [], []
    for k in range(400):
        t = texts[k % len(texts)]
        names.append(f"segment_ph{1 + (k // 9) % 17}"); words.append(t); pauses.append(0 if t else int(rng.integers(0, 900)))
    pauses[0] = 300; pauses[3] = 49; pauses[13] = 50; pauses[6] = 777
    p = rng.normal(0, 8, 400); r = rng.normal(0, 10, 400); v = np.clip(rng.normal(0, 6, 400), -7, 7)
    p[5] = -0.0; r[6] = 0.004999; v[7] = 12.345; p[8] = 0.005; p[9] = 0.015; p[10] = -0.025; r[11] = 1e-9; p[12] = float("nan"); r[12] = float("inf"); v[12] = -float("inf")
    p[13] = 123456.789; r[14] = -0.004999999
    for factor in (1, 0.5, 1.7):
        final, syn_rows, synth_rows = SSML.build(names, words, pauses, p, r, v, "fr-FR-HenriNeural", factor)
        SSML.write_csvs(final, syn_rows, synth_rows, tmp_path / "a.csv", tmp_path / "b.csv", tmp_path / "c.csv")
        got = SSML.build_csv_bytes(SSML.TextPools(names, words), pauses, p, r, v, "fr-FR-HenriNeural", factor, lib=native_lib)
        for g, f in zip(got, ("a.csv", "b.csv", "c.csv")):
            assert g == (tmp_path / f).read_bytes(), (factor, f)
    # empty input: headers only would need a frame with columns; the reference never gets here (KeyError on the empty frame, :593)
    one = SSML.build_csv_bytes(SSML.TextPools(["s"], ["mot"]), [0], [1.0], [2.0], [3.0], "v", 1, lib=native_lib)
    assert one[0].startswith(b"segment,ssml\ns,") and one[1].count(b"\n") == 2


def test_native_two_decimal_formatter_on_the_rounding_lattice(native_lib):
    """pb_ssml_csv formats its percentages without the C library (exact integer arithmetic on the double's mantissa): every multiple
    of 0.005 in [-40, 40] (the ties and near-ties of the :+.2f lattice), both neighbouring doubles of each, tiny, huge and subnormal
    values must print exactly like Python's f'{x:+.2f}' (audioPipeline.py:610-612)."""
    import re
    from prosody_b200 import ssml as SSML
    base = np.arange(-8000, 8001) * 0.005
    vals = np.concatenate([base, np.nextafter(base, 1e9), np.nextafter(base, -1e9), np.arange(-8000, 8001) / 800.0,
                           [0.0, -0.0, 5e-324, -5e-324, 1e-300, 0.125, 0.375, 2.675, 1.005, 999999999999999.0, -123456789.125, 1e15, -3e18, 7.5e22]])
    n = len(vals)
    tables = SSML.build_csv_bytes(SSML.TextPools(["s"] * n, ["mot"] * n), [0] * n, vals, vals[::-1].copy(), -vals, "v", 1, lib=native_lib)
    rows = tables[1].decode().splitlines()[1:]
    assert len(rows) == n
    pat = re.compile(r'pitch=""([^"]*)%"" rate=""([^"]*)%"" volume=""([^"]*)%""')
    for k, row in enumerate(rows):
        got = pat.search(row).groups()
        assert got == (f"{vals[k]:+.2f}", f"{vals[n - 1 - k]:+.2f}", f"{-vals[k]:+.2f}"), (k, vals[k], got)
