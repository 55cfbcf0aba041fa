"""BASELINE config 1: the reference's ten example clips (real French speech, 44.1 kHz) through the step, with the reference's
own pitch parameters (floor 150 / ceiling 600, Code/audioPipeline.py:329-332).

CPU: the fixture builder is deterministic and the committed golden files are what the oracle's restatement of the step produces.
GPU: the FILE-LEVEL drop-in (prosody_b200.pipeline.measure_prosody_and_build_ssml: WAV + TextGrid on disk -> three CSVs) against
those golden files: text / break tags / rate / volume strings identical, pitch strings identical except `:+.2f` lattice flips,
which are counted and listed (tests/parity_report.py)."""
import hashlib
import json
import re
import sys
import types
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden"
sys.path.insert(0, str(GOLD))

NUM = re.compile(r'pitch="([+-][\d.]+)%" rate="([+-][\d.]+)%" volume="([+-][\d.]+)%"')


def _self(v, res):
    return types.SimpleNamespace(
        voice_dir=v["voice_dir"], raw_audio_dir=v["raw_audio_dir"], textgrid_dir=v["textgrid_dir"], p_st=1.3, pitch_lower_clip_factor=0.7,
        v_pct=7.0, r_pct_clamp=15.0, alpha=0.2, max_jump=5.0, end_pause_ms=400, baseline_window=None, inter_syntagme_pause_factor=1,
        threshold_duration_before_slowing_down=1.0, slow_floor_per_sec=2.0, azure_voice="fr-FR-HenriNeural",
        bdd_ssml_csv=res / "BDD_ssml.csv", bdd_syntagme_ssml_csv=res / "BDD_syntagme_ssml.csv", bdd_syntagme_synth_csv=res / "BDD_syntagme_for_synth.csv")


def test_c1_fixture_layout_and_clips(tmp_path):
    import make_c1_fixture as C1
    clips = C1.clip_paths()
    assert [p.name for p in clips] == [f"segment_ph{k}.wav" for k in range(2, 12)]          # sorted by number, like :364-367
    total = 0.0
    for p in clips:
        pcm, sr = C1.read_wav(p)
        assert sr == 44100
        total += len(pcm) / sr
    assert abs(total - 161.85) < 0.01
    manifest = json.loads((GOLD / "c1_oracle" / "manifest.json").read_text())
    v = C1.build_voice(tmp_path)
    for f in sorted(tmp_path.rglob("*")):
        if f.is_file():
            assert hashlib.sha1(f.read_bytes()).hexdigest() == manifest[str(f.relative_to(tmp_path))], f
    assert len(v["segments"]) == 10


def test_c1_oracle_reproduces_the_golden_files(tmp_path, oracle):
    """The committed golden CSVs ARE the oracle's output (regenerated here and compared byte for byte)."""
    import pandas as pd
    import make_c1_oracle_golden as G
    v, ref = G.run(tmp_path)
    for key, name in (("bdd_ssml", "BDD_ssml.csv"), ("bdd_syntagme_ssml", "BDD_syntagme_ssml.csv"), ("bdd_syntagme_synth", "BDD_syntagme_for_synth.csv")):
        pd.DataFrame(ref[key]).to_csv(tmp_path / name, index=False)
        assert (tmp_path / name).read_bytes() == (GOLD / "c1_oracle" / name).read_bytes(), name
    vals = json.loads((GOLD / "c1_oracle" / "values.json").read_text())
    assert [s["p_nat"] for s in vals["seg_stats"]] == [s["p_nat"] for s in ref["seg_stats"]]
    # real speech sanity: every clip is voiced, medians in a speaking range, loudness near the corpus' -17 LUFS
    assert all(120.0 < s["p_nat"] < 260.0 for s in ref["seg_stats"])
    assert all(-20.0 < s["l_nat"] < -14.0 for s in ref["seg_stats"])


@pytest.mark.gpu
def test_c1_file_level_dropin_on_real_speech(tmp_path, gpu_extractor):
    import pandas as pd
    import make_c1_fixture as C1
    import make_c1_oracle_golden as G
    from parity_report import column_report
    from prosody_b200 import pipeline as P
    v = C1.build_voice(tmp_path / "data")
    res = tmp_path / "Out"
    self = _self(v, res)
    out = P.measure_prosody_and_build_ssml(self, extractor=gpu_extractor, pos_of=G.pos_of)
    vals = json.loads((GOLD / "c1_oracle" / "values.json").read_text())
    # ---- per segment (pass 1)
    p_ref = np.array([s["p_nat"] for s in vals["seg_stats"]]); l_ref = np.array([s["l_nat"] for s in vals["seg_stats"]])
    ls_ref = np.array([s["l_syn"] for s in vals["seg_stats"]])
    assert np.max(np.abs(out["seg_stats"]["p_nat"] - p_ref) / p_ref) < 1e-4
    assert np.max(np.abs(out["seg_stats"]["l_nat"] - l_ref)) < 1e-9 and np.max(np.abs(out["seg_stats"]["l_syn"] - ls_ref)) < 1e-9
    assert [float(s["d_nat"]) for s in vals["seg_stats"]] == list(out["seg_stats"]["d_nat"])
    # ---- per syntagme (pass 2)
    pu = np.array([u["p_nat"] for u in vals["units"]]); got = out["syn"]["p_nat"]
    assert np.array_equal(pu > 0, got > 0)                                    # same voiced / unvoiced units
    both = pu > 0
    assert np.max(np.abs(got[both] - pu[both]) / pu[both]) < 2e-4
    assert np.max(np.abs(out["syn"]["l_syn"] - np.array([u["l_syn"] for u in vals["units"]]))) < 1e-9
    assert list(out["syn"]["nat_total"]) == [u["nat_total"] for u in vals["units"]] and list(out["syn"]["syn_total"]) == [u["syn_total"] for u in vals["units"]]
    # ---- the three CSVs
    strip = lambda s: NUM.sub("P", s)
    flips = 0
    for name in ("BDD_syntagme_ssml.csv", "BDD_syntagme_for_synth.csv"):
        got_df = pd.read_csv(res / name, keep_default_na=False); ref_df = pd.read_csv(GOLD / "c1_oracle" / name, keep_default_na=False)
        assert list(got_df.columns) == list(ref_df.columns) and len(got_df) == len(ref_df) == 115
        assert list(got_df["segment"]) == list(ref_df["segment"]) and list(got_df["syntagme"]) == list(ref_df["syntagme"]) and list(got_df["pause"]) == list(ref_df["pause"])
        for g, w in zip(got_df["ssml"], ref_df["ssml"]):
            assert strip(g) == strip(w)
            (gp, gr, gv), (wp, wr, wv) = NUM.search(g).groups(), NUM.search(w).groups()
            assert gr == wr and gv == wv                                      # float64 paths: identical strings
            assert abs(float(gp) - float(wp)) <= 0.0101                       # pitch: at most one lattice step
            flips += gp != wp
    seg_got = pd.read_csv(res / "BDD_ssml.csv"); seg_ref = pd.read_csv(GOLD / "c1_oracle" / "BDD_ssml.csv")
    assert [strip(s) for s in seg_got["ssml"]] == [strip(s) for s in seg_ref["ssml"]]
    rep = column_report(out["sm_pitch"], vals["sm_pitch"])
    print("C1 pitch strings:", {k: rep[k] for k in ("rows", "identical", "flipped", "max_abs_delta")}, rep["listed"][:5])
    assert rep["flipped"] <= 6, rep["listed"]                                 # documented rounding boundary: a few of 115 rows
    assert column_report(out["sm_rate"], vals["sm_rate"])["flipped"] == 0 and column_report(out["raw_volume"], [r["raw_volume"] for r in vals["raw_rows"]])["flipped"] == 0


@pytest.mark.gpu
def test_c1_frame_level_tolerances_on_real_speech(tmp_path, gpu_extractor):
    """BASELINE north_star states its tolerances per FRAME: F0 within 0.5 % on frames voiced in both implementations, voicing
    decisions agreeing on >= 99.5 % of frames, intensity within 0.05 dB.  Every frame of the reference's ten clips (44.1 kHz, floor
    150 / ceiling 600: the reference's own parameters) against the oracle — and far inside those bars."""
    sys.path.insert(0, str(GOLD.parent.parent))
    import bench
    import make_c1_fixture as C1
    from prosody_b200 import pipeline as P
    from prosody_b200 import step as S
    v = C1.build_voice(tmp_path / "data")
    pcm, segs = P.load_voice(v["voice_dir"] / "audio", v["raw_audio_dir"], v["textgrid_dir"])
    rep = bench.frame_level_parity(gpu_extractor, pcm, segs, dict(S.REFERENCE_PITCH))
    assert rep["frames"] > 30000 and rep["voiced_in_both"] > 5000      # floor 150 Hz: most of this speaker's frames are below it
    assert rep["voicing_agreement"] >= 0.9999                 # north_star: 0.995
    assert rep["frames_over_0p5_percent"] == 0 and rep["f0_rel_err_max"] < 5e-3
    assert rep["f0_rel_err_p99"] < 2e-4 and rep["strength_abs_err_max"] < 2e-3
    assert rep["frame_intensity_err_max_db"] < 0.01           # north_star: 0.05 dB
