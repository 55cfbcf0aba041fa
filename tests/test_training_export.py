"""Training-JSON export against golden files produced by the REFERENCE's own create_training_data.py (tests/golden/
make_training_json_golden.py, run in the build container) from a CSV written by our ssml.write_csvs."""
import json
import shutil
from pathlib import Path

GOLD = Path(__file__).resolve().parent / "golden"


def test_ssml_csv_is_reproduced_and_export_matches_reference_bytes(tmp_path):
    import prosody_b200  # noqa: F401
    from prosody_b200 import ssml, training_export as TE
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", GOLD / "make_training_json_golden.py")
    src = (GOLD / "make_training_json_golden.py").read_text()
    rows = eval(src[src.index("ROWS = ["):src.index("]\n\n\ndef main")].split("=", 1)[1] + "]")      # the fixed rows only
    segs, words, pauses, p, r, v = zip(*rows)
    final, syn_rows, synth_rows = ssml.build(segs, words, pauses, p, r, v, "fr-FR-HenriNeural", factor=0.8)
    ssml.write_csvs(final, syn_rows, synth_rows, tmp_path / "a.csv", tmp_path / "b.csv", tmp_path / "c.csv")
    assert (tmp_path / "b.csv").read_bytes() == (GOLD / "bdd_syntagme_ssml_golden.csv").read_bytes()
    assert (tmp_path / "a.csv").read_bytes() == (GOLD / "bdd_ssml_golden.csv").read_bytes()
    out = tmp_path / "results" / "VoiceA" / "training_data_VoiceA.json"
    TE.create_training_data(str(GOLD / "bdd_syntagme_ssml_golden.csv"), str(out))
    assert out.read_bytes() == (GOLD / "training_data_golden.json").read_bytes()
    TE.combine_training_jsons(str(tmp_path / "results"), str(tmp_path / "results" / "bdd.json"))
    assert (tmp_path / "results" / "bdd.json").read_bytes() == (GOLD / "bdd_golden.json").read_bytes()
    # the consumers' parse (Code/baseline_models/bilstm.py:46-48) reads back what the step computed, to 2 decimals
    data = json.loads(out.read_text(encoding="utf-8"))
    texts = [e for e in data["y"]["parsed_sequence"] if e["type"] == "text"]
    want = [(pp, rr, vv) for (_, w, _, pp, rr, vv) in rows if w]
    assert len(texts) == len(want)
    for e, (pp, rr, vv) in zip(texts, want):
        assert abs(TE.parse_percent(e["prosody"]["pitch"]) - pp) <= 0.005 + 1e-12
        assert abs(TE.parse_percent(e["prosody"]["rate"]) - rr) <= 0.005 + 1e-12
        assert abs(TE.parse_percent(e["prosody"]["volume"]) - vv) <= 0.005 + 1e-12


def test_export_errors_like_the_reference(tmp_path):
    import pytest
    import prosody_b200  # noqa: F401
    from prosody_b200 import training_export as TE
    with pytest.raises(FileNotFoundError):
        TE.create_training_data(str(tmp_path / "missing.csv"), str(tmp_path / "o" / "x.json"))
    (tmp_path / "empty.csv").write_text("segment,syntagme,pause,ssml\nseg,txt,0,no ssml here\n", encoding="utf-8")
    with pytest.raises(ValueError):
        TE.create_training_data(str(tmp_path / "empty.csv"), str(tmp_path / "o" / "x.json"))
