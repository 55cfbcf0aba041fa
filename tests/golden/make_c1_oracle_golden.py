"""Writes tests/golden/c1_oracle/: what the ORACLE's loop-by-loop restatement of the step (oracle/flow.py <- Code/audioPipeline.py:261-711)
produces for BASELINE config 1 (the ten repo clips + the make_c1_fixture.py TextGrids / raw twins), with the YAML's prosody
settings (config.yaml:24-44: pitch_semitones 1.3, smoothing_alpha 0.2, end_punctuation_pause_ms 400) and a fixed POS table
standing in for spaCy.  The three CSVs are written with the reference's own writer call (pandas to_csv(index=False)); values.json
keeps the numbers behind them (per segment, per syntagme) at full precision.

The real parselmouth / pyloudnorm / pydub cannot be installed here (profiles/r02_pip_real_packages.log), so this pins the GPU
path and the oracle to EACH OTHER on real speech, not to Praat itself: parity stays "unpinned" in the sense of DESIGN.md §2."""
from __future__ import annotations

import json
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent)); sys.path.insert(0, str(HERE))

PROSODY = dict(pitch_semitones=1.3, smoothing_alpha=0.2, end_punctuation_pause_ms=400)
VOICE = "fr-FR-HenriNeural"
POS = {"le": "DET", "la": "DET", "les": "DET", "une": "DET", "son": "DET", "de": "ADP", "dans": "ADP", "sur": "ADP", "avec": "ADP",
       "pendant": "ADP", "et": "CCONJ", "puis": "CCONJ", "que": "SCONJ", "qui": "PRON", "il": "PRON"}


def pos_of(w):
    return POS.get(w.lower().strip(",.?!"), "NOUN")


def run(tmp_root):
    import make_c1_fixture as C1
    from oracle import flow as F
    v = C1.build_voice(tmp_root)
    fsegs = [F.Segment(n, nat, nsr, syn, ssr, grid) for n, nat, nsr, syn, ssr, grid in v["segments"]]
    return v, F.measure_and_build(fsegs, PROSODY, pos_of, VOICE)


if __name__ == "__main__":
    import pandas as pd
    out = HERE / "c1_oracle"; out.mkdir(exist_ok=True)
    v, ref = run(sys.argv[1] if len(sys.argv) > 1 else "/tmp/c1_voice")
    pd.DataFrame(ref["bdd_ssml"]).to_csv(out / "BDD_ssml.csv", index=False)
    pd.DataFrame(ref["bdd_syntagme_ssml"]).to_csv(out / "BDD_syntagme_ssml.csv", index=False)
    pd.DataFrame(ref["bdd_syntagme_synth"]).to_csv(out / "BDD_syntagme_for_synth.csv", index=False)
    vals = dict(seg_stats=ref["seg_stats"], baselines=ref["baselines"], units=ref["units"],
                raw_rows=[{k: r[k] for k in ("segment", "syntagme", "pause", "raw_pitch", "raw_volume", "raw_rate")} for r in ref["raw_rows"]],
                sm_pitch=list(map(float, ref["sm_p"])), sm_rate=list(map(float, ref["sm_r"])))
    (out / "values.json").write_text(json.dumps(vals, indent=1))
    print(f"{len(ref['raw_rows'])} syntagme rows, {len(ref['seg_stats'])} segments -> {out}")
