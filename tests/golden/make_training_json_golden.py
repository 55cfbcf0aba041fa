"""Generates the training-JSON golden files.  Run ONCE in the build container (it imports the reference from
/root/reference, which does not exist on the GPU box); the outputs are committed next to this script.

  bdd_syntagme_ssml_golden.csv   written by OUR ssml.build / ssml.write_csvs from the fixed rows below
  training_data_golden.json      what the REFERENCE's Code/Pipeline/create_training_data.py makes of that CSV
  bdd_golden.json                what its combine_training_jsons makes of a results folder holding that JSON
"""
import shutil
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, "/root/reference/Code")

import prosody_b200  # noqa: E402,F401
from prosody_b200 import ssml  # noqa: E402
import Pipeline.create_training_data as ref  # noqa: E402  (stdlib only: csv, json, re, xml.etree)

ROWS = [  # segment, syntagme text, pause ms, pitch %, rate %, volume %
    ("segment_ph1", "Bonjour à tous,", 320, 1.234, -3.5, 2.0),
    ("segment_ph1", "c'est l'été & l'hiver <ensemble>.", 612, -0.004, 0.0, -7.0),
    ("segment_ph1", 'il a dit "non"', 20, 12.25, 14.999, 0.005),
    ("segment_ph2", "Où est-ce ?", 800, -5.0, -22.5, 7.0),
    ("segment_ph2", "", 49, 0.0, 0.0, 0.0),
    ("segment_ph3", "fin", 50, 3.14159, 2.71828, -1.41421),
]


def main():
    segs, words, pauses, p, r, v = zip(*ROWS)
    final, syn_rows, synth_rows = ssml.build(segs, words, pauses, p, r, v, "fr-FR-HenriNeural", factor=0.8)
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        ssml.write_csvs(final, syn_rows, synth_rows, td / "a.csv", td / "b.csv", td / "c.csv")
        shutil.copy(td / "b.csv", HERE / "bdd_syntagme_ssml_golden.csv")
        shutil.copy(td / "a.csv", HERE / "bdd_ssml_golden.csv")
        out = td / "results" / "VoiceA" / "training_data_VoiceA.json"
        ref.create_training_data(str(td / "b.csv"), str(out))
        ref.combine_training_jsons(str(td / "results"), str(td / "results" / "bdd.json"))
        shutil.copy(out, HERE / "training_data_golden.json")
        shutil.copy(td / "results" / "bdd.json", HERE / "bdd_golden.json")
    print("written:", sorted(x.name for x in HERE.iterdir()))


if __name__ == "__main__":
    main()
