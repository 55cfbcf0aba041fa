"""Training-JSON export: the consumer of the step's syntagme-level CSV.

Same entry points and output bytes as /root/reference/Code/Pipeline/create_training_data.py
(`create_training_data` :26-118, `combine_training_jsons` :120-156), which `AudioPipeline.export_training_json`
(Code/audioPipeline.py:840-854) runs on `BDD_syntagme_ssml.csv`.  The JSON pins the byte format of the SSML this
package emits: downstream models read it back with ``float(pitch.strip('%'))`` (Code/baseline_models/bilstm.py:44-50).
tests/golden/ holds a CSV written by `ssml.write_csvs` and the JSON the reference's own script produced from it.
"""
from __future__ import annotations

import csv
import json
import os
import re
import xml.etree.ElementTree as ET
from pathlib import Path

_NS = "{http://www.w3.org/2001/10/synthesis}"
_SPEAK = re.compile(r"<speak.*?</speak>", re.DOTALL)
_XMLNS = re.compile(r'\sxmlns(:\w+)?="[^"]+"')
_PREFIX = re.compile(r"\w+:(prosody|break)")


def _plain(element) -> str:
    """Serialised element without namespace declarations or prefixes (ns0:prosody -> prosody)."""
    return _PREFIX.sub(r"\1", _XMLNS.sub("", ET.tostring(element, encoding="unicode", method="xml")))


def _entries_of_block(block: str, segment: str):
    """One <speak> block -> (parsed_sequence entries, stripped SSML strings) in document order."""
    voice = ET.fromstring(block).find(f".//{_NS}voice")
    prosody = None if voice is None else voice.find(f".//{_NS}prosody")
    if prosody is None:
        return [], []
    attrs = {k: prosody.get(k, "") for k in ("pitch", "rate", "volume")}

    def text_entry(t):
        return {"segment": segment, "type": "text", "text": t, "prosody": dict(attrs)}

    seq, stripped = [], []
    lead = (prosody.text or "").strip()
    if lead:
        seq.append(text_entry(lead))
        stripped.append(_plain(prosody))                 # the whole prosody element, once
    for child in prosody:
        if child.tag.rsplit("}", 1)[-1] == "break":
            seq.append({"segment": segment, "type": "break", "time": child.get("time", "")})
            stripped.append(_plain(child))
        tail = (child.tail or "").strip()
        if tail:
            seq.append(text_entry(tail))
    return seq, stripped


def create_training_data(bdd_ssml_path, output_path) -> None:
    if not os.path.exists(bdd_ssml_path):
        raise FileNotFoundError(f"CSV not found: {bdd_ssml_path}")
    os.makedirs(os.path.dirname(output_path), exist_ok=True)
    texts, sequence, raw, stripped = [], [], {}, {}
    with open(bdd_ssml_path, "r", encoding="utf-8") as f:
        for row in csv.DictReader(f):
            seg, syntagme, cell = row["segment"].strip(), row["syntagme"].strip(), row["ssml"].strip()
            if syntagme:
                texts.append(syntagme)
            raw.setdefault(seg, []).append(cell)
            bucket = stripped.setdefault(seg, [])
            for block in _SPEAK.findall(cell):
                entries, plain = _entries_of_block(block, seg)
                sequence.extend(entries)
                bucket.extend(plain)
    if not sequence:
        raise ValueError("No SSML elements found in CSV.")
    doc = {"x": " ".join(texts).strip(), "y": {"parsed_sequence": sequence, "stripped_ssml": stripped, "raw_ssml": raw}}
    with open(output_path, "w", encoding="utf-8") as jf:
        json.dump(doc, jf, ensure_ascii=False, indent=2)


def combine_training_jsons(results_folder, combined_json_path) -> None:
    """bdd.json: per voice folder, the concatenation of its training_data_*.json files (os.listdir order, like the reference)."""
    if not os.path.isdir(results_folder):
        return
    combined = {}
    for name in os.listdir(results_folder):
        folder = Path(results_folder) / name
        if not folder.is_dir():
            continue
        x_parts, seq, stripped, raw = [], [], {}, {}
        for fn in os.listdir(folder):
            if not (fn.startswith("training_data_") and fn.endswith(".json")) or fn == "bdd.json":
                continue
            with open(folder / fn, "r", encoding="utf-8") as jf:
                data = json.load(jf)
            x_parts.append(data.get("x", "") + " ")
            y = data["y"]
            seq.extend(y.get("parsed_sequence", []))
            for seg, items in y.get("stripped_ssml", {}).items():
                stripped.setdefault(seg, []).extend(items)
            for seg, items in y.get("raw_ssml", {}).items():
                raw.setdefault(seg, []).extend(items)
        combined[name] = {"x": "".join(x_parts).strip(), "y": {"parsed_sequence": seq, "stripped_ssml": stripped, "raw_ssml": raw}}
    with open(combined_json_path, "w", encoding="utf-8") as jf:
        json.dump(combined, jf, ensure_ascii=False, indent=2)


def parse_percent(value: str) -> float:
    """How the reference's models read a prosody attribute back (bilstm.py:46-48)."""
    return float(value.strip("%"))
