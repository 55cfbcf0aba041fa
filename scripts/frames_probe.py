"""Development aid: K1 / K2 / K3 timings of the pitch path on a config-2 shaped resident batch (one process per setting:
the library reads PB_* environment switches once)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import prosody_b200 as pb
from prosody_b200 import synth

n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
sr, dur = int(os.environ.get("PB_PROBE_SR", 16000)), float(os.environ.get("PB_PROBE_DUR", 5.0))
pcm = synth.make_corpus(n_utt, dur, sr, seed=1234, device="cuda")
n = pcm.shape[1]
units = pb.Units.from_list([(i * n, n, sr, 0.0, None, float(sr)) for i in range(n_utt)])
p = pb.pitch_params(75.0, 600.0)
ex = pb.Extractor(0, lib=pb._native.load(os.environ["PB_LIB"])) if os.environ.get("PB_LIB") else pb.Extractor(0)
flat = pcm.reshape(-1)
best = None
for it in range(5):
    r = ex.extract(flat, units, p, lufs=False, durations=False)
    t = ex.timings()
    if it >= 2 and (best is None or t["frames_ms"] < best["frames_ms"]):
        best = t
fr = best["n_frames"]
print({k: os.environ.get(k) for k in ("PB_LIB", "PB_CAND_CTAS", "PB_ACF_CTAS", "PB_ACF_WSYNC", "PB_RACF_BYTES") if os.environ.get(k)},
      f"frames {fr}  acf {best['acf_ms']:.3f} ms  cand {best['cand_ms']:.3f} ms  frames(K1+K2) {best['frames_ms']:.3f} ms  path {best['path_ms']:.3f} ms  "
      f"-> {best['frames_ms'] * 1e6 / fr:.2f} ns/frame; voiced {r['n_voiced'].sum() / fr:.3f}")
