"""Development aid: the split 2048-point K1 (PB_ACF_SPLIT=1) against the general kernel and the oracle on a real clip, frame by frame.
One process per setting (the switch is read once)."""
import sys, os, wave
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import prosody_b200 as pb
from oracle import oracle as O
ex = pb.Extractor(0)
with wave.open(str(ROOT / "tests/golden/clips/segment_ph2.wav"), "rb") as w:
    sr = w.getframerate(); pcm = np.frombuffer(w.readframes(w.getnframes()), np.int16).copy()
n = len(pcm)
items = [(0, n, sr, 0.0, None)] + [(0, n, sr, 0.37 * k, 0.37 * k + 0.9 + 0.05 * k) for k in range(1, 40) if 0.37 * k + 0.9 + 0.05 * k < n / sr]
units = pb.Units.from_list(items)
r = ex.median_pitch(pcm, units, pb.pitch_params(150.0, 600.0), frames=True)
np.savez(sys.argv[1], f0=r["frame_f0"], st=r["frame_strength"], med=r["median_f0"], off=r["frame_off"])
worst = 0.0
for i, it in enumerate(items):
    o = O.pitch_track(pcm, sr, it[3], it[4], params=O.pitch_params(150.0, 600.0))
    a, b = r["frame_off"][i], r["frame_off"][i + 1]
    f = r["frame_f0"][a:b]; both = (f > 0) & (o["frequency"] > 0)
    e = np.abs(f[both] - o["frequency"][both]) / o["frequency"][both]
    mism = int(np.sum((f > 0) != (o["frequency"] > 0)))
    se = np.max(np.abs(r["frame_strength"][a:b] - o["strength"]))
    if e.max() > 1e-3 or mism or se > 1e-4:
        k = int(np.argmax(np.abs(r["frame_strength"][a:b] - o["strength"])))
        print("unit", i, it[3:], "frames", b - a, "max rel", e.max(), "voicing mismatches", mism, "strength err", se, "at frame", k, "of", b - a)
    worst = max(worst, e.max())
print(sys.argv[1], "worst rel err", worst, "median", r["median_f0"][:4])
