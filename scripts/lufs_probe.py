"""Development aid: the loudness kernels by themselves on a config-2 shaped batch (whole files + 2 s slices, 16 kHz)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import prosody_b200 as pb
from prosody_b200 import synth

n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
sr, dur = int(os.environ.get("PB_PROBE_SR", 16000)), float(os.environ.get("PB_PROBE_DUR", 5.0))
pcm = synth.make_corpus(n_utt, dur, sr, seed=1234, device="cuda")
n = pcm.shape[1]
items = []
for i in range(n_utt):
    items.append((i * n, n, sr, 0.0, None, float(sr)))
    items.append((i * n, n, sr, 0.5, 2.5, float(sr)))
units = pb.Units.from_list(items)
ex = pb.Extractor(0, lib=pb._native.load(os.environ["PB_LIB"])) if os.environ.get("PB_LIB") else pb.Extractor(0)
flat = pcm.reshape(-1)
best = None
for it in range(6):
    out, st = ex.lufs(flat, units)
    t = ex.timings()
    if it >= 2 and (best is None or t["lufs_ms"] < best):
        best = t["lufs_ms"]
samples = n_utt * (n + 2 * sr)
print({k: os.environ.get(k) for k in ("PB_LIB",) if os.environ.get(k)}, f"lufs {best:.3f} ms for {samples / 1e6:.0f} M samples -> {samples / best / 1e6:.0f} G samples/s; lufs[:3] {out[:3]}")
