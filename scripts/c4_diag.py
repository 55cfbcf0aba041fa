"""Development aid: where the 1-hour config-4 recording differs from the oracle (run under gpurun)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import prosody_b200 as pb
from prosody_b200 import synth
from oracle import oracle

sr = 22050
pcm = synth.make_corpus(720, 5.0, sr, seed=3456, device="cuda")
x = pcm.reshape(-1).clone()
n = x.numel()
t = np.arange(0, 3600, 19.0)
for k, a in enumerate(t[1:]):
    i0 = int(a * sr); i1 = i0 + int((1.1 + 0.1 * (k % 13)) * sr)
    x[i0:i1] = (x[i0:i1].float() * 0.004).to(x.dtype)
host = x.cpu().numpy()
ex = pb.Extractor(0)
whole = pb.Units.from_list([(0, n, sr, 0.0, None, float(sr))])
r = ex.median_pitch(x, whole, pb.pitch_params(75.0, 600.0), frames=True)
o = oracle.pitch_track(host, sr, params=oracle.pitch_params(75.0, 600.0), want_candidates=True)
fg, fo = r["frame_f0"].astype(float), o["frequency"]
both = (fg > 0) & (fo > 0)
rel = np.zeros_like(fo); rel[both] = np.abs(fg[both] - fo[both]) / fo[both]
print("frames", len(fo), "voiced both", both.sum(), "disagree", int(((fg > 0) != (fo > 0)).sum()))
for thr in (1e-3, 2e-3, 5e-3, 1e-2, 5e-2):
    print("rel >", thr, int((rel > thr).sum()))
bad = np.nonzero(rel > 2e-3)[0]
tt = o["t1"] + bad * o["dt"]
in_gap = [any(a <= ti <= a + 2.4 for a in t[1:]) for ti in tt]
print("bad frames in attenuated gaps:", sum(in_gap), "of", len(bad))
for b in bad[:12]:
    print(b, round(float(o["t1"] + b * o["dt"]), 3), fg[b], fo[b], rel[b], "strength", r["frame_strength"][b], o["strength"][b])
s = ex.split_on_silence(x, whole, 1000, -50, 300)
ref = oracle.split_on_silence(host, sr, 1000, -50, 300)
got = list(zip(s["start_ms"].tolist(), s["end_ms"].tolist()))
print("segments gpu", len(got), "oracle", len(ref), "pauses", len(t) - 1, "equal", got == ref)
if got != ref:
    for k, (g_, r_) in enumerate(zip(got, ref)):
        if g_ != r_:
            print("first diff at", k, g_, r_); break
print(got[:3], ref[:3])
