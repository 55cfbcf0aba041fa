"""Quick GPU probe (development aid): times the hot path on a config-2 shaped batch and prints kernel timings."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import prosody_b200 as pb
from prosody_b200 import synth

n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
sr, dur = 16000, 5.0
t0 = time.time()
pcm = synth.make_corpus(n_utt, dur, sr, seed=1234, device="cuda")
torch.cuda.synchronize(); print("gen", time.time() - t0, "s", pcm.shape)
n = pcm.shape[1]
items = []
for i in range(n_utt):
    items.append((i * n, n, sr, 0.0, None, float(sr)))
    items.append((i * n, n, sr, 0.5, 2.5, float(sr)))
    items.append((i * n, n, sr, 2.5, 4.5, float(sr)))
units = pb.Units.from_list(items)
p = pb.pitch_params(75.0, 600.0)
import os
ex = pb.Extractor(0, lib=pb._native.load(os.environ["PB_LIB"])) if os.environ.get("PB_LIB") else pb.Extractor(0)
print(ex.device_info())
flat = pcm.reshape(-1)
for it in range(4):
    torch.cuda.synchronize(); t0 = time.time()
    r = ex.extract(flat, units, p)
    dt = time.time() - t0
    t = ex.timings()
    audio_s = n_utt * dur
    print(f"iter {it}: wall {dt*1e3:.1f} ms  xRT(file audio) {audio_s/dt:.0f}  timings {t}")
print("median f0 sample", r["median_f0"][:6], "lufs", r["lufs"][:6], "voiced frac", r["n_voiced"].sum() / max(1, r["n_frames"].sum()))
host = flat.cpu().pin_memory()
for it in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    r2 = ex.extract(host, units, p)
    dt = time.time() - t0
    print(f"host-pcm iter {it}: wall {dt*1e3:.1f} ms  {ex.timings()}")
print("host==dev", np.array_equal(r["median_f0"], r2["median_f0"]), np.array_equal(r["lufs"], r2["lufs"], equal_nan=True))
