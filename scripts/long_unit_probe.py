"""Development aid: one unsegmented recording of PB_PROBE_HOURS hours (22.05 kHz) through F0 + loudness as ONE unit, so that a launch
list (`ncu --metrics gpu__time_duration.sum`) shows the long-unit kernels of K0 / K3 / K4 by themselves."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench_workloads as W
import prosody_b200 as pb

n_h = int(os.environ.get("PB_PROBE_HOURS", 1))
dev = torch.device("cuda", 0)
pcm, sr, per = W.c4_recordings(dev, n_h)
ex = pb.Extractor(0)
whole = pb.Units.from_list([(h * per, per, sr, 0.0, None, float(sr)) for h in range(n_h)])
p = pb.pitch_params(75.0, 600.0)
for it in range(3):
    r = ex.extract(pcm, whole, p)
    t = ex.timings()
print({k: round(float(t[k]), 3) for k in ("unit_stats_ms", "acf_ms", "cand_ms", "path_ms", "lufs_ms", "total_ms", "n_launches")}, "frames", int(r["n_frames"].sum()),
      "median", r["median_f0"][:2], "lufs", r["lufs"][:2])
