"""GPU probe for K5 (silence segmentation): HBM throughput of pb_silence_runs_kernel on N x 1 h of 22.05 kHz audio
(BASELINE config C4's segmentation stage).  Prints one JSON line; run under gpurun."""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import prosody_b200 as pb  # noqa: E402


def main(n_files=16, sr=22050, hours=1.0, seed=3456):
    g = torch.Generator(device="cuda").manual_seed(seed)
    n = int(sr * 3600 * hours)
    pcm = torch.empty(n_files * n, dtype=torch.int16, device="cuda")
    for f in range(n_files):
        x = torch.randn(n, generator=g, device="cuda") * 3000.0
        # a pause every ~12 s, 0.4-2.5 s long, at the noise floor
        t = torch.arange(n, device="cuda", dtype=torch.float32) / sr
        period = 9.0 + 6.0 * torch.rand(1, generator=g, device="cuda").item()
        gap = 0.4 + 2.1 * ((t / period).floor() * 0.61803 % 1.0)
        x = torch.where((t % period) < gap, x * 0.02, x)
        pcm[f * n:(f + 1) * n] = x.clamp(-32768, 32767).to(torch.int16)
        del x, t, gap
    units = pb.Units.from_list([(f * n, n, sr, 0.0, None) for f in range(n_files)])
    ex = pb.Extractor(0)
    best = None
    for it in range(6):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = ex.split_on_silence(pcm, units, 1000, -50, 300)
        wall = time.perf_counter() - t0
        ms = ex.timings()["intensity_ms"]
        if it >= 2:
            best = ms if best is None else min(best, ms)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    gb = pcm.numel() * 2 / 1e9
    print(json.dumps(dict(kernel="pb_silence_runs_kernel", files=n_files, hours_each=hours, rate=sr, segments=int(r["seg_off"][-1]),
                          kernel_ms=best, call_wall_ms=wall * 1e3, algorithmic_gb=gb, achieved_gbs=gb / (best * 1e-3),
                          hbm_peak_gbs=peaks.get("hbm_gbs"), frac=(gb / (best * 1e-3) / peaks["hbm_gbs"]) if peaks.get("hbm_gbs") else None,
                          audio_hours_per_s=n_files * hours / (wall))))


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:2]))
