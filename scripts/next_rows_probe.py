"""GPU probe for the "next" rows: throughput of intensity (K4b), legacy loudness / pitch, interval reduction (K6)
and the batch TextGrid reader on bench-sized inputs.  Prints one JSON line; run under gpurun."""
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import prosody_b200 as pb  # noqa: E402
from prosody_b200 import legacy, synth, textgrid as TG  # noqa: E402


def best(fn, n=4):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts), r


def main(n_utt=4000):
    sr, dur = 16000, 5.0
    pcm = synth.make_corpus(n_utt, dur, sr, seed=1234, device="cuda")
    n = pcm.shape[1]
    flat = pcm.reshape(-1)
    ex = pb.Extractor(0)
    whole = pb.Units.from_list([(i * n, n, sr, 0.0, None, float(sr)) for i in range(n_utt)])
    audio_s = n_utt * dur
    out = {}
    t, r = best(lambda: ex.intensity(flat, whole))
    out["intensity"] = dict(call_ms=t * 1e3, kernel_ms=ex.timings()["intensity_ms"], frames=int(r["frame_off"][-1]), audio_s_per_s=audio_s / t)
    rng = np.random.default_rng(1)
    rows = []
    for i in range(n_utt):
        cuts = np.sort(rng.uniform(0.0, dur, 7))
        rows += [(i * n, n, sr, float(a), float(b)) for a, b in zip(cuts[:-1], cuts[1:])]
    syn = pb.Units.from_list(rows)
    t, _ = best(lambda: legacy.loudness_segments(ex, flat, syn))
    out["legacy_loudness"] = dict(call_ms=t * 1e3, rows=len(rows), rows_per_s=len(rows) / t, gb_per_s=flat.numel() * 2 / 1e9 / t)
    t, _ = best(lambda: legacy.pitch_segments(ex, flat, syn), n=2)
    out["legacy_pitch"] = dict(call_ms=t * 1e3, rows=len(rows), rows_per_s=len(rows) / t)
    p = pb.pitch_params(75.0, 600.0)
    r = ex.median_pitch(flat, whole, p, frames=True)
    t1, dt = pb.pitch_frame_times(whole, p)
    grids = synth.make_word_grid(n_utt, dur, seed=1234)
    ivs = [(i, a, b) for i, g in enumerate(grids) for (a, b, mark) in g if mark.strip()]
    iv = tuple(np.asarray(c) for c in zip(*ivs))
    f0d, ind = torch.from_numpy(r["frame_f0"]).cuda(), torch.from_numpy(r["frame_intensity"]).cuda()
    t, _ = best(lambda: ex.reduce_intervals(r["frame_off"], t1, dt, f0d, iv, track2=ind))
    out["interval_reduction"] = dict(call_ms=t * 1e3, kernel_ms=ex.timings()["intensity_ms"], intervals=len(ivs), frames=int(r["frame_off"][-1]),
                                     intervals_per_s=len(ivs) / t)
    with tempfile.TemporaryDirectory() as td:
        paths = []
        for i, g in enumerate(grids[:2000]):
            pth = Path(td) / f"segment_ph{i + 1}.TextGrid"
            TG.write(pth, {"words": g})
            paths.append(pth)
        t0 = time.perf_counter(); a = [TG.read(q).tiers[0].intervals for q in paths]; t_py = time.perf_counter() - t0
        t0 = time.perf_counter(); st, b = TG.read_tier_batch(paths); t_nat = time.perf_counter() - t0
        out["textgrid_batch"] = dict(files=len(paths), python_ms=t_py * 1e3, native_ms=t_nat * 1e3, equal=bool(a == b))
    print(json.dumps(out))


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:2]))
