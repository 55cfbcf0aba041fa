"""Development aid: where the host time of a pipelined config-3 step goes (run under gpurun)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bench_workloads as W
import prosody_b200 as pb
from prosody_b200 import step as S, ssml as SSML

dev = torch.device("cuda", 0)
wl = W.c3(dev, n_utt=int(sys.argv[1]) if len(sys.argv) > 1 else 2000, seed=2345)
pcm, segs, prosody, pitch = wl.pcm, wl.segments, wl.prosody, wl.pitch
pl = S.plan(segs, prosody)
pools = SSML.TextPools([segs[i].name for i in pl.syn_seg], pl.syn_words)
exs = [pb.Extractor(0), pb.Extractor(0)]
T = {}
def lap(k, t0):
    T[k] = T.get(k, 0.0) + time.perf_counter() - t0
def finish(ex):
    t0 = time.perf_counter(); r = ex.wait(); lap("wait", t0)
    t0 = time.perf_counter(); S._raise_like_the_reference(r["status"], pl, True); lap("raise_check", t0)
    prm = dict(S.DEFAULT_PROSODY); prm.update(prosody or {})
    t0 = time.perf_counter(); out = S.finish(pl, r["median_f0"], r["lufs"], r["duration_s"], prm, ex._lib); lap("finish_math", t0)
    t0 = time.perf_counter(); out["ssml"] = SSML.build_csv_bytes(pools, pl.syn_pause_ms, out["sm_pitch"], out["sm_rate"], out["raw_volume"], "fr-FR-HenriNeural", prosody["inter_syntagme_pause_factor"], lib=ex._lib); lap("csv", t0)
    t0 = time.perf_counter(); out["timings"] = ex.timings(); lap("timings", t0)
    return out
for rep in range(2):
    T.clear()
    torch.cuda.synchronize(); tA = time.perf_counter()
    pend = []
    steps = 6
    for k in range(steps):
        ex = exs[k % 2]
        if len(pend) == 2: finish(pend.pop(0))
        t0 = time.perf_counter(); S.submit(ex, pcm, pl, pitch); lap("submit", t0)
        pend.append(ex)
    while pend: out = finish(pend.pop(0))
    torch.cuda.synchronize(); dt = time.perf_counter() - tA
print("ms per step", 1e3 * dt / steps, {k: round(1e3 * v / steps, 2) for k, v in T.items()}, "gpu total", out["timings"]["total_ms"])
