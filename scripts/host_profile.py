"""Development aid: where the host time of one bench step goes.  Run under gpurun; PB_PLAN_PROFILE=1 adds the planning laps."""
import cProfile
import pstats
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import bench as B  # noqa: E402
import prosody_b200 as pb  # noqa: E402
from prosody_b200 import step as S  # noqa: E402

n_utt = 10000
pcm, nat_n, syn_n = B.make_pcm(n_utt, 1234, "cuda")
segs = B.build_segments(n_utt, 1234, nat_n, syn_n)
prosody = dict(S.DEFAULT_PROSODY)
pitch = dict(pitch_floor=B.FLOOR, pitch_ceiling=B.CEILING)
pl = S.plan(segs, prosody)
ex = pb.Extractor(0)
for _ in range(3):
    out = S.measure(ex, pcm, pl, prosody, pitch)
torch.cuda.synchronize()
t0 = time.perf_counter()
out = S.measure(ex, pcm, pl, prosody, pitch)
dt = time.perf_counter() - t0
print(f"step wall {dt * 1e3:.2f} ms; gpu total {out['timings']['total_ms']:.2f} ms; host plan {out['timings']['host_plan_ms']:.2f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    S.measure(ex, pcm, pl, prosody, pitch)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
