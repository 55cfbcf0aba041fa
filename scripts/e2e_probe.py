"""Development aid: where does a bench step spend its wall time (host planning, GPU, python post-processing)?"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bench
import prosody_b200 as pb
from prosody_b200 import step as S

n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
pcm, nat_n, syn_n = bench.make_pcm(n_utt, 1234, "cuda")
segs = bench.build_segments(n_utt, 1234, nat_n, syn_n)
pl = S.plan(segs)
host = torch.empty(pcm.shape, dtype=torch.int16, pin_memory=True); host.copy_(pcm); torch.cuda.synchronize()
ex = pb.Extractor(0)
pp = pb.pitch_params(bench.FLOOR, bench.CEILING)
for name, src in (("device", pcm), ("host-pinned", host)):
    for it in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = ex.extract(src, pl.units, pp, want_pitch=pl.want_pitch, want_lufs=pl.want_lufs)
        t1 = time.perf_counter()
        t = ex.timings()
        print(f"{name} iter {it}: extract wall {1e3*(t1-t0):.1f} ms | plan {t['host_plan_ms']:.1f} total_ev {t['total_ms']:.1f} h2d {t['h2d_ms']:.1f} "
              f"stats {t['unit_stats_ms']:.1f} frames {t['frames_ms']:.1f} path {t['path_ms']:.1f} lufs {t['lufs_ms']:.1f} d2h {t['d2h_ms']:.2f} launches {t['n_launches']}")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = S.measure(ex, src, pl, None, dict(pitch_floor=bench.FLOOR, pitch_ceiling=bench.CEILING))
    print(f"{name}: full measure wall {1e3*(time.perf_counter()-t0):.1f} ms")
