"""Development aid: re-run the same batch many times and report any unit whose results are not bit-identical
(a race in a kernel shows up here long before it shows up as a parity failure)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import bench
import prosody_b200 as pb
from prosody_b200 import step as S

n_utt = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
pcm, nat_n, syn_n = bench.make_pcm(n_utt, 1234, "cuda")
segs = bench.build_segments(n_utt, 1234, nat_n, syn_n)
pl = S.plan(segs)
host = torch.empty(pcm.shape, dtype=torch.int16, pin_memory=True); host.copy_(pcm); torch.cuda.synchronize()
ex = pb.Extractor(0)
pp = pb.pitch_params(bench.FLOOR, bench.CEILING)
ref = ex.extract(pcm, pl.units, pp, want_pitch=pl.want_pitch, want_lufs=pl.want_lufs)
bad = 0
for it in range(reps):
    src = pcm if it % 2 == 0 else host
    r = ex.extract(src, pl.units, pp, want_pitch=pl.want_pitch, want_lufs=pl.want_lufs)
    for k in ("median_f0", "n_voiced", "lufs", "status", "n_frames"):
        neq = ~((r[k] == ref[k]) | (np.isnan(r[k].astype(float)) & np.isnan(ref[k].astype(float))))
        if neq.any():
            bad += 1
            idx = np.nonzero(neq)[0]
            print(f"iter {it} ({'dev' if it % 2 == 0 else 'host'}): {k} differs on {len(idx)} units, first {idx[:5]}, "
                  f"got {r[k][idx[:3]]} want {ref[k][idx[:3]]} t0 {pl.units.t0[idx[:3]]} t1 {pl.units.t1[idx[:3]]}")
print("mismatching iterations:", bad, "of", reps)
