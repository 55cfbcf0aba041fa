"""Per-run listing of `:+.2f` lattice flips (SURVEY.md §8d "flips listed per run"): runs a config on the GPU and
its units through the CPU oracle, pushes both through the SAME float64 host math (baselines, deltas, EMA) and compares the
strings the SSML emitters would print.  -> profiles/r02_flips_<config>.json
    python scripts/flip_listing.py c2 [--utts N]     (the oracle runs on all host cores: C2 in full is ~20 s on 16 cores)"""
import json, os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import torch
import bench_workloads as W
import prosody_b200 as pb
from prosody_b200 import step as S
from parity_report import flip_report, oracle_units

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
utts = int(sys.argv[sys.argv.index("--utts") + 1]) if "--utts" in sys.argv else 0
dev = torch.device("cuda", 0)
wl = W.c2(dev, n_utt=utts or 10000) if cfg == "c2" else W.c3(dev, n_utt=utts or 2000)
pl = S.plan(wl.segments, wl.prosody)
with pb.Extractor(0) as ex:
    out = S.measure(ex, wl.pcm, pl, wl.prosody, wl.pitch)
host = wl.pcm.cpu().numpy()
t0 = time.perf_counter()
med, lufs, dur = oracle_units(host, pl, wl.pitch["pitch_floor"], wl.pitch["pitch_ceiling"])
t_cpu = time.perf_counter() - t0
prm = dict(S.DEFAULT_PROSODY); prm.update(wl.prosody)
ref = S.finish(pl, med, lufs, dur, prm)
rep = flip_report(out, ref, max_listed=400)
rep["config"] = dict(name=cfg, workload=wl.description, units=len(pl.units), rows=pl.n_syn, oracle_seconds=t_cpu, cores=os.cpu_count())
p = ROOT / "profiles" / f"r02_flips_{cfg}.json"
p.write_text(json.dumps(rep, indent=1))
brief = {k: {q: rep[k][q] for q in ("rows", "identical", "flipped", "flip_rate", "flips_one_lattice_step", "max_abs_delta")} for k in ("pitch", "rate", "volume")}
print(json.dumps(dict(brief, median_f0=rep["median_f0"], lufs=rep["lufs"], durations_identical=rep["durations_identical"])))
