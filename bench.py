#!/usr/bin/env python
"""bench.py — audio-seconds processed per second (xRT) for F0 + loudness + word/syntagme aggregation.

Workload (BASELINE.json configs[1]): 10 000 synthetic 5 s utterances per GPU, 16 kHz mono s16, pitch floor 75 Hz /
ceiling 600 Hz (10 ms hop), each with a synthetic word grid and a paired "raw synth" utterance (4.65 s).  One STEP is
one pass of the reference's "Measure & Build SSML" measurements over the whole batch: per utterance a whole-file F0
track + median and two whole-file loudness values, then per syntagme a fresh F0 analysis of the natural slice, the
loudness of the synthetic slice and both slice durations, followed by baselines, %-deltas and EMA smoothing on the host.

    value  : natural-audio seconds per second, PCM already resident in HBM (kernels + descriptor traffic + host math)
    e2e    : same through the host-buffer API (pinned host PCM -> H2D inside the timed region, records D2H)
    roofline / cpu_baseline : see DESIGN.md "Measurement"

`--impl reference` times the CPU restatement of the reference's own libraries (oracle/, all host threads) on a bounded
sample of the same workload: the reference's real dependencies (parselmouth, pyloudnorm, pydub) are not installable here.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SR, DUR, SYN_DUR = 16000, 5.0, 4.65
FLOOR, CEILING = 75.0, 600.0
METRIC = "audio-sec processed/sec (xRT) for F0+intensity+word aggregation"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--utts", type=int, default=10000, help="utterances per GPU (default: the BASELINE config)")
    ap.add_argument("--cpu-sample", type=int, default=1024, help="utterances in the bounded CPU-baseline sample")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------- workload
def build_segments(n_utt, seed, nat_n, syn_n):
    """Word grids + segment descriptors; natural utterance i at i*nat_n, its synth twin at n_utt*nat_n + i*syn_n."""
    from prosody_b200 import step as S
    from prosody_b200 import synth
    grids = synth.make_word_grid(n_utt, DUR, seed=seed)
    base = n_utt * nat_n
    return [S.Segment(f"segment_ph{i + 1}", i * nat_n, nat_n, SR, grids[i], base + i * syn_n, syn_n, SR) for i in range(n_utt)]


def make_pcm(n_utt, seed, device):
    import torch
    from prosody_b200 import synth
    nat_n, syn_n = int(round(DUR * SR)), int(round(SYN_DUR * SR))
    pcm = torch.empty(n_utt * (nat_n + syn_n), dtype=torch.int16, device=device)
    synth.make_corpus(n_utt, DUR, SR, seed=seed, device=device, out=pcm[:n_utt * nat_n].view(n_utt, nat_n))
    synth.make_corpus(n_utt, SYN_DUR, SR, seed=seed + 7919, device=device, out=pcm[n_utt * nat_n:].view(n_utt, syn_n))
    return pcm, nat_n, syn_n


# ----------------------------------------------------------------------------------------------------------- CPU arm
def cpu_units(pl, n_seg_sample):
    """The units of the first n_seg_sample segments of a plan, in oracle terms."""
    import numpy as np
    S_ = pl.n_seg
    u = pl.units
    keep = [i for i in range(n_seg_sample)] + [S_ + i for i in range(n_seg_sample)]
    syn_rows = np.nonzero(pl.syn_seg < n_seg_sample)[0]
    for k in syn_rows:
        keep += [2 * S_ + 2 * int(k), 2 * S_ + 2 * int(k) + 1]
    return np.asarray(keep, np.int64)


def run_cpu_baseline(pcm_host, pl, n_seg_sample, threads=0):
    """Times the oracle (CPU restatement of parselmouth / pyloudnorm / pydub) on a bounded sample. -> dict"""
    import numpy as np
    from oracle import oracle as O
    O.build()
    keep = cpu_units(pl, n_seg_sample)
    u = pl.units
    wp, wl = pl.want_pitch[keep] != 0, pl.want_lufs[keep] != 0
    pk, lk = keep[wp], keep[wl]
    a = np.zeros(len(lk), np.int64); b = np.zeros(len(lk), np.int64); npad = np.zeros(len(lk), np.int64)
    for j, i in enumerate(lk):
        a[j], b[j], npad[j], _ = O.lufs_resolve(int(u.file_nx[i]), int(u.rate[i]), float(u.meter_rate[i]), float(u.t0[i]),
                                                float(u.t1[i]) if u.has_t1[i] else None)
    params = O.pitch_params(FLOOR, CEILING)
    nthreads = threads or os.cpu_count() or O.max_threads()      # explicit: the box may export OMP_NUM_THREADS=1
    O.lib().po_counters_reset()
    t0 = time.perf_counter()
    med, nv, nf, st = O.batch_median_pitch(pcm_host, u.file_off[pk], u.file_nx[pk], u.rate[pk], u.has_t1[pk], u.t0[pk], u.t1[pk],
                                           params, nthreads)
    t1 = time.perf_counter()
    lufs, lst = O.batch_lufs(pcm_host, u.file_off[lk], a, b, npad, u.meter_rate[lk], nthreads)
    t2 = time.perf_counter()
    audio_s = n_seg_sample * DUR
    # work model for the roofline: single-threaded counters on a few utterances (thread-private in the OpenMP run)
    O.lib().po_counters_reset()
    sub = pk[:min(len(pk), 12)]
    O.batch_median_pitch(pcm_host, u.file_off[sub], u.file_nx[sub], u.rate[sub], u.has_t1[sub], u.t0[sub], u.t1[sub], params, 1)
    cnt = O.counters()
    return dict(value=audio_s / (t2 - t0), seconds=t2 - t0, pitch_s=t1 - t0, lufs_s=t2 - t1, cores=nthreads, audio_s=audio_s,
                n_pitch_units=int(len(pk)), n_lufs_units=int(len(lk)), frames=int(nf.sum()), counters=cnt,
                sample=f"first {n_seg_sample} utterances of the workload ({audio_s:.0f} s natural audio, {len(pk)} pitch units, "
                       f"{len(lk)} loudness units), OpenMP over units")


def algorithmic_flops_per_frame(counters, geom):
    """F = 4 nw + 2 (2.5 N log2 N) + 1.5 N + 2 B + 3 L + 8 (sinc terms) + 5 K^2 — the REFERENCE
    algorithm's work per frame (Praat's FFT size, Brent's sinc evaluations as counted by the oracle on this input).
    -> (autocorrelation part: window, two FFTs, power spectrum, normalisation = kernel K1;
        candidate part: peak scan, sinc interpolation / Brent refinement = kernel K2;  the path finder's 5 K^2 = K3)"""
    nw, N, B, L = geom["nw"], geom["nfft"], geom["brent_ixmax"], geom["max_lag"]
    fr = max(counters["frames"], 1)
    terms = counters["sinc_terms"] / fr
    K = counters["candidates"] / fr + 1.0
    return 4 * nw + 2 * 2.5 * N * math.log2(N) + 1.5 * N + 2 * B, 3 * L + 8.0 * terms, 5.0 * K * K


# ----------------------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm))


# ----------------------------------------------------------------------------------------------------------- main
def reference_arm(args, rank):
    """The reference's CPU implementation of the path (its libraries restated in oracle/), all host threads."""
    if rank != 0:
        return
    import numpy as np
    import torch
    from prosody_b200 import step as S
    n_s = max(4, args.cpu_sample)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    pcm, nat_n, syn_n = make_pcm(n_s, 1234, dev)
    segs = build_segments(n_s, 1234, nat_n, syn_n)
    pl = S.plan(segs)
    host = pcm.cpu().numpy()
    vals = []
    for it in range(args.warmup + args.steps):
        r = run_cpu_baseline(host, pl, n_s)
        if it >= args.warmup:
            vals.append(r)
    tot_audio = sum(r["audio_s"] for r in vals); tot_s = sum(r["seconds"] for r in vals)
    v = tot_audio / tot_s
    line = dict(metric=METRIC, value=v, unit="audio-s/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * tot_s / len(vals), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic", impl="reference",
                config=dict(workload=f"{n_s} synthetic {DUR:g} s utterances per step (bounded sample of the 10k-utterance config), "
                                     f"16 kHz mono, F0 {FLOOR:g}-{CEILING:g} Hz, 10 ms hop, word grids + paired raw-synth stream",
                            parallelism="cpu-openmp"),
                cpu_baseline=dict(value=v, unit="audio-s/s", cores=vals[-1]["cores"], kind="port", sample=vals[-1]["sample"]),
                e2e=dict(value=v, unit="audio-s/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    import prosody_b200 as pb
    from prosody_b200 import shard
    from prosody_b200 import step as S
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    # ---- synthetic shard of this rank (weak scaling: every GPU gets the full BASELINE config)
    n_utt = args.utts
    pcm, nat_n, syn_n = make_pcm(n_utt, 1234 + 1000 * rank, dev)
    segs = build_segments(n_utt, 1234 + 1000 * rank, nat_n, syn_n)
    prosody = dict(S.DEFAULT_PROSODY)
    pitch = dict(pitch_floor=FLOOR, pitch_ceiling=CEILING)
    pl = S.plan(segs, prosody)
    host_pcm = torch.empty(pcm.shape, dtype=torch.int16, pin_memory=True)
    host_pcm.copy_(pcm)
    torch.cuda.synchronize()
    ex = pb.Extractor(local)
    audio_s = n_utt * DUR
    # the shards are fixed for the whole run: their row counts are exchanged once, not in every step
    row_sizes = shard.row_counts(pl.n_syn, dev) if world > 1 else None

    def step(src):
        out = S.measure(ex, src, pl, prosody, pitch)
        if world > 1:
            # final gather of the per-syntagme results on rank 0 (the path's only exchange)
            rows = torch.from_numpy(np.stack([out["raw_pitch"], out["raw_volume"], out["raw_rate"], out["sm_pitch"], out["sm_rate"]], 1)).to(dev)
            ids = torch.arange(rows.shape[0], device=dev, dtype=torch.int64) + rank * (1 << 32)
            gathered = shard.gather_rows(rows, ids, dst=0, sizes=row_sizes)
            assert rank != 0 or gathered.shape[1] == 5
        return out

    def timed(src, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        acc = dict(frames_ms=0.0, acf_ms=0.0, cand_ms=0.0, lufs_ms=0.0, path_ms=0.0, unit_stats_ms=0.0, h2d_ms=0.0, total_ms=0.0, host_plan_ms=0.0, n_launches=0, n_frames=0)
        for _ in range(steps):
            out = step(src)
            for k in acc:
                acc[k] += out["timings"][k]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt, acc, out

    for _ in range(max(args.warmup, 3)):
        step(pcm)
    sampler = ClockSampler(local); sampler.start()
    dt_dev, acc, out = timed(pcm, args.steps)
    for _ in range(2):
        step(host_pcm)
    dt_e2e, acc_e2e, _ = timed(host_pcm, args.steps)
    sampler.stop_flag.set(); sampler.join(timeout=2)

    if rank == 0:
        value = world * audio_s * args.steps / dt_dev
        e2e = world * audio_s * args.steps / dt_e2e
        n_units = len(pl.units)
        # ---- CPU baseline on a bounded sample + the reference algorithm's work model
        cpu = run_cpu_baseline(host_pcm.numpy(), pl, min(args.cpu_sample, n_utt))
        from oracle import oracle as O
        _, g, *_ = O.pitch_geometry(nat_n, float(SR), params=O.pitch_params(FLOOR, CEILING))
        geom = dict(nw=g.nsamp_window, nfft=g.nsampFFT, brent_ixmax=g.brent_ixmax, max_lag=g.maximumLag)
        f_acf, f_cand, f_path = algorithmic_flops_per_frame(cpu["counters"], geom)
        fpf = f_acf + f_cand + f_path
        frames_per_launch = acc["n_frames"] / args.steps
        kernel_ms = acc["acf_ms"] / args.steps                      # K1, the dominant kernel
        cand_ms = acc["cand_ms"] / args.steps
        frames_ms = acc["frames_ms"] / args.steps                   # K1 + K2 (+ the pair-position kernel), the round-1 kernel's job
        achieved_tflops = f_acf * frames_per_launch / (kernel_ms * 1e-3) / 1e12
        # context only: the same launches with nothing beside them (inside the step the loudness kernels share the SMs)
        alone = []
        for _ in range(2):
            ex.extract(pcm, pl.units, pb.pitch_params(FLOOR, CEILING), want_pitch=pl.want_pitch, want_lufs=np.zeros(n_units, np.uint8),
                       lufs=False, durations=False)
            t_ = ex.timings()
            alone.append((t_["acf_ms"], t_["cand_ms"], t_["frames_ms"]))
        kernel_ms_alone = min(a_[0] for a_ in alone)
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        info = ex.device_info()
        # FP32 (non-tensor) peak: MEASURED_PEAKS.json carries none, so it was measured on this pool with a dependent-free
        # FFMA loop on every SM (scripts/fp32_peak_probe.cu -> profiles/r02_fp32_peak.json); nominal only if that is missing
        try:
            fp32_peak = float(json.loads((ROOT / "profiles" / "r02_fp32_peak.json").read_text())["best_burst_tflops"])
            fp32_src = "measured: profiles/r02_fp32_peak.json (FFMA loop, all SMs)"
        except Exception:
            fp32_peak = info["sm_count"] * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
            fp32_src = "nominal: SMs x 128 lanes x 2 x max SM clock"
        alg_bytes = 2.0 * SR * 0.01 + 8.0      # per frame: s16 in once (10 ms hop) + f32 F0 + f32 strength
        hbm_gbs = alg_bytes * frames_per_launch / (frames_ms * 1e-3) / 1e9
        traffic = None       # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
        try:
            tr = json.loads((ROOT / "profiles" / "r02_acf_traffic.json").read_text())
            traffic = tr["bytes_per_frame"] * frames_per_launch
        except Exception:
            pass
        line = dict(
            metric=METRIC, value=value, unit="audio-s/s", n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
            ms_per_step=1e3 * dt_dev / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
            data="synthetic",
            config=dict(workload=f"{n_utt} synthetic {DUR:g} s utterances per GPU, 16 kHz mono s16, F0 {FLOOR:g}-{CEILING:g} Hz, 10 ms hop, "
                                 f"word grids + paired {SYN_DUR:g} s raw-synth stream: {n_units} measurement units per GPU "
                                 f"({pl.n_seg} utterances, {pl.n_syn} syntagme rows)",
                        units_per_gpu=n_units, pitch_frames_per_step=int(frames_per_launch), parallelism=f"units sharded over {world} GPU(s)",
                        l2="inputs (3.1 GB PCM per GPU) exceed the 126 MB L2; no flush needed",
                        timing="wall clock between barrier+synchronize, max over ranks; kernel times from CUDA events on the launch stream",
                        audio_hours_per_s=value / 3600.0),
            e2e=dict(value=e2e, unit="audio-s/s", h2d_bytes_per_step=int(host_pcm.numel() * 2 + n_units * 120),
                     d2h_bytes_per_step=int(n_units * 20), ms_per_step=1e3 * dt_e2e / args.steps),
            gpu_launches=int(acc["n_launches"]),
            kernels_ms_per_step={k: acc[k] / args.steps for k in ("unit_stats_ms", "frames_ms", "acf_ms", "cand_ms", "path_ms", "lufs_ms", "h2d_ms", "total_ms", "host_plan_ms")},
            roofline=dict(bound="fp32", kernel="pb_pitch_acf_kernel<10>", achieved=achieved_tflops, peak=fp32_peak, unit="TFLOP/s",
                          frac=achieved_tflops / fp32_peak, traffic=traffic, peak_source=fp32_src,
                          note="non-tensor FP32 pipe: no stage is a dense contraction. Round 2 split the round-1 frames kernel in two: K1 "
                               "pb_pitch_acf_kernel (window, two FFTs, power spectrum, normalisation; dominant) and K2 pb_pitch_cand_kernel (peak scan, "
                               "sinc refinement). achieved = the REFERENCE algorithm's flops for K1's part of a frame "
                               f"({f_acf:.0f} of {fpf:.0f}, oracle-counted on this input) x frames per launch / CUDA-event time of K1 inside the step, where "
                               "the loudness kernels run beside it on another stream. `path` is K1+K2 together against all of the reference's "
                               "per-frame flops: the figure comparable with round 1's frames kernel (0.19).",
                          flops_per_frame=f_acf, frames_per_launch=int(frames_per_launch), kernel_ms=kernel_ms,
                          kernel_ms_alone=kernel_ms_alone, frac_alone=f_acf * frames_per_launch / (kernel_ms_alone * 1e-3) / 1e12 / fp32_peak,
                          cand=dict(kernel="pb_pitch_cand_kernel", kernel_ms=cand_ms, flops_per_frame=f_cand,
                                    achieved=f_cand * frames_per_launch / (cand_ms * 1e-3) / 1e12,
                                    frac=f_cand * frames_per_launch / (cand_ms * 1e-3) / 1e12 / fp32_peak),
                          path=dict(kernels="K1+K2", kernel_ms=frames_ms, flops_per_frame=f_acf + f_cand,
                                    achieved=(f_acf + f_cand) * frames_per_launch / (frames_ms * 1e-3) / 1e12,
                                    frac=(f_acf + f_cand) * frames_per_launch / (frames_ms * 1e-3) / 1e12 / fp32_peak,
                                    kernel_ms_alone=min(a_[2] for a_ in alone)),
                          hbm=dict(achieved=hbm_gbs, peak=peaks.get("hbm_gbs"), unit="GB/s",
                                   frac=(hbm_gbs / peaks["hbm_gbs"]) if peaks.get("hbm_gbs") else None, bytes_per_frame=alg_bytes,
                                   peak_source="MEASURED_PEAKS.json" if peaks.get("hbm_gbs") else "absent")),
            cpu_baseline=dict(value=cpu["value"], unit="audio-s/s", cores=cpu["cores"], kind="port", sample=cpu["sample"],
                              pitch_s=cpu["pitch_s"], lufs_s=cpu["lufs_s"]),
            clocks=sampler.summary())
        print(json.dumps(line))
    ex.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
