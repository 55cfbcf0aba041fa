#!/usr/bin/env python
"""bench.py — audio-seconds processed per second (xRT) for F0 + loudness + word/syntagme aggregation.

Default workload (BASELINE.json configs[1], `--config c2`): 10 000 synthetic 5 s utterances per GPU, 16 kHz mono s16, pitch floor
75 Hz / ceiling 600 Hz (10 ms hop), each with a synthetic word grid and a paired "raw synth" utterance (4.65 s).  One STEP is one
pass of the reference's "Measure & Build SSML" measurements over the whole batch: per utterance a whole-file F0 track + median and
two whole-file loudness values, then per syntagme a fresh F0 analysis of the natural slice, the loudness of the synthetic slice
and both slice durations, followed by baselines, %-deltas and EMA smoothing on the host.

    value  : natural-audio seconds per second, PCM already resident in HBM (kernels + descriptor traffic + host math)
    e2e    : same through the host-buffer API (pinned host PCM -> H2D inside the timed region, records D2H)
    roofline / cpu_baseline : see DESIGN.md "Measurement"

Steps are PIPELINED (`--in-flight 2`, the library's pb_extract_submit / pb_extract_wait on two handles): while the GPU works on
step k the host plans and enqueues step k+1 and post-processes step k-1.  `serial` in the JSON line is the same loop with one
step at a time (its kernel times feed the roofline, its ms_per_step is the latency of a step).

Other configs (`--config c1|c3|c4|c5`, bench_workloads.py) are not the driver's headline; their lines are kept under profiles/.
`--impl reference` times the CPU restatement of the reference's own libraries (oracle/, all host threads) on a bounded sample of
the same workload: the reference's real dependencies (parselmouth, pyloudnorm, pydub) are not installable here
(profiles/r02_pip_real_packages.log).
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

METRIC = "audio-sec processed/sec (xRT) for F0+intensity+word aggregation"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--utts", type=int, default=0, help="utterances per GPU (default: the config's own: 10 000 for c2, 2 000 for c3)")
    ap.add_argument("--hours", type=float, default=100.0, help="c5: corpus size in hours (all ranks together); c4: recordings of 1 h")
    ap.add_argument("--in-flight", type=int, default=2, help="steps in flight (1 = one at a time)")
    ap.add_argument("--cpu-sample", type=int, default=1024, help="utterances in the bounded CPU-baseline sample (c2; scaled for longer ones)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------- CPU arm
def cpu_units(pl, n_seg_sample):
    """The units of the first n_seg_sample segments of a plan, in oracle terms."""
    import numpy as np
    S_ = pl.n_seg
    keep = [i for i in range(n_seg_sample)] + [S_ + i for i in range(n_seg_sample)]
    syn_rows = np.nonzero(pl.syn_seg < n_seg_sample)[0]
    for k in syn_rows:
        keep += [2 * S_ + 2 * int(k), 2 * S_ + 2 * int(k) + 1]
    return np.asarray(keep, np.int64)


def run_cpu_baseline(pcm_host, pl, n_seg_sample, pitch, threads=0, count_work=True):
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
    params = O.pitch_params(pitch["pitch_floor"], pitch["pitch_ceiling"])
    nthreads = threads or os.cpu_count() or O.max_threads()      # explicit: the box may export OMP_NUM_THREADS=1
    O.lib().po_counters_reset()
    t0 = time.perf_counter()
    med, nv, nf, st = O.batch_median_pitch(pcm_host, u.file_off[pk], u.file_nx[pk], u.rate[pk], u.has_t1[pk], u.t0[pk], u.t1[pk],
                                           params, nthreads)
    t1 = time.perf_counter()
    lufs, lst = O.batch_lufs(pcm_host, u.file_off[lk], a, b, npad, u.meter_rate[lk], nthreads)
    t2 = time.perf_counter()
    audio_s = float(sum(pl.segments[i].nat_nx / pl.segments[i].nat_sr for i in range(n_seg_sample)))
    cnt = None
    if count_work:
        # work model for the roofline: single-threaded counters on a few utterances (thread-private in the OpenMP run)
        O.lib().po_counters_reset()
        sub = pk[:min(len(pk), 12)]
        O.batch_median_pitch(pcm_host, u.file_off[sub], u.file_nx[sub], u.rate[sub], u.has_t1[sub], u.t0[sub], u.t1[sub], params, 1)
        cnt = O.counters()
    return dict(value=audio_s / (t2 - t0), seconds=t2 - t0, pitch_s=t1 - t0, lufs_s=t2 - t1, cores=nthreads, audio_s=audio_s,
                n_pitch_units=int(len(pk)), n_lufs_units=int(len(lk)), frames=int(nf.sum()), counters=cnt,
                sample=f"first {n_seg_sample} utterances of the workload ({audio_s:.0f} s natural audio, {len(pk)} pitch units, "
                       f"{len(lk)} loudness units); OpenMP over UNITS with the PCM already in memory — no file I/O, no per-call WAV decoding and a "
                       f"finer-grained parallelism than the reference's one process per voice (Code/audioPipeline.py:1141-1150): it flatters the CPU")


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


# ----------------------------------------------------------------------------------------------------------- clocks / placement
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


def bind_to_gpu_numa(local: int, world: int) -> dict:
    """Pin this process (and the pinned-memory pages it is about to touch) to the NUMA node its GPU hangs off, sharing that node's
    cores with the other ranks on it.  Round 1 left every rank on node 0: 8 ranks planning on 4 cores each and all H2D traffic
    crossing one socket's memory controllers.  Best effort: returns what it did."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if not bus:
            return dict(bound=False, why="no bus id")
        dom, rest = bus.split(":", 1)
        dev = f"{dom[-4:]}:{rest}"
        node = int(Path(f"/sys/bus/pci/devices/{dev}/numa_node").read_text().strip())
        if node < 0:
            return dict(bound=False, why="no numa node reported", bus=dev)
        cpus = []
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return dict(bound=False, why="node cpus not in this process' affinity mask", node=node)
        # ranks on the same node share its cores evenly (ranks are placed on GPUs in order: neighbours share a node)
        same = [r for r in range(world) if _numa_of(r) == node] if world > 1 else [local]
        if len(same) > 1 and local in same:
            per = max(1, len(allowed) // len(same)); k = same.index(local)
            mine = allowed[k * per:(k + 1) * per] or allowed
        else:
            mine = allowed
        os.sched_setaffinity(0, mine)
        return dict(bound=True, node=node, cpus=len(mine), bus=dev)
    except Exception as e:      # noqa: BLE001 — placement is an optimisation, never a failure
        return dict(bound=False, why=f"{type(e).__name__}: {e}")


def _numa_of(idx: int) -> int:
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(idx)],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        dom, rest = bus.split(":", 1)
        return int(Path(f"/sys/bus/pci/devices/{dom[-4:]}:{rest}/numa_node").read_text().strip())
    except Exception:           # noqa: BLE001
        return -1


def load_peaks():
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:           # noqa: BLE001
        pass
    # FP32 (non-tensor) peak: MEASURED_PEAKS.json carries none, so it was measured on this pool with a dependent-free FFMA loop on
    # every SM (scripts/fp32_peak_probe.cu -> profiles/r02_fp32_peak.json); nominal only if that file is missing
    try:
        fp32 = float(json.loads((ROOT / "profiles" / "r02_fp32_peak.json").read_text())["best_burst_tflops"])
        src = "measured: profiles/r02_fp32_peak.json (dependent-free FFMA loop on every SM, this pool's B200)"
    except Exception:           # noqa: BLE001
        fp32 = 148 * 128 * 2 * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
        src = "nominal: SMs x 128 lanes x 2 x max SM clock"
    return peaks, fp32, src


# ----------------------------------------------------------------------------------------------------------- reference arm
def reference_arm(args, rank):
    """The reference's CPU implementation of the path (its libraries restated in oracle/), all host threads."""
    if rank != 0:
        return
    import torch
    import bench_workloads as W
    from prosody_b200 import step as S
    n_s = max(4, args.cpu_sample)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    wl = W.c2(dev, n_utt=n_s, seed=1234)
    pl = S.plan(wl.segments, wl.prosody)
    host = wl.pcm.cpu().numpy()
    vals = []
    for it in range(args.warmup + args.steps):
        r = run_cpu_baseline(host, pl, n_s, wl.pitch, count_work=False)
        if it >= args.warmup:
            vals.append(r)
    tot_audio = sum(r["audio_s"] for r in vals); tot_s = sum(r["seconds"] for r in vals)
    v = tot_audio / tot_s
    line = dict(metric=METRIC, value=v, unit="audio-s/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * tot_s / len(vals), higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic", impl="reference",
                config=dict(workload=f"{n_s} synthetic 5 s utterances per step (bounded sample of the 10k-utterance config), "
                                     f"16 kHz mono, F0 75-600 Hz, 10 ms hop, word grids + paired raw-synth stream",
                            parallelism="cpu-openmp"),
                cpu_baseline=dict(value=v, unit="audio-s/s", cores=vals[-1]["cores"], kind="port", sample=vals[-1]["sample"]),
                e2e=dict(value=v, unit="audio-s/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------- c2 / c3 / c5
def bench_steps(args, rank, world, local):
    import numpy as np
    import torch
    import torch.distributed as dist
    import bench_workloads as W
    import prosody_b200 as pb
    from prosody_b200 import shard
    from prosody_b200 import ssml as SSML
    from prosody_b200 import step as S
    dev = torch.device("cuda", local)
    placement = bind_to_gpu_numa(local, world)
    cfg = args.config
    if cfg == "c2":
        wl = W.c2(dev, n_utt=args.utts or 10000, seed=1234 + 1000 * rank)
    elif cfg == "c3":
        wl = W.c3(dev, n_utt=args.utts or 2000, seed=2345 + 1000 * rank)
    else:
        wl = W.c5(dev, rank, world, hours=args.hours)
    strong = cfg == "c5"
    pcm, segs, prosody, pitch = wl.pcm, wl.segments, wl.prosody, wl.pitch
    t_plan = time.perf_counter()
    pl = S.plan(segs, prosody)
    t_plan = time.perf_counter() - t_plan
    host_pcm = torch.empty(pcm.shape, dtype=torch.int16, pin_memory=True)
    host_pcm.copy_(pcm)
    torch.cuda.synchronize()
    F = max(1, args.in_flight)
    exs = [pb.Extractor(local) for _ in range(F)]
    pools = SSML.TextPools([segs[i].name for i in pl.syn_seg], pl.syn_words) if cfg == "c3" else None
    # the shards are fixed for the whole run: their row counts (and, for the interleaved c5 partition, the global order of the
    # gathered rows) are exchanged once, not in every step
    # The rows of a step are computed on the HOST (float64 baselines / deltas / EMA), so the gather runs over a gloo group: pushing
    # them to the GPU for a NCCL send would queue a NCCL kernel behind the persistent pitch kernels that hold every SM for tens of
    # milliseconds (round-2 measurement: that wait was the weak-scaling loss, 41.9 -> 46.8 ms per step at 8 GPUs).  NCCL stays for
    # what lives on the GPUs: the barriers and the timing reductions.  PB_BENCH_GATHER=nccl switches back.
    row_sizes, perm, ggroup = None, None, None
    use_nccl_gather = os.environ.get("PB_BENCH_GATHER", "gloo") == "nccl"
    gdev = dev if use_nccl_gather else torch.device("cpu")
    if world > 1:
        ggroup = None if use_nccl_gather else dist.new_group(backend="gloo")
        row_sizes = shard.row_counts(pl.n_syn, gdev, ggroup)
        if strong:
            gid = np.asarray(wl.extra["global_ids"], np.int64)[pl.syn_seg]
            first = np.concatenate([[0], np.cumsum(np.bincount(pl.syn_seg, minlength=pl.n_seg))])
            keys = gid * (1 << 20) + (np.arange(pl.n_syn) - first[pl.syn_seg])
            gathered = [None] * world if rank == 0 else None
            dist.gather_object(keys, gathered, dst=0)
            if rank == 0:
                perm = torch.from_numpy(np.argsort(np.concatenate(gathered), kind="stable")).to(gdev)
    gathered_host = torch.empty(sum(row_sizes), 5, dtype=torch.float64, pin_memory=True) if (world > 1 and rank == 0 and use_nccl_gather) else None

    def finish_step(ex):
        out = S.collect(ex, pl, prosody)
        if pools is not None:            # c3: full SSML-delta output, the three CSV tables built (as bytes, natively) every step
            out["ssml"] = SSML.build_csv_bytes(pools, pl.syn_pause_ms, out["sm_pitch"], out["sm_rate"], out["raw_volume"], "fr-FR-HenriNeural",
                                               prosody["inter_syntagme_pause_factor"], lib=ex._lib)
        if world > 1:
            # final gather of the per-syntagme results on rank 0 (the path's only exchange): ragged, true counts, no ids on the wire
            rows = torch.from_numpy(np.stack([out["raw_pitch"], out["raw_volume"], out["raw_rate"], out["sm_pitch"], out["sm_rate"]], 1)).to(gdev)
            # (rank 0 puts them in global order with a permutation worked out once)
            g = shard.gather_rows(rows, None, dst=0, sizes=row_sizes, perm=perm, out=gathered_host, group=ggroup)
            if rank == 0:
                out["gathered"] = g
                assert g.shape == (sum(row_sizes), 5)
        return out

    keys_t = ("frames_ms", "acf_ms", "cand_ms", "lufs_ms", "path_ms", "unit_stats_ms", "h2d_ms", "total_ms", "host_plan_ms", "n_launches", "n_frames")

    def timed(src, steps, in_flight):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        acc = {k: 0.0 for k in keys_t}
        pend, out = [], None
        for k in range(steps):
            ex = exs[k % in_flight]
            if len(pend) == in_flight:
                out = finish_step(pend.pop(0))
                for kk in acc:
                    acc[kk] += out["timings"][kk]
            S.submit(ex, src, pl, pitch)
            pend.append(ex)
        while pend:
            out = finish_step(pend.pop(0))
            for kk in acc:
                acc[kk] += out["timings"][kk]
        torch.cuda.synchronize()
        busy = time.perf_counter() - t0                 # this rank's own time, before it waits for the others
        if world > 1:
            dist.barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt, busy], dtype=torch.float64, device=dev)
            allt = torch.zeros(world, 2, dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(allt, t)
            dt = float(allt[:, 0].max().item())
            busy = [float(v) for v in allt[:, 1].tolist()]
        else:
            busy = [busy]
        return dt, acc, out, busy

    W_ = max(args.warmup, 3)
    timed(pcm, W_, F)
    sampler = ClockSampler(local)
    if rank == 0:                   # one nvidia-smi poller per node, not one per rank
        sampler.start()
    dt_dev, _, out, busy_dev = timed(pcm, args.steps, F)
    timed(host_pcm, 2, F)
    dt_e2e, _, _, busy_e2e = timed(host_pcm, args.steps, F)
    dt_ser, acc, _, _ = timed(pcm, args.steps, 1)
    sampler.stop_flag.set()
    if rank == 0:
        sampler.join(timeout=2)

    if rank == 0:
        total_audio = wl.total_audio_s if strong else world * wl.audio_s
        value = total_audio * args.steps / dt_dev
        e2e = total_audio * args.steps / dt_e2e
        n_units = len(pl.units)
        peaks, fp32_peak, fp32_src = load_peaks()
        frames_per_launch = acc["n_frames"] / args.steps
        kernel_ms = acc["acf_ms"] / args.steps                      # K1, the dominant kernel
        cand_ms = acc["cand_ms"] / args.steps
        frames_ms = acc["frames_ms"] / args.steps                   # K1 + K2 (+ the pair-position kernel): the round-1 kernel's job
        # ---- CPU baseline on a bounded sample (all cores, and one core on a smaller one) + the reference algorithm's work model
        n_cpu = max(4, min(pl.n_seg, int(args.cpu_sample * 5.0 / max(1.0, wl.audio_s / max(1, pl.n_seg)))))
        cpu = run_cpu_baseline(host_pcm.numpy(), pl, n_cpu, pitch, count_work=not strong)
        cpu1 = run_cpu_baseline(host_pcm.numpy(), pl, max(4, n_cpu // 16), pitch, threads=1, count_work=False)
        roof = None
        if not strong:
            from oracle import oracle as O
            _, g, *_ = O.pitch_geometry(segs[0].nat_nx, float(segs[0].nat_sr), params=O.pitch_params(pitch["pitch_floor"], pitch["pitch_ceiling"]))
            geom = dict(nw=g.nsamp_window, nfft=g.nsampFFT, brent_ixmax=g.brent_ixmax, max_lag=g.maximumLag)
            f_acf, f_cand, f_path = algorithmic_flops_per_frame(cpu["counters"], geom)
            fpf = f_acf + f_cand + f_path
            achieved = f_acf * frames_per_launch / (kernel_ms * 1e-3) / 1e12
            # context only: the same launches with nothing beside them (inside the step the loudness kernels share the SMs)
            alone = []
            for _ in range(2):
                exs[0].extract(pcm, pl.units, pb.pitch_params(**pitch), want_pitch=pl.want_pitch, want_lufs=np.zeros(n_units, np.uint8), lufs=False, durations=False)
                t_ = exs[0].timings()
                alone.append((t_["acf_ms"], t_["cand_ms"], t_["frames_ms"]))
            k1_alone = min(a_[0] for a_ in alone)
            sr0 = float(segs[0].nat_sr)
            alg_bytes = 2.0 * sr0 * g.dt + 8.0      # per frame: s16 in once (one hop) + f32 F0 + f32 strength
            hbm_gbs = alg_bytes * frames_per_launch / (frames_ms * 1e-3) / 1e9
            traffic = None                           # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
            try:
                tr = json.loads((ROOT / "profiles" / "r02b_acf_traffic.json").read_text())
            except Exception:       # noqa: BLE001
                tr = {}
            log2n = max(8, int(math.ceil(math.log2(g.nsamp_window + g.brent_ixmax + 1))))
            kname = f"pb_pitch_acf_kernel<{log2n}>"
            if log2n == 11 and g.nsamp_window <= 1024 and g.brent_ixmax + 2 <= 512 and os.environ.get("PB_ACF_SPLIT", "1") != "0":
                kname = "pb_pitch_acf_split_kernel (N = 2048)"       # two independent 1024-point pipelines per frame pair
            if kname in tr and cfg in ("c2", "c3"):          # captured on this geometry
                traffic = tr[kname]["bytes_per_frame"] * frames_per_launch
            roof = dict(bound="fp32", kernel=kname, achieved=achieved, peak=fp32_peak, unit="TFLOP/s", frac=achieved / fp32_peak, traffic=traffic,
                        peak_source=fp32_src,
                        note="non-tensor FP32 pipe: no stage is a dense contraction. Round 2 split the round-1 frames kernel in two: K1 pb_pitch_acf_kernel "
                             "(window, two FFTs, power spectrum, normalisation; dominant) and K2 pb_pitch_cand_kernel (peak scan, sinc refinement). achieved = "
                             f"the REFERENCE algorithm's flops for K1's part of a frame ({f_acf:.0f} of {fpf:.0f}, oracle-counted on this input) x frames per "
                             "launch / CUDA-event time of K1 inside a (serial) step, where the loudness kernels run beside it on another stream. `path` is K1+K2 "
                             "together against all of the reference's per-frame flops: the figure comparable with round 1's single frames kernel (0.19 of nominal).",
                        flops_per_frame=f_acf, frames_per_launch=int(frames_per_launch), kernel_ms=kernel_ms, kernel_ms_alone=k1_alone,
                        frac_alone=f_acf * frames_per_launch / (k1_alone * 1e-3) / 1e12 / fp32_peak,
                        cand=dict(kernel="pb_pitch_cand_kernel", kernel_ms=cand_ms, flops_per_frame=f_cand,
                                  achieved=f_cand * frames_per_launch / (cand_ms * 1e-3) / 1e12, frac=f_cand * frames_per_launch / (cand_ms * 1e-3) / 1e12 / fp32_peak),
                        path=dict(kernels="K1+K2", kernel_ms=frames_ms, flops_per_frame=f_acf + f_cand,
                                  achieved=(f_acf + f_cand) * frames_per_launch / (frames_ms * 1e-3) / 1e12,
                                  frac=(f_acf + f_cand) * frames_per_launch / (frames_ms * 1e-3) / 1e12 / fp32_peak,
                                  kernel_ms_alone=min(a_[2] for a_ in alone),
                                  frac_alone=(f_acf + f_cand) * frames_per_launch / (min(a_[2] for a_ in alone) * 1e-3) / 1e12 / fp32_peak),
                        hbm=dict(achieved=hbm_gbs, peak=peaks.get("hbm_gbs"), unit="GB/s", frac=(hbm_gbs / peaks["hbm_gbs"]) if peaks.get("hbm_gbs") else None,
                                 bytes_per_frame=alg_bytes, peak_source="MEASURED_PEAKS.json" if peaks.get("hbm_gbs") else "absent"))
        h2d_bytes = int(host_pcm.numel() * 2 + n_units * 120)
        e2e_obj = dict(value=e2e, unit="audio-s/s", h2d_bytes_per_step=h2d_bytes, d2h_bytes_per_step=int(n_units * 20), ms_per_step=1e3 * dt_e2e / args.steps,
                       h2d_gbps_per_gpu=h2d_bytes / (dt_e2e / args.steps) / 1e9)
        try:
            ceil = json.loads((ROOT / "profiles" / "r02_h2d_ceiling.json").read_text())
            c = ceil.get(str(world))
            if c:
                e2e_obj["h2d_ceiling_gbps_per_gpu"] = c["per_gpu_gbps"]
                e2e_obj["h2d_frac_of_ceiling"] = e2e_obj["h2d_gbps_per_gpu"] / c["per_gpu_gbps"]
        except Exception:           # noqa: BLE001
            pass
        line = dict(
            metric=METRIC, value=value, unit="audio-s/s", n_gpus=world, steps=args.steps, warmup=W_,
            ms_per_step=1e3 * dt_dev / args.steps, higher_is_better=True, scaling="strong" if strong else "weak", vs_baseline=None, dtype="f32",
            data="synthetic",
            config=dict(workload=wl.description + f": {n_units} measurement units on rank 0 ({pl.n_seg} utterances, {pl.n_syn} syntagme rows)",
                        name=cfg, units_rank0=n_units, pitch_frames_per_step_rank0=int(frames_per_launch), parallelism=f"units sharded over {world} GPU(s)",
                        in_flight=F, l2=f"inputs ({host_pcm.numel() * 2 / 1e9:.1f} GB PCM per GPU) exceed the 126 MB L2; no flush needed",
                        timing="wall clock between barrier+synchronize, max over ranks; kernel times from CUDA events on the launch stream (serial pass)",
                        audio_hours_per_s=value / 3600.0, placement=placement, host_plan_setup_s=t_plan),
            e2e=e2e_obj,
            serial=dict(ms_per_step=1e3 * dt_ser / args.steps, value=total_audio * args.steps / dt_ser,
                        note="one step at a time (no overlap of host planning / post-processing with GPU work): the latency of a step"),
            gpu_launches=int(acc["n_launches"]),
            kernels_ms_per_step={k: acc[k] / args.steps for k in ("unit_stats_ms", "frames_ms", "acf_ms", "cand_ms", "path_ms", "lufs_ms", "h2d_ms", "total_ms", "host_plan_ms")},
            roofline=roof,
            cpu_baseline=dict(value=cpu["value"], unit="audio-s/s", cores=cpu["cores"], kind="port", sample=cpu["sample"], pitch_s=cpu["pitch_s"], lufs_s=cpu["lufs_s"],
                              single_core=dict(value=cpu1["value"], unit="audio-s/s", cores=1, audio_s=cpu1["audio_s"])),
            clocks=sampler.summary())
        # north_star's per-frame tolerances on a bounded sample of this workload: the first 64 natural utterances, every frame, against
        # the oracle (outside every timed region)
        try:
            line["parity"] = dict(frames=frame_level_parity(exs[0], host_pcm.numpy(), segs[:64], pitch),
                                  sample="first 64 natural utterances of rank 0, whole files, every frame against the float64 oracle; bars: voicing "
                                         "agreement >= 0.995, F0 within 5e-3 on frames voiced in both, intensity within 0.05 dB")
        except Exception as e:          # noqa: BLE001  (the measurement stands without it)
            line["parity"] = dict(error=repr(e))
        if world > 1 or strong:
            mean_b = sum(busy_dev) / len(busy_dev)
            line["ranks"] = dict(busy_s_resident=busy_dev, busy_s_e2e=busy_e2e, imbalance_resident=max(busy_dev) / mean_b if mean_b else None,
                                 planned_imbalance=wl.extra.get("planned_imbalance"), audio_s_rank0=wl.audio_s, total_audio_s=total_audio)
        print(json.dumps(line))
    for ex in exs:
        ex.close()


# ----------------------------------------------------------------------------------------------------------- c1: the repo clips, file level
def frame_level_parity(ex, pcm, segs, pitch):
    """north_star's tolerances are stated per FRAME (F0 within 0.5 % on frames voiced in both, voicing decisions agree on >= 99.5 % of
    frames, intensity within 0.05 dB): every whole natural clip, frame by frame, against the oracle."""
    import numpy as np
    import prosody_b200 as pb
    from oracle import oracle as O
    items = [(s.nat_off, s.nat_nx, s.nat_sr, 0.0, None) for s in segs]
    r = ex.median_pitch(pcm, pb.Units.from_list(items), pb.pitch_params(**pitch), frames=True)
    n = agree = n_both = 0
    rel, st_err, in_err = [], 0.0, 0.0
    for i, s in enumerate(segs):
        o = O.pitch_track(pcm[s.nat_off:s.nat_off + s.nat_nx], s.nat_sr, params=O.pitch_params(pitch["pitch_floor"], pitch["pitch_ceiling"]))
        a, b = r["frame_off"][i], r["frame_off"][i + 1]
        f, g = r["frame_f0"][a:b].astype(np.float64), o["frequency"]
        n += len(g); agree += int(np.sum((f > 0) == (g > 0)))
        both = (f > 0) & (g > 0); n_both += int(both.sum())
        rel.append(np.abs(f[both] - g[both]) / g[both])
        st_err = max(st_err, float(np.max(np.abs(r["frame_strength"][a:b] - o["strength"]))))
        gi, oi = r["frame_intensity"][a:b].astype(np.float64), o["intensity"]
        pos = (gi > 0) & (oi > 0)
        if pos.any():
            in_err = max(in_err, float(np.max(np.abs(20.0 * np.log10(gi[pos] / oi[pos])))))
    rel = np.concatenate(rel) if rel else np.zeros(1)
    return dict(frames=int(n), voicing_agreement=agree / max(1, n), voiced_in_both=int(n_both), f0_rel_err_p50=float(np.median(rel)),
                f0_rel_err_p99=float(np.quantile(rel, 0.99)), f0_rel_err_max=float(rel.max()), frames_over_0p5_percent=int(np.sum(rel > 5e-3)),
                strength_abs_err_max=st_err, frame_intensity_err_max_db=in_err)


def bench_c1(args, rank, world, local):
    """BASELINE configs[0]: the ten repo clips through the file-level drop-in (WAV + TextGrid read from disk, three CSVs written),
    next to the oracle's loop-by-loop restatement of the reference step on the same files, plus the flip listing against it."""
    if rank != 0:
        return
    import types
    import numpy as np
    import torch
    import bench_workloads as W
    import prosody_b200 as pb
    from prosody_b200 import pipeline as P
    from prosody_b200 import step as S
    sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "tests" / "golden"))
    import make_c1_oracle_golden as G
    from parity_report import column_report
    wl = W.c1(torch.device("cuda", local))
    v, root = wl.extra["voice"], wl.extra["root"]
    res = root / "Out"
    self = types.SimpleNamespace(
        voice_dir=v["voice_dir"], raw_audio_dir=v["raw_audio_dir"], textgrid_dir=v["textgrid_dir"], p_st=1.3, pitch_lower_clip_factor=0.7, v_pct=7.0,
        r_pct_clamp=15.0, alpha=0.2, max_jump=5.0, end_pause_ms=400, baseline_window=None, inter_syntagme_pause_factor=1,
        threshold_duration_before_slowing_down=1.0, slow_floor_per_sec=2.0, azure_voice="fr-FR-HenriNeural",
        bdd_ssml_csv=res / "BDD_ssml.csv", bdd_syntagme_ssml_csv=res / "BDD_syntagme_ssml.csv", bdd_syntagme_synth_csv=res / "BDD_syntagme_for_synth.csv")
    ex = pb.Extractor(local)
    for _ in range(max(3, args.warmup)):
        out = P.measure_prosody_and_build_ssml(self, extractor=ex, pos_of=G.pos_of)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(args.steps):
        out = P.measure_prosody_and_build_ssml(self, extractor=ex, pos_of=G.pos_of)
    torch.cuda.synchronize(); dt_e2e = time.perf_counter() - t0
    # resident: the same units with the PCM already in HBM (no file reading, no CSV writing)
    pcm, segs = P.load_voice(self.voice_dir / "audio", self.raw_audio_dir, self.textgrid_dir)
    pl = S.plan(segs, wl.prosody, G.pos_of)
    d_pcm = torch.from_numpy(pcm).to(f"cuda:{local}")
    for _ in range(3):
        S.measure(ex, d_pcm, pl, wl.prosody)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(args.steps):
        o2 = S.measure(ex, d_pcm, pl, wl.prosody)
    torch.cuda.synchronize(); dt_dev = time.perf_counter() - t0
    # the reference step restated (oracle/flow.py), single Python thread like the reference's own loop
    t0 = time.perf_counter(); _, ref = G.run(root / "oracle_copy"); dt_cpu = time.perf_counter() - t0
    rep = column_report(out["sm_pitch"], ref["sm_p"])
    frames = frame_level_parity(ex, pcm, segs, wl.pitch)
    pu = np.array([u["p_nat"] for u in ref["units"]]); got = out["syn"]["p_nat"]; both = (pu > 0) & (got > 0)
    line = dict(metric=METRIC, value=wl.audio_s * args.steps / dt_dev, unit="audio-s/s", n_gpus=1, steps=args.steps, warmup=max(3, args.warmup),
                ms_per_step=1e3 * dt_dev / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="reference clips + synthetic alignments",
                config=dict(workload=wl.description, name="c1", units=len(pl.units), rows=pl.n_syn),
                e2e=dict(value=wl.audio_s * args.steps / dt_e2e, unit="audio-s/s", ms_per_step=1e3 * dt_e2e / args.steps,
                         h2d_bytes_per_step=int(pcm.nbytes), d2h_bytes_per_step=int(len(pl.units) * 20), note="files read from disk and CSVs written inside the timed region"),
                gpu_launches=int(o2["timings"]["n_launches"]),
                kernels_ms_per_step={k: float(o2["timings"][k]) for k in ("frames_ms", "acf_ms", "cand_ms", "path_ms", "lufs_ms", "total_ms", "host_plan_ms")},
                cpu_baseline=dict(value=wl.audio_s / dt_cpu, unit="audio-s/s", cores=1, kind="port", sample="the whole config: oracle/flow.py, one Python thread"),
                parity=dict(frames=frames, pitch_strings=dict(rows=rep["rows"], identical=rep["identical"], flipped=rep["flipped"], max_abs_delta=rep["max_abs_delta"], listed=rep["listed"]),
                            rate_strings_flipped=column_report(out["sm_rate"], ref["sm_r"])["flipped"],
                            volume_strings_flipped=column_report(out["raw_volume"], [r["raw_volume"] for r in ref["raw_rows"]])["flipped"],
                            median_f0_rel_err_max=float(np.max(np.abs(got[both] - pu[both]) / pu[both])), voicing_mismatch_units=int(np.sum((pu > 0) != (got > 0))),
                            lufs_abs_err_max_db=float(np.max(np.abs(out["syn"]["l_syn"] - np.array([u["l_syn"] for u in ref["units"]]))))))
    print(json.dumps(line))
    ex.close()


# ----------------------------------------------------------------------------------------------------------- c4: long-form recordings
def bench_c4(args, rank, world, local):
    """BASELINE configs[3]: N x 1 h @ 22.05 kHz — GPU silence segmentation, every segment analysed as its own file, and the
    unsegmented hours through one Viterbi chain each (359 997 frames)."""
    if rank != 0:
        return
    import numpy as np
    import torch
    import bench_workloads as W
    import prosody_b200 as pb
    n_h = max(1, int(min(args.hours, 8)))
    dev = torch.device("cuda", local)
    pcm, sr, per = W.c4_recordings(dev, n_h)
    ex = pb.Extractor(local)
    whole = pb.Units.from_list([(h * per, per, sr, 0.0, None, float(sr)) for h in range(n_h)])
    p = pb.pitch_params(75.0, 600.0)

    def one():
        t = {}
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s = ex.split_on_silence(pcm, whole, 1000, -50, 300)
        torch.cuda.synchronize(); t["segmentation_ms"] = 1e3 * (time.perf_counter() - t0)
        base = np.repeat(np.arange(n_h) * per, np.diff(s["seg_off"]))
        seg_units = pb.Units(base + s["first_sample"], s["n_samples"].astype(np.int64), np.full(len(base), float(sr)), np.zeros(len(base), np.int32),
                             np.zeros(len(base)), np.zeros(len(base)), np.full(len(base), float(sr)))
        t0 = time.perf_counter()
        e = ex.extract(pcm, seg_units, p)
        torch.cuda.synchronize(); t["segments_ms"] = 1e3 * (time.perf_counter() - t0); t["segments_kernels"] = ex.timings()
        t0 = time.perf_counter()
        w = ex.extract(pcm, whole, p)
        torch.cuda.synchronize(); t["unsegmented_ms"] = 1e3 * (time.perf_counter() - t0); t["unsegmented_kernels"] = ex.timings()
        t["n_segments"] = int(len(base)); t["frames_segments"] = int(e["n_frames"].sum()); t["frames_unsegmented"] = int(w["n_frames"].sum())
        return t
    for _ in range(max(3, args.warmup)):
        one()
    runs = [one() for _ in range(args.steps)]
    med = lambda k: float(np.median([r[k] for r in runs]))
    audio_s = n_h * 3600.0
    step_ms = med("segmentation_ms") + med("segments_ms")
    kk = lambda r, k: {q: float(r[k][q]) for q in ("unit_stats_ms", "frames_ms", "acf_ms", "cand_ms", "path_ms", "lufs_ms", "total_ms", "host_plan_ms")}
    line = dict(metric=METRIC, value=audio_s / (step_ms * 1e-3), unit="audio-s/s", n_gpus=1, steps=args.steps, warmup=max(3, args.warmup), ms_per_step=step_ms,
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=f"{n_h} x 1 h synthetic recordings, 22.05 kHz mono s16, a 1.1-2.3 s pause every 19 s: split_on_silence(1000 ms, -50 dBFS, keep 300) on the GPU, "
                                     f"then every segment analysed as its own file (F0 75-600 Hz + loudness); PCM resident", name="c4",
                            n_segments=runs[-1]["n_segments"], frames_segments=runs[-1]["frames_segments"]),
                segmentation_ms=med("segmentation_ms"), segments_ms=med("segments_ms"), segments_kernels=kk(runs[-1], "segments_kernels"),
                unsegmented=dict(ms=med("unsegmented_ms"), value=audio_s / (med("unsegmented_ms") * 1e-3), frames=runs[-1]["frames_unsegmented"],
                                 kernels=kk(runs[-1], "unsegmented_kernels"),
                                 note="one Viterbi chain per hour (359 997 frames each): K3 cuts it into 512-frame blocks ((max, +) transfer matrices, back-maps), K0 and the loudness peak / state scan / gates of such a unit are spread over the grid as well (round 2; one warp per recording took 230 ms in K3 and 86 ms in K4)"),
                gpu_launches=int(runs[-1]["segments_kernels"]["n_launches"]))
    print(json.dumps(line))
    ex.close()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.config == "c1":
        bench_c1(args, rank, world, local)
    elif args.config == "c4":
        bench_c4(args, rank, world, local)
    else:
        bench_steps(args, rank, world, local)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
