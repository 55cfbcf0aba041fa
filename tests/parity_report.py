"""TEST INFRASTRUCTURE: per-run parity listing of the step's outputs against the oracle ("flips listed per run").

The reference prints every delta with f'{x:+.2f}%' (Code/audioPipeline.py:610-612), i.e. on a 0.01 lattice whose rounding
boundaries sit at odd multiples of 0.005.  The loudness, duration and rate paths are float64 end to end and reproduce the
oracle's strings exactly; the pitch path carries the F0 of the FP32 kernels, so a value that lies within the F0 error of a
rounding boundary can print the neighbouring lattice point: a "flip".  This module measures that per run: rows, identical
strings, max |delta|, and every flipped row with both values.
"""
from __future__ import annotations

import os

import numpy as np


def oracle_units(pcm_host: np.ndarray, pl, pitch_floor: float, pitch_ceiling: float, threads: int = 0):
    """Per-unit oracle measurements laid out like the plan's units -> (median_f0, lufs, duration_s)."""
    from oracle import oracle as O
    O.build()
    u = pl.units
    n = len(u)
    threads = threads or os.cpu_count() or 1
    med = np.zeros(n); lufs = np.full(n, np.nan); dur = np.zeros(n)
    pk = np.nonzero(pl.want_pitch)[0]; lk = np.nonzero(pl.want_lufs)[0]
    m, nv, nf, st = O.batch_median_pitch(pcm_host, u.file_off[pk], u.file_nx[pk], u.rate[pk], u.has_t1[pk], u.t0[pk], u.t1[pk],
                                         O.pitch_params(pitch_floor, pitch_ceiling), threads)
    med[pk] = m
    a = np.zeros(len(lk), np.int64); b = np.zeros(len(lk), np.int64); npad = np.zeros(len(lk), np.int64)
    for j, i in enumerate(lk):
        a[j], b[j], npad[j], _ = O.lufs_resolve(int(u.file_nx[i]), int(u.rate[i]), float(u.meter_rate[i]), float(u.t0[i]),
                                                float(u.t1[i]) if u.has_t1[i] else None)
    lufs[lk], _ = O.batch_lufs(pcm_host, u.file_off[lk], a, b, npad, u.meter_rate[lk], threads)
    for i in range(n):
        dur[i] = O.part_duration(int(u.file_nx[i]), int(u.rate[i]), float(u.t0[i]), float(u.t1[i]) if u.has_t1[i] else None)
    return med, lufs, dur


def _fmt(x):
    return f"{float(x):+.2f}"


def column_report(ours, ref, max_listed=200) -> dict:
    ours = np.asarray(ours, np.float64); ref = np.asarray(ref, np.float64)
    so = [_fmt(v) for v in ours]; sr = [_fmt(v) for v in ref]
    flips = [i for i in range(len(so)) if so[i] != sr[i]]
    d = np.abs(ours - ref)
    one_step = sum(1 for i in flips if abs(float(so[i]) - float(sr[i])) <= 0.0101)
    return dict(rows=len(so), identical=len(so) - len(flips), flipped=len(flips), flip_rate=len(flips) / max(1, len(so)),
                flips_one_lattice_step=one_step, max_abs_delta=float(d.max()) if len(d) else 0.0,
                bit_identical=int(np.sum(ours == ref)),
                listed=[dict(row=int(i), ours=so[i], ref=sr[i], ours_value=float(ours[i]), ref_value=float(ref[i])) for i in flips[:max_listed]])


def flip_report(out, ref, max_listed=200) -> dict:
    """out / ref: dicts as returned by prosody_b200.step.measure / finish (ours, and the same host math on oracle units)."""
    rep = dict(pitch=column_report(out["sm_pitch"], ref["sm_pitch"], max_listed),
               rate=column_report(out["sm_rate"], ref["sm_rate"], max_listed),
               volume=column_report(out["raw_volume"], ref["raw_volume"], max_listed))
    p, q = np.asarray(out["syn"]["p_nat"]), np.asarray(ref["syn"]["p_nat"])
    both = (p > 0) & (q > 0)
    rel = np.abs(p[both] - q[both]) / q[both]
    rep["median_f0"] = dict(units=int(len(p)), voiced_in_both=int(both.sum()), voicing_mismatch=int(np.sum((p > 0) != (q > 0))),
                            rel_err_max=float(rel.max()) if len(rel) else 0.0,
                            rel_err_p50=float(np.percentile(rel, 50)) if len(rel) else 0.0,
                            rel_err_p99=float(np.percentile(rel, 99)) if len(rel) else 0.0,
                            bit_identical=int(np.sum(p == q)))
    l, m = np.asarray(out["syn"]["l_syn"]), np.asarray(ref["syn"]["l_syn"])
    fin = np.isfinite(l) & np.isfinite(m)
    rep["lufs"] = dict(max_abs_db=float(np.max(np.abs(l[fin] - m[fin]))) if fin.any() else 0.0, nonfinite_mismatch=int(np.sum(np.isfinite(l) != np.isfinite(m))))
    rep["durations_identical"] = bool(np.array_equal(out["syn"]["nat_total"], ref["syn"]["nat_total"]) and np.array_equal(out["syn"]["syn_total"], ref["syn"]["syn_total"]))
    return rep
