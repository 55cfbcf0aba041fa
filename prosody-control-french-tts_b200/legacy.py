"""Batched forms of the legacy DataFrame pipeline's two per-syntagme measurements ("next" rows).

  calculate_pitch_segment(path, start, end)   /root/reference/Code/Pipeline/compute_pitch_adjustments.py:166-208
  _calculate_loudness(path, start, end)       /root/reference/Code/Pipeline/compute_loudness_adjustments.py:8-25

The reference applies them row by row with ``df.apply``; here every row of the frame is one unit of a GPU batch.
"""
from __future__ import annotations

import numpy as np

from . import _native as N
from .batch import Extractor, Units, pitch_params

PITCH_FLOORS = (75.0, 100.0, 150.0, 200.0)   # the floors the reference tries in turn


def pitch_segments(ex: Extractor, pcm, units: Units) -> np.ndarray:
    """calculate_pitch_segment for every unit: 0 for invalid times or when no floor yields a voiced frame, else the
    geometric mean of the voiced frequencies found with the first floor that has any.  units.t0/t1 are start/end."""
    n = len(units)
    out = np.zeros(n)
    total = units.file_nx.astype(np.float64) * (1.0 / units.rate)          # Sound.get_total_duration() = nx * dx
    todo = np.flatnonzero(~((units.t0 >= units.t1) | (units.t0 < 0) | (units.t1 > total)))
    for floor in PITCH_FLOORS:
        if len(todo) == 0:
            break
        sub = units.select(todo)
        sub.has_t1 = np.full(len(todo), 2, np.int32)                         # extract_part(from_time, to_time): times not preserved
        r = ex.median_pitch(pcm, sub, pitch_params(pitch_floor=floor, pitch_ceiling=600.0), frames=True)
        fo, f0 = r["frame_off"], r["frame_f0"].astype(np.float64)
        voiced = f0 > 0
        cnt = np.add.reduceat(np.append(voiced, False).astype(np.int64), fo[:-1])
        cnt[np.diff(fo) == 0] = 0
        logs = np.where(voiced, np.log(np.where(voiced, f0, 1.0)), 0.0)
        sums = np.add.reduceat(np.append(logs, 0.0), fo[:-1])
        done = (r["status"] == N.PB_UNIT_OK) & (cnt > 0)
        out[todo[done]] = np.exp(sums[done] / cnt[done])
        todo = todo[~done]
    return out


def loudness_segments(ex: Extractor, pcm, units: Units) -> np.ndarray:
    """_calculate_loudness for every unit (RMS dB with numpy's int16 wrap-around of ``samples ** 2``)."""
    return ex.legacy_loudness(pcm, units)
