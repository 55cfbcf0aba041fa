"""Batch API over the C ABI: many (file, t0, t1) units per call, PCM resident on the GPU or staged from the host.

One `Extractor` = one PbHandle = one GPU + one stream.  Units mirror the arguments of the reference closures
(/root/reference/Code/audioPipeline.py:314-361): a file, optional slice times in seconds, and for loudness the
rate the caller's pyln.Meter was built with.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _native as N


def pitch_params(pitch_floor: float = 150.0, pitch_ceiling: float = 600.0, time_step: float = 0.0, **kw) -> N.PbPitchParams:
    """parselmouth ``Sound.to_pitch(time_step, pitch_floor, pitch_ceiling)``; defaults are the reference's (150, 600)."""
    p = N.PbPitchParams(time_step, pitch_floor, pitch_ceiling, 3.0, 0.03, 0.45, 0.01, 0.35, 0.14, 15, 0)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError(f"unknown pitch parameter {k}")
        setattr(p, k, v)
    return p


@dataclass
class Units:
    """SoA description of the units of one call (all arrays length n)."""
    file_off: np.ndarray      # int64  sample offset of the unit's file inside the pcm buffer
    file_nx: np.ndarray       # int64  samples in that file
    rate: np.ndarray          # float64
    has_t1: np.ndarray        # int32  0 = whole file
    t0: np.ndarray            # float64 seconds
    t1: np.ndarray            # float64 seconds
    meter_rate: np.ndarray | None = None   # float64, rate of the pyln.Meter used for this call (default: file rate)

    def __post_init__(self):
        self.file_off = np.ascontiguousarray(self.file_off, np.int64)
        self.file_nx = np.ascontiguousarray(self.file_nx, np.int64)
        self.rate = np.ascontiguousarray(self.rate, np.float64)
        self.has_t1 = np.ascontiguousarray(self.has_t1, np.int32)
        self.t0 = np.ascontiguousarray(self.t0, np.float64)
        self.t1 = np.ascontiguousarray(self.t1, np.float64)
        if self.meter_rate is not None:
            self.meter_rate = np.ascontiguousarray(self.meter_rate, np.float64)
        n = len(self.file_off)
        for a in (self.file_nx, self.rate, self.has_t1, self.t0, self.t1) + (() if self.meter_rate is None else (self.meter_rate,)):
            if len(a) != n:
                raise ValueError("unit arrays must have the same length")

    def __len__(self) -> int:
        return len(self.file_off)

    @classmethod
    def from_list(cls, items) -> "Units":
        """items: iterable of (file_off, file_nx, rate, t0, t1_or_None[, meter_rate])."""
        items = list(items)
        has_mr = any(len(it) > 5 for it in items)
        return cls(file_off=[it[0] for it in items], file_nx=[it[1] for it in items], rate=[it[2] for it in items],
                   has_t1=[0 if it[4] is None else 1 for it in items], t0=[it[3] for it in items],
                   t1=[0.0 if it[4] is None else it[4] for it in items],
                   meter_rate=[(it[5] if len(it) > 5 else it[2]) for it in items] if has_mr else None)

    def select(self, idx) -> "Units":
        return Units(self.file_off[idx], self.file_nx[idx], self.rate[idx], self.has_t1[idx], self.t0[idx], self.t1[idx],
                     None if self.meter_rate is None else self.meter_rate[idx])

    def c_struct(self) -> N.PbUnits:
        def p(a, t):
            return a.ctypes.data_as(C.POINTER(t))
        return N.PbUnits(len(self), p(self.file_off, C.c_int64), p(self.file_nx, C.c_int64), p(self.rate, C.c_double),
                         p(self.has_t1, C.c_int32), p(self.t0, C.c_double), p(self.t1, C.c_double),
                         p(self.meter_rate, C.c_double) if self.meter_rate is not None else None)


def _pcm_pointer(pcm):
    """-> (address, n_samples, on_device, keepalive). Accepts numpy int16, torch CPU (pinned or not) or CUDA int16 tensors."""
    if isinstance(pcm, np.ndarray):
        if pcm.dtype != np.int16 or not pcm.flags.c_contiguous:
            pcm = np.ascontiguousarray(pcm, np.int16)
        return pcm.ctypes.data, pcm.size, 0, pcm
    try:
        import torch
    except ImportError:  # pragma: no cover
        torch = None
    if torch is not None and isinstance(pcm, torch.Tensor):
        if pcm.dtype != torch.int16:
            raise TypeError("pcm tensor must be int16")
        pcm = pcm.contiguous()
        if pcm.is_cuda:
            # the library reads the tensor on its own stream: whatever torch still has queued to produce it must be done
            torch.cuda.current_stream(pcm.device).synchronize()
        return pcm.data_ptr(), pcm.numel(), int(pcm.is_cuda), pcm
    raise TypeError("pcm must be a numpy int16 array or a torch int16 tensor")


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Extractor:
    """Owns one native handle. Not thread-safe (one per GPU / stream)."""

    def __init__(self, device: int = 0, lib=None):
        self._lib = lib if lib is not None else N.load()
        self._h = C.c_void_p()
        rc = self._lib.pb_create(int(device), C.byref(self._h))
        if rc != N.PB_OK:
            self._h = None
            raise N.NativeError(f"pb_create(device={device}) failed: {N._ERR_NAMES.get(rc, rc)} — the CUDA path is the only path")
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ------------------------------------------------------------------ info
    def set_stream(self, cuda_stream_ptr: int | None):
        N.check(self._lib, self._h, self._lib.pb_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)), "pb_set_stream")

    def timings(self) -> dict:
        t = N.PbTimings()
        N.check(self._lib, self._h, self._lib.pb_get_timings(self._h, C.byref(t)), "pb_get_timings")
        return {n: getattr(t, n) for n, _ in N.PbTimings._fields_ if n != "reserved"}

    def device_info(self) -> dict:
        sm, ma, mi, mem = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
        N.check(self._lib, self._h, self._lib.pb_device_info(self._h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)), "pb_device_info")
        return dict(sm_count=sm.value, cc=(ma.value, mi.value), total_mem=mem.value)

    # ------------------------------------------------------------------ planning (host only)
    def pitch_plan(self, units: Units, params: N.PbPitchParams | None = None):
        return pitch_plan(units, params, self._lib)

    # ------------------------------------------------------------------ batched closures
    def median_pitch(self, pcm, units: Units, params: N.PbPitchParams | None = None, frames: bool = False) -> dict:
        """get_median_pitch for every unit. frames=True also returns parselmouth's selected_array per frame."""
        params = params or pitch_params()
        addr, n_samp, on_dev, keep = _pcm_pointer(pcm)
        n = len(units)
        med = np.zeros(n); nv = np.zeros(n, np.int32); nf = np.zeros(n, np.int32); st = np.zeros(n, np.int32)
        f0 = fs = fi = None
        frame_off = None
        if frames:
            st0, nf0, frame_off = self.pitch_plan(units, params)
            tot = int(frame_off[-1])
            f0 = np.zeros(tot, np.float32); fs = np.zeros(tot, np.float32); fi = np.zeros(tot, np.float32)
        cu = units.c_struct()
        rc = self._lib.pb_median_pitch_batch(self._h, C.c_void_p(addr), n_samp, on_dev, C.byref(cu), C.byref(params),
                                             _ptr(med, C.c_double), _ptr(nv, C.c_int32), _ptr(nf, C.c_int32), _ptr(st, C.c_int32),
                                             _ptr(f0, C.c_float) if frames else None, _ptr(fs, C.c_float) if frames else None,
                                             _ptr(fi, C.c_float) if frames else None)
        N.check(self._lib, self._h, rc, "pb_median_pitch_batch")
        del keep
        return dict(median_f0=med, n_voiced=nv, n_frames=nf, status=st, frame_off=frame_off, frame_f0=f0,
                    frame_strength=fs, frame_intensity=fi)

    def lufs(self, pcm, units: Units):
        addr, n_samp, on_dev, keep = _pcm_pointer(pcm)
        n = len(units)
        out = np.zeros(n); st = np.zeros(n, np.int32)
        cu = units.c_struct()
        rc = self._lib.pb_lufs_batch(self._h, C.c_void_p(addr), n_samp, on_dev, C.byref(cu), _ptr(out, C.c_double), _ptr(st, C.c_int32))
        N.check(self._lib, self._h, rc, "pb_lufs_batch")
        del keep
        return out, st

    def extract(self, pcm, units: Units, params: N.PbPitchParams | None = None, want_pitch=None, want_lufs=None,
                pitch: bool = True, lufs: bool = True, durations: bool = True) -> dict:
        """All per-unit measurements of one pass in one call (the end-to-end entry point)."""
        params = params or pitch_params()
        addr, n_samp, on_dev, keep = _pcm_pointer(pcm)
        n = len(units)
        med = np.zeros(n); nv = np.zeros(n, np.int32); nf = np.zeros(n, np.int32); st = np.zeros(n, np.int32)
        lu = np.full(n, np.nan); du = np.zeros(n)
        wp = None if want_pitch is None else np.ascontiguousarray(want_pitch, np.uint8)
        wl = None if want_lufs is None else np.ascontiguousarray(want_lufs, np.uint8)
        cu = units.c_struct()
        rc = self._lib.pb_extract_batch(self._h, C.c_void_p(addr), n_samp, on_dev, C.byref(cu), C.byref(params),
                                        _ptr(wp, C.c_uint8) if wp is not None else None, _ptr(wl, C.c_uint8) if wl is not None else None,
                                        _ptr(med, C.c_double) if pitch else None, _ptr(nv, C.c_int32) if pitch else None,
                                        _ptr(nf, C.c_int32) if pitch else None, _ptr(lu, C.c_double) if lufs else None,
                                        _ptr(du, C.c_double) if durations else None, _ptr(st, C.c_int32))
        N.check(self._lib, self._h, rc, "pb_extract_batch")
        del keep
        return dict(median_f0=med, n_voiced=nv, n_frames=nf, lufs=lu, duration_s=du, status=st)

    def submit(self, pcm, units: Units, params: N.PbPitchParams | None = None, want_pitch=None, want_lufs=None) -> None:
        """First half of extract(): plan the units, enqueue uploads / kernels / result download, return at once.  The host is free
        (to plan the next batch on ANOTHER Extractor, to post-process the previous one) while the GPU works; wait() collects."""
        params = params or pitch_params()
        addr, n_samp, on_dev, keep = _pcm_pointer(pcm)
        n = len(units)
        out = dict(median_f0=np.zeros(n), n_voiced=np.zeros(n, np.int32), n_frames=np.zeros(n, np.int32), lufs=np.full(n, np.nan),
                   duration_s=np.zeros(n), status=np.zeros(n, np.int32))
        wp = None if want_pitch is None else np.ascontiguousarray(want_pitch, np.uint8)
        wl = None if want_lufs is None else np.ascontiguousarray(want_lufs, np.uint8)
        cu = units.c_struct()
        rc = self._lib.pb_extract_submit(self._h, C.c_void_p(addr), n_samp, on_dev, C.byref(cu), C.byref(params),
                                         _ptr(wp, C.c_uint8) if wp is not None else None, _ptr(wl, C.c_uint8) if wl is not None else None,
                                         _ptr(out["median_f0"], C.c_double), _ptr(out["n_voiced"], C.c_int32), _ptr(out["n_frames"], C.c_int32),
                                         _ptr(out["lufs"], C.c_double), _ptr(out["duration_s"], C.c_double), _ptr(out["status"], C.c_int32))
        N.check(self._lib, self._h, rc, "pb_extract_submit")
        self._pending = (out, keep, units, wp, wl)          # keeps the PCM and the output arrays alive until wait()

    def wait(self) -> dict:
        """Second half of extract(): blocks until the submitted batch is done, returns its per-unit results."""
        if getattr(self, "_pending", None) is None:
            raise RuntimeError("no submitted batch to wait for")
        out = self._pending[0]
        rc = self._lib.pb_extract_wait(self._h)
        self._pending = None
        N.check(self._lib, self._h, rc, "pb_extract_wait")
        return out


    def intensity(self, pcm, units: Units, minimum_pitch: float = 100.0, time_step: float = 0.0, subtract_mean: bool = True) -> dict:
        """parselmouth ``Sound.to_intensity(minimum_pitch, time_step, subtract_mean)`` for every (whole-file) unit.
        Returns the dB values of all files back to back plus frame_off / t_first / dt to place them in time."""
        addr, n_samp, on_dev, keep = _pcm_pointer(pcm)
        st, nf, fo, t1, dt = intensity_plan(units, minimum_pitch, time_step, self._lib)
        out = np.zeros(int(fo[-1]), np.float32)
        st2 = np.zeros(len(units), np.int32)
        cu = units.c_struct()
        rc = self._lib.pb_intensity_batch(self._h, C.c_void_p(addr), n_samp, on_dev, C.byref(cu), float(minimum_pitch), float(time_step),
                                          int(bool(subtract_mean)), _ptr(out, C.c_float), _ptr(st2, C.c_int32))
        N.check(self._lib, self._h, rc, "pb_intensity_batch")
        del keep
        return dict(intensity_db=out, frame_off=fo, n_frames=nf, t_first=t1, dt=dt, status=st2)

    def legacy_loudness(self, pcm, units: Units) -> np.ndarray:
        """_calculate_loudness(path, start, end) of the legacy DataFrame pipeline for every unit (t0/t1 = start/end seconds)."""
        addr, n_samp, on_dev, keep = _pcm_pointer(pcm)
        out = np.zeros(len(units)); sums = np.zeros(len(units), np.int64); cnt = np.zeros(len(units), np.int64)
        cu = units.c_struct()
        rc = self._lib.pb_legacy_loudness_batch(self._h, C.c_void_p(addr), n_samp, on_dev, C.byref(cu), _ptr(out, C.c_double),
                                                _ptr(sums, C.c_int64), _ptr(cnt, C.c_int64))
        N.check(self._lib, self._h, rc, "pb_legacy_loudness_batch")
        del keep
        # the reference finishes with numpy (np.mean -> np.sqrt -> np.log10); do the same on the exact integer sums so
        # the last bit matches numpy's log10 rather than libm's
        with np.errstate(divide="ignore", invalid="ignore"):
            mean = np.where(cnt > 0, sums.astype(np.float64) / np.maximum(cnt, 1), np.nan)
            return 20 * np.log10(np.sqrt(np.abs(mean)))


    def split_on_silence(self, pcm, units: Units, min_silence_len: int = 1000, silence_thresh: float = -16, keep_silence=100) -> dict:
        """pydub.silence.split_on_silence for every (whole, mono) file of the batch; defaults are pydub's.
        Returns seg_off (n+1), start_ms / end_ms of every segment and its sample range (first, n_samples, n_pad) in the file."""
        addr, n_samp, on_dev, keep = _pcm_pointer(pcm)
        n = len(units)
        keep_ms = (-1 if keep_silence else 0) if isinstance(keep_silence, bool) else int(keep_silence)
        cu = units.c_struct()
        cap = int(self._lib.pb_split_on_silence_bound(C.byref(cu), int(min_silence_len)))
        if cap < 0:
            raise ValueError("min_silence_len must be >= 1 ms")
        so = np.zeros(n + 1, np.int64); s_ms = np.zeros(cap, np.int32); e_ms = np.zeros(cap, np.int32)
        first = np.zeros(cap, np.int64); ns = np.zeros(cap, np.int32); npad = np.zeros(cap, np.int32)
        rc = self._lib.pb_split_on_silence_batch(self._h, C.c_void_p(addr), n_samp, on_dev, C.byref(cu), int(min_silence_len),
                                                 float(silence_thresh), keep_ms, cap, _ptr(so, C.c_int64), _ptr(s_ms, C.c_int32),
                                                 _ptr(e_ms, C.c_int32), _ptr(first, C.c_int64), _ptr(ns, C.c_int32), _ptr(npad, C.c_int32))
        N.check(self._lib, self._h, rc, "pb_split_on_silence_batch")
        del keep
        m = int(so[-1])
        return dict(seg_off=so, start_ms=s_ms[:m], end_ms=e_ms[:m], first_sample=first[:m], n_samples=ns[:m], n_pad=npad[:m])


    def reduce_intervals(self, frame_off, t_first, dt, f0, intervals, track2=None) -> dict:
        """Per-frame tracks -> per-interval statistics on the GPU (pb_reduce_intervals).
        frame_off / t_first / dt describe the series (e.g. from median_pitch(frames=True) and pitch_frame_times);
        intervals = (series index, tmin, tmax) arrays or a list of such triples; f0 / track2 numpy float32 or CUDA tensors."""
        if isinstance(intervals, (list, tuple)) and len(intervals) and not isinstance(intervals[0], np.ndarray):
            intervals = tuple(np.asarray(c) for c in zip(*intervals)) if len(intervals) else (np.zeros(0, np.int64),) * 3
        ser = np.ascontiguousarray(intervals[0], np.int64); tmin = np.ascontiguousarray(intervals[1], np.float64)
        tmax = np.ascontiguousarray(intervals[2], np.float64)
        fo = np.ascontiguousarray(frame_off, np.int64); tf = np.ascontiguousarray(t_first, np.float64); dd = np.ascontiguousarray(dt, np.float64)
        on_dev = 0
        keep = []

        def ptr(a):
            nonlocal on_dev
            if a is None:
                return None
            if isinstance(a, np.ndarray):
                a = np.ascontiguousarray(a, np.float32); keep.append(a)
                return C.c_void_p(a.ctypes.data)
            import torch
            a = a.contiguous()
            if a.dtype != torch.float32:
                raise TypeError("tracks must be float32")
            if a.is_cuda:
                torch.cuda.current_stream(a.device).synchronize(); on_dev = 1
            keep.append(a)
            return C.c_void_p(a.data_ptr())

        p_f0, p_t2 = ptr(f0), ptr(track2)
        m = len(ser)
        nf = np.zeros(m, np.int32); nv = np.zeros(m, np.int32); med = np.zeros(m); mean = np.zeros(m); m2 = np.zeros(m)
        rc = self._lib.pb_reduce_intervals(self._h, len(fo) - 1, _ptr(fo, C.c_int64), _ptr(tf, C.c_double), _ptr(dd, C.c_double), p_f0, p_t2, on_dev,
                                           m, _ptr(ser, C.c_int64), _ptr(tmin, C.c_double), _ptr(tmax, C.c_double), _ptr(nf, C.c_int32),
                                           _ptr(nv, C.c_int32), _ptr(med, C.c_double), _ptr(mean, C.c_double), _ptr(m2, C.c_double) if track2 is not None else None)
        N.check(self._lib, self._h, rc, "pb_reduce_intervals")
        return dict(n_frames=nf, n_voiced=nv, median_f0=med, mean_f0=mean, mean_track2=m2 if track2 is not None else None)


def pitch_frame_times(units: Units, params: N.PbPitchParams | None = None, lib=None):
    """Host-only: (t_first, dt) of the pitch frames of every unit."""
    lib = lib if lib is not None else N.load()
    params = params or pitch_params()
    n = len(units)
    t1 = np.zeros(n); dt = np.zeros(n)
    cu = units.c_struct()
    N.check(lib, None, lib.pb_pitch_frame_times(C.byref(params), C.byref(cu), _ptr(t1, C.c_double), _ptr(dt, C.c_double)), "pb_pitch_frame_times")
    return t1, dt


def intensity_plan(units: Units, minimum_pitch: float = 100.0, time_step: float = 0.0, lib=None):
    """Host-only: (status, n_frames, frame_off, t_first, dt) of Praat's Sound_to_Intensity per whole-file unit."""
    lib = lib if lib is not None else N.load()
    n = len(units)
    st = np.zeros(n, np.int32); nf = np.zeros(n, np.int32); fo = np.zeros(n + 1, np.int64); t1 = np.zeros(n); dt = np.zeros(n)
    cu = units.c_struct()
    rc = lib.pb_intensity_plan(C.byref(cu), float(minimum_pitch), float(time_step), _ptr(st, C.c_int32), _ptr(nf, C.c_int32),
                               _ptr(fo, C.c_int64), _ptr(t1, C.c_double), _ptr(dt, C.c_double))
    N.check(lib, None, rc, "pb_intensity_plan")
    return st, nf, fo, t1, dt


def pitch_plan(units: Units, params: N.PbPitchParams | None = None, lib=None):
    """Host-only: (status, n_frames, frame_off) per unit, as Praat would see them."""
    lib = lib if lib is not None else N.load()
    params = params or pitch_params()
    n = len(units)
    st = np.zeros(n, np.int32); nf = np.zeros(n, np.int32); fo = np.zeros(n + 1, np.int64)
    cu = units.c_struct()
    rc = lib.pb_pitch_plan(C.byref(params), C.byref(cu), _ptr(st, C.c_int32), _ptr(nf, C.c_int32), _ptr(fo, C.c_int64))
    N.check(lib, None, rc, "pb_pitch_plan")
    return st, nf, fo


def part_durations(units: Units, lib=None):
    """Host-only: get_part_duration / get_duration for every unit."""
    lib = lib if lib is not None else N.load()
    n = len(units)
    out = np.zeros(n); st = np.zeros(n, np.int32)
    cu = units.c_struct()
    N.check(lib, None, lib.pb_part_duration_batch(C.byref(cu), _ptr(out, C.c_double), _ptr(st, C.c_int32)), "pb_part_duration_batch")
    return out, st
