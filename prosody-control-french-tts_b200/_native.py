"""ctypes binding of libprosody_b200.so (include/prosody_b200.h).

The library is the product: there is no Python or CPU fallback.  If it is missing, or no CUDA device can be
opened, the error is raised to the caller.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libprosody_b200.so"

PB_OK, PB_EINVAL, PB_ENODEVICE, PB_ECUDA, PB_ENOMEM, PB_EUNSUPPORTED = range(6)
PB_UNIT_OK, PB_UNIT_TOO_SHORT, PB_UNIT_NO_SAMPLES, PB_UNIT_WINDOW = 0, 1, 2, 3
PB_UNIT_PITCH_MASK = 0x0F
PB_UNIT_LUFS_FALLBACK, PB_UNIT_LUFS_ERROR, PB_UNIT_SLICE_ERROR = 16, 32, 64
_ERR_NAMES = {1: "invalid argument", 2: "no usable CUDA device", 3: "CUDA error", 4: "out of memory", 5: "unsupported"}


class NativeError(RuntimeError):
    pass


class PbPitchParams(C.Structure):
    _fields_ = [("time_step", C.c_double), ("pitch_floor", C.c_double), ("pitch_ceiling", C.c_double),
                ("periods_per_window", C.c_double), ("silence_threshold", C.c_double), ("voicing_threshold", C.c_double),
                ("octave_cost", C.c_double), ("octave_jump_cost", C.c_double), ("voiced_unvoiced_cost", C.c_double),
                ("max_candidates", C.c_int32), ("reserved", C.c_int32)]


class PbUnits(C.Structure):
    _fields_ = [("n_units", C.c_int64), ("file_off", C.POINTER(C.c_int64)), ("file_nx", C.POINTER(C.c_int64)),
                ("rate", C.POINTER(C.c_double)), ("has_t1", C.POINTER(C.c_int32)), ("t0", C.POINTER(C.c_double)),
                ("t1", C.POINTER(C.c_double)), ("meter_rate", C.POINTER(C.c_double))]


class PbTimings(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("h2d_ms", C.c_float), ("unit_stats_ms", C.c_float), ("frames_ms", C.c_float),
                ("path_ms", C.c_float), ("lufs_ms", C.c_float), ("intensity_ms", C.c_float), ("d2h_ms", C.c_float),
                ("n_frames", C.c_int64), ("n_lufs_samples", C.c_int64), ("n_launches", C.c_int32), ("host_plan_ms", C.c_float),
                ("acf_ms", C.c_float), ("cand_ms", C.c_float)]


class PbDeltaParams(C.Structure):
    _fields_ = [("pitch_semitones", C.c_double), ("pitch_lower_clip_factor", C.c_double), ("volume_pct", C.c_double),
                ("rate_percent", C.c_double), ("threshold_duration_before_slowing_down", C.c_double),
                ("slow_floor_per_sec", C.c_double)]


EXPORTS = ["pb_syntagme_deltas", "pb_ema_clamp", "pb_abi_version", "pb_create", "pb_destroy", "pb_last_error", "pb_set_stream", "pb_get_timings",
           "pb_device_info", "pb_pitch_params_default", "pb_pitch_plan", "pb_median_pitch_batch", "pb_lufs_batch",
           "pb_part_duration_batch", "pb_extract_batch", "pb_intensity_plan", "pb_intensity_batch", "pb_legacy_loudness_batch", "pb_split_on_silence_bound",
           "pb_split_on_silence_batch", "pb_segment_baselines", "pb_textgrid_parse_files",
           "pb_textgrid_sizes", "pb_textgrid_copy", "pb_textgrid_free", "pb_pitch_frame_times", "pb_reduce_intervals",
           "pb_extract_submit", "pb_extract_wait", "pb_ssml_csv", "pb_ssml_free"]


def bind(lib: C.CDLL) -> C.CDLL:
    vp, i64p, i32p, dp, fp, u8p = C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_uint8)
    U, P = C.POINTER(PbUnits), C.POINTER(PbPitchParams)
    lib.pb_abi_version.restype = C.c_int
    lib.pb_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.pb_destroy.argtypes = [vp]; lib.pb_destroy.restype = None
    lib.pb_last_error.argtypes = [vp]; lib.pb_last_error.restype = C.c_char_p
    lib.pb_set_stream.argtypes = [vp, vp]
    lib.pb_get_timings.argtypes = [vp, C.POINTER(PbTimings)]
    lib.pb_device_info.argtypes = [vp, i32p, i32p, i32p, i64p]
    lib.pb_pitch_params_default.argtypes = [P]; lib.pb_pitch_params_default.restype = None
    lib.pb_pitch_plan.argtypes = [P, U, i32p, i32p, i64p]
    lib.pb_median_pitch_batch.argtypes = [vp, vp, C.c_int64, C.c_int, U, P, dp, i32p, i32p, i32p, fp, fp, fp]
    lib.pb_lufs_batch.argtypes = [vp, vp, C.c_int64, C.c_int, U, dp, i32p]
    lib.pb_part_duration_batch.argtypes = [U, dp, i32p]
    lib.pb_extract_batch.argtypes = [vp, vp, C.c_int64, C.c_int, U, P, u8p, u8p, dp, i32p, i32p, dp, dp, i32p]
    lib.pb_extract_submit.argtypes = [vp, vp, C.c_int64, C.c_int, U, P, u8p, u8p, dp, i32p, i32p, dp, dp, i32p]
    lib.pb_extract_wait.argtypes = [vp]
    lib.pb_ssml_csv.argtypes = [C.c_int64, C.c_char_p, i64p, C.c_char_p, i64p, i32p, dp, dp, dp, C.c_double, C.c_char_p, C.c_int,
                                C.POINTER(vp), i64p, C.POINTER(vp), i64p, C.POINTER(vp), i64p]
    lib.pb_ssml_free.argtypes = [vp]; lib.pb_ssml_free.restype = None
    lib.pb_intensity_plan.argtypes = [U, C.c_double, C.c_double, i32p, i32p, i64p, dp, dp]
    lib.pb_intensity_batch.argtypes = [vp, vp, C.c_int64, C.c_int, U, C.c_double, C.c_double, C.c_int, fp, i32p]
    lib.pb_legacy_loudness_batch.argtypes = [vp, vp, C.c_int64, C.c_int, U, dp, i64p, i64p]
    lib.pb_split_on_silence_bound.argtypes = [U, C.c_int]; lib.pb_split_on_silence_bound.restype = C.c_int64
    lib.pb_split_on_silence_batch.argtypes = [vp, vp, C.c_int64, C.c_int, U, C.c_int, C.c_double, C.c_int, C.c_int64, i64p, i32p, i32p, i64p, i32p, i32p]
    lib.pb_syntagme_deltas.argtypes = [C.c_int64, dp, dp, dp, dp, i32p, dp, dp, i32p, C.POINTER(PbDeltaParams), dp, dp, dp]
    lib.pb_segment_baselines.argtypes = [C.c_int64, dp, dp, dp, C.c_int32, dp, dp, dp]
    lib.pb_textgrid_parse_files.argtypes = [C.POINTER(C.c_char_p), C.c_int64, C.c_int, C.c_int, C.POINTER(vp)]
    lib.pb_textgrid_sizes.argtypes = [vp, i64p, i64p, i64p]
    lib.pb_textgrid_copy.argtypes = [vp, i32p, dp, dp, i64p, dp, dp, i64p, C.c_char_p]
    lib.pb_textgrid_free.argtypes = [vp]; lib.pb_textgrid_free.restype = None
    lib.pb_pitch_frame_times.argtypes = [P, U, dp, dp]
    lib.pb_reduce_intervals.argtypes = [vp, C.c_int64, i64p, dp, dp, vp, vp, C.c_int, C.c_int64, i64p, dp, dp, i32p, i32p, dp, dp, dp]
    lib.pb_ema_clamp.argtypes = [dp, C.c_int64, C.c_double, C.c_double, dp]
    for name in EXPORTS:
        if name not in ("pb_destroy", "pb_last_error", "pb_pitch_params_default", "pb_split_on_silence_bound", "pb_textgrid_free"):
            getattr(lib, name).restype = C.c_int
    return lib


_lib = None


def load(path: str | Path | None = None) -> C.CDLL:
    """Load the CUDA library. `path` is for tests that bind another build of the same ABI."""
    global _lib
    if path is not None:
        return bind(C.CDLL(str(path)))
    if _lib is None:
        if not LIB_PATH.exists():
            raise NativeError(f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                              "(nvcc, sm_100a). There is no CPU fallback.")
        _lib = bind(C.CDLL(str(LIB_PATH)))
        if _lib.pb_abi_version() != 2:
            raise NativeError("libprosody_b200.so has an unexpected ABI version")
    return _lib


def check(lib: C.CDLL, handle, rc: int, what: str) -> None:
    if rc != PB_OK:
        msg = lib.pb_last_error(handle).decode() if handle else ""
        raise NativeError(f"{what}: {_ERR_NAMES.get(rc, rc)}{': ' + msg if msg else ''}")
