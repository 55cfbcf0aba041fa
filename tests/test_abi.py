"""The C-ABI library: builds with nvcc for sm_100a, loads without a GPU, exports every symbol the header declares,
fails loudly (no CPU fallback) when no device is present, and its host-only planners agree with the oracle."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def header_symbols():
    text = (ROOT / "include" / "prosody_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(native_lib):
    syms = header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(native_lib, s), s
    assert native_lib.pb_abi_version() == 2


def test_library_is_sm100a_cuda(native_lib):
    import prosody_b200 as pb
    out = subprocess.run(["cuobjdump", "-lelf", str(pb._native.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_gpu_means_loud_failure(native_lib):
    import torch
    import prosody_b200 as pb
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pb._native.NativeError):
        pb.Extractor(0)


def test_product_never_imports_the_oracle_or_emulator():
    pkg = ROOT / "prosody-control-french-tts_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")):
        src = p.read_text()
        assert "oracle" not in src.replace("pitch_floor", "") or "import oracle" not in src and "from oracle" not in src, p
        assert "simt_emu/_build" not in src and "libprosody_b200_emu" not in src, p


def _random_units(rng, n):
    rates = rng.choice([8000.0, 16000.0, 22050.0, 24000.0, 44100.0, 48000.0], n)
    nx = rng.integers(200, 200000, n)
    has = rng.integers(0, 2, n)
    t0 = np.round(rng.uniform(0, 4, n), 3)
    t1 = t0 + np.round(rng.uniform(0, 3, n), 3)
    return rates, nx, has, t0, t1


def test_pitch_plan_matches_oracle_geometry(native_lib, oracle):
    import prosody_b200 as pb
    rng = np.random.default_rng(5)
    rates, nx, has, t0, t1 = _random_units(rng, 400)
    for floor in (75.0, 150.0):
        units = pb.Units(np.zeros(400, np.int64), nx, rates, has, t0, t1)
        st, nf, fo = pb.pitch_plan(units, pb.pitch_params(floor, 600.0), native_lib)
        for i in range(400):
            ost, g, ix1, n, x1 = oracle.pitch_geometry(int(nx[i]), float(rates[i]), float(t0[i]), float(t1[i]) if has[i] else None,
                                                       oracle.pitch_params(floor, 600.0))
            assert st[i] == ost, (i, st[i], ost)
            assert nf[i] == (g.nFrames if ost == 0 else 0)
        assert fo[-1] == nf[st == 0].sum()


def test_part_duration_matches_oracle(native_lib, oracle):
    import prosody_b200 as pb
    rng = np.random.default_rng(6)
    rates, nx, has, t0, t1 = _random_units(rng, 2000)
    units = pb.Units(np.zeros(2000, np.int64), nx, rates, has, t0, t1)
    d, st = pb.part_durations(units, native_lib)
    for i in range(2000):
        try:
            ref = oracle.part_duration(int(nx[i]), int(rates[i]), float(t0[i]), float(t1[i]) if has[i] else None)
        except ValueError:
            assert st[i] == 64
            continue
        assert d[i] == ref and st[i] == 0


def test_frame_times_and_intensity_plan_match_oracle(native_lib, oracle):
    """Host-only planning of the other entry points on random units: pitch frame times (both extract_part modes) and the
    intensity frame grid, bit for bit against the oracle's float64 arithmetic."""
    import prosody_b200 as pb
    rng = np.random.default_rng(7)
    rates, nx, has, t0, t1 = _random_units(rng, 600)
    for mode in (1, 2):
        h = np.where(has != 0, mode, 0).astype(np.int32)
        units = pb.Units(np.zeros(600, np.int64), nx, rates, h, t0, t1)
        p = pb.pitch_params(75.0, 600.0)
        st, nf, _ = pb.pitch_plan(units, p, native_lib)
        tf, dt = pb.pitch_frame_times(units, p, native_lib)
        for i in range(600):
            ost, g, ix1, n, x1 = oracle.pitch_geometry(int(nx[i]), float(rates[i]), float(t0[i]), float(t1[i]) if has[i] else None,
                                                       oracle.pitch_params(75.0, 600.0), preserve_times=(mode == 1))
            assert st[i] == ost
            if ost == 0:
                assert tf[i] == g.t1 and dt[i] == g.dt and nf[i] == g.nFrames, (mode, i)
    whole = pb.Units(np.zeros(600, np.int64), nx, rates, np.zeros(600, np.int32), t0, t1)
    ist, inf, ifo, it1, idt = pb.intensity_plan(whole, 100.0, 0.0, native_lib)
    for i in range(600):
        ost, n_frames, t_first, dts, half = oracle.intensity_geometry(int(nx[i]), float(rates[i]))
        assert (ist[i] == 0) == (ost == 0)
        if ost == 0:
            assert inf[i] == n_frames and it1[i] == t_first and idt[i] == dts
