"""CPU-side check of the kernel sources themselves: the SIMT-emulated build (tests/simt_emu, test infrastructure, never
loaded by the product) against the oracle, on inputs small enough for one-thread-per-CUDA-thread emulation."""
import math

import numpy as np
import pytest

from conftest import compare_tracks, speechlike


@pytest.fixture(scope="module")
def emu(emu_lib):
    import prosody_b200 as pb
    ex = pb.Extractor(0, lib=emu_lib)
    yield ex
    ex.close()


@pytest.mark.parametrize("sr,floor,dur", [(16000, 75.0, 0.6), (16000, 150.0, 0.4), (44100, 150.0, 0.25), (44100, 75.0, 0.25), (8000, 150.0, 0.5)])
def test_emulated_pitch_kernels_match_oracle(emu, oracle, sr, floor, dur):
    import prosody_b200 as pb
    x = speechlike(2, dur, sr, seed=3)
    n = x.shape[1]
    units = pb.Units.from_list([(0, n, sr, 0.0, None), (n, n, sr, 0.05, dur - 0.03)])
    r = emu.median_pitch(x.reshape(-1), units, pb.pitch_params(floor, 600.0), frames=True)
    for i, (t0, t1) in enumerate(((0.0, None), (0.05, dur - 0.03))):
        o = oracle.pitch_track(x[i], sr, t0, t1, params=oracle.pitch_params(floor, 600.0))
        a, b = r["frame_off"][i], r["frame_off"][i + 1]
        assert b - a == o["n_frames"]
        agree, rel = compare_tracks(r["frame_f0"][a:b], o["frequency"])
        assert agree >= 1.0 - 1.0 / (b - a) and rel < 5e-3
        assert np.max(np.abs(r["frame_intensity"][a:b] - o["intensity"])) < 1e-5
        if o["median"] > 0:
            assert abs(r["median_f0"][i] - o["median"]) / o["median"] < 5e-3


def test_emulated_lufs_kernels_match_oracle(emu, oracle):
    import prosody_b200 as pb
    x = speechlike(1, 1.2, 16000, seed=4)[0]
    items = [(0, len(x), 16000, 0.0, None, 16000.0), (0, len(x), 16000, 0.1, 0.9, 44100.0), (0, len(x), 16000, 0.2, 0.4, 16000.0),
             (0, len(x), 16000, 0.7, 1.2015, 16000.0)]
    out, st = emu.lufs(x, pb.Units.from_list(items))
    for k, it in enumerate(items):
        ref = oracle.lufs(x, it[2], it[5], it[3], it[4])
        assert abs(out[k] - ref) < 1e-9
    assert st[2] & 16


def test_emulated_intensity_matches_oracle(emu, oracle):
    import prosody_b200 as pb
    a = speechlike(1, 0.5, 16000, seed=5)[0]
    b = speechlike(1, 0.3, 8000, seed=6)[0]
    short = speechlike(1, 0.05, 16000, seed=7)[0]            # shorter than the 64 ms window: Praat refuses
    pcm = np.concatenate([a, b, short])
    units = pb.Units.from_list([(0, len(a), 16000, 0.0, None), (len(a), len(b), 8000, 0.0, None), (len(a) + len(b), len(short), 16000, 0.0, None)])
    for subtract in (True, False):
        r = emu.intensity(pcm, units, subtract_mean=subtract)
        assert list(r["status"]) == [0, 0, 1]
        for i, (x, sr) in enumerate(((a, 16000), (b, 8000))):
            o = oracle.intensity(x, sr, subtract_mean=subtract)
            st, nfr, t_first, dt, _ = oracle.intensity_geometry(len(x), sr)
            lo, hi = r["frame_off"][i], r["frame_off"][i + 1]
            assert hi - lo == len(o) == nfr and abs(r["t_first"][i] - t_first) < 1e-12 and abs(r["dt"][i] - dt) < 1e-15
            assert np.max(np.abs(r["intensity_db"][lo:hi] - o)) < 2e-4


def test_emulated_legacy_measurements_match_oracle(emu, oracle):
    import prosody_b200 as pb
    from prosody_b200 import legacy
    sr = 16000
    x = speechlike(1, 0.9, sr, seed=8)[0]
    loud = (x.astype(np.int32) * 3).clip(-32768, 32767).astype(np.int16)     # squares overflow int16 almost everywhere
    pcm = np.concatenate([x, loud])
    n = len(x)
    spans = [(0.0, 0.9), (0.1033, 0.5071), (0.25, 0.25), (0.3, 0.2), (0.5, 0.9004), (0.85, 5.0)]
    units = pb.Units.from_list([(off, n, sr, s, e) for off in (0, n) for s, e in spans])
    got = legacy.loudness_segments(emu, pcm, units)
    for k, (off, (s, e)) in enumerate((off, se) for off in (0, n) for se in spans):
        ref = oracle.legacy_loudness(pcm[off:off + n], sr, s, e)
        assert (math.isnan(ref) and math.isnan(got[k])) or got[k] == ref, (k, got[k], ref)
    spans = [(0.0, 0.9), (0.1033, 0.5071), (0.3, 0.2), (0.2, 0.95), (0.41, 0.43), (-0.1, 0.5)]
    units = pb.Units.from_list([(0, n, sr, s, e) for s, e in spans])
    got = legacy.pitch_segments(emu, pcm, units)
    for k, (s, e) in enumerate(spans):
        ref = oracle.legacy_pitch_segment(x, sr, s, e)
        assert (ref == 0 and got[k] == 0) or abs(got[k] - ref) / ref < 5e-3, (k, got[k], ref)


def _gappy(sr, dur, gaps, seed, level=3000, floor=60):
    rng = np.random.default_rng(seed)
    n = int(sr * dur) + 7
    x = rng.normal(0, level, n).clip(-32768, 32767).astype(np.int16)
    for a, b in gaps:
        i, j = int(a * sr), min(int(b * sr), n)
        x[i:j] = rng.normal(0, floor, j - i).astype(np.int16)
    return x


def test_emulated_silence_split_matches_oracle(emu, oracle):
    import prosody_b200 as pb
    files = [(_gappy(8000, 2.3, ((0.0, 0.25), (0.6, 0.95), (1.3, 1.42), (1.9, 2.4)), 0), 8000),
             (_gappy(22050, 1.7, ((0.3, 0.62), (0.9, 1.25)), 1), 22050),
             (_gappy(16000, 0.15, (), 2), 16000),                              # shorter than the window: one segment
             (np.zeros(16000, np.int16), 16000),                              # all silent: no segment
             (_gappy(16000, 1.0, ((0.0, 0.4), (0.7, 1.1)), 3), 16000)]          # silent at both ends
    pcm = np.concatenate([f for f, _ in files])
    off = np.cumsum([0] + [len(f) for f, _ in files])
    units = pb.Units.from_list([(int(off[i]), len(f), sr, 0.0, None) for i, (f, sr) in enumerate(files)])
    for W, th, keep in ((200, -50, 30), (100, -45, 300), (300, -50, True), (250, -50, False)):
        r = emu.split_on_silence(pcm, units, W, th, keep)
        for i, (x, sr) in enumerate(files):
            ref = oracle.split_on_silence(x, sr, W, th, keep)
            lo, hi = r["seg_off"][i], r["seg_off"][i + 1]
            got = list(zip(r["start_ms"][lo:hi].tolist(), r["end_ms"][lo:hi].tolist()))
            assert got == ref, (W, th, keep, i, got, ref)
            for k, (s, e) in enumerate(ref):
                seg = oracle._pydub_slice_samples(x, sr, s, e)
                assert r["n_samples"][lo + k] + r["n_pad"][lo + k] == len(seg)
                a = r["first_sample"][lo + k]
                assert np.array_equal(x[a:a + r["n_samples"][lo + k]], seg[:r["n_samples"][lo + k]])


def test_emulated_empty_and_degenerate_batches(emu):
    """Every batched entry point with nothing to do, with empty files and with units it must refuse."""
    import prosody_b200 as pb
    from prosody_b200 import legacy
    none = pb.Units.from_list([])
    pcm = np.zeros(16, np.int16)
    r = emu.extract(pcm, none)
    assert len(r["median_f0"]) == 0 and len(r["lufs"]) == 0
    assert len(emu.intensity(pcm, none)["intensity_db"]) == 0
    assert len(legacy.loudness_segments(emu, pcm, none)) == 0 and len(legacy.pitch_segments(emu, pcm, none)) == 0
    s = emu.split_on_silence(pcm, none)
    assert list(s["seg_off"]) == [0] and len(s["start_ms"]) == 0
    empty_file = pb.Units.from_list([(0, 0, 16000, 0.0, None), (0, 16, 16000, 0.0, None)])
    s = emu.split_on_silence(pcm, empty_file, 1000, -50, 300)
    assert list(s["seg_off"]) == [0, 1, 2]                   # shorter than the window: one (possibly empty) segment per file
    assert list(zip(s["start_ms"], s["end_ms"])) == [(0, 0), (0, 1)]
    r = emu.extract(pcm, empty_file)
    assert list(r["status"] & 3) == [1, 1] and list(r["n_frames"]) == [0, 0]      # Praat refuses both (too short)
    sliced = pb.Units.from_list([(0, 16, 16000, 0.0, 0.001)])
    import pytest
    with pytest.raises(pb._native.NativeError):
        emu.split_on_silence(pcm, sliced)                     # whole files only
    assert list(emu.intensity(pcm, sliced)["status"]) == [64]


def _reduce_ref(frame_off, t_first, dt, f0, t2, intervals):
    """numpy restatement of pb_reduce_intervals (Praat window rule, np.median / np.mean of the voiced frames)."""
    out = []
    for s, a, b in intervals:
        v = f0[frame_off[s]:frame_off[s + 1]]; w = t2[frame_off[s]:frame_off[s + 1]]
        nf = len(v)
        k0 = int(min(max(math.ceil((a - t_first[s]) / dt[s]), 0), nf)) if nf else 0
        k1 = int(min(max(math.floor((b - t_first[s]) / dt[s]), -1), nf - 1)) if nf else -1
        if b < a or k1 < k0:
            out.append((0, 0, 0.0, 0.0, 0.0)); continue
        seg = v[k0:k1 + 1].astype(np.float64); voiced = seg[seg > 0]
        out.append((len(seg), len(voiced), float(np.median(voiced)) if len(voiced) else 0.0,
                    float(np.mean(voiced)) if len(voiced) else 0.0, float(np.mean(w[k0:k1 + 1].astype(np.float64)))))
    return out


def test_emulated_interval_reduction_matches_numpy(emu, oracle):
    import prosody_b200 as pb
    sr = 16000
    x = speechlike(3, 1.0, sr, seed=12)
    n = x.shape[1]
    units = pb.Units.from_list([(i * n, n, sr, 0.0, None) for i in range(3)] + [(0, n, sr, 0.2, 0.9)])
    p = pb.pitch_params(75.0, 600.0)
    r = emu.median_pitch(x.reshape(-1), units, p, frames=True)
    t_first, dt = pb.pitch_frame_times(units, p)
    for i, (t0, t1) in enumerate(((0.0, None), (0.0, None), (0.0, None), (0.2, 0.9))):
        g = oracle.pitch_track(x[i % 3 if i < 3 else 0], sr, t0, t1, params=oracle.pitch_params(75.0, 600.0))
        assert abs(t_first[i] - g["t1"]) < 1e-12 and abs(dt[i] - g["dt"]) < 1e-15
    rng = np.random.default_rng(2)
    ivs = [(int(rng.integers(0, 4)), float(a), float(a + rng.uniform(0.0, 0.5))) for a in rng.uniform(-0.1, 1.0, 60)]
    ivs += [(0, 0.0, 1.0), (1, 0.5, 0.4), (2, 5.0, 6.0), (3, 0.2, 0.9), (0, t_first[0], t_first[0])]
    got = emu.reduce_intervals(r["frame_off"], t_first, dt, r["frame_f0"], ivs, track2=r["frame_intensity"])
    ref = _reduce_ref(r["frame_off"], t_first, dt, r["frame_f0"], r["frame_intensity"], ivs)
    for j, (nf, nv, med, mean, m2) in enumerate(ref):
        assert got["n_frames"][j] == nf and got["n_voiced"][j] == nv, (j, ivs[j])
        assert got["median_f0"][j] == med
        assert abs(got["mean_f0"][j] - mean) <= 1e-12 * max(1.0, abs(mean)) and abs(got["mean_track2"][j] - m2) <= 1e-12
    # the whole-unit interval reproduces get_median_pitch
    whole = emu.reduce_intervals(r["frame_off"], t_first, dt, r["frame_f0"], [(i, -1.0, 99.0) for i in range(4)])
    assert np.array_equal(whole["median_f0"], r["median_f0"]) and np.array_equal(whole["n_voiced"], r["n_voiced"])


def _many_maxima(sr, dur, seed):
    """Voiced speech-like fundamental plus strong tonal high-frequency content: more autocorrelation maxima than candidate slots."""
    rng = np.random.default_rng(seed)
    t = np.arange(int(sr * dur)) / sr
    x = 0.25 * np.sin(2 * np.pi * 190.0 * t) + 0.3 * np.sin(2 * np.pi * 2500.0 * t + 0.3) + 0.2 * np.sin(2 * np.pi * 3100.0 * t) + 0.01 * rng.standard_normal(len(t))
    return (x * 20000).astype(np.int16)


@pytest.mark.parametrize("floor,ceiling", [(75.0, 600.0), (150.0, 4800.0)])
def test_emulated_overflow_and_deep_refinement_paths(emu, oracle, floor, ceiling):
    """More maxima than candidate slots (Praat's replace-the-weakest insertion), and candidates above 0.3 / dx that are refined at
    interpolation depth 700 (reach beyond the mirrored part of r): both are rare on speech, so they get their own input."""
    import prosody_b200 as pb
    sr = 16000
    x = _many_maxima(sr, 0.3, 5)
    units = pb.Units.from_list([(0, len(x), sr, 0.0, None)])
    r = emu.median_pitch(x, units, pb.pitch_params(floor, ceiling), frames=True)
    o = oracle.pitch_track(x, sr, params=oracle.pitch_params(floor, ceiling))
    assert r["n_frames"][0] == o["n_frames"]
    agree, rel = compare_tracks(r["frame_f0"], o["frequency"])
    assert agree >= 1.0 - 1.0 / o["n_frames"] and rel < 5e-3
    assert np.max(np.abs(r["frame_strength"] - o["strength"])) < 2e-3


def test_emulated_blocked_path_finder_matches_one_warp_walk(emu, monkeypatch):
    """K3 for long chains (pb_path_block_*_kernel: blocks, (max, +) transfer matrices, back-maps) against the one-warp kernel on the
    same candidate lattice: selected frequencies / strengths frame by frame, medians and voiced counts identical."""
    import prosody_b200 as pb
    sr, dur = 16000, 0.9
    x = speechlike(3, dur, sr, seed=11)
    x[1, 3000:7000] = 0                                # a stretch of single-candidate frames inside a unit
    n = x.shape[1]
    units = pb.Units.from_list([(0, n, sr, 0.0, None), (n, n, sr, 0.0, None), (2 * n, n, sr, 0.1, 0.8), (0, n, sr, 0.0, 0.3)])
    p = pb.pitch_params(75.0, 600.0)
    monkeypatch.delenv("PB_PATH_LONG", raising=False)
    ref = emu.median_pitch(x.reshape(-1), units, p, frames=True)
    for long_thresh, block in ((40, 16), (40, 7), (60, 512), (10, 2)):
        monkeypatch.setenv("PB_PATH_LONG", str(long_thresh))
        monkeypatch.setenv("PB_PATH_BLOCK", str(block))
        monkeypatch.setenv("PB_STATS_LONG", str(50 * block))       # K0 for long units too (piecewise, exact integer merges)
        r = emu.median_pitch(x.reshape(-1), units, p, frames=True)
        assert np.array_equal(r["frame_f0"], ref["frame_f0"]) and np.array_equal(r["frame_strength"], ref["frame_strength"])
        assert np.array_equal(r["median_f0"], ref["median_f0"]) and np.array_equal(r["n_voiced"], ref["n_voiced"])
    assert (ref["n_voiced"] > 0).all()


def test_emulated_long_unit_loudness_matches_chained_scan(emu, oracle, monkeypatch):
    """K4 for long units (piecewise peak, grouped state scan, CTA-wide gates: pb_lufs.cuh "long units") against the per-unit
    chain and against the oracle; the thresholds are lowered so a 1.2 s clip counts as long."""
    import prosody_b200 as pb
    x = speechlike(1, 1.7, 16000, seed=4)[0]
    y = speechlike(1, 0.9, 11025, seed=8)[0]            # 1102.5 samples per chunk: chunk lengths alternate
    pcm = np.concatenate([x, y])
    items = [(0, len(x), 16000, 0.0, None, 16000.0), (0, len(x), 16000, 0.1, 1.3, 44100.0), (0, len(x), 16000, 0.2, 0.4, 16000.0),
             (len(x), len(y), 11025, 0.0, None, 11025.0), (0, len(x), 16000, 0.7, 1.7015, 16000.0)]
    units = pb.Units.from_list(items)
    monkeypatch.delenv("PB_LUFS_LONG", raising=False)
    ref, st_ref = emu.lufs(pcm, units)
    for long_chunks, group in ((3, 2), (5, 3), (1, 1), (6, 64)):
        monkeypatch.setenv("PB_LUFS_LONG", str(long_chunks)); monkeypatch.setenv("PB_LUFS_GROUP", str(group))
        out, st = emu.lufs(pcm, units)
        assert np.array_equal(st, st_ref)
        assert np.max(np.abs(out - ref)) < 1e-10, (long_chunks, group, out, ref)
        for k, it in enumerate(items):
            src = x if it[0] == 0 else y
            assert abs(out[k] - oracle.lufs(src, it[2], it[5], it[3], it[4])) < 1e-9


@pytest.mark.parametrize("floor,nfft", [(150.0, 2048), (75.0, 4096)])
def test_emulated_split_2048_kernel_on_odd_and_edge_frames(emu, oracle, floor, nfft):
    """The split K1 (two / four 1024-point pipelines per frame pair at 2048 / 4096 points) on slices whose frame counts are odd (an unpaired last
    frame: the path where the two frames' power-of-two scales differ most) and whose first / last frames are zero-filled: every
    frame's strength and frequency against the oracle, tighter than the tolerance gates (a scale applied after the W^n rotation
    once produced 3 % strength errors on exactly these frames and still passed the 0.5 % F0 gate)."""
    import prosody_b200 as pb
    sr = 44100
    x = speechlike(1, 0.9, sr, seed=31)[0]
    n = len(x)
    items = [(0, n, sr, 0.0, 0.1), (0, n, sr, 0.05, 0.33), (0, n, sr, 0.2, 0.565), (0, n, sr, 0.0, None), (0, n, sr, 0.41, 0.9)]
    if nfft == 4096:
        items = items[1:]                              # 0.1 s is shorter than three periods of 75 Hz plus a frame
    r = emu.median_pitch(x, pb.Units.from_list(items), pb.pitch_params(floor, 600.0), frames=True)
    counts = np.diff(r["frame_off"])
    assert (counts % 2 == 1).sum() >= 1
    for i, it in enumerate(items):
        o = oracle.pitch_track(x, sr, it[3], it[4], params=oracle.pitch_params(floor, 600.0))
        a, b = r["frame_off"][i], r["frame_off"][i + 1]
        assert b - a == o["n_frames"] and o["geom"].nsampFFT == nfft
        f = r["frame_f0"][a:b]
        assert np.array_equal(f > 0, o["frequency"] > 0)
        both = f > 0
        assert np.max(np.abs(r["frame_strength"][a:b] - o["strength"])) < 5e-4
        if both.any():
            assert np.max(np.abs(f[both] - o["frequency"][both]) / o["frequency"][both]) < 1e-3


def test_emulated_four_pipeline_split_in_a_subprocess(oracle):
    """PB_ACF_SPLIT=4 (N = 4096 as four 1024-point pipelines; not the default: no faster than the general kernel on the GPU) keeps
    working: the switch is read once per process, so this case runs the emulated library in a process of its own."""
    import subprocess, sys, os
    from pathlib import Path
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r); sys.path.insert(0, %r)
import build_emu, prosody_b200 as pb
from conftest import speechlike
from oracle import oracle as O
ex = pb.Extractor(0, lib=pb._native.load(build_emu.build()))
sr = 44100
x = speechlike(1, 0.6, sr, seed=31)[0]
items = [(0, len(x), sr, 0.05, 0.33), (0, len(x), sr, 0.0, None)]
r = ex.median_pitch(x, pb.Units.from_list(items), pb.pitch_params(75.0, 600.0), frames=True)
for i, it in enumerate(items):
    o = O.pitch_track(x, sr, it[3], it[4], params=O.pitch_params(75.0, 600.0))
    a, b = r["frame_off"][i], r["frame_off"][i + 1]
    assert b - a == o["n_frames"] and o["geom"].nsampFFT == 4096
    f = r["frame_f0"][a:b]
    assert np.array_equal(f > 0, o["frequency"] > 0)
    assert np.max(np.abs(r["frame_strength"][a:b] - o["strength"])) < 5e-4
print("ok")
"""
    root = Path(__file__).resolve().parent.parent
    env = dict(os.environ, PB_ACF_SPLIT="4")
    res = subprocess.run([sys.executable, "-c", code % (str(root), str(root / "tests"), str(root / "tests" / "simt_emu"))], env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and "ok" in res.stdout, res.stderr[-2000:]
