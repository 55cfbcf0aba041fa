"""Builds libprosody_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libprosody_b200.so"
SOURCES = [CSRC / "pb_api.cu"]
HEADERS = [CSRC / "pb_api_host.inc", CSRC / "pb_api_next.inc", CSRC / "pb_api_silence.inc", CSRC / "pb_textgrid.inc", CSRC / "pb_silence.cuh", CSRC / "pb_intervals.cuh", CSRC / "pb_stream.cuh", CSRC / "pb_rt.h", CSRC / "pb_plan.h", CSRC / "pb_pitch.cuh", CSRC / "pb_pitch_acf.cuh", CSRC / "pb_pitch_cand.cuh", CSRC / "pb_async.cuh", CSRC / "pb_pitch_path.cuh", CSRC / "pb_lufs.cuh",
           HERE.parent / "include" / "prosody_b200.h"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def nvcc_command(out: Path = LIB, extra: list[str] | None = None) -> list[str]:
    return [nvcc_path(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
            "-Xcompiler", "-fPIC,-ffp-contract=off,-pthread", "-shared", "-o", str(out)] + (extra or []) + [str(s) for s in SOURCES]


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    return any(p.stat().st_mtime > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, out: Path = LIB, defines: list[str] | None = None) -> Path:
    """defines: extra -D macros (development: variant libraries for A/B timing, loaded with PB_LIB / Extractor(lib=...))."""
    if force or is_stale() or out != LIB:
        cmd = nvcc_command(out=out, extra=(["-Xptxas", "-v"] if verbose else []) + [f"-D{d}" for d in (defines or [])])
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        if verbose:
            print(res.stderr)
    return out


if __name__ == "__main__":
    import sys
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[2:] for a in sys.argv[1:] if a.startswith("-o")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, out=Path(outs[0]) if outs else LIB, defines=defs))
