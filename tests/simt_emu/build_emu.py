"""TEST INFRASTRUCTURE ONLY: compiles the kernel sources with g++ against the SIMT emulator (pb_simt_emu.h) so the
exact device code can be exercised on a machine without a GPU.  Output goes to tests/simt_emu/_build/ and is loaded
only by tests; the product never looks there."""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "prosody-control-french-tts_b200" / "csrc"
OUT = HERE / "_build" / "libprosody_b200_emu.so"


def build(force: bool = False) -> Path:
    srcs = [CSRC / "pb_api.cu", HERE / "pb_simt_emu.cpp"]
    deps = srcs + list(CSRC.glob("*.h")) + list(CSRC.glob("*.cuh")) + [HERE / "pb_simt_emu.h", ROOT / "include" / "prosody_b200.h"]
    if force or not OUT.exists() or any(d.stat().st_mtime > OUT.stat().st_mtime for d in deps):
        OUT.parent.mkdir(exist_ok=True)
        gxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
        cmd = [gxx, "-O2", "-std=c++17", "-DPB_SIMT_EMU", "-ffp-contract=off", "-x", "c++", str(srcs[0]), "-x", "c++", str(srcs[1]),
               f"-I{HERE}", "-shared", "-fPIC", "-pthread", "-o", str(OUT)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("emulator build failed:\n" + res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
