import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "simt_emu"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def native_lib():
    """The nvcc-built product library (loads without a GPU; compute calls need one)."""
    import prosody_b200 as pb
    from prosody_b200 import build as B
    B.build()
    return pb._native.load()


@pytest.fixture(scope="session")
def emu_lib():
    """TEST-ONLY: the same kernel sources compiled against the SIMT emulator."""
    import build_emu
    import prosody_b200 as pb
    return pb._native.load(build_emu.build())


@pytest.fixture(scope="session")
def gpu_extractor():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import prosody_b200 as pb
    ex = pb.Extractor(0)
    yield ex
    ex.close()


def speechlike(n_utt, dur_s, sr, seed):
    """Small synthetic corpus as a numpy int16 [n_utt, n] array (torch CPU generator)."""
    from prosody_b200 import synth
    return synth.make_corpus(n_utt, dur_s, sr, seed=seed, device="cpu").numpy()


def compare_tracks(f_gpu, f_ref):
    """-> (voicing agreement, max relative F0 error on frames voiced in both)."""
    f_gpu = np.asarray(f_gpu, float); f_ref = np.asarray(f_ref, float)
    both = (f_gpu > 0) & (f_ref > 0)
    agree = float(np.mean((f_gpu > 0) == (f_ref > 0))) if len(f_ref) else 1.0
    rel = float(np.max(np.abs(f_gpu[both] - f_ref[both]) / f_ref[both])) if both.any() else 0.0
    return agree, rel
