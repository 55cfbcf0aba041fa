"""Several processes on ONE GPU, as the reference runs under `multiprocessing: True` (config.yaml:57-58; one spawn-pool worker per
voice, Code/audioPipeline.py:1143-1150): every process creates its own handle on device 0 and measures concurrently; results must
equal a lone run, and a handle's footprint must stay small enough for several to coexist."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent

WORKER = r"""
import sys, json
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np, torch
import prosody_b200 as pb
from conftest import speechlike
seed = int(sys.argv[1])
x = speechlike(24, 2.0, 16000, seed=seed)
n = x.shape[1]
items = []
for i in range(x.shape[0]):
    items += [(i * n, n, 16000, 0.0, None, 16000.0), (i * n, n, 16000, 0.25, 1.75, 16000.0)]
units = pb.Units.from_list(items)
free0, total = torch.cuda.mem_get_info(0)
with pb.Extractor(0) as ex:
    outs = [ex.extract(x.reshape(-1), units, pb.pitch_params(75.0, 600.0)) for _ in range(6)]
    free1, _ = torch.cuda.mem_get_info(0)
for o in outs[1:]:
    assert np.array_equal(o["median_f0"], outs[0]["median_f0"]) and np.array_equal(o["lufs"], outs[0]["lufs"], equal_nan=True)
print(json.dumps(dict(median=outs[0]["median_f0"].tolist(), lufs=outs[0]["lufs"].tolist())))
"""


@pytest.mark.gpu
def test_two_processes_share_one_gpu(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % (str(ROOT), str(ROOT / "tests")))
    env = dict(os.environ, PB_RACF_BYTES=str(256 << 20))            # a worker sharing the device keeps its scratch small
    lone = subprocess.run([sys.executable, str(script), "11"], capture_output=True, text=True, env=env, timeout=600)
    assert lone.returncode == 0, lone.stderr[-2000:]
    procs = [subprocess.Popen([sys.executable, str(script), str(s)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env) for s in (11, 11, 12)]
    outs = [p.communicate(timeout=900) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-2000:]
    import json
    want = json.loads(lone.stdout.strip().splitlines()[-1])
    for so, _ in outs[:2]:                                          # the two workers on the lone run's data: identical results
        got = json.loads(so.strip().splitlines()[-1])
        assert got["median"] == want["median"] and np.array_equal(np.array(got["lufs"]), np.array(want["lufs"]), equal_nan=True)
