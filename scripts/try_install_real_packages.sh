#!/bin/bash
# Tries to install the reference's real numerics packages on the box this runs on and logs the outcome.
# The GPU boxes have no network and /opt/wheelhouse does not carry these wheels, so this is expected to fail; the log is
# the evidence either way.  If it ever succeeds, tests/golden/make_praat_golden.py can pin the oracle against the real thing.
out=${1:-gpurun_out/pip_real_packages.log}
{
  echo "== $(date -u) host=$(hostname)"
  python -c "import sys; print(sys.version)"
  for spec in "praat-parselmouth==0.4.5" "pyloudnorm" "pydub==0.25.1" "textgrid==1.6.1"; do
    echo "---- pip install $spec (index)"; timeout 60 python -m pip install --disable-pip-version-check --target baseline/_ref "$spec" 2>&1 | tail -5
    echo "---- pip install $spec (--no-index --find-links /opt/wheelhouse)"; timeout 60 python -m pip install --disable-pip-version-check --no-index --find-links /opt/wheelhouse --target baseline/_ref "$spec" 2>&1 | tail -3
  done
  echo "---- importable afterwards?"
  PYTHONPATH=baseline/_ref python - <<'PY'
for m in ("parselmouth", "pyloudnorm", "pydub", "textgrid", "spacy"):
    try:
        __import__(m); print(m, "OK")
    except Exception as e:
        print(m, "MISSING:", type(e).__name__, e)
PY
} > "$out" 2>&1
exit 0
