#!/bin/bash
# The GPU test-suite and the sanitizer smoke script against libpxb_dbg.so (kernel index / capacity
# assertions compiled in: python -m pyxsim_b200.build --debug-checks).  $1: optional pytest selection.
export PXB_LIBRARY=$PWD/pyxsim_b200/libpxb_dbg.so
python - <<PY
from pyxsim_b200 import _lib
print("library:", _lib.LIB_PATH)
PY
if [ -n "$1" ]; then
  timeout 600 python -m pytest "$1" -x -q > gpurun_out/r2_dbg_one.log 2>&1; grep -n "Assertion\|assert" gpurun_out/r2_dbg_one.log | head -5; tail -3 gpurun_out/r2_dbg_one.log
  exit 0
fi
timeout 400 python scripts/sanitize_smoke.py > gpurun_out/r2_dbg_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r2_dbg_smoke.log
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_dbg_tests.log; tail -4 gpurun_out/r2_dbg_tests.log
