#!/bin/bash
# compute-sanitizer on a small invocation of every pipeline mode ($1 = memcheck | racecheck | synccheck | initcheck):
# the script runs plain first, then under the tool.
timeout 300 python scripts/sanitize_smoke.py > gpurun_out/r2_sanitize_plain.log 2>&1 || { tail -5 gpurun_out/r2_sanitize_plain.log; exit 1; }
tail -2 gpurun_out/r2_sanitize_plain.log
timeout 1500 compute-sanitizer --tool $1 --error-exitcode 7 python scripts/sanitize_smoke.py > gpurun_out/r2_sanitize_$1.log 2>&1; echo "exit $?" >> gpurun_out/r2_sanitize_$1.log
tail -6 gpurun_out/r2_sanitize_$1.log
