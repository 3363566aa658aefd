#!/bin/bash
# Final check A (one GPU): the whole GPU test-suite, smoke(), a plain short bench, then the ncu
# launch list of that same command (gpu__time_duration only).
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_final_tests.log; tail -4 gpurun_out/r2_final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; tail -2 gpurun_out/r2_final_smoke.log
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 300 $B > gpurun_out/r2_final_short.json 2> gpurun_out/r2_final_short.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_final_launches.csv $B > gpurun_out/r2_final_launches.log 2>&1
tail -2 gpurun_out/r2_final_launches.log | head -c 300
