#!/bin/bash
# Final check of the round (one GPU): the whole GPU test-suite, smoke(), the full default bench
# with the DRAM traffic of this code's own ncu capture (scripts/gpu_prof.sh), then the ncu launch
# list of the short bench (gpu__time_duration only).
T=r2_end
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/${T}_tests.log; tail -4 gpurun_out/${T}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py --traffic-from profiles/r2_end_traffic.json > gpurun_out/bench_${T}.json 2> gpurun_out/bench_${T}.err; tail -c 300 gpurun_out/bench_${T}.err; head -c 700 gpurun_out/bench_${T}.json; echo
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv $B > gpurun_out/${T}_launches.log 2>&1
tail -2 gpurun_out/${T}_launches.log | head -c 300
