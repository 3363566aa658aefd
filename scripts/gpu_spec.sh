#!/bin/bash
# Development check of the spectrum mode + a kernel-level look at the amr (var_elem) configuration.
timeout 600 python -m pytest tests/test_gpu_spectrum.py -q 2>&1 | tail -15 > gpurun_out/r2_spec_tests.log; tail -5 gpurun_out/r2_spec_tests.log
timeout 300 python scripts/bench_spectrum.py > gpurun_out/r2_spectrum.jsonl 2> gpurun_out/r2_spectrum.err; tail -c 600 gpurun_out/r2_spectrum.err; cat gpurun_out/r2_spectrum.jsonl
B="python bench.py --config amr --grid 256 --photons 2e9 --no-cpu-baseline --no-e2e --steps 2 --warmup 1"
timeout 400 $B > gpurun_out/r2_amr1.json 2> gpurun_out/r2_amr1.err; tail -c 300 gpurun_out/r2_amr1.err; head -c 1500 gpurun_out/r2_amr1.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_amr1_launches.csv $B > gpurun_out/r2_amr1_ncu.log 2>&1; tail -2 gpurun_out/r2_amr1_ncu.log | head -c 300
