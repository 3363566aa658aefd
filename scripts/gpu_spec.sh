#!/bin/bash
# Development check of the spectrum mode + one full ncu capture of the amr (var_elem) sampler.
timeout 600 python -m pytest tests/test_gpu_spectrum.py tests/test_gpu_e2e.py -q 2>&1 | tail -15 > gpurun_out/r2_spec_tests.log; tail -5 gpurun_out/r2_spec_tests.log
timeout 300 python scripts/bench_spectrum.py > gpurun_out/r2_spectrum.jsonl 2> gpurun_out/r2_spectrum.err; tail -c 600 gpurun_out/r2_spectrum.err; cat gpurun_out/r2_spectrum.jsonl
B="python bench.py --config amr --grid 192 --photons 8e8 --no-cpu-baseline --no-e2e --no-fused --steps 1 --warmup 1"
timeout 400 $B > gpurun_out/r2_amr2.json 2> gpurun_out/r2_amr2.err; tail -c 300 gpurun_out/r2_amr2.err; head -c 300 gpurun_out/r2_amr2.json; echo
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_pipeline|k_prep_split' -c 3 -f -o gpurun_out/r2_amr_sampler $B > gpurun_out/r2_amr2_ncu.log 2>&1; tail -3 gpurun_out/r2_amr2_ncu.log | head -c 400
