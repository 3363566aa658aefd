#!/bin/bash
# Development check after the sampler change: thermal / e2e / multirank / bench-shape tests, amr
# bench at the per-GPU load of the 2-GPU named run, launch list.
timeout 900 python -m pytest tests/test_gpu_thermal.py tests/test_gpu_e2e.py tests/test_gpu_multirank.py tests/test_gpu_benchshape.py -q 2>&1 | tail -40 > gpurun_out/r2_amr_tests.log; tail -5 gpurun_out/r2_amr_tests.log
B="python bench.py --config amr --grid 256 --photons 2e9 --no-cpu-baseline --no-e2e --steps 2 --warmup 1"
timeout 400 $B > gpurun_out/r2_amr4.json 2> gpurun_out/r2_amr4.err; tail -c 200 gpurun_out/r2_amr4.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_amr4.json"))
print(d["ms_per_step"], d["roofline"]["stage_ms"])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_amr4_launches.csv $B > gpurun_out/r2_amr4_ncu.log 2>&1; tail -2 gpurun_out/r2_amr4_ncu.log | head -c 300
