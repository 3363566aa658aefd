#!/bin/bash
# After the SPH-projection change: projection / e2e / event tests, then the sph configuration at
# one eighth of its named size (the per-GPU load of the 8-GPU run).
timeout 900 python -m pytest tests/test_gpu_project.py tests/test_gpu_e2e.py tests/test_gpu_events.py tests/test_gpu_seams.py -q 2>&1 | tail -30 > gpurun_out/r2_sph_tests.log; tail -4 gpurun_out/r2_sph_tests.log
B="python bench.py --config sph --particles 16777216 --photons 5e8 --no-cpu-baseline --no-e2e --steps 3 --warmup 2"
timeout 400 $B > gpurun_out/r2_sph2.json 2> gpurun_out/r2_sph2.err || { tail -5 gpurun_out/r2_sph2.err; exit 1; }
python - <<PY
import json
d=json.load(open("gpurun_out/r2_sph2.json"))
print(d["ms_per_step"], d["roofline"]["stage_ms"])
PY
