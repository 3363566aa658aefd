#!/bin/bash
# One full ncu capture of the sph configuration's pipeline kernels at one eighth of its named size
# (the per-GPU load of the 8-GPU run).
B="python bench.py --config sph --particles 16777216 --photons 5e8 --no-cpu-baseline --no-e2e --steps 1 --warmup 1"
timeout 400 $B > gpurun_out/r2_sph1.json 2> gpurun_out/r2_sph1.err || { tail -5 gpurun_out/r2_sph1.err; exit 1; }
python - <<PY
import json
d=json.load(open("gpurun_out/r2_sph1.json"))
print(d["ms_per_step"], d["roofline"]["stage_ms"])
PY
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_pipeline' -c 2 -f -o gpurun_out/r2_sph $B > gpurun_out/r2_sph_ncu.log 2>&1; tail -2 gpurun_out/r2_sph_ncu.log | head -c 300
