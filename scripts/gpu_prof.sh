#!/bin/bash
# One ncu --set full capture of the photon kernels of the short bench (after the same command has
# run plainly), and its summaries.   scripts/gpu_prof.sh TAG ["pytest selection"]
TAG=$1; SEL=$2
if [ -n "$SEL" ]; then
  timeout 600 python -m pytest $SEL -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_tests.log
fi
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_short.json 2> gpurun_out/${TAG}_short.err || exit 1
python - "$TAG" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/{sys.argv[1]}_short.json"))
print("step %.2f" % d["ms_per_step"], d["roofline"]["stage_ms"], "two-stage", d["two_stage"]["roofline"]["stage_ms"])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_pipeline|k_prep' -c 11 -f -o gpurun_out/${TAG} $B > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log | head -c 300
python scripts/summarise_profile.py --tag ${TAG} --report gpurun_out/${TAG}.ncu-rep --outdir gpurun_out --command "$B" > gpurun_out/${TAG}_traffic_stdout.json 2> gpurun_out/${TAG}_summarise.err
python scripts/prof_regions.py gpurun_out/${TAG}.ncu-rep --only k_pipeline --per 1000019607 > gpurun_out/${TAG}_regions.txt 2>&1
grep -A12 "k_pipeline<2" gpurun_out/${TAG}_regions.txt | head -30
