#!/bin/bash
# A/B of library variants on one GPU: a pytest selection with the default library, then the short
# bench once per library (libpxb.so and every pyxsim_b200/libpxb_<variant>.so named as argument).
#   scripts/gpu_ab.sh TAG "pytest selection" variant...
TAG=$1; SEL=$2; shift 2
if [ -n "$SEL" ]; then
  timeout 600 python -m pytest $SEL -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_tests.log
fi
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 200 $B > gpurun_out/${TAG}_main.json 2> gpurun_out/${TAG}_main.err
for v in "$@"; do
  PXB_LIBRARY=$PWD/pyxsim_b200/libpxb_$v.so timeout 200 $B > gpurun_out/${TAG}_$v.json 2> gpurun_out/${TAG}_$v.err
done
python - "$TAG" main "$@" <<'PY'
import json, sys
tag = sys.argv[1]
for v in sys.argv[2:]:
    try:
        d = json.load(open(f"gpurun_out/{tag}_{v}.json"))
        print(v, "step %.2f" % d["ms_per_step"], d["roofline"]["stage_ms"], "two-stage", d["two_stage"]["roofline"]["stage_ms"],
              d["fused_same_counts_as_two_stage"], d["clocks"]["sm_mhz"])
    except Exception as e:
        print(v, "failed", e, open(f"gpurun_out/{tag}_{v}.err").read()[-400:])
PY
