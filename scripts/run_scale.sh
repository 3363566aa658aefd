#!/bin/bash
# Run bench.py for the named BASELINE configs on N GPUs of this box and collect the JSON lines.
#   scripts/run_scale.sh N "cluster amr sph plline" [extra bench args]
# (N = 1 runs plain python, N > 1 the same torchrun line the driver uses)
N=$1; CONFIGS=${2:-cluster}; shift 2
OUT=gpurun_out/r2_scale_n${N}.jsonl
: > $OUT
for c in $CONFIGS; do
  if [ "$N" = "1" ]; then
    LAUNCH="python"
  else
    LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
  fi
  echo "== $c on $N GPUs" >&2
  timeout 900 $LAUNCH bench.py --gpus $N --config $c --steps 3 --warmup 2 --no-cpu-baseline "$@" \
      >> $OUT 2> gpurun_out/r2_scale_n${N}_${c}.err || echo "{\"config\": \"$c\", \"n_gpus\": $N, \"failed\": true}" >> $OUT
  tail -c 300 gpurun_out/r2_scale_n${N}_${c}.err >&2
done
cat $OUT | cut -c1-400
