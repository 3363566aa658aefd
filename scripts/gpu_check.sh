#!/bin/bash
# One-GPU check used during development: full GPU test-suite, small-size runs of the named bench
# configs (1 rank and 2 ranks sharing the GPU through gloo), smoke().
timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_t12.log; tail -4 gpurun_out/r2_t12.log
B="timeout 300 python bench.py --no-cpu-baseline"
$B --config amr --grid 64 --photons 3e7 --steps 2 --warmup 1 > gpurun_out/r2_c_amr.json 2> gpurun_out/r2_c_amr.err
$B --config plline --grid 96 --photons 3e7 --steps 2 --warmup 1 > gpurun_out/r2_c_pl.json 2> gpurun_out/r2_c_pl.err
T="timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --no-cpu-baseline --steps 2 --warmup 1"
PXB_DIST_BACKEND=gloo $T --config amr --grid 64 --photons 3e7 > gpurun_out/r2_c2_amr.json 2> gpurun_out/r2_c2_amr.err
PXB_DIST_BACKEND=gloo $T --config plline --grid 96 --photons 3e7 > gpurun_out/r2_c2_pl.json 2> gpurun_out/r2_c2_pl.err
for f in gpurun_out/r2_c_amr gpurun_out/r2_c_pl gpurun_out/r2_c2_amr gpurun_out/r2_c2_pl; do echo "== $f"; tail -c 300 $f.err; head -c 300 $f.json; echo; done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_b12.json 2> gpurun_out/r2_b12.err; tail -c 300 gpurun_out/r2_b12.err; head -c 400 gpurun_out/r2_b12.json
