#!/bin/bash
# Final check B (one GPU): plain short bench, ONE ncu --set full capture of the pipeline kernels
# of that command, its summaries, then the full default bench with the measured traffic, and the
# reference arm.
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 300 $B > gpurun_out/r2_final_short2.json 2> gpurun_out/r2_final_short2.err || exit 1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_pipeline|k_prep' -c 11 -f -o gpurun_out/r2_final $B > gpurun_out/r2_final_ncu.log 2>&1
tail -2 gpurun_out/r2_final_ncu.log | head -c 300
python scripts/summarise_profile.py --tag r2_final --report gpurun_out/r2_final.ncu-rep --outdir gpurun_out --command "$B" > gpurun_out/r2_final_traffic_stdout.json 2> gpurun_out/r2_final_summarise.err
python scripts/prof_regions.py gpurun_out/r2_final.ncu-rep > gpurun_out/r2_pipeline_regions.txt 2>&1
timeout 900 python bench.py --traffic-from gpurun_out/r2_final_traffic.json > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; tail -c 300 gpurun_out/bench_r2_final.err; head -c 600 gpurun_out/bench_r2_final.json; echo
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err; head -c 600 gpurun_out/bench_r2_reference.json; echo
