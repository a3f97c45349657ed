#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:pat_main -c 1 -o gpurun_out/prof_gbs16_v2 -f \
    python bench.py --workload gbs16 --batch 20000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_gbs.log 2>&1; tail -1 gpurun_out/ncu_gbs.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:haf_dmma -c 1 -o gpurun_out/prof_haf40 -f \
    python tools/gpu_one_hafnian.py hafnian40 > gpurun_out/ncu_haf.log 2>&1; tail -1 gpurun_out/ncu_haf.log | cut -c1-200
