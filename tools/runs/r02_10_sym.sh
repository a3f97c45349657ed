#!/bin/bash
# round 2, GPU call 10: symmetric-half hafnian kernel: parity against the row-panel kernel, timing, ncu --set full
mkdir -p gpurun_out
timeout 200 python tools/gpu_haf_sym.py 20 2>&1 | tee gpurun_out/r02_haf_sym_b.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:haf_sym_kernel -c 1 -o gpurun_out/r02_prof_haf50_sym -f python tools/gpu_range.py hafnian50 18 > gpurun_out/r02_ncu_haf50_sym.log 2>&1
tail -2 gpurun_out/r02_ncu_haf50_sym.log
