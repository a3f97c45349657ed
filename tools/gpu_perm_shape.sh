#!/bin/bash
mkdir -p gpurun_out
L="thewalrus_b200/libwalrus_b200.so tools/variants/libwb_perm_t256.so tools/variants/libwb_perm_t64.so tools/variants/libwb_perm_t32.so"
python tools/gpu_perm_shape.py $L 2>&1 | tee gpurun_out/perm_shape.txt
for l in $L; do
  ncu --metrics sm__warps_active.avg.per_cycle_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,gpu__time_duration.sum --clock-control none -k regex:perm_kernel -c 1 python tools/gpu_perm_shape.py $l 2>&1 | grep -E "perm_kernel|warps_active|pipe_fp64|time_duration" | tee -a gpurun_out/perm_shape.txt
done
