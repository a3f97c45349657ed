#!/bin/bash
# round 2, GPU call 15: racecheck of the symmetric-half kernel with the split-phase barrier replaced by an ordinary team
# barrier at the point of the wait (build with WB200_NVCC_EXTRA=-DWB_HS_PLAINBAR): the tool does not model mbarrier ordering
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 --error-exitcode 99 python tools/gpu_sanitize_sym.py > gpurun_out/r02_sanitizer_sym_racecheck_plainbar.txt 2>&1; echo "racecheck plainbar rc=$?"
grep -E "rel err|RACECHECK SUMMARY|Race reported" gpurun_out/r02_sanitizer_sym_racecheck_plainbar.txt | tail -14

