#!/bin/bash
# round 2, GPU call 14: compute-sanitizer memcheck + racecheck + synccheck over every shape of the symmetric-half kernel
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 10 --error-exitcode 99 python tools/gpu_sanitize_sym.py > gpurun_out/r02_sanitizer_sym_$tool.txt 2>&1; echo "$tool rc=$?"
  grep -E "rel err|ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|error" gpurun_out/r02_sanitizer_sym_$tool.txt | tail -12
done
