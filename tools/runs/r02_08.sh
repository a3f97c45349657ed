#!/bin/bash
# round 2, GPU call 8: compute-sanitizer memcheck + racecheck over the kernels added in round 2
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/gpu_sanitize_r02.py > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r02_sanitizer_memcheck.txt
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 1 python tools/gpu_sanitize_r02.py > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r02_sanitizer_racecheck.txt
tail -5 gpurun_out/r02_sanitizer_memcheck.txt gpurun_out/r02_sanitizer_racecheck.txt
echo finished
