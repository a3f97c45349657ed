#!/bin/bash
# round 2, GPU call 18: TF = 8 shape (n = 60, 62, 64) and the re-indexed series: timing, parity tests, memcheck
mkdir -p gpurun_out
timeout 400 python tools/gpu_haf_sym.py 18 60 62 64 50 2>&1 | grep -v "20540" | tee gpurun_out/r02_haf_sym_n.txt
python -m pytest tests/test_gpu_haf_sym.py -m gpu -q -x 2>&1 | tail -3
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 --error-exitcode 99 python tools/gpu_sanitize_sym.py > gpurun_out/r02_sanitizer_sym_memcheck.txt 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid" gpurun_out/r02_sanitizer_sym_memcheck.txt | tail -4
