#!/bin/bash
# round 2, GPU call 17: padded sizes of the symmetric-half kernel: parity tests, memcheck over every shape
mkdir -p gpurun_out
python -m pytest tests/test_gpu_haf_sym.py -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -70 > gpurun_out/r02_pytest_sym_pad.log; tail -3 gpurun_out/r02_pytest_sym_pad.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 --error-exitcode 99 python tools/gpu_sanitize_sym.py > gpurun_out/r02_sanitizer_sym_memcheck.txt 2>&1; echo "memcheck rc=$?"
grep -E "rel err|ERROR SUMMARY|Invalid" gpurun_out/r02_sanitizer_sym_memcheck.txt | tail -16
