#!/bin/bash
# round 2, GPU call 21: real-symmetric torontonian: parity tests, the reference's own test_torontonian.py, timing, memcheck
mkdir -p gpurun_out
python -m pytest tests/test_gpu_tor_real.py tests/golden/ref_suite/test_torontonian.py tests/test_gpu_parity.py -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -25 > gpurun_out/r02_pytest_tor_real.log; tail -3 gpurun_out/r02_pytest_tor_real.log
python tools/gpu_tor_real.py 2>&1 | tee gpurun_out/r02_tor_real.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 10 --error-exitcode 99 python -m pytest tests/test_gpu_tor_real.py -m gpu -q -x -k "not 20" > gpurun_out/r02_sanitizer_tor_real.txt 2>&1; echo "memcheck rc=$?"
grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r02_sanitizer_tor_real.txt | tail -3
