#!/bin/bash
# compute-sanitizer memcheck over the parity tests of the newer kernels (batched, montrealer, Bristolian, loop torontonian, samplers)
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --print-limit 10 --error-exitcode 99 python -m pytest tests/test_gpu_next.py tests/test_samples.py -m gpu -x -q > gpurun_out/memcheck_next.log 2>&1; echo "memcheck rc=$?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|error" gpurun_out/memcheck_next.log | tail -8
timeout 600 compute-sanitizer --tool memcheck --print-limit 10 --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not full_size and not sampled and not n50 and not 56 and not 64" > gpurun_out/memcheck_parity.log 2>&1; echo "memcheck parity rc=$?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|error" gpurun_out/memcheck_parity.log | tail -8
