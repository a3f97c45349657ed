#!/bin/bash
# tor v3c (operand prefetch): parity, racecheck, timing
mkdir -p gpurun_out
python tools/gpu_tor_small.py 2>&1 | tail -5
timeout 300 compute-sanitizer --tool racecheck --print-limit 5 python tools/gpu_tor_small.py > gpurun_out/racecheck_tor.log 2>&1; tail -1 gpurun_out/racecheck_tor.log
python -m pytest tests -m gpu -x -q -k "tor or threshold" > gpurun_out/pytest_gpu_tor.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_gpu_tor.log
python tools/gpu_tor_e2e.py 2>&1 | grep "host wall" | tail -6
