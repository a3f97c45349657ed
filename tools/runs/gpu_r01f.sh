#!/bin/bash
# tor kernel v3 check: parity, sanitizer, timing
mkdir -p gpurun_out
python tools/gpu_tor_small.py 2>&1 | tail -5
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python tools/gpu_tor_small.py > gpurun_out/racecheck_tor.log 2>&1; tail -2 gpurun_out/racecheck_tor.log
python -m pytest tests -m gpu -x -q -k "tor or threshold" > gpurun_out/pytest_gpu_tor.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_tor.log
python tools/gpu_tor_e2e.py 2>&1 | grep "host wall\|tor call\|ltor(O" | tail -12
ncu --set full --clock-control none --import-source on -k regex:tor_kernel -c 1 -o gpurun_out/prof_tor48_v3 -f \
    python tools/gpu_one_hafnian.py tor48 > gpurun_out/ncu_tor.log 2>&1; tail -1 gpurun_out/ncu_tor.log
