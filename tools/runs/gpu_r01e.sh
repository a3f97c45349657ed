#!/bin/bash
# round 1, re-entry call 4: torontonian kernel v2 — sanitizer, parity, timing, ncu
mkdir -p gpurun_out
python tools/gpu_tor_small.py 2>&1 | tail -8
timeout 600 compute-sanitizer --tool racecheck --print-limit 5 python tools/gpu_tor_small.py > gpurun_out/racecheck_tor.log 2>&1; tail -6 gpurun_out/racecheck_tor.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/gpu_tor_small.py > gpurun_out/memcheck_tor.log 2>&1; tail -4 gpurun_out/memcheck_tor.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
python tools/gpu_tor_e2e.py 2>&1 | grep -v "^ \|^$\|ncalls\|Ordered\|List reduced\|function calls" | tail -12
ncu --set full --clock-control none --import-source on -k regex:tor_kernel -c 1 -o gpurun_out/prof_tor48_v2 -f \
    python tools/gpu_one_hafnian.py tor48 > gpurun_out/ncu_tor.log 2>&1; tail -2 gpurun_out/ncu_tor.log
for w in tor48 ltor48 tor60; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "== $w rc=$?"; head -c 400 gpurun_out/bench_$w.json; echo; tail -3 gpurun_out/bench_$w.err
done
