#!/bin/bash
# Round-1 GPU pass: parity suite, bench lines for every workload, ncu launch lists and full captures.
set -x
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for w in hafnian50 perm32 tor48 gbs16 hafnian24; do
  timeout 600 python bench.py --workload $w --steps 2 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
done
# launch lists (shares of the step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_hafnian40.csv python bench.py --workload hafnian40 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_perm28.csv python bench.py --workload perm28 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l2.log 2>&1
# full captures of the two top kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:haf_dmma -s 3 -c 1 -f -o gpurun_out/prof_haf40 python bench.py --workload hafnian40 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:perm_kernel -s 3 -c 1 -f -o gpurun_out/prof_perm28 python bench.py --workload perm28 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f2.log 2>&1
ls -la gpurun_out
