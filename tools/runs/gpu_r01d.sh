#!/bin/bash
# round 1, re-entry call 3: ncu --set full of the torontonian and patterns kernels, launch lists, bench re-runs
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:tor_kernel -c 1 -o gpurun_out/prof_tor48 -f \
    python tools/gpu_one_hafnian.py tor48 > gpurun_out/ncu_tor.log 2>&1; tail -2 gpurun_out/ncu_tor.log
ncu --set full --clock-control none --import-source on -k regex:pat_main -c 1 -o gpurun_out/prof_gbs16 -f \
    python bench.py --workload gbs16 --batch 20000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_gbs.log 2>&1; tail -2 gpurun_out/ncu_gbs.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_hafnian50.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/launches_run.log 2>&1; tail -2 gpurun_out/launches_run.log
for w in hsample8 brs12 mtl14; do
  timeout 600 python bench.py --workload $w --steps 2 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "== $w rc=$?"; tail -c 600 gpurun_out/bench_$w.json; tail -3 gpurun_out/bench_$w.err
done
