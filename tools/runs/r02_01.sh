#!/bin/bash
# round 2, GPU call 1: full GPU suite (incl. full-size goldens and the reference's own test files), default bench line,
# ncu --set full of the headline instantiation haf_dmma_kernel<6,1,12> and the final perm_kernel shape, launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rs -s 2>&1 | tail -150 > gpurun_out/r02_pytest_gpu_a.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_default_a.json 2> gpurun_out/r02_bench_default_a.err
ncu --set full --clock-control none --import-source on -k regex:haf_dmma_kernel -c 1 -o gpurun_out/r02_prof_haf50 -f python tools/gpu_range.py hafnian50 20 > gpurun_out/r02_ncu_haf50.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:perm_kernel -c 1 -o gpurun_out/r02_prof_perm32 -f python tools/gpu_range.py perm32 31 > gpurun_out/r02_ncu_perm32.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_hafnian50.csv python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/r02_launches_run.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_hafnian24.csv python bench.py --workload hafnian24 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches24_run.log 2>&1
echo finished
