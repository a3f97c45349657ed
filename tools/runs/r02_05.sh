#!/bin/bash
# round 2, GPU call 5: _dev entries, device chain sampler, shard-independent perm sums, tightened tolerances
mkdir -p gpurun_out
export WB200_SKIP_SLOW=1
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/golden/ref_suite -s 2>&1 | grep -v "^\[fullsize\]\|^$" | tail -40 > gpurun_out/r02_pytest_gpu_d.log
python bench.py --workload hsample8 --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_hsample8_c.json 2>&1
python bench.py --workload perm32 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_perm32_c.json 2>&1
python bench.py --workload perm40 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_perm40_c.json 2>&1
python bench.py --workload hafnian24 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_hafnian24_d.json 2>&1
echo finished
