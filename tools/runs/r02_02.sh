#!/bin/bash
# round 2, GPU call 2: the tensor-core batched loop-hafnian kernel (pat_dmma): parity, then gbs16 bench + ncu
mkdir -p gpurun_out
export WB200_SKIP_SLOW=1
timeout 900 python -m pytest tests/test_gpu_pat_dmma.py -x -q -m gpu 2>&1 | tail -30 > gpurun_out/r02_pytest_pat_dmma.log
timeout 1200 python -m pytest tests -q -m gpu -x --deselect tests/golden/ref_suite 2>&1 | tail -30 > gpurun_out/r02_pytest_gpu_b.log
python bench.py --workload gbs16 --steps 5 --warmup 2 > gpurun_out/r02_bench_gbs16_dmma.json 2> gpurun_out/r02_bench_gbs16_dmma.err
WB200_PAT_DFMA=1 python bench.py --workload gbs16 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_gbs16_dfma.json 2>&1
python bench.py --workload hafnian24 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_hafnian24_b.json 2>&1
python bench.py --workload hsample8 --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_hsample8_b.json 2>&1
ncu --set full --clock-control none --import-source on -k regex:pat_dmma_kernel -c 4 -o gpurun_out/r02_prof_gbs16_dmma -f python bench.py --workload gbs16 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02_ncu_gbs16.log 2>&1
echo finished
