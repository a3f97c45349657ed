#!/bin/bash
# round 2, GPU call 3: pat_dmma all classes, hafnian panel-split shapes + fused prep, lean host path
mkdir -p gpurun_out
export WB200_SKIP_SLOW=1
timeout 900 python -m pytest tests/test_gpu_pat_dmma.py -q -m gpu 2>&1 | tail -30 > gpurun_out/r02_pytest_pat_dmma_b.log
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -30 > gpurun_out/r02_pytest_gpu_c.log
python tools/gpu_haf_small.py > gpurun_out/r02_haf_small_sweep.txt 2>&1
python bench.py --workload hafnian24 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_hafnian24_c.json 2>&1
python bench.py --workload tor48 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_tor48_c.json 2>&1
python bench.py --workload hafnian50 --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r02_bench_hafnian50_c.json 2>&1
echo finished
