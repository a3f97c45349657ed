#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 120 python bench.py --workload tor48 --steps 3 --warmup 3 > gpurun_out/bench_tor48.json 2> gpurun_out/bench_tor48.err; tail -c 300 gpurun_out/bench_tor48.json
