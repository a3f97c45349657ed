#!/bin/bash
# round 1, re-entry call 2: sampler GPU tests + bench lines of the SURVEY 8(f) components
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for w in ltor48 mtl14 brs12 hsample8 gbs16; do
  timeout 600 python bench.py --workload $w --steps 2 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "== $w rc=$?"; tail -c 1500 gpurun_out/bench_$w.json; tail -3 gpurun_out/bench_$w.err
done
