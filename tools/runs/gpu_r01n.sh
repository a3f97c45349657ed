#!/bin/bash
# K-packed tail chunk in the hafnian DMMA kernel + batched-kernel micro-optimisations: parity and timing
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
python tools/gpu_haf_sweep.py 10 18 26 34 42 50 52 2>&1 | tail -14
for w in hafnian50 lhaf50 gbs16 hsample8 mtl14; do
  timeout 900 python bench.py --workload $w --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "== $w rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1])
print('value %.4g %s ms/step %.3f e2e %.4g (%.3f ms) frac %.4f' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac']))
PY
  tail -2 gpurun_out/bench_$w.err
done
