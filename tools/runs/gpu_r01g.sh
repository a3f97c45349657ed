#!/bin/bash
# tor v3b + 2x2-tiled batched kernels: parity, timing; perm32 ncu capture with source
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
python tools/gpu_tor_e2e.py 2>&1 | grep "host wall\|ltor(O" | tail -7
for w in gbs16 tor48 ltor48 mtl14 hsample8; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "== $w rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1])
print('value %.4g %s ms/step %.3f e2e %.4g (%.3f ms) frac %.4f cpu %.4g' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value']))
PY
  tail -2 gpurun_out/bench_$w.err
done
ncu --set full --clock-control none --import-source on -k regex:perm_kernel -c 1 -o gpurun_out/prof_perm32 -f \
    python tools/gpu_one_hafnian.py perm32 > gpurun_out/ncu_perm.log 2>&1; tail -1 gpurun_out/ncu_perm.log
ncu --set full --clock-control none --import-source on -k regex:tor_kernel -c 1 -o gpurun_out/prof_tor48_v3b -f \
    python tools/gpu_one_hafnian.py tor48 > gpurun_out/ncu_tor.log 2>&1; tail -1 gpurun_out/ncu_tor.log
