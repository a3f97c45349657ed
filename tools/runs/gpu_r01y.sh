#!/bin/bash
mkdir -p gpurun_out
for w in gbs16 hsample8; do
  timeout 100 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "== $w rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1])
print('value %.4g %s ms/step %.3f e2e %.4g (%.3f ms) frac %.4f cpu %.4g (%d cores)' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['cores']))
PY
done
