#!/bin/bash
# adaptive warps per CTA for small hafnian problems: parity + timing
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "haf or loop or repeated or matching" > gpurun_out/pytest_gpu_haf.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_gpu_haf.log
python tools/gpu_haf_sweep.py 18 22 24 26 28 30 32 2>&1 | tail -14
WB200_HAF_WARPS=12 python tools/gpu_haf_sweep.py 24 28 2>&1 | grep haf | sed 's/^/W12 /'
python bench.py --workload hafnian24 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_hafnian24.json 2> gpurun_out/bench_hafnian24.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_hafnian24.json').read().strip().splitlines()[-1])
print('hafnian24 value %.4g %s ms/step %.4f e2e %.4g (%.3f ms) frac %.4f' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac']))
PY
