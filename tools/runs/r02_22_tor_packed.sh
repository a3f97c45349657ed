#!/bin/bash
# round 2, GPU call 22: torontonian kernel on packed lower triangles (two CTAs per SM): parity, timing
mkdir -p gpurun_out
WB200_SKIP_SLOW=1 python -m pytest tests/test_gpu_tor_real.py tests/golden/ref_suite/test_torontonian.py tests/test_gpu_parity.py tests/test_gpu_next.py tests/test_gpu_fullsize.py tests/test_gpu_pat_dmma.py -m gpu -q -x -k "tor or ltor or threshold" 2>&1 | tail -4
python tools/gpu_tor_real.py 2>&1 | tee gpurun_out/r02_tor_packed.txt
for w in tor48 ltor48; do python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith('{')][-1])
print('$w', 'ms %.4g e2e ms %.4g err %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], d.get('result_rel_err')))" | tee -a gpurun_out/r02_tor_packed.txt; done
