#!/bin/bash
# round 2, 8-GPU call 3: north-star sizes with the symmetric-half kernel (hafnian 56 complete results against the long-double goldens, bench line)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR tools/gpu_fullsize_x8.py > gpurun_out/r02_fullsize_sym_x$N.log 2>&1; grep "^\[x\|FULLSIZE" gpurun_out/r02_fullsize_sym_x$N.log | cut -c1-220
timeout 300 $TR bench.py --gpus $N --workload hafnian56 --steps 2 --warmup 1 > gpurun_out/r02_bench_hafnian56_sym_x$N.json 2> gpurun_out/r02_bench_hafnian56_sym_x$N.err; echo "bench hafnian56 x$N rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_hafnian56_sym_x$N.json').read().strip().splitlines() if l.startswith('{')][-1])
print('value %.4g %s ms/step %.3f e2e %.4g (%.3f ms) frac %.4f' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac']))
PY
