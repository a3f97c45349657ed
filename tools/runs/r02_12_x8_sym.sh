#!/bin/bash
# round 2, 8-GPU call 2: headline bench line with the symmetric-half kernel at 8 ranks (complete result against the golden)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --workload hafnian50 --steps 5 --warmup 3 > gpurun_out/r02_bench_hafnian50_sym_x$N.json 2> gpurun_out/r02_bench_hafnian50_sym_x$N.err; echo "bench hafnian50 x$N rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_hafnian50_sym_x$N.json').read().strip().splitlines() if l.startswith('{')][-1])
print('value %.4g %s ms/step %.3f e2e %.4g (%.3f ms) frac %.4f err %s' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d.get('result_rel_err')))
PY
