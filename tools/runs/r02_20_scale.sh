#!/bin/bash
# round 2, multi-GPU call: hafnian50 bench line with the symmetric-half kernel at N ranks (scaling table)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --workload hafnian50 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_hafnian50_sym_x$N.json 2> gpurun_out/r02_bench_hafnian50_sym_x$N.err; echo "bench hafnian50 x$N rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_hafnian50_sym_x$N.json').read().strip().splitlines() if l.startswith('{')][-1])
print('value %.4g %s ms/step %.3f e2e %.4g (%.3f ms) frac %.4f err %s' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d.get('result_rel_err')))
PY
