#!/bin/bash
# 8-GPU closing check (run under gpurun --gpus 8): NCCL parity of every sharded entry point, headline bench lines
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/gpu_dist_check.py > gpurun_out/dist_check_$N.log 2>&1; tail -2 gpurun_out/dist_check_$N.log | cut -c1-1200
for w in hafnian50 perm40 hafnian56; do
  timeout 300 $TR bench.py --gpus $N --workload $w --steps 1 --warmup 1 > gpurun_out/bench_${w}_x$N.json 2> gpurun_out/bench_${w}_x$N.err; echo "bench $w x$N rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_${w}_x$N.json').read().strip().splitlines() if l.startswith('{')][-1])
print('value %.4g %s ms/step %.3f e2e %.4g (%.3f ms) frac %.4f' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac']))
PY
done
