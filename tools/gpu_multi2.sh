#!/bin/bash
# 2-GPU check (run under gpurun --gpus 2): NCCL parity of every sharded entry point, bench lines at N = 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
python -m pytest tests -m gpu -x -q -k "hafnian_batch or batch_gamma or multi_gamma" > gpurun_out/pytest_gpu_new.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_new.log
timeout 600 $TR tools/gpu_dist_check.py > gpurun_out/dist_check_2.log 2>&1; tail -3 gpurun_out/dist_check_2.log | cut -c1-1500
for w in hafnian50 perm36 tor48 gbs16 hsample8; do
  timeout 600 $TR bench.py --gpus 2 --workload $w --steps 2 --warmup 3 > gpurun_out/bench_${w}_x2.json 2> gpurun_out/bench_${w}_x2.err; echo "bench $w x2 rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_${w}_x2.json').read().strip().splitlines() if l.startswith('{')][-1])
print('value %.4g %s ms/step %.3f e2e %.4g (%.3f ms) frac %.4f' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac']))
PY
done
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>&1 | tail -1 | cut -c1-200
