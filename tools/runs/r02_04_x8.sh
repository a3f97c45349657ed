#!/bin/bash
# round 2, 8-GPU call: NCCL parity of every sharded entry point, full-size goldens at 8 ranks, scaling bench lines
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/gpu_dist_check.py > gpurun_out/r02_dist_check_x$N.log 2>&1; tail -2 gpurun_out/r02_dist_check_x$N.log | cut -c1-1500
timeout 400 $TR tools/gpu_fullsize_x8.py > gpurun_out/r02_fullsize_x$N.log 2>&1; grep "^\[x\|FULLSIZE" gpurun_out/r02_fullsize_x$N.log | cut -c1-220
for w in hafnian50 hafnian56 perm40 tor48 tor60 gbs16; do
  timeout 300 $TR bench.py --gpus $N --workload $w --steps 2 --warmup 1 > gpurun_out/r02_bench_${w}_x$N.json 2> gpurun_out/r02_bench_${w}_x$N.err; echo "bench $w x$N rc=$?"
done
timeout 200 python bench.py --workload tor60 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_tor60_x1.json 2>&1
echo finished
