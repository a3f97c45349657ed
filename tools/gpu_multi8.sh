N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/gpu_dist_check.py > gpurun_out/dist_check_$N.log 2>&1; tail -3 gpurun_out/dist_check_$N.log
for w in hafnian50 hafnian56 perm40 tor48; do
  timeout 900 $TR bench.py --gpus $N --workload $w --steps 2 --warmup 3 > gpurun_out/bench_${w}_x$N.json 2> gpurun_out/bench_${w}_x$N.err; echo "bench $w x$N rc=$?"
done
