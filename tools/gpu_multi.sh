# usage: bash tools/gpu_multi.sh N   (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_multi.log 2>&1; tail -2 gpurun_out/pytest_gpu_multi.log; fi
timeout 600 $TR tools/gpu_dist_check.py > gpurun_out/dist_check_$N.log 2>&1; tail -3 gpurun_out/dist_check_$N.log
for w in hafnian50 perm36 tor48 gbs16; do
  timeout 900 $TR bench.py --gpus $N --workload $w --steps 2 --warmup 3 > gpurun_out/bench_${w}_x$N.json 2> gpurun_out/bench_${w}_x$N.err; echo "bench $w x$N rc=$?"; tail -c 600 gpurun_out/bench_${w}_x$N.json | head -c 400; echo
done
for w in hafnian56 perm40; do
  timeout 1500 $TR bench.py --gpus $N --workload $w --steps 1 --warmup 3 > gpurun_out/bench_${w}_x$N.json 2> gpurun_out/bench_${w}_x$N.err; echo "bench $w x$N rc=$?"
done
