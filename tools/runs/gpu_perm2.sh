set -x
mkdir -p gpurun_out
python tools/gpu_perm_sweep.py > gpurun_out/perm_sweep_auto.log 2>&1; tail -8 gpurun_out/perm_sweep_auto.log
WB200_PERM_SL=4 python tools/gpu_perm_sweep.py > gpurun_out/perm_sweep_sl4.log 2>&1; tail -8 gpurun_out/perm_sweep_sl4.log
timeout 600 python -m pytest tests -m gpu -x -q -k perm > gpurun_out/pytest_perm.log 2>&1; tail -3 gpurun_out/pytest_perm.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:perm_kernel -s 3 -c 1 -f -o gpurun_out/prof_perm32_v3 python bench.py --workload perm32 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f3.log 2>&1
