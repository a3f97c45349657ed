mkdir -p gpurun_out
WB200_PERM_SL=2 WB200_PERM_NS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:perm_kernel -s 3 -c 1 -f -o gpurun_out/prof_perm32_v4 python bench.py --workload perm32 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f4.log 2>&1
tail -3 gpurun_out/ncu_f4.log
