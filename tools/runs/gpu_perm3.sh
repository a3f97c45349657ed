mkdir -p gpurun_out
for ncu in 0 8 10 12 16 20; do
  echo "=== NCU=$ncu"
  WB200_PERM_NCU=$ncu python tools/gpu_perm_sweep.py --fast 2>&1 | tail -4
done > gpurun_out/perm_sweep_v5.log 2>&1
cat gpurun_out/perm_sweep_v5.log
