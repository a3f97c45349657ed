#!/bin/bash
mkdir -p gpurun_out
T=tools/variants/libwb_perm_t256.so
B=thewalrus_b200/libwalrus_b200.so
{
for cfg in "2 2" "4 2" "2 1" "4 1"; do set -- $cfg
  echo "== T256 SL=$1 NS=$2"; WB200_PERM_SL=$1 WB200_PERM_NS=$2 PERM_SIZES=24,28,32,36,40 python tools/gpu_perm_shape.py $T
done
echo "== default shapes, T128 vs T256, other sizes"
PERM_SIZES=16,20,24,28,36,44,48,56,64 python tools/gpu_perm_shape.py $B $T
} 2>&1 | tee gpurun_out/perm_shape2.txt
