#!/bin/bash
# round 2, GPU call 6: tor v4 (warp-autonomous tensor-core expansion): parity, shape sweep, ncu
mkdir -p gpurun_out
export WB200_SKIP_SLOW=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next.py tests/test_gpu_fullsize.py -q -m gpu -k "tor" -x 2>&1 | tail -25 > gpurun_out/r02_pytest_tor4.log
tail -3 gpurun_out/r02_pytest_tor4.log
for cfg in "4 8" "5 8" "3 8" "2 8" "4 6" "4 4" "5 6"; do
  set -- $cfg
  echo "== G=$1 W=$2" >> gpurun_out/r02_tor4_shapes.txt
  WB200_TOR4_G=$1 WB200_TOR4_W=$2 python bench.py --workload tor48 --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('tor48 ms %.4f e2e %.4f err %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d.get('result_rel_err')))
" >> gpurun_out/r02_tor4_shapes.txt
done
echo "== v3" >> gpurun_out/r02_tor4_shapes.txt
WB200_TOR_V3=1 python bench.py --workload tor48 --steps 100 --warmup 5 --no-cpu-baseline 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1 >> gpurun_out/r02_tor4_shapes.txt
python bench.py --workload tor60 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_bench_tor60_v4.json 2>&1
ncu --set full --clock-control none --import-source on -k regex:tor4_kernel -c 1 -o gpurun_out/r02_prof_tor48_v4 -f python tools/gpu_one_hafnian.py tor48 > gpurun_out/r02_ncu_tor4.log 2>&1
cat gpurun_out/r02_tor4_shapes.txt
echo finished
