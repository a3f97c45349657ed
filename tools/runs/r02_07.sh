#!/bin/bash
# round 2, GPU call 7: odd totals on the tensor-core pattern kernel, lane-parallel series; sampler + gbs16 bench
mkdir -p gpurun_out
export WB200_SKIP_SLOW=1
timeout 900 python -m pytest tests/test_gpu_pat_dmma.py tests/test_samples.py tests/test_gpu_dev_entries.py -q -m gpu -x 2>&1 | tail -25 > gpurun_out/r02_pytest_pat_odd.log
tail -4 gpurun_out/r02_pytest_pat_odd.log
timeout 1200 python -m pytest tests -q -m gpu -x --deselect tests/golden/ref_suite 2>&1 | tail -8 > gpurun_out/r02_pytest_gpu_e.log
tail -3 gpurun_out/r02_pytest_gpu_e.log
python bench.py --workload gbs16 --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_gbs16_e.json 2>&1
python bench.py --workload hsample8 --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_hsample8_e.json 2>&1
python - <<'PY'
import json
for f in ["r02_bench_gbs16_e.json","r02_bench_hsample8_e.json"]:
    d=json.loads([l for l in open('gpurun_out/'+f).read().strip().splitlines() if l.startswith('{')][-1])
    print(f, "value %.4g"%d["value"], "ms %.4g"%d["ms_per_step"], "e2e ms %.4g"%d["e2e"]["ms_per_step"], "roof %.3f"%d["roofline"]["frac"], d.get("result_rel_err"))
PY
echo finished
