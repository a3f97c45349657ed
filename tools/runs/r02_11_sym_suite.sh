#!/bin/bash
# round 2, GPU call 11: symmetric-half kernel through the test suite (own tests + the complete n = 50 goldens), smoke, headline bench
mkdir -p gpurun_out
WB200_SKIP_SLOW=1 python -m pytest tests/test_gpu_haf_sym.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py tests/test_gpu_dev_entries.py -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r02_pytest_sym.log
tail -4 gpurun_out/r02_pytest_sym.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_sym.log 2>&1; tail -2 gpurun_out/r02_smoke_sym.log
python bench.py --steps 3 --warmup 3 --no-secondary > gpurun_out/r02_bench_hafnian50_sym.json 2> gpurun_out/r02_bench_hafnian50_sym.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_hafnian50_sym.json').read().strip().splitlines() if l.startswith('{')][-1])
print("hafnian50 value %.4g ms %.5g e2e ms %.5g roof %.3f err %s clk %s cpu %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d.get("result_rel_err"), d["clocks"]["sm_mhz"], (d.get("cpu_baseline") or {}).get("value")))
PY
echo finished
