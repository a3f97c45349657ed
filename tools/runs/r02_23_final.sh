#!/bin/bash
# round 2, closing GPU call on the final tree (after the packed / real-symmetric torontonian): full GPU suite, smoke(), default bench
# line (headline + secondary), reference arm.  (The ncu launch lists of r02_19_final.sh were not repeated.)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rs -s 2>&1 | grep -v "^$" | tail -120 > gpurun_out/r02_pytest_gpu_final4.log
tail -3 gpurun_out/r02_pytest_gpu_final4.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02_smoke_final4.log 2>&1; tail -2 gpurun_out/r02_smoke_final4.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_default_final4.json 2> gpurun_out/r02_bench_default_final4.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_final4.json 2> gpurun_out/r02_bench_reference_final4.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_default_final4.json').read().strip().splitlines() if l.startswith('{')][-1])
def show(n,x): print(n, "value %.4g"%x["value"], "ms %.4g"%x["ms_per_step"], "e2e ms %.4g"%x["e2e"]["ms_per_step"], "roof %.3f"%x["roofline"]["frac"], "err", x.get("result_rel_err"), "clk", x["clocks"]["sm_mhz"], x["clocks"]["samples"], "cpu", (x.get("cpu_baseline") or {}).get("kind"), (x.get("cpu_baseline") or {}).get("value"))
show("hafnian50", d)
for k,v in d.get("secondary",{}).items():
    show(k,v) if "error" not in v else print(k, v)
PY
echo finished
