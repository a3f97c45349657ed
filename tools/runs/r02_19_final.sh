#!/bin/bash
# round 2, closing GPU call (after the symmetric-half kernel): full GPU suite (incl. n = 56 cases and the reference's own test files), smoke(), the default
# bench line (headline + secondary), launch list + DRAM traffic of the dominant kernels of the same commands
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rs -s 2>&1 | grep -v "^$" | tail -120 > gpurun_out/r02_pytest_gpu_final3.log
tail -3 gpurun_out/r02_pytest_gpu_final3.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02_smoke_final3.log 2>&1; tail -2 gpurun_out/r02_smoke_final3.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_default_final3.json 2> gpurun_out/r02_bench_default_final3.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_final3.json 2> gpurun_out/r02_bench_reference_final3.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_hafnian50_final3.csv python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/r02_launches_final_run.log 2>&1
for w in hafnian24 perm32 tor48 gbs16; do
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_${w}_final3.csv python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
done
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_default_final3.json').read().strip().splitlines() if l.startswith('{')][-1])
def show(n,x): print(n, "value %.4g"%x["value"], "ms %.4g"%x["ms_per_step"], "e2e ms %.4g"%x["e2e"]["ms_per_step"], "roof %.3f"%x["roofline"]["frac"], "err", x.get("result_rel_err"), "clk", x["clocks"]["sm_mhz"], x["clocks"]["samples"], "cpu", (x.get("cpu_baseline") or {}).get("kind"), (x.get("cpu_baseline") or {}).get("value"))
show("hafnian50", d)
for k,v in d.get("secondary",{}).items():
    show(k,v) if "error" not in v else print(k, v)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:haf_sym_kernel -c 1 -o gpurun_out/r02_prof_haf50_sym_final3 -f python tools/gpu_range.py hafnian50 18 > gpurun_out/r02_ncu_haf50_sym_final3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:haf_sym_kernel -c 1 -o gpurun_out/r02_prof_haf56_sym_final3 -f python tools/gpu_range.py hafnian56 18 > gpurun_out/r02_ncu_haf56_sym_final3.log 2>&1
echo finished
