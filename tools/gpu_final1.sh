#!/bin/bash
# round 1 closing run on one B200: full parity suite, smoke, every bench workload, launch list of the default bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_hafnian50.json 2> gpurun_out/bench_hafnian50.err; tail -c 400 gpurun_out/bench_hafnian50.json; tail -2 gpurun_out/bench_hafnian50.err
for w in hafnian24 lhaf50 perm32 perm40 tor48 gbs16 ltor48 mtl14 brs12 hsample8; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  echo "== $w rc=$?"; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1])
print('value %.4g %s ms/step %.3f e2e %.4g (%.3f ms) frac %.4f cpu %.4g' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['cpu_baseline']['value']))
PY
  tail -2 gpurun_out/bench_$w.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_hafnian50.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/launches_run.log 2>&1; tail -1 gpurun_out/launches_run.log | head -c 300
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_hafnian50.json 2>&1; tail -c 300 gpurun_out/bench_ref_hafnian50.json
