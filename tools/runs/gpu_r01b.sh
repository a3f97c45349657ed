#!/bin/bash
# round 1, re-entry call 1: parity suite, tor e2e breakdown, DRAM traffic of the headline kernel, bench lines
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python tools/gpu_tor_e2e.py > gpurun_out/tor_e2e.log 2>&1; tail -40 gpurun_out/tor_e2e.log
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:haf_dmma -c 1 \
    --csv --log-file gpurun_out/traffic_hafnian50.csv python tools/gpu_one_hafnian.py hafnian50 > gpurun_out/traffic_run.log 2>&1
tail -3 gpurun_out/traffic_hafnian50.csv
python bench.py > gpurun_out/bench_hafnian50.json 2> gpurun_out/bench_hafnian50.err; cat gpurun_out/bench_hafnian50.json
python bench.py --workload tor48 > gpurun_out/bench_tor48.json 2> gpurun_out/bench_tor48.err; cat gpurun_out/bench_tor48.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_hafnian50.json 2>&1; cat gpurun_out/bench_ref_hafnian50.json
