#!/bin/bash
# torontonian: fewer prefixes per CTA for small problems — parity, racecheck, timing at N = 12..24
mkdir -p gpurun_out
python tools/gpu_tor_small.py 2>&1 | tail -5
timeout 300 compute-sanitizer --tool racecheck --print-limit 5 python tools/gpu_tor_small.py > gpurun_out/racecheck_tor.log 2>&1; tail -1 gpurun_out/racecheck_tor.log
python -m pytest tests -m gpu -x -q -k "tor or threshold" > gpurun_out/pytest_gpu_tor.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_gpu_tor.log
python - <<'PY'
import sys, ctypes, numpy as np
sys.path.insert(0, ".")
import bench
from thewalrus_b200 import _lib, _engine
lib = _lib.load()
for n in (24, 28, 32, 36, 40, 48):
    _, _, O = bench.make_input(f"tor{n}")
    Oc, pO = _lib.as_c128(O)
    out = np.zeros(2); ms = ctypes.c_double(0)
    total = _engine.tor_num_prefixes(n // 2)
    best = 1e9
    for rep in range(4):
        lib.wb200_tor_host(0, pO, n // 2, 0, total, _lib.dptr(out), ctypes.byref(ms)); best = min(best, ms.value)
    print(f"tor 2N={n}: prefixes {total} kernel {best:.4f} ms  {2.0 ** (n // 2) / best * 1e-6:.3f} Gsubsets/s", flush=True)
PY
