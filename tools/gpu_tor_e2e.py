"""Where does the time of thewalrus_b200.tor(O) go?  (bench tor48 showed e2e >> kernel.)  Also times ltor."""
import cProfile
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import torch

import bench
import thewalrus_b200 as wb
from thewalrus_b200 import _engine

kind, n, O = bench.make_input("tor48")
N = n // 2
for name, fn in (("tor", lambda: wb.tor(O)),):
    fn()
    torch.cuda.synchronize()
    for rep in range(3):
        t0 = time.perf_counter()
        v = fn()
        torch.cuda.synchronize()
        print(name, "call", rep, f"{(time.perf_counter() - t0) * 1e3:.3f} ms", v, flush=True)
pr = cProfile.Profile()
pr.enable()
wb.tor(O)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)

# raw C-ABI host entry for comparison
import ctypes
from thewalrus_b200 import _lib
lib = _lib.load()
Oc, pO = _lib.as_c128(O)
out = np.zeros(2)
ms = ctypes.c_double(0)
total = _engine.tor_num_prefixes(N)
for rep in range(3):
    t0 = time.perf_counter()
    lib.wb200_tor_host(0, pO, N, 0, total, _lib.dptr(out), ctypes.byref(ms))
    print("wb200_tor_host wall", f"{(time.perf_counter() - t0) * 1e3:.3f} ms kernel {ms.value:.3f} ms", out[0] + out[1], flush=True)

rng = np.random.default_rng(5)
g = 0.2 * (rng.standard_normal(N) + 1j * rng.standard_normal(N))
gamma = np.concatenate([g, g.conj()])
gp, pg = _lib.as_c128(gamma)
for rep in range(3):
    t0 = time.perf_counter()
    lib.wb200_ltor_host(0, pO, pg, N, 0, total, _lib.dptr(out), ctypes.byref(ms))
    print("wb200_ltor_host wall", f"{(time.perf_counter() - t0) * 1e3:.3f} ms kernel {ms.value:.3f} ms", out[0] + out[1], flush=True)
print("ltor(O, 0) vs tor(O):", wb.ltor(O, np.zeros(2 * N)), wb.tor(O))
