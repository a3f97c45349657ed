"""Shape experiment for the torontonian kernel: the same source built with different breadth-first depth (DC),
prefixes per CTA (2^G), CTA size and shared-memory cap (-DWB_TOR_*); prints kernel ms for 2N = 40 / 48 and checks
the values against each other."""
import ctypes
import sys

import numpy as np

sys.path.insert(0, ".")
import bench

dp = ctypes.POINTER(ctypes.c_double)
ref = {}
for path in sys.argv[1:]:
    lib = ctypes.CDLL(path)
    lib.wb200_tor_host.restype = ctypes.c_int
    lib.wb200_tor_host.argtypes = [ctypes.c_int, dp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, dp, dp]
    lib.wb200_tor_num_prefixes.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_uint64)]
    for w in ("tor32", "tor40", "tor48"):
        _, n, O = bench.make_input(w)
        O = np.ascontiguousarray(O, dtype=np.complex128)
        total = ctypes.c_uint64(0)
        lib.wb200_tor_num_prefixes(n // 2, ctypes.byref(total))
        out, ms, best = np.zeros(2), ctypes.c_double(0), 1e30
        for rep in range(4):
            rc = lib.wb200_tor_host(0, O.view(np.float64).ctypes.data_as(dp), n // 2, 0, total.value, out.ctypes.data_as(dp), ctypes.byref(ms))
            assert rc == 0, rc
            best = min(best, ms.value)
        val = out[0] + out[1]
        ref.setdefault(w, val)
        print(f"{path} {w}: prefixes {total.value} kernel {best:.4f} ms  value {val:.12e} (rel diff to first {abs(val - ref[w]) / abs(ref[w]):.1e})", flush=True)
