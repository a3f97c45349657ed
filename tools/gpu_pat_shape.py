"""Work-item size experiment for the patterns kernel (-DWB_BW_CHUNK): kernel ms of the gbs16 workload."""
import ctypes
import sys

import numpy as np

sys.path.insert(0, ".")
import bench

M, mu, cov, pats, A, gamma, rpt = bench.gbs_inputs("gbs16", 100000)
dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
A = np.ascontiguousarray(A, dtype=np.complex128)
gamma = np.ascontiguousarray(gamma, dtype=np.complex128)
ref = None
for path in sys.argv[1:]:
    lib = ctypes.CDLL(path)
    lib.wb200_lhaf_patterns_host.restype = ctypes.c_int
    lib.wb200_lhaf_patterns_host.argtypes = [ctypes.c_int, dp, dp, ctypes.c_int, ip, ctypes.c_int64, ctypes.c_int, dp, dp]
    out, ms, best = np.zeros(len(rpt), dtype=np.complex128), ctypes.c_double(0), 1e30
    for rep in range(3):
        rc = lib.wb200_lhaf_patterns_host(0, A.view(np.float64).ctypes.data_as(dp), gamma.view(np.float64).ctypes.data_as(dp),
                                          rpt.shape[1], rpt.ctypes.data_as(ip), len(rpt), 1, out.view(np.float64).ctypes.data_as(dp),
                                          ctypes.byref(ms))
        assert rc == 0, rc
        best = min(best, ms.value)
    ref = out.copy() if ref is None else ref
    print(f"{path}: gbs16 kernel {best:.2f} ms, max rel diff to first {np.max(np.abs(out - ref) / np.maximum(np.abs(ref), 1e-300)):.1e}", flush=True)
