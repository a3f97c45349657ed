"""Block-shape experiment for the permanent kernel: same kernel built with different CTA sizes
(tools/variants/libwb_perm_t*.so, -DWB_PERM_THREADS / -DWB_PERM_MIN_CTAS); prints kernel ms for n = 32 / 40."""
import ctypes
import sys

import numpy as np

sys.path.insert(0, ".")
import bench

libs = sys.argv[1:]
dp = ctypes.POINTER(ctypes.c_double)
for path in libs:
    lib = ctypes.CDLL(path)
    lib.wb200_perm_host.restype = ctypes.c_int
    lib.wb200_perm_host.argtypes = [ctypes.c_int, dp, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, dp, dp]
    import os
    sizes = [int(x) for x in os.environ.get("PERM_SIZES", "32,40").split(",")]
    for w, steps in [(f"perm{n}", 1 << min(n - 1, 34)) for n in sizes]:
        _, n, U = bench.make_input(w)
        U = np.ascontiguousarray(U, dtype=np.complex128)
        out = np.zeros(4)
        ms = ctypes.c_double(0)
        best = 1e30
        for rep in range(3):
            rc = lib.wb200_perm_host(0, U.view(np.float64).ctypes.data_as(dp), n, 0, 0, steps, out.ctypes.data_as(dp), ctypes.byref(ms))
            assert rc == 0, rc
            best = min(best, ms.value)
        print(f"{path} {w} steps 2^{int(np.log2(steps))}: {best:.3f} ms -> {steps / best * 1e-6:.4g} Gsteps/s, {(8 * n - 4) * steps / best * 1e-9:.2f} TFLOP/s, sum {out[0]:.6e}", flush=True)
