"""Development timing sweep of the hafnian DMMA kernel (kernel ms via the host C ABI)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import thewalrus_b200 as wb
from thewalrus_b200 import _lib
from oracle import c_oracle as co
lib = _lib.load()
rng = np.random.default_rng(1)
sizes = [int(a) for a in sys.argv[1:]] or [24, 32, 40, 42, 44, 46, 48, 50, 52, 56]
for n in sizes:
    G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)); A = G + G.T
    x = co.matched_order(A); Ax = np.ascontiguousarray(A[np.ix_(x, x)]); Dx = np.ascontiguousarray(np.diag(A)[x])
    m = n // 2
    steps = min(1 << (m - 1), 1 << 21)
    for D in (None, Dx):
        out = np.zeros(4); ms = ctypes.c_double(0)
        for rep in range(2):
            rc = lib.wb200_hafnian_host(0, _lib.dptr(Ax.view(np.float64)), None if D is None else _lib.dptr(D.view(np.float64)), n, 0, steps, _lib.dptr(out), ctypes.byref(ms))
            assert rc == 0, lib.wb200_last_error()
        got = complex(out[0] + out[1], out[2] + out[3])
        chk = co.hafnian_range(Ax, 0, 64, D); out2 = np.zeros(4)
        lib.wb200_hafnian_host(0, _lib.dptr(Ax.view(np.float64)), None if D is None else _lib.dptr(D.view(np.float64)), n, 0, 64, _lib.dptr(out2), None)
        g2 = complex(out2[0] + out2[1], out2[2] + out2[3])
        nprod = (m - 1) // 2
        print("n=%d %s steps=%d kernel %.2f ms  %.3e subsets/s  useful %.2f TFLOP/s  relerr(64 subsets) %.1e" % (
            n, "loop" if D is not None else "haf ", steps, ms.value, steps / ms.value * 1e3, steps * 8.0 * n**3 * nprod / ms.value * 1e-9, abs(g2 - chk) / abs(chk)), flush=True)
