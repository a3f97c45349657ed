"""Symmetric-half hafnian kernel (hafnian_sym.cu) against the row-panel kernel (hafnian_dmma.cu) on the same ranges:
partial sums, kernel time; WB200_HAF_SYM = 0 row-panel kernel, 4 one team of 12 warps, 1 two teams of 6.  usage: python tools/gpu_haf_sym.py [log2 steps, default 20] [sizes ...]"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thewalrus_b200 import _lib
from oracle import c_oracle as co
lib = _lib.load()
rng = np.random.default_rng(7)
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
sizes = [int(a) for a in sys.argv[2:]] or [48, 50]


def run(Ax, n, j0, j1, sym):
    os.environ["WB200_HAF_SYM"] = str(sym)
    out = np.zeros(4); ms = ctypes.c_double(0)
    for _ in range(2):
        rc = lib.wb200_hafnian_host(0, _lib.dptr(Ax.view(np.float64)), None, n, j0, j1, _lib.dptr(out), ctypes.byref(ms))
        assert rc == 0, lib.wb200_last_error()
    return complex(out[0] + out[1], out[2] + out[3]), ms.value


for n in sizes:
    G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)); A = G + G.T
    x = co.matched_order(A); Ax = np.ascontiguousarray(A[np.ix_(x, x)])
    m = n // 2
    for (j0, j1) in ((0, 8192), (12345, 12345 + 8192 + 3), (0, min(1 << lg, 1 << (m - 1)))):
        a, tp = run(Ax, n, j0, j1, 0)
        c, t4 = run(Ax, n, j0, j1, 4) if n in (48, 50) else (a, float("nan"))
        b, ts = run(Ax, n, j0, j1, 1)
        print("n=%d [%d, %d): panel %.3f ms  sym one team %.3f ms (%.1e)  sym two teams %.3f ms  (x%.3f)  rel diff %.2e  %.3e subsets/s" % (
            n, j0, j1, tp, t4, abs(a - c) / abs(a), ts, tp / ts, abs(a - b) / abs(a), (j1 - j0) / ts * 1e3), flush=True)
