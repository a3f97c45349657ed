"""Two-team symmetric-half kernel: kernel time against the start skew of the second team (cycles)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thewalrus_b200 import _lib
from oracle import c_oracle as co
lib = _lib.load()
rng = np.random.default_rng(7)
for n in (50, 48):
    G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)); A = G + G.T
    x = co.matched_order(A); Ax = np.ascontiguousarray(A[np.ix_(x, x)])
    for skew in [0, 3000, 6000, 9000, 12000, 15000, 18000, 21000, 24000]:
        os.environ["WB200_HS_SKEW"] = str(skew)
        out = np.zeros(4); ms = ctypes.c_double(0)
        for _ in range(2):
            rc = lib.wb200_hafnian_host(0, _lib.dptr(Ax.view(np.float64)), None, n, 0, 1 << 19, _lib.dptr(out), ctypes.byref(ms))
        print("n=%d skew %5d: %.2f ms" % (n, skew, ms.value), flush=True)
