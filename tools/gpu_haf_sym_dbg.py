import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thewalrus_b200 import _lib
from oracle import c_oracle as co
lib = _lib.load()
rng = np.random.default_rng(7)
n = 50
G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)); A = G + G.T
x = co.matched_order(A); Ax = np.ascontiguousarray(A[np.ix_(x, x)])
for dbg in [0, 1, 2, 3, 4, 8, 16, 7, 15, 31]:
    os.environ["WB200_HS_DBG"] = str(dbg)
    out = np.zeros(4); ms = ctypes.c_double(0)
    for _ in range(2):
        rc = lib.wb200_hafnian_host(0, _lib.dptr(Ax.view(np.float64)), None, n, 0, 1 << 19, _lib.dptr(out), ctypes.byref(ms))
    print("dbg %2d: %.2f ms" % (dbg, ms.value), flush=True)
