"""Real-symmetric torontonian kernel against the complex kernel on the same real O: kernel time per call."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thewalrus_b200 import _lib, _engine
lib = _lib.load()
for N in (16, 20, 24, 28):
    rng = np.random.default_rng(N)
    B = rng.standard_normal((2 * N, 2 * N)); H = B @ B.T
    O = np.ascontiguousarray(0.9 * H / np.linalg.norm(H, 2)); Oc = np.ascontiguousarray(O.astype(np.complex128))
    total = _engine.tor_num_prefixes(N)
    outr, outc, ms = np.zeros(2), np.zeros(2), ctypes.c_double(0)
    tr = tc = 0.0
    for _ in range(3):
        assert lib.wb200_tor_f64_host(0, _lib.dptr(O), N, 0, total, _lib.dptr(outr), ctypes.byref(ms)) == 0; tr = ms.value
        assert lib.wb200_tor_host(0, _lib.dptr(Oc.view(np.float64)), N, 0, total, _lib.dptr(outc), ctypes.byref(ms)) == 0; tc = ms.value
    a, b = outr[0] + outr[1], outc[0] + outc[1]
    print("2N = %d: complex kernel %.3f ms, real kernel %.3f ms (x%.2f), rel diff %.1e" % (2 * N, tc, tr, tc / tr, abs(a - b) / abs(b)), flush=True)
