"""Permanent kernel sweep on one B200: parity against the long-double C oracle on sub-ranges, then timings of
full or truncated step ranges through wb200_perm_host (kernel ms from CUDA events inside the call)."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thewalrus_b200 import _lib
from oracle import c_oracle as co

lib = _lib.load()
rng = np.random.default_rng(11)

def haar_block(n):
    Z = (rng.standard_normal((2 * n, 2 * n)) + 1j * rng.standard_normal((2 * n, 2 * n))) / np.sqrt(2)
    Q, R = np.linalg.qr(Z)
    return np.ascontiguousarray((Q * (np.diag(R) / np.abs(np.diag(R))))[:n, :n])

def gpu(U, method, k0, k1):
    out = np.zeros(4); ms = ctypes.c_double(0)
    rc = lib.wb200_perm_host(0, _lib.dptr(U.view(np.float64)), U.shape[0], method, k0, k1, _lib.dptr(out), ctypes.byref(ms))
    _lib.check(rc, "perm")
    return complex(out[0] + out[1], out[2] + out[3]), ms.value

worst = 0.0
FAST = "--fast" in sys.argv
for n in ((32, 36, 40) if FAST else (2, 3, 5, 8, 9, 12, 16, 17, 20, 23, 26, 29, 32, 33, 36, 40, 44)):
    U = haar_block(n)
    for method in (0, 1):
        steps = 1 << (n - 1 if method == 0 else n)
        for (k0, k1) in ((0, min(steps, 1 << 16)), (min(steps, 12345) // 3, min(steps, 70001)), (max(0, steps - 33333), steps)):
            if k1 <= k0: continue
            g, _ = gpu(U, method, k0, k1)
            w = co.perm_range(U, method, k0, k1)
            w = complex(w)
            e = abs(g - w) / max(abs(w), 1e-300)
            worst = max(worst, e)
            if e > 1e-9: print("MISMATCH n", n, "method", method, k0, k1, g, w, e)
print("perm parity worst rel err vs oracle: %.3e" % worst)
for n, cap in (((32, None), (36, 1 << 33), (40, 1 << 33)) if FAST else ((24, None), (28, None), (30, None), (32, None), (36, 1 << 33), (40, 1 << 33))):
    U = haar_block(n)
    steps = 1 << (n - 1)
    k1 = steps if cap is None else min(steps, cap)
    gpu(U, 0, 0, k1)
    _, ms = gpu(U, 0, 0, k1)
    print("perm n=%d steps=%d kernel %.3f ms -> %.4e subsets/s, %.2f TFLOP/s (8n-4 flops/subset), SL=%s" % (
        n, k1, ms, k1 / ms * 1e3, k1 * (8.0 * n - 4) / ms * 1e-9, os.environ.get("WB200_PERM_SL", "auto")))
