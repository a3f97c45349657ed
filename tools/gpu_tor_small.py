"""Small torontonian / loop torontonian calls (target of compute-sanitizer runs) checked against the oracle."""
import sys

import numpy as np

sys.path.insert(0, ".")
import thewalrus_b200 as wb
from oracle import walrus_oracle as wo

rng = np.random.default_rng(3)
for N in (2, 3, 9, 11, 13):
    B = rng.standard_normal((2 * N, 2 * N)) + 1j * rng.standard_normal((2 * N, 2 * N))
    H = B @ B.conj().T
    O = 0.8 * H / np.linalg.norm(H, 2)
    g = 0.3 * (rng.standard_normal(N) + 1j * rng.standard_normal(N))
    gam = np.concatenate([g, g.conj()])
    t, lt = wb.tor(O), wb.ltor(O, gam)
    wt, wl = wo.tor_direct(O), wo.ltor_direct(O, gam)
    print(N, "tor rel err %.2e  ltor rel err %.2e" % (abs(t - wt) / abs(wt), abs(lt - wl) / abs(wl)), flush=True)
