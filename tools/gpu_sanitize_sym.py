"""Small ranges through every shape of the symmetric-half hafnian kernel (hafnian_sym.cu: named team barriers, mbarrier
split-phase barrier, column-wise reads of other panels' rows) - the target of compute-sanitizer memcheck / racecheck."""
import os
import sys

sys.path.insert(0, ".")
import numpy as np

from oracle import c_oracle as co
from thewalrus_b200 import _engine

rng = np.random.default_rng(5)
grid = int(sys.argv[1]) if len(sys.argv) > 1 else 8          # a few CTAs' worth of groups is enough for the tools
for n, mode in ((38, "1"), (40, "1"), (42, "1"), (44, "1"), (46, "1"), (48, "1"), (48, "4"), (50, "1"), (50, "4"), (54, "1"), (56, "1"), (58, "1"), (60, "1"), (62, "1"), (64, "1")):
    os.environ["WB200_HAF_SYM"] = mode
    G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    A = G + G.T
    x = co.matched_order(A)
    Ax = np.ascontiguousarray(A[np.ix_(x, x)])
    j0, j1 = 1000, 1000 + 8 * 148 * 4 + 3                      # the smallest range the dispatcher gives to this kernel
    got = _engine.combine4([_engine.hafnian_range(Ax, None, j0, j1)])
    want = co.hafnian_range(Ax, j0, j1)
    print("n", n, "mode", mode, "rel err", abs(got - want) / abs(want), flush=True)
    assert abs(got - want) <= 1e-10 * abs(want)
