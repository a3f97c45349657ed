"""One kernel call over a SUB-RANGE of the index space (ncu target: same template instance as the full run, bounded
time).  usage: python tools/gpu_range.py hafnian50 20   -> first 2^20 subsets of the bench input."""
import sys

sys.path.insert(0, ".")
import numpy as np

import bench
from thewalrus_b200 import _engine
from thewalrus_b200._prep import matched_reps

w, lg = sys.argv[1], int(sys.argv[2])
kind, n, A = bench.make_input(w)
if kind == "hafnian":
    x, _, _ = matched_reps([1] * n)
    Ax = np.ascontiguousarray(A[np.ix_(x, x)].astype(np.complex128))
    print(w, _engine.combine4([_engine.hafnian_range(Ax, None, 0, 1 << lg)]))
elif kind == "perm":
    print(w, _engine.combine4([_engine.perm_range(A.astype(np.complex128), 0, 0, 1 << lg)]))
