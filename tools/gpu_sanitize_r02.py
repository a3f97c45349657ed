"""Small calls through every kernel added in round 2 — the target of compute-sanitizer (memcheck / racecheck):
pat_dmma_kernel (all tile classes, even and odd totals), haf_dmma_kernel with panel-split teams (named barriers),
tor4_kernel (opt-in experiment), the chain sampler kernels and the *_dev entries' stream-ordered scratch."""
import os
import sys

sys.path.insert(0, ".")
import numpy as np

import bench
import thewalrus_b200 as wb
from thewalrus_b200 import samples as ws

rng = np.random.default_rng(3)


def mat(nv):
    G = rng.standard_normal((nv, nv)) + 1j * rng.standard_normal((nv, nv))
    return (G + G.T) / np.sqrt(2.0 * nv), (rng.standard_normal(nv) + 1j * rng.standard_normal(nv)) / np.sqrt(nv)


for E in (2, 5, 6, 8, 10, 12, 13, 16):                       # every tile class of pat_dmma_kernel
    nv = 2 * E + 2
    A, D = mat(nv)
    rpt = np.zeros((3, nv), dtype=np.int32)
    for b in range(3):
        rpt[b, rng.permutation(nv)[: 2 * E + (b & 1)]] = 1     # even and odd totals
    print("pat E", E, wb.quantum.lhaf_patterns(A, D, rpt)[:2])
A, D = mat(8)
rpt = rng.integers(0, 3, (40, 8)).astype(np.int32)
print("pat reps", wb.quantum.lhaf_patterns(A, D, rpt)[:2])
for n in (6, 12, 16, 20, 24, 28):                            # PS = 3 team shapes (3-warp and 12-warp CTAs)
    A, _ = mat(n)
    print("haf", n, wb.hafnian(A), wb.hafnian(A, loop=True))
_, _, O = bench.make_input("tor24")
os.environ["WB200_TOR_V4"] = "1"
print("tor4", wb.tor(O))
del os.environ["WB200_TOR_V4"]
print("tor", wb.tor(O))
_, M, (mu, cov) = bench.make_input("hsample5")
np.random.seed(1)
print("chains", ws.hafnian_sample_state(cov, 64, mean=mu, cutoff=4).sum())
_, _, U = bench.make_input("perm14")
print("perm", wb.perm(U), wb.perm(U, "ryser"))
