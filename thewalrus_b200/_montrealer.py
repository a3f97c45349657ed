"""Montrealer front end — drop-in for thewalrus.mtl / thewalrus.lmtl (thewalrus/_montrealer.py:105-135).

The 2^n subset sum of power traces (``montrealer`` :37-57, ``lmontrealer`` :78-102; serial in the reference)
runs on the GPU with one warp per subset (``wb200_mtl_host``), reusing the shared-memory power-trace chain of the
batched loop-hafnian kernels.
"""
import numpy as np

from . import _engine
from ._prep import dd_sum
from ._torontonian import tor_input_checks

__all__ = ["mtl", "lmtl"]


def _run(A, zeta, group, device):
    n = len(A) // 2
    if n == 0:
        return np.complex128(0.0)
    table = _engine.run_sharded(1 << n, lambda lo, hi: _engine.mtl_range(A, zeta, lo, hi, device), group, width=8)

    def col(i):
        hi, lo = dd_sum([(t[i], t[i + 1]) for t in table])
        return hi + lo

    V = complex(col(0), col(2))
    W = complex(col(4), col(6))
    return np.complex128((-1) ** (n + 1) * (V / (2 * n) + W / 2))


def mtl(A, *, group=None, device=None):
    """Montrealer of a 2n x 2n matrix (thewalrus/_montrealer.py:121-135)."""
    tor_input_checks(A)
    return _run(np.asarray(A, dtype=np.complex128), None, group, device)


def lmtl(A, zeta, *, group=None, device=None):
    """Loop montrealer of a 2n x 2n matrix and a 2n vector (thewalrus/_montrealer.py:105-118)."""
    tor_input_checks(A, zeta)
    return _run(np.asarray(A, dtype=np.complex128), np.asarray(zeta, dtype=np.complex128), group, device)
