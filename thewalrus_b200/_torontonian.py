"""Torontonian front end — drop-in for thewalrus.tor (thewalrus/_torontonian.py:23-58)."""
import numpy as np

from . import _engine
from ._prep import dd_sum

__all__ = ["tor", "tor_input_checks"]


def tor_input_checks(A, loops=None):
    """Same checks and messages as thewalrus/_torontonian.py:23-44."""
    if not isinstance(A, np.ndarray):
        raise TypeError("Input matrix must be a NumPy array.")
    matshape = A.shape
    if matshape[0] != matshape[1]:
        raise ValueError("Input matrix must be square.")
    if matshape[0] % 2 != 0:
        raise ValueError("matrix dimension must be even")
    if loops is not None:
        if not isinstance(loops, np.ndarray):
            raise TypeError("Input matrix must be a NumPy array.")
        if matshape[0] != len(loops):
            raise ValueError("gamma must be a vector matching the dimension of A")


def tor(A, recursive=True, *, group=None, device=None):
    """Torontonian  sum_S (-1)^(N-|S|) / sqrt(det(I - A_S))  (thewalrus/_torontonian.py:47-58).

    ``recursive`` is accepted for signature compatibility: the reference's two variants
    (``rec_torontonian`` :227-247 and ``numba_tor`` :123-154) compute the same number; the GPU kernel
    evaluates the subset tree with shared Schur-complement prefixes either way.
    Returns ``np.float64`` for real input and ``np.complex128`` for complex input, like the reference.
    """
    tor_input_checks(A)
    del recursive
    N = A.shape[0] // 2
    is_complex = np.iscomplexobj(A)
    if N == 0:
        return np.complex128(1.0) if is_complex else np.float64(1.0)
    if N == 1:
        # single mode: closed form  -1 + 1/sqrt(det(I - A))  (both terms of the two-subset sum)
        det = ((1 - A[0, 0]) * (1 - A[1, 1]) - A[0, 1] * A[1, 0]).real
        val = -1.0 + 1.0 / np.sqrt(det)
        return np.complex128(val) if is_complex else np.float64(val)
    total = _engine.tor_num_prefixes(N)
    table = _engine.run_sharded(total, lambda lo, hi: _engine.tor_range(A, lo, hi, device), group, width=2)
    hi, lo = dd_sum([(p[0], p[1]) for p in table])
    val = hi + lo
    return np.complex128(val) if is_complex else np.float64(val)
