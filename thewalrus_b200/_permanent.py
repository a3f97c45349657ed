"""Permanent front end — drop-in for thewalrus.perm (thewalrus/_permanent.py:34-83)."""
import numpy as np

from . import _engine

__all__ = ["perm", "perm_bbfg", "perm_ryser"]


def _steps(n, ryser):
    return 1 << (n if ryser else n - 1)


def _run(A, ryser, group, device):
    n = len(A)
    method = 1 if ryser else 0
    total = _steps(n, ryser)
    if np.issubdtype(A.dtype, np.integer):
        # numba specialises the reference on int64 and sums in exact (wrapping) int64
        # [SURVEY.md 3.3]; bbfg then performs a true division by 2^(n-1).
        rank, world = _engine._rank_world(group, None)
        from ._prep import shard_range

        lo, hi = shard_range(total, rank, world)
        part = _engine.perm_int64_range(A.astype(np.int64), method, lo, hi, device) if hi > lo else 0
        if world > 1:
            table = _engine.allreduce_partials(np.array([float(part >> 32), float(part & 0xFFFFFFFF)]),
                                               None if group is True else group)
            tot = sum((int(h) << 32) + int(l) for h, l in table)
            tot = (tot + (1 << 63)) % (1 << 64) - (1 << 63)
        else:
            tot = part
        return tot if ryser else tot / total
    if np.iscomplexobj(A):
        def runner(lo, hi):
            return _engine.perm_range(A.astype(np.complex128), method, lo, hi, device)
    else:
        def runner(lo, hi):
            return _engine.perm_f64_range(A.astype(np.float64), method, lo, hi, device)
    val = _engine.combine4(_engine.run_sharded(total, runner, group))
    if not ryser:
        val = val / total
    return val if np.iscomplexobj(A) else val.real


def perm_bbfg(M, group=None, device=None):
    """BBFG/Glynn permanent in Gray-code order (thewalrus/_permanent.py:130-168)."""
    if len(M) == 0:
        return M.dtype.type(1.0)
    return _run(M, False, group, device)


def perm_ryser(M, group=None, device=None):
    """Ryser permanent in Gray-code order (thewalrus/_permanent.py:86-127)."""
    if len(M) == 0:
        return M.dtype.type(1.0)
    return _run(M, True, group, device)


def perm(A, method="bbfg", *, group=None, device=None):
    """Permanent of a square matrix.  Same checks and closed forms as thewalrus/_permanent.py:34-83.

    ``method``: ``"bbfg"`` (default) or its synonym ``"glynn"`` run the BBFG/Glynn formula, ``"ryser"``
    the Ryser formula.  (The reference silently sends every string other than ``"bbfg"`` to Ryser,
    _permanent.py:81; here ``"glynn"`` means Glynn and unknown strings are rejected.)
    """
    if not isinstance(A, np.ndarray):
        raise TypeError("Input matrix must be a NumPy array.")
    matshape = A.shape
    if matshape[0] != matshape[1]:
        raise ValueError("Input matrix must be square.")
    if np.isnan(A).any():
        raise ValueError("Input matrix must not contain NaNs.")
    if matshape[0] == 0:
        return A.dtype.type(1.0)
    if matshape[0] == 1:
        return A[0, 0]
    if matshape[0] == 2:
        return A[0, 0] * A[1, 1] + A[0, 1] * A[1, 0]
    if matshape[0] == 3:
        return (A[0, 2] * A[1, 1] * A[2, 0] + A[0, 1] * A[1, 2] * A[2, 0] + A[0, 2] * A[1, 0] * A[2, 1]
                + A[0, 0] * A[1, 2] * A[2, 1] + A[0, 1] * A[1, 0] * A[2, 2] + A[0, 0] * A[1, 1] * A[2, 2])
    if method in ("bbfg", "glynn"):
        return perm_bbfg(A, group=group, device=device)
    if method == "ryser":
        return perm_ryser(A, group=group, device=device)
    raise ValueError("method must be 'bbfg', 'glynn' or 'ryser'")
