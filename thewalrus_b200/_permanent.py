"""Permanent front end — drop-in for thewalrus.perm (thewalrus/_permanent.py:34-83)."""
import numpy as np

from . import _engine

__all__ = ["perm", "perm_bbfg", "perm_ryser", "permanent_repeated", "brs", "ubrs", "fock_prob", "fock_threshold_prob"]


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


def permanent_repeated(A, rpt, *, group=None, device=None):
    """Permanent of ``A`` with row/column ``i`` repeated ``rpt[i]`` times (thewalrus/_permanent.py:171-195):
    ``perm(A_rpt) = haf([[0, A], [A^T, 0]]_(rpt, rpt))``, evaluated by the repeated-edge hafnian kernel."""
    from ._hafnian import hafnian_repeated

    A = np.asarray(A)
    n = A.shape[0]
    Z = np.zeros((n, n), dtype=A.dtype)
    B = np.vstack([np.hstack([Z, A]), np.hstack([A.T, Z])])
    return hafnian_repeated(B, list(rpt) * 2, loop=False, group=group, device=device)


# ---------------------------------------------------------------------------------------------------
# Bristolian (thewalrus/_permanent.py:198-321)
# ---------------------------------------------------------------------------------------------------
def _brs_sum(A, E, first, group, device):
    A = np.ascontiguousarray(A, dtype=np.complex128)
    m, n = A.shape
    if n == 0:                      # perm of a 0 x 0 matrix is 1 for every subset
        return complex(sum((-1) ** ((m - bin(j).count("1")) % 2) for j in range(first, 1 << m)))
    if m == 0:
        return complex(perm(np.asarray(E, dtype=np.complex128))) if first == 0 and E is not None else 0j
    En = None if E is None else np.ascontiguousarray(E, dtype=np.complex128)
    total = 1 << m
    rank, world = _engine._rank_world(group, None)
    from ._prep import shard_range

    lo, hi = shard_range(total - first, rank, world)
    lo, hi = lo + first, hi + first
    part = _engine.brs_range(A, En, lo, hi, device) if hi > lo else np.zeros(4)
    table = part.reshape(1, 4) if world == 1 else _engine.allreduce_partials(part, None if group is True else group)
    return _engine.combine4(table) / 2.0 ** (n - 1)       # perm_bbfg's final scale (:167)


def brs(A, E, *, group=None, device=None):
    """Bristolian of an m x n matrix A and an n x n matrix E (thewalrus/_permanent.py:198-223):
    ``sum_Y (-1)^(m-|Y|) perm(A_Y^H A_Y + E)`` over the row subsets Y of A."""
    A, E = np.asarray(A), np.asarray(E)
    if A.ndim != 2 or E.shape != (A.shape[1], A.shape[1]):
        raise ValueError("A must be m x n and E n x n")
    return _brs_sum(A, E, 0, group, device)


def ubrs(A, *, group=None, device=None):
    """Unitary Bristolian (thewalrus/_permanent.py:226-249): the same sum with E = 0, empty subset excluded."""
    A = np.asarray(A)
    if A.ndim != 2:
        raise ValueError("A must be a matrix")
    return _brs_sum(A, None, 1, group, device)


def _expand_modes(occ):
    return np.array([i for i, c in enumerate(occ) for _ in range(int(c))], dtype=int)


def fock_prob(n, m, U, *, group=None, device=None):
    """Probability of the Fock state ``n`` scattering to ``m`` through ``U`` (thewalrus/_permanent.py:252-279)."""
    if sum(n) != sum(m):
        raise ValueError("number of input photons must equal number of output photons")
    from math import factorial

    Umn = np.asarray(U)[np.ix_(_expand_modes(m), _expand_modes(n))]
    norm = float(np.prod([factorial(int(x)) for x in n])) * float(np.prod([factorial(int(x)) for x in m]))
    return abs(perm(np.ascontiguousarray(Umn), group=group, device=device)) ** 2 / norm


def fock_threshold_prob(n, d, T, *, group=None, device=None):
    """Probability that the Fock state ``n`` sent through ``T`` (M_out x M_in, M_out <= M_in) gives the threshold
    detector outcome ``d`` (thewalrus/_permanent.py:282-321)."""
    n, d, T = np.array(n), np.array(d), np.asarray(T)
    if len(n) != T.shape[1]:
        raise ValueError("length of n must matrix number of input modes of T")
    if len(d) != T.shape[0]:
        raise ValueError("length of d must match number of output modes of T")
    if T.shape[0] > T.shape[1]:
        raise ValueError("number of output modes cannot be larger than number of input modes")
    from math import factorial

    fac_prod = float(np.prod([factorial(int(x)) for x in n]))
    in_modes = _expand_modes(n)
    C = np.where(d > 0)[0]
    A = T[np.ix_(C, in_modes)]
    E = np.eye(T.shape[1]) - T.conj().T @ T
    if np.allclose(E, np.zeros((T.shape[1], T.shape[1]))):
        return ubrs(A, group=group, device=device).real / fac_prod
    return brs(A, E[np.ix_(in_modes, in_modes)], group=group, device=device).real / fac_prod
