"""Hafnian / loop hafnian front end — drop-in for thewalrus._hafnian (public names and semantics).

Validation, early exits and the small-n closed forms follow thewalrus/_hafnian.py:634-665 and :750-818;
the exponential subset sum itself (thewalrus/_hafnian.py:416-577) runs on the GPU through the C ABI.
"""
import warnings

import numpy as np

from . import _engine
from ._prep import glynn_steps, matched_reps

__all__ = ["hafnian", "loop_hafnian", "hafnian_repeated", "hafnian_batch", "reduction", "input_validation", "_haf",
           "matched_reps", "find_kept_edges", "recursive_hafnian"]

_DMMA_MAX_N = 64


def _allclose(a, b, rtol, atol):
    """``np.allclose(a, b, rtol, atol)`` for arrays already known to be free of NaNs: the same predicate
    |a - b| <= atol + rtol |b| without allclose's per-call bookkeeping (40-50 us on a 24 x 24 matrix, a quarter of a
    whole n = 24 GPU call).  Infinite entries take the NumPy path."""
    if not np.isfinite(a).all() or not np.isfinite(b).all():
        return bool(np.allclose(a, b, rtol=rtol, atol=atol))
    return bool((np.abs(a - b) <= atol + rtol * np.abs(b)).all())


def input_validation(A, rtol=1e-05, atol=1e-08):
    """Same checks and messages as thewalrus/_hafnian.py:634-665."""
    if not isinstance(A, np.ndarray):
        raise TypeError("Input matrix must be a NumPy array.")
    n = A.shape
    if n[0] != n[1]:
        raise ValueError("Input matrix must be square.")
    if np.isnan(A).any():
        raise ValueError("Input matrix must not contain NaNs.")
    if not _allclose(A, A.T, rtol, atol):
        raise ValueError("Input matrix must be symmetric.")
    return True


def _ones_matching(n):
    """``matched_reps([1] * n)`` in closed form: (x, edge_reps, oddmode)."""
    first = np.arange(n - 1, 0, -2, dtype=np.int64)
    return np.concatenate([first, first - 1]), np.ones(len(first), dtype=np.int64), (0 if n % 2 else None)


def reduction(A, rpt):
    """Repeat row/column ``i`` of ``A`` ``rpt[i]`` times (thewalrus/_hafnian.py:697-714)."""
    rows = [i for i, r in enumerate(rpt) for _ in range(int(r))]
    if A.ndim == 1:
        return A[rows]
    return A[:, rows][rows]


def find_kept_edges(j, reps):
    """Mixed-radix digits of ``j`` in bases ``reps + 1``, most significant first
    (thewalrus/_hafnian.py:162-180)."""
    reps = np.asarray(reps)
    out = np.zeros(len(reps), dtype=reps.dtype)
    num = int(j)
    for i in range(len(reps) - 1, -1, -1):
        base = int(reps[i]) + 1
        out[i] = num % base
        num //= base
    return out


def recursive_hafnian(A):
    """Hafnian by the recursive algorithm of Bjorklund, Gupt and Quesada (arXiv:1805.12498, Algorithm 2), in the
    arithmetic of ``A.dtype`` — exact for integer matrices, which is what the reference's ``hafnian(A,
    method="recursive")`` returns for them (thewalrus/_hafnian.py:815-816, 978-1045).  Host code: an exact-integer
    convenience beside the GPU path (2^(n/2) contraction steps of small polynomial arrays; the floating-point
    subset sums run on the GPU).

    The remaining vertices carry a symmetric matrix of polynomials in one variable, truncated at degree n/2.  Each level
    removes vertices 0 and 1: either without the edge (0, 1) (sign flips), or contracting them, which multiplies the
    running polynomial g by (1 + x P[0,1]) and adds x (P[j,0] P[k,1] + P[k,0] P[j,1]) to every remaining pair (j, k).
    """
    A = np.asarray(A)
    nb = A.shape[0]
    if nb == 0:
        return A.dtype.type(1)
    if nb % 2:
        return A.dtype.type(0)
    n = nb // 2
    P = np.zeros((nb, nb, n + 1), dtype=A.dtype)
    P[:, :, 0] = A
    g = np.zeros(n + 1, dtype=A.dtype)
    g[0] = 1

    def conv_shift(a, b):
        """x * a * b truncated at degree n, broadcasting over leading axes."""
        out = np.zeros(np.broadcast_shapes(a.shape, b.shape), dtype=A.dtype)
        for u in range(n):
            out[..., u + 1:] += a[..., u:u + 1] * b[..., :n - u]
        return out

    def solve(P, w, g):
        s = P.shape[0]
        if s == 0:
            return w * g[n]
        rest = P[2:, 2:]
        h = solve(rest, -w, g)
        e = g + conv_shift(g, P[0, 1])
        c0, c1 = P[2:, 0], P[2:, 1]                       # polynomials of the pairs (j, 0) and (j, 1)
        t = conv_shift(c0[:, None, :], c1[None, :, :])
        return h + solve(rest + t + t.transpose(1, 0, 2), w, e)

    return solve(P, A.dtype.type(1), g)


def _all_ones(edge_reps):
    return len(edge_reps) > 0 and bool(np.all(np.asarray(edge_reps) == 1))


def _subset_sum(Ax, Dx, edge_reps, oddloop, oddV, glynn, group, device):
    """Run the subset sum on the GPU (sharded over ``group`` if given) and apply the final scale
    (thewalrus/_hafnian.py:464-465, 571-575)."""
    n = Ax.shape[0]
    has_odd = oddloop is not None
    steps = glynn_steps(edge_reps, glynn, has_odd)
    fast = glynn and not has_odd and _all_ones(edge_reps) and n <= _DMMA_MAX_N

    if fast:
        def runner(lo, hi):
            return _engine.hafnian_range(Ax, Dx, lo, hi, device)
    else:
        def runner(lo, hi):
            return _engine.lhaf_general_range(Ax, Dx, oddV, oddloop, edge_reps, glynn, lo, hi, device)

    H = _engine.combine4(_engine.run_sharded(steps, runner, group))
    if glynn:
        N = 2 * int(np.sum(edge_reps)) + (1 if has_odd else 0)
        H *= 0.5 ** (N // 2 if has_odd else N // 2 - 1)
    return H


def _haf(A, reps=None, glynn=True, group=None, device=None):
    """Hafnian with optional repeated rows/columns (thewalrus/_hafnian.py:470-508)."""
    n = A.shape[0]
    ones = reps is None
    if ones:
        reps = [1] * n
    N = sum(reps)
    if N == 0:
        return 1.0
    if N % 2 == 1:
        return 0.0
    assert n == len(reps)
    x, edge_reps, _ = _ones_matching(n) if ones else matched_reps(reps)
    Ax = A[np.ix_(x, x)].astype(np.complex128)
    return _subset_sum(Ax, None, edge_reps, None, None, glynn, group, device)


def loop_hafnian(A, D=None, reps=None, glynn=True, group=None, device=None):
    """Loop hafnian with optional repeated rows/columns (thewalrus/_hafnian.py:581-631)."""
    n = A.shape[0]
    ones = reps is None
    if ones:
        reps = [1] * n
    if D is None:
        D = A.diagonal()
    N = sum(reps)
    if N == 0:
        return 1.0
    if N == 1:
        return D[np.where(np.array(reps) == 1)[0][0]]
    assert n == len(reps)
    assert D.shape[0] == n
    x, edge_reps, oddmode = _ones_matching(n) if (ones and n > 1) else matched_reps(reps)
    if oddmode is not None:
        oddloop = np.complex128(D[oddmode])
        oddV = A[oddmode, x].astype(np.complex128)
    else:
        oddloop = None
        oddV = None
    Ax = A[np.ix_(x, x)].astype(np.complex128)
    Dx = np.asarray(D)[x].astype(np.complex128)
    return _subset_sum(Ax, Dx, edge_reps, oddloop, oddV, glynn, group, device)


def hafnian(A, loop=False, rtol=1e-05, atol=1e-08, approx=False, num_samples=1000, method="glynn", *,
            group=None, device=None):  # pylint: disable=too-many-arguments,too-many-return-statements
    """Hafnian of a symmetric matrix; same signature and early exits as thewalrus.hafnian
    (thewalrus/_hafnian.py:718-818).

    Extra keyword-only arguments: ``group`` (``True`` or a ``torch.distributed`` process group: shard
    the subset index over its ranks and combine with one all-reduce) and ``device``.
    ``method="recursive"`` (a different, non-subset-sum algorithm in the reference) returns the exact integer count
    for integer matrices (host recursion, :func:`recursive_hafnian`) and is evaluated with the Glynn kernel, which
    returns the same value, for floating-point ones.  ``approx=True`` (Barvinok sampling) is outside the scope of
    this package.
    """
    input_validation(A, rtol=rtol, atol=atol)
    matshape = A.shape
    if method not in ("glynn", "inclexcl", "recursive"):
        raise ValueError("method must be 'glynn', 'inclexcl' or 'recursive'")
    glynn = method != "inclexcl"

    if matshape == (0, 0):
        return 1
    if matshape[0] % 2 != 0 and not loop:
        return 0.0
    if _allclose(np.diag(np.diag(A)), A, rtol, atol):
        if loop:
            return np.prod(np.diag(A))
        return 0
    if matshape[0] == 2:
        if loop:
            return A[0, 1] + A[0, 0] * A[1, 1]
        return A[0][1]
    if matshape[0] == 3 and loop:
        return A[0, 0] * A[1, 2] + A[1, 1] * A[0, 2] + A[2, 2] * A[0, 1] + A[0, 0] * A[1, 1] * A[2, 2]
    if matshape[0] == 4:
        if loop:
            return (A[0, 1] * A[2, 3] + A[0, 2] * A[1, 3] + A[0, 3] * A[1, 2]
                    + A[0, 0] * A[1, 1] * A[2, 3] + A[0, 1] * A[2, 2] * A[3, 3] + A[0, 2] * A[1, 1] * A[3, 3]
                    + A[0, 0] * A[2, 2] * A[1, 3] + A[0, 0] * A[3, 3] * A[1, 2] + A[0, 3] * A[1, 1] * A[2, 2]
                    + A[0, 0] * A[1, 1] * A[2, 2] * A[3, 3])
        return A[0, 1] * A[2, 3] + A[0, 2] * A[1, 3] + A[0, 3] * A[1, 2]

    if approx:
        if np.any(np.iscomplex(A)):
            raise ValueError("Input matrix must be real")
        if np.any(A < 0):
            raise ValueError("Input matrix must not have negative entries")
        raise NotImplementedError("hafnian_approx (Barvinok sampling) is outside the B200 hot-path scope")

    if loop:
        if method == "recursive":
            warnings.warn("Recursive algorithm does not support the loop hafnian")
        return loop_hafnian(A, D=None, reps=None, glynn=True, group=group, device=device)
    if method == "recursive" and np.issubdtype(A.dtype, np.integer):
        return recursive_hafnian(A)           # exact integer count, as the reference returns (:815-816)
    return _haf(A, reps=None, glynn=glynn, group=group, device=device)


def hafnian_repeated(A, rpt, mu=None, loop=False, rtol=1e-05, atol=1e-08, glynn=True, *, group=None,
                     device=None):  # pylint: disable=too-many-arguments,too-many-return-statements
    """Hafnian with repeated rows/columns; same checks as thewalrus/_hafnian.py:864-936."""
    input_validation(A, atol=atol, rtol=rtol)
    if len(rpt) != len(A):
        raise ValueError("the rpt argument must be 1-dimensional sequence of length len(A).")
    nud = np.array(rpt, dtype=np.int32)
    if not np.all(np.mod(rpt, 1) == 0) or np.any(nud < 0):
        raise ValueError("the rpt argument must contain non-negative integers.")
    if np.all(nud == 0):
        return 1.0
    if np.sum(nud) % 2 != 0 and not loop:
        return 0.0
    if mu is None:
        mu = A.diagonal().copy()
    if np.allclose(A, 0, rtol=rtol, atol=atol):
        if loop:
            return np.prod(mu**rpt)
        return 0
    if len(mu) != len(A):
        raise ValueError("Length of means vector must be the same length as the matrix A.")
    rpt = [int(r) for r in rpt]
    if loop:
        return loop_hafnian(A, D=np.asarray(mu), reps=rpt, glynn=glynn, group=group, device=device)
    return _haf(A, reps=rpt, glynn=glynn, group=group, device=device)


def hafnian_batch(As, loop=False, rtol=1e-05, atol=1e-08, *, group=None, device=None):
    """Hafnians (``loop=True``: loop hafnians, loops on the diagonals) of a stack of matrices ``As[B, n, n]`` ->
    ``complex128[B]``: what ``[hafnian(A, loop=loop) for A in As]`` returns (thewalrus/_hafnian.py:718-861), in
    ONE GPU call of the batched-matrix front end (``wb200_lhaf_matrices_host``: one warp per (matrix, subset)).
    With ``group`` the matrices are sharded over the ranks in contiguous blocks and the results all-gathered.
    Meant for many small matrices; beyond n = 28 the matrices go one by one through the DMMA kernel."""
    if not isinstance(As, np.ndarray):
        raise TypeError("Input matrix must be a NumPy array.")
    if As.ndim != 3 or As.shape[1] != As.shape[2]:
        raise ValueError("Input matrix must be square.")
    if np.isnan(As).any():
        raise ValueError("Input matrix must not contain NaNs.")
    if not np.allclose(As, np.swapaxes(As, 1, 2), rtol=rtol, atol=atol):
        raise ValueError("Input matrix must be symmetric.")
    B, n = As.shape[0], As.shape[1]
    if B == 0:
        return np.zeros(0, dtype=np.complex128)
    if n == 0:
        return np.ones(B, dtype=np.complex128)
    if n % 2 == 1 and not loop:
        return np.zeros(B, dtype=np.complex128)
    if n > 28:
        return np.array([complex(hafnian(np.ascontiguousarray(A), loop=loop, group=group, device=device)) for A in As])
    from .quantum import lhaf_patterns

    Ac = np.ascontiguousarray(As, dtype=np.complex128)
    idx = np.arange(B, dtype=np.int32)
    gamma = np.ascontiguousarray(np.diagonal(Ac, axis1=1, axis2=2)) if loop else None
    return lhaf_patterns(Ac, gamma, np.ones((B, n), dtype=np.int32), A_index=idx,
                         gamma_index=idx if loop else None, group=group, device=device)
