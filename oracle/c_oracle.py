"""ctypes access to the C oracle.  TEST INFRASTRUCTURE ONLY (see oracle/walrus_oracle.c)."""
import ctypes

import numpy as np

from . import build

_dp = ctypes.POINTER(ctypes.c_double)
_libs = {}


def _lib(long_double=False):
    key = bool(long_double)
    if key not in _libs:
        so, so_ld = build.ensure()
        lib = ctypes.CDLL(so_ld if long_double else so)
        lib.oracle_hafnian_range.argtypes = [_dp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, _dp]
        lib.oracle_loop_hafnian_range.argtypes = [_dp, _dp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64,
                                                  ctypes.c_int, _dp]
        lib.oracle_perm_range.argtypes = [_dp, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64,
                                          ctypes.c_int, _dp]
        lib.oracle_tor_recursive.argtypes = [_dp, ctypes.c_int]
        lib.oracle_tor_recursive.restype = ctypes.c_double
        lib.oracle_tor_direct_range.argtypes = [_dp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int]
        lib.oracle_tor_direct_range.restype = ctypes.c_double
        lib.oracle_ltor_direct_range.argtypes = [_dp, _dp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, _dp]
        lib.oracle_brs_range.argtypes = [_dp, _dp, ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64,
                                         ctypes.c_int, _dp]
        lib.oracle_lhaf_patterns.argtypes = [_dp, _dp, ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.c_int64,
                                             ctypes.c_int, ctypes.c_int, _dp]
        lib.oracle_max_threads.restype = ctypes.c_int
        _libs[key] = lib
    return _libs[key]


def _c128(a):
    a = np.ascontiguousarray(a, dtype=np.complex128)
    return a, a.view(np.float64).ctypes.data_as(_dp)


def max_threads():
    return int(_lib().oracle_max_threads())


def matched_order(A):
    """Permutation the reference applies for all-ones reps: x = [n-1, n-3, ..., 1, n-2, ..., 0]
    (matched_reps, thewalrus/_hafnian.py:80-159, evaluated for reps = [1]*n)."""
    n = A.shape[0]
    x = list(range(n - 1, -1, -2)) + list(range(n - 2, -1, -2))
    return np.asarray(x)


def hafnian_range(Ax, j0, j1, D=None, threads=0, long_double=False):
    """Unscaled partial Glynn sum over [j0, j1) of the MATCHED-order matrix Ax (and Dx)."""
    lib = _lib(long_double)
    Ax, pA = _c128(Ax)
    out = np.zeros(4)
    po = out.ctypes.data_as(_dp)
    if D is None:
        lib.oracle_hafnian_range(pA, Ax.shape[0], j0, j1, threads, po)
    else:
        D, pD = _c128(D)
        lib.oracle_loop_hafnian_range(pA, pD, Ax.shape[0], j0, j1, threads, po)
    return complex(out[0] + out[2], out[1] + out[3])


def hafnian(A, loop=False, threads=0, long_double=False):
    """Full (loop) hafnian of an even-dimensional symmetric matrix, reps all 1."""
    n = A.shape[0]
    x = matched_order(A)
    Ax = A[np.ix_(x, x)]
    D = np.diag(A)[x] if loop else None
    return hafnian_range(Ax, 0, 1 << (n // 2 - 1), D, threads, long_double) * 0.5 ** (n // 2 - 1)


def perm_range(M, method, k0, k1, threads=0, long_double=False):
    lib = _lib(long_double)
    M, pM = _c128(M)
    out = np.zeros(4)
    lib.oracle_perm_range(pM, M.shape[0], method, k0, k1, threads, out.ctypes.data_as(_dp))
    return complex(out[0] + out[2], out[1] + out[3])


def perm(M, method="bbfg", threads=0, long_double=False):
    n = M.shape[0]
    if method in ("bbfg", "glynn"):
        return perm_range(M, 0, 0, 1 << (n - 1), threads, long_double) / (1 << (n - 1))
    return perm_range(M, 1, 0, 1 << n, threads, long_double)


def tor_recursive(O, long_double=False):
    O, pO = _c128(O)
    return float(_lib(long_double).oracle_tor_recursive(pO, O.shape[0] // 2))


def tor_direct(O, j0=0, j1=None, threads=0, long_double=False):
    O, pO = _c128(O)
    N = O.shape[0] // 2
    return float(_lib(long_double).oracle_tor_direct_range(pO, N, j0, (1 << N) if j1 is None else j1, threads))


def ltor_direct(O, gamma, j0=0, j1=None, threads=0, long_double=False):
    """Loop torontonian over subset indices [j0, j1) (numba_ltor, thewalrus/_torontonian.py:369-412)."""
    O, pO = _c128(O)
    gamma, pg = _c128(gamma)
    N = O.shape[0] // 2
    out = np.zeros(2)
    _lib(long_double).oracle_ltor_direct_range(pO, pg, N, j0, (1 << N) if j1 is None else j1, threads,
                                               out.ctypes.data_as(_dp))
    return complex(out[0], out[1])


def brs(A, E=None, j0=0, j1=None, threads=0, long_double=False):
    """Bristolian over row-subset labels [j0, j1) (brs / ubrs, thewalrus/_permanent.py:198-249); ``E=None`` and
    ``j0=1`` give ubrs."""
    A, pA = _c128(A)
    pE = None
    if E is not None:
        E, pE = _c128(E)
    m, n = A.shape
    out = np.zeros(2)
    _lib(long_double).oracle_brs_range(pA, pE, m, n, j0, (1 << m) if j1 is None else j1, threads,
                                       out.ctypes.data_as(_dp))
    return complex(out[0], out[1])


def lhaf_patterns(A, gamma, rpt, glynn=True, threads=0, long_double=False):
    """``[loop_hafnian(A, gamma, reps=r) for r in rpt]`` (``gamma=None``: ``hafnian_repeated``), patterns in parallel
    (thewalrus/_hafnian.py:581-631 per pattern, as quantum/fock_tensors.py:191-232 calls it)."""
    A, pA = _c128(A)
    pG = None
    if gamma is not None:
        gamma, pG = _c128(gamma)
    rpt = np.ascontiguousarray(rpt, dtype=np.int32)
    out = np.zeros(len(rpt), dtype=np.complex128)
    _lib(long_double).oracle_lhaf_patterns(pA, pG, A.shape[0], rpt.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                           len(rpt), 1 if glynn else 0, threads, out.view(np.float64).ctypes.data_as(_dp))
    return out
