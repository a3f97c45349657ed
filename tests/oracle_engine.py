"""TEST INFRASTRUCTURE: stand-ins for the `_engine.*_range` kernel runners built on the CPU oracle.

With ``install(monkeypatch)`` the package's host logic (validation, early exits, matching, scaling, dtype handling)
can be exercised on a box without a GPU: every kernel call is answered by `oracle/` instead of
libwalrus_b200.so.  Only tests use this (the CPU twins of the reference test-suite run, tests/test_ref_suite_cpu.py);
the product never imports it.
"""
import numpy as np

from oracle import c_oracle as co
from oracle import walrus_oracle as wo


def _c4(z):
    z = complex(z)
    return np.array([z.real, 0.0, z.imag, 0.0])


def hafnian_range(Ax, Dx, j0, j1, device=None):
    return _c4(co.hafnian_range(Ax, j0, j1, Dx))


def lhaf_general_range(Ax, Dx, oddV, oddloop, edge_reps, glynn, j0, j1, device=None):
    er = np.asarray(edge_reps)
    if Dx is None:
        return _c4(wo.calc_hafnian(Ax, er, glynn, j0, j1, scale=False))
    return _c4(wo.calc_loop_hafnian(Ax, Dx, er, oddloop, oddV, glynn, j0, j1, scale=False))


def perm_range(M, method, k0, k1, device=None):
    return _c4(co.perm_range(M, method, k0, k1))


def perm_f64_range(M, method, k0, k1, device=None):
    return _c4(co.perm_range(np.asarray(M, dtype=np.complex128), method, k0, k1).real)


def perm_int64_range(M, method, k0, k1, device=None):
    M = np.asarray(M, dtype=np.int64)
    v = wo.perm_ryser(M, k0, k1) if method else wo.perm_bbfg(M, k0, k1, scale=False)
    return int(v)


def tor_range(O, p0, p1, device=None, gamma=None):
    """The kernel's range unit is a prefix; the stand-in only answers whole-problem calls."""
    from thewalrus_b200 import _engine

    assert p0 == 0 and p1 == _engine.tor_num_prefixes(O.shape[0] // 2)
    if gamma is None:
        return np.array([co.tor_recursive(np.asarray(O, dtype=np.complex128)), 0.0])
    return np.array([co.ltor_direct(O, gamma).real, 0.0])


def mtl_range(A, zeta, p0, p1, device=None):
    """V, W partial sums of thewalrus_b200/_montrealer.py (montrealer, _montrealer.py:37-102 of the reference)."""
    A = np.asarray(A, dtype=np.complex128)
    n = len(A) // 2
    X = np.block([[np.zeros((n, n)), np.eye(n)], [np.eye(n), np.zeros((n, n))]])
    S = X @ A
    V = W = 0j
    for p in range(max(p0, 1), p1):
        modes = [i for i in range(n) if (p >> (n - 1 - i)) & 1]
        pos = modes + [i + n for i in modes]
        sub = S[np.ix_(pos, pos)]
        sign = (-1) ** (len(modes) + 1)
        V += sign * np.trace(np.linalg.matrix_power(sub, n))
        if zeta is not None:
            z = np.asarray(zeta, dtype=np.complex128)[pos]
            W += sign * (z.conj() @ np.linalg.matrix_power(sub, n - 1) @ z)
    return np.concatenate([_c4(V), _c4(W)])


def brs_range(A, E, j0, j1, device=None):
    n = A.shape[1]
    return _c4(co.brs(A, E, j0, j1) * 2.0 ** (n - 1))


def lhaf_patterns_local(A, gamma, rpt, glynn=True, device=None, want_ms=False, gamma_index=None, A_index=None):
    A = np.asarray(A, dtype=np.complex128)
    rpt = np.asarray(rpt, dtype=np.int32)
    out = np.zeros(len(rpt), dtype=np.complex128)
    for i, r in enumerate(rpt):
        Ai = A if A.ndim == 2 else A[0 if A_index is None else A_index[i]]
        g = None if gamma is None else (np.atleast_2d(gamma)[0 if gamma_index is None else gamma_index[i]])
        out[i] = co.lhaf_patterns(Ai, g, r[None, :], glynn)[0]
    return (out, 0.0) if want_ms else out


def install(monkeypatch):
    from thewalrus_b200 import _engine

    for name in ("hafnian_range", "lhaf_general_range", "perm_range", "perm_f64_range", "perm_int64_range", "tor_range",
                 "mtl_range", "brs_range", "lhaf_patterns_local"):
        monkeypatch.setattr(_engine, name, globals()[name])
