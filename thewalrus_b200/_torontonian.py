"""Torontonian front ends — drop-ins for thewalrus.tor, ltor and threshold_detection_prob
(thewalrus/_torontonian.py:23-120)."""
import numpy as np

from . import _engine
from ._prep import dd_sum

__all__ = ["tor", "ltor", "threshold_detection_prob", "numba_vac_prob", "tor_input_checks", "numba_tor", "rec_torontonian",
           "numba_ltor", "rec_ltorontonian"]


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


def ltor(A, gamma, recursive=True, *, group=None, device=None):
    """Loop torontonian  sum_S (-1)^(N-|S|) exp(gamma_S (I - A_S)^-1 gamma_S^* / 2) / sqrt(det(I - A_S))
    (thewalrus/_torontonian.py:61-74; rec_ltorontonian :320-345, numba_ltor :369-412).

    Like ``rec_ltorontonian`` (the reference default) the exponent is evaluated as the squared norm of the
    forward-substituted vector, so the value is real; it is returned as ``np.complex128`` as the reference does.
    ``recursive`` is accepted for signature compatibility (both reference variants compute the same number).
    """
    tor_input_checks(A, gamma)
    del recursive
    N = A.shape[0] // 2
    if N == 0:
        return np.complex128(1.0)
    A = np.asarray(A, dtype=np.complex128)
    gamma = np.asarray(gamma, dtype=np.complex128)
    if N == 1:
        # two subsets: {} -> -1, {0} -> exp(x^H B^-1 x / 2) / sqrt(det B), B = I - A, x = conj(gamma)
        B = np.identity(2) - A
        d1 = B[0, 0].real
        l10 = B[1, 0] / d1
        d2 = (B[1, 1] - l10 * B[0, 1]).real
        x = gamma.conj()
        z1 = x[1] - l10 * x[0]
        q = abs(x[0]) ** 2 / d1 + abs(z1) ** 2 / d2
        return np.complex128(-1.0 + np.exp(0.5 * q) / np.sqrt(d1 * d2))
    total = _engine.tor_num_prefixes(N)
    table = _engine.run_sharded(total, lambda lo, hi: _engine.tor_range(A, lo, hi, device, gamma=gamma), group,
                                width=2)
    hi, lo = dd_sum([(p[0], p[1]) for p in table])
    return np.complex128(hi + lo)


def numba_vac_prob(alpha, sigma):
    """Vacuum probability  exp(-alpha^H sigma^-1 alpha / 2) / sqrt(det sigma)  (thewalrus/_torontonian.py:348-366)."""
    alpha = np.asarray(alpha, dtype=np.complex128)
    sigma = np.asarray(sigma, dtype=np.complex128)
    return (np.exp(-0.5 * alpha.conj() @ np.linalg.solve(sigma, alpha)).real / np.sqrt(np.linalg.det(sigma))).real


def threshold_detection_prob(mu, cov, det_pattern, hbar=2, atol=1e-10, rtol=1e-10, *, group=None, device=None):
    """Probability of the click pattern ``det_pattern`` for the Gaussian state (mu, cov)
    (thewalrus/_torontonian.py:77-120): ``tor`` of the clicked block of ``O = I - Q^-1`` for zero-mean states,
    ``ltor`` with ``gamma = (sigma^-1 alpha)^*`` times the vacuum probability for displaced ones."""
    from .quantum import Qmat

    mu, cov = np.asarray(mu), np.asarray(cov)
    n = cov.shape[0] // 2
    clicked = np.where(np.asarray(det_pattern) == 1)[0]
    rows = np.concatenate([clicked, clicked + n])
    if np.allclose(mu, 0, atol=atol, rtol=rtol):
        Q = Qmat(cov, hbar)
        O = np.identity(2 * n) - np.linalg.inv(Q)
        Os = np.ascontiguousarray(O[np.ix_(rows, rows)])
        return tor(Os, group=group, device=device) / np.sqrt(np.linalg.det(Q))
    alpha = np.concatenate((mu[:n] + 1j * mu[n:], mu[:n] - 1j * mu[n:])) / np.sqrt(2 * hbar)
    sigma = Qmat(cov, hbar=hbar).conj()
    inv_sigma = np.linalg.inv(sigma)
    O = np.identity(2 * n) - inv_sigma
    gamma = (inv_sigma @ alpha).conj()
    O_red = np.ascontiguousarray(O[np.ix_(rows, rows)])
    return numba_vac_prob(alpha, sigma) * ltor(O_red, gamma[rows], group=group, device=device).real


# The reference exports its four numba drivers as well (thewalrus/__init__.py:126-134); callers that use them
# directly (threshold_detection_prob does, _torontonian.py:120) get the same GPU evaluation.
def numba_tor(O, **kw):
    """thewalrus/_torontonian.py:123-154."""
    return tor(np.asarray(O), recursive=False, **kw)


def rec_torontonian(A, **kw):
    """thewalrus/_torontonian.py:227-247."""
    return tor(np.asarray(A), recursive=True, **kw)


def numba_ltor(O, gamma, **kw):
    """thewalrus/_torontonian.py:369-412."""
    return ltor(np.asarray(O), np.asarray(gamma), recursive=False, **kw)


def rec_ltorontonian(A, gamma, **kw):
    """thewalrus/_torontonian.py:320-345."""
    return ltor(np.asarray(A), np.asarray(gamma), recursive=True, **kw)
