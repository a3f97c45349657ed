"""GBS photon-number probabilities — the batched caller of the loop-hafnian path.

Drop-in for ``thewalrus.quantum.density_matrix_element`` (thewalrus/quantum/fock_tensors.py:191-232) and
``probabilities`` (:392-430) for mixed states, plus ``probabilities_batch``: all patterns of one Gaussian state
in ONE GPU call (``wb200_lhaf_patterns_host``) instead of one Python call per pattern.  The host side is the
O(M^3) state preparation: ``Qmat``/``Amat`` (thewalrus/quantum/conversions.py:70-150), gamma and the
prefactor (fock_tensors.py:566-581), written from the formulas in docs/gbs.rst.
"""
from itertools import product
from math import lgamma

import numpy as np

from . import _engine

__all__ = ["Qmat", "Amat", "complex_to_real_displacements", "density_matrix_element", "probabilities",
           "probabilities_batch", "lhaf_patterns"]


def Qmat(cov, hbar=2):
    """Husimi covariance  Q = sigma_complex + I/2  in the (a, a^dagger) basis (conversions.py:70-96)."""
    cov = np.asarray(cov)
    N = len(cov) // 2
    x, xp, p = (cov[:N, :N] * 2 / hbar, cov[:N, N:] * 2 / hbar, cov[N:, N:] * 2 / hbar)
    ada = (x + p + 1j * (xp - xp.T) - 2 * np.identity(N)) / 4   # <a_i^dagger a_j>
    aa = (x - p + 1j * (xp + xp.T)) / 4                          # <a_i a_j>
    return np.block([[ada, aa.conj()], [aa, ada.conj()]]) + np.identity(2 * N)


def Amat(cov, hbar=2, cov_is_qmat=False):
    """A = X (I - Q^-1)^*  (conversions.py:124-150)."""
    N = len(cov) // 2
    Q = np.asarray(cov) if cov_is_qmat else Qmat(cov, hbar=hbar)
    B = (np.identity(2 * N) - np.linalg.inv(Q)).conj()
    return np.vstack([B[N:], B[:N]])


def complex_to_real_displacements(mu, hbar=2):
    """(alpha, alpha^*) from the xp means (conversions.py:153-169)."""
    mu = np.asarray(mu)
    N = len(mu) // 2
    alpha = (mu[:N] + 1j * mu[N:]) / np.sqrt(2 * hbar)
    return np.concatenate([alpha, alpha.conj()])


def _prefactor(mu, cov, hbar=2):
    """exp(-beta Q^-1 beta^* / 2) / sqrt(det Q)  (fock_tensors.py:566-581)."""
    Q = Qmat(cov, hbar=hbar)
    beta = complex_to_real_displacements(mu, hbar=hbar)
    return np.exp(-0.5 * beta @ np.linalg.inv(Q) @ beta.conj()) / np.sqrt(np.linalg.det(Q))


def lhaf_patterns(A, gamma, rpt, glynn=True, *, group=None, device=None):
    """``[loop_hafnian(A, gamma, reps=r) for r in rpt]`` (``gamma=None``: ``hafnian_repeated(A, r)``) on the GPU.

    ``rpt``: integer array ``[B, len(A)]``.  With ``group`` the patterns are sharded over the ranks in
    contiguous blocks and the results all-gathered (one collective).
    """
    rpt = np.ascontiguousarray(rpt, dtype=np.int32)
    if rpt.ndim != 2 or rpt.shape[1] != len(A):
        raise ValueError("rpt must have shape [batch, len(A)]")
    return _engine.run_sharded_patterns(A, gamma, rpt, glynn, group, device)


def _state(mu, cov, hbar, tol):
    A = Amat(cov, hbar=hbar)
    beta = complex_to_real_displacements(mu, hbar=hbar)
    gamma = None if np.linalg.norm(beta) < tol else beta.conj() - A @ beta
    return A, gamma


def density_matrix_element(mu, cov, i, j, include_prefactor=True, tol=1e-10, hbar=2, *, device=None):
    """<i| rho |j> of a Gaussian state (fock_tensors.py:191-232)."""
    rpt = list(i) + list(j)
    A, gamma = _state(mu, cov, hbar, tol)
    haf = complex(lhaf_patterns(A, gamma, np.array([rpt]), device=device)[0])
    if include_prefactor:
        haf *= _prefactor(mu, cov, hbar=hbar)
    return haf / np.sqrt(np.prod([np.exp(lgamma(k + 1)) for k in rpt]))


def probabilities_batch(mu, cov, patterns, hbar=2, tol=1e-10, *, group=None, device=None):
    """Probabilities p(n) = <n| rho |n> of many photon-number patterns of one Gaussian state.

    ``patterns``: integer array ``[B, M]``.  Equals
    ``[density_matrix_element(mu, cov, n, n).real for n in patterns]`` of the reference, clipped at 0 as
    ``probabilities`` does (fock_tensors.py:424-428).
    """
    patterns = np.ascontiguousarray(patterns, dtype=np.int32)
    A, gamma = _state(mu, cov, hbar, tol)
    if patterns.ndim != 2 or 2 * patterns.shape[1] != len(A):
        raise ValueError("patterns must have shape [batch, n_modes]")
    rpt = np.concatenate([patterns, patterns], axis=1)
    lh = lhaf_patterns(A, gamma, rpt, group=group, device=device)
    pref = _prefactor(mu, cov, hbar=hbar)
    from scipy.special import gammaln

    logfac = gammaln(patterns + 1.0).sum(axis=1)       # sqrt(prod rpt!) = prod n!
    return np.maximum(0.0, (lh * pref).real * np.exp(-logfac))


def probabilities(mu, cov, cutoff, parallel=False, hbar=2.0, rtol=1e-05, atol=1e-08, *, group=None, device=None):
    """Fock probabilities up to ``cutoff`` per mode, shape ``[cutoff] * n_modes`` (fock_tensors.py:392-430).
    Every pattern goes through the batched GPU path (the reference's pure-state shortcut through
    ``state_vector`` returns the same numbers)."""
    del parallel, rtol, atol
    M = len(mu) // 2
    pats = np.array(list(product(range(cutoff), repeat=M)), dtype=np.int32).reshape(-1, M)
    return probabilities_batch(mu, cov, pats, hbar=hbar, group=group, device=device).reshape([cutoff] * M)
