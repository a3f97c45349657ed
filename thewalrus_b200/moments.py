"""Photon-number and click statistics of Gaussian states — callers of the hot path
(thewalrus/quantum/means_and_variances.py:31-457).

The s-ordered expectation values are loop hafnians of ``V = (Q - (s + 1)/2 I) X`` with repeated rows
(:212-249); a photon-number moment is a weighted sum of many of them (:329-367).  The reference evaluates them one
``hafnian`` call at a time; here all terms of a moment go through ONE call of the batched front end
(``quantum.lhaf_patterns``: repetition patterns of one matrix).  Click cumulants are sums of threshold
probabilities (torontonian / loop torontonian kernel).  First and second moments have closed forms and stay on
the host.
"""
from itertools import product
from math import factorial

import numpy as np

from . import quantum as _q

__all__ = ["reduced_gaussian", "photon_number_mean", "photon_number_mean_vector", "photon_number_covar",
           "photon_number_covmat", "s_ordered_expectation", "normal_ordered_expectation",
           "photon_number_expectation", "photon_number_squared_expectation", "photon_number_moment",
           "photon_number_cumulant", "click_cumulant", "mean_clicks", "variance_clicks", "partition"]


def reduced_gaussian(mu, cov, modes):
    """Means and covariance of the listed modes (quantum/conversions.py:22-51)."""
    mu, cov = np.asarray(mu), np.asarray(cov)
    N = len(mu) // 2
    modes = [modes] if isinstance(modes, (int, np.integer)) else list(modes)
    if np.any(np.array(modes) > N):
        raise ValueError("Provided mode is larger than the number of subsystems.")
    if len(modes) == N:
        return mu, cov
    ind = np.concatenate([np.array(modes, dtype=int), np.array(modes, dtype=int) + N])
    return mu[ind], cov[np.ix_(ind, ind)]


def photon_number_mean(mu, cov, j, hbar=2):
    """<n_j> (means_and_variances.py:31-48)."""
    return _q.photon_number_mean_vector(mu, cov, hbar=hbar)[j]


photon_number_mean_vector = _q.photon_number_mean_vector


def photon_number_covar(mu, cov, j, k, hbar=2):
    """Cov(n_j, n_k) from the closed forms of Dodonov et al. (means_and_variances.py:69-123)."""
    mu, cov = np.asarray(mu), np.asarray(cov)
    N = len(mu) // 2
    if j == k:
        idx = [j, j + N]
        c, m = cov[np.ix_(idx, idx)], mu[idx]
        return ((0.5 * np.trace(c) ** 2 - np.linalg.det(c)) + m @ c @ m) / hbar**2 - 0.25
    quad_j, quad_k = (j, j + N), (k, k + N)
    sq = sum(cov[a, b] ** 2 for a in quad_j for b in quad_k)
    cross = sum(cov[a, b] * mu[a] * mu[b] for a in quad_j for b in quad_k)
    return (sq + 2 * cross) / (2 * hbar**2)


def photon_number_covmat(mu, cov, hbar=2):
    """Covariance matrix of the photon-number distribution (means_and_variances.py:126-146)."""
    N = len(mu) // 2
    out = np.zeros((N, N))
    for i in range(N):
        for j in range(i + 1):
            out[i, j] = out[j, i] = photon_number_covar(mu, cov, i, j, hbar=hbar)
    return out


def _s_ordered_batch(mu, cov, rpts, hbar=2, s=0, *, group=None, device=None):
    """s-ordered expectation values for many derivative patterns ``rpts[B, 2M]`` at once: one batched GPU call.
    Modes with no derivatives simply have zero repetitions (their rows drop out), which is what the reference's
    ``reduced_gaussian`` step does explicitly (:233-238)."""
    mu, cov = np.asarray(mu), np.asarray(cov)
    rpts = np.ascontiguousarray(rpts, dtype=np.int32)
    n = len(cov)
    V = (_q.Qmat(cov, hbar=hbar) - 0.5 * (s + 1) * np.identity(n)) @ _q.Xmat(n // 2)
    gamma = None if np.allclose(mu, 0) else _q.complex_to_real_displacements(mu, hbar=hbar).conj()
    return _q.lhaf_patterns(np.ascontiguousarray(V), gamma, rpts, group=group, device=device)


def _real_scalar(z):
    z = complex(z)
    return z.real if abs(z.imag) <= 100 * np.finfo(float).eps * max(1.0, abs(z.real)) else z


def s_ordered_expectation(mu, cov, rpt, hbar=2, s=0, *, device=None):
    """Expectation of the s-ordered product prod a_i^dagger^(n_i) a_j^(m_j), rpt = (n, m)
    (means_and_variances.py:212-249): the loop hafnian of ``reduction((Q - (s+1)/2 I) X, rpt)`` with
    ``alpha^*`` on the diagonal."""
    if np.allclose(rpt, 0):
        return 1.0
    return complex(_s_ordered_batch(mu, cov, np.array([list(rpt)]), hbar=hbar, s=s, device=device)[0])


def normal_ordered_expectation(mu, cov, rpt, hbar=2, *, device=None):
    """s = 1 (means_and_variances.py:195-209)."""
    return s_ordered_expectation(mu, cov, rpt, hbar=hbar, s=1, device=device)


def photon_number_expectation(mu, cov, modes, hbar=2, *, device=None):
    """<prod_{j in modes} n_j> (means_and_variances.py:149-168)."""
    N = len(cov) // 2
    rpt = np.zeros(2 * N, dtype=int)
    for i in modes:
        rpt[i] = rpt[i + N] = 1
    return normal_ordered_expectation(mu, cov, rpt, hbar=hbar, device=device)


def photon_number_squared_expectation(mu, cov, modes, hbar=2, *, group=None, device=None):
    """<prod_{j in modes} n_j^2> = sum over k_j in {1, 2} of <: prod (a_j^dagger a_j)^(k_j) :>
    (means_and_variances.py:171-192), all 2^len(modes) terms in one batched call."""
    modes = list(modes)
    mu_r, cov_r = reduced_gaussian(mu, cov, modes)
    items = np.array(list(product([1, 2], repeat=len(modes))), dtype=np.int32)
    vals = _s_ordered_batch(mu_r, cov_r, np.concatenate([items, items], axis=1), hbar=hbar, s=1, group=group, device=device)
    return _real_scalar(np.sum(vals))


def _coeff_normal_ordered(m, k):
    """Coefficient of a^dagger^k a^k in (a^dagger a)^m: sum_mu (-1)^(k-mu) mu^m / (mu! (k-mu)!) (:309-326)."""
    return sum((-1) ** (k - u) * (u**m) / (factorial(u) * factorial(k - u)) for u in range(k + 1))


def photon_number_moment(mu, cov, indices, hbar=2, *, group=None, device=None):
    """<prod_j n_j^(p_j)> for ``indices = {mode: power}`` (means_and_variances.py:329-367): every power is expanded
    in normal-ordered terms and all prod_j p_j products are evaluated in one batched call."""
    N = len(cov) // 2
    modes = list(indices)
    powers = [int(indices[m]) for m in modes]
    coeff = [[_coeff_normal_ordered(p, k) for k in range(1, p + 1)] for p in powers]
    items = list(product(*[range(p) for p in powers]))
    rpts = np.zeros((len(items), 2 * N), dtype=np.int32)
    weights = np.ones(len(items))
    for t, item in enumerate(items):
        for i, m in enumerate(modes):
            rpts[t, m] = rpts[t, m + N] = item[i] + 1
            weights[t] *= coeff[i][item[i]]
    vals = _s_ordered_batch(mu, cov, rpts, hbar=hbar, s=1, group=group, device=device)
    return np.real_if_close(np.sum(weights * vals))


def partition(collection):
    """All set partitions of a list (means_and_variances.py:370-389)."""
    if len(collection) == 1:
        yield [collection]
        return
    head, rest = collection[0], collection[1:]
    for smaller in partition(rest):
        for i, block in enumerate(smaller):
            yield smaller[:i] + [[head] + block] + smaller[i + 1:]
        yield [[head]] + smaller


def _joint_cumulant(modes, block_value):
    """kappa = sum over set partitions pi of (|pi| - 1)! (-1)^(|pi| - 1) prod_{B in pi} block_value(B)."""
    kappa = 0
    for pi in partition(list(modes)):
        term = factorial(len(pi) - 1) * (-1) ** (len(pi) - 1)
        for block in pi:
            term = term * block_value(block)
        kappa += term
    return kappa


def photon_number_cumulant(mu, cov, modes, hbar=2, *, device=None):
    """Joint photon-number cumulant of the listed modes, repetitions allowed (means_and_variances.py:404-428)."""
    cache = {}

    def moment(block):
        key = tuple(sorted(block))
        if key not in cache:
            cache[key] = photon_number_moment(mu, cov, {m: key.count(m) for m in set(key)}, hbar=hbar, device=device)
        return cache[key]

    return _joint_cumulant(modes, moment)


def click_cumulant(mu, cov, modes, hbar=2, *, device=None):
    """Joint click cumulant of the listed modes (means_and_variances.py:431-457): the block values are the
    probabilities that all detectors of the block click (torontonian / loop torontonian kernel)."""
    from ._torontonian import threshold_detection_prob

    cache = {}

    def all_click(block):
        key = tuple(sorted(set(block)))
        if key not in cache:
            mu_r, cov_r = reduced_gaussian(mu, cov, list(key))
            cache[key] = threshold_detection_prob(mu_r, cov_r, np.ones(len(key), dtype=int), hbar=hbar, device=device)
        return cache[key]

    return _joint_cumulant(modes, all_click)


def _single_mode_vacuum_probs(cov, hbar):
    Q = _q.Qmat(cov, hbar=hbar)
    N = len(cov) // 2
    i = np.arange(N)
    return 1.0 / np.sqrt(np.real(Q[i, i] * Q[i + N, i + N] - Q[i + N, i] * Q[i, i + N])), Q


def mean_clicks(cov, hbar=2):
    """Mean total number of clicks of a zero-mean state under threshold detection (means_and_variances.py:252-271)."""
    p0, _ = _single_mode_vacuum_probs(np.asarray(cov), hbar)
    return float(len(p0) - p0.sum())


def variance_clicks(cov, hbar=2):
    """Variance of the total number of clicks of a zero-mean state (means_and_variances.py:274-306)."""
    cov = np.asarray(cov)
    p0, Q = _single_mode_vacuum_probs(cov, hbar)
    N = len(p0)
    total = float(np.sum(p0 * (1 - p0)))
    for i in range(N):
        for j in range(i):
            idx = [i, j, i + N, j + N]
            total += 2 * (1.0 / np.sqrt(np.linalg.det(Q[np.ix_(idx, idx)]).real) - p0[i] * p0[j])
    return total
