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

__all__ = ["Qmat", "Amat", "Covmat", "Xmat", "sympmat", "complex_to_real_displacements", "density_matrix_element",
           "density_matrix", "pure_state_amplitude", "state_vector", "fock_tensor", "is_pure_cov", "is_symplectic",
           "loss_mat", "update_probabilities_with_loss", "update_probabilities_with_noise", "tvd_cutoff_bounds",
           "n_body_marginals", "find_classical_subsystem", "real_to_complex_displacements",
           "probabilities", "probabilities_batch", "lhaf_patterns", "photon_number_mean_vector", "adj_scaling",
           "adj_to_qmat", "gen_Qmat_from_graph", "is_valid_cov", "is_classical_cov", "williamson"]


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


def lhaf_patterns(A, gamma, rpt, glynn=True, *, gamma_index=None, A_index=None, group=None, device=None):
    """``[loop_hafnian(A, gamma, reps=r) for r in rpt]`` (``gamma=None``: ``hafnian_repeated(A, r)``) on the GPU.

    ``rpt``: integer array ``[B, len(A)]``.  ``A`` may be a stack ``[n_A, nv, nv]`` with ``A_index[B]`` naming the
    matrix of each problem (the batched-matrix front end, see :func:`thewalrus_b200.hafnian_batch`).  ``gamma`` may also be a table ``[G, len(A)]`` of loop vectors with
    ``gamma_index[B]`` naming the row each pattern uses (the batched samplers' call).  With ``group`` the
    patterns are sharded over the ranks in contiguous blocks and the results all-gathered (one collective).
    """
    rpt = np.ascontiguousarray(rpt, dtype=np.int32)
    nv = np.shape(A)[-1]
    if rpt.ndim != 2 or rpt.shape[1] != nv:
        raise ValueError("rpt must have shape [batch, len(A)]")
    if gamma is not None and np.ndim(gamma) == 2 and gamma_index is None and len(gamma) != 1:
        raise ValueError("a table of loop vectors needs gamma_index")
    if np.ndim(A) == 3 and A_index is None and len(A) != 1:
        raise ValueError("a stack of matrices needs A_index")
    return _engine.run_sharded_patterns(A, gamma, rpt, glynn, group, device, gamma_index=gamma_index, A_index=A_index)


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


def _log_factorial_sums(patterns):
    """sum_i log(n_i!) per row (sqrt(prod rpt!) = prod n_i! for rpt = n + n), from a table of log-factorials:
    the photon numbers are small integers, and a table lookup is several times cheaper than 10^6 gammaln calls."""
    patterns = np.asarray(patterns)
    if patterns.size == 0:
        return np.zeros(patterns.shape[0])
    top = int(patterns.max())
    table = np.concatenate([[0.0], np.cumsum(np.log(np.arange(1, max(top, 1) + 1, dtype=np.float64)))])
    return table[patterns].sum(axis=1)


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
    return np.maximum(0.0, (lh * pref).real * np.exp(-_log_factorial_sums(patterns)))


def probabilities(mu, cov, cutoff, parallel=False, hbar=2.0, rtol=1e-05, atol=1e-08, *, group=None, device=None):
    """Fock probabilities up to ``cutoff`` per mode, shape ``[cutoff] * n_modes`` (fock_tensors.py:392-430).
    Every pattern goes through the batched GPU path (the reference's pure-state shortcut through
    ``state_vector`` returns the same numbers)."""
    del parallel, rtol, atol
    M = len(mu) // 2
    pats = np.array(list(product(range(cutoff), repeat=M)), dtype=np.int32).reshape(cutoff ** M, M)
    return probabilities_batch(mu, cov, pats, hbar=hbar, group=group, device=device).reshape([cutoff] * M)


# ---------------------------------------------------------------------------------------------------
# Fock-space tensors: every element is one loop hafnian, all of them in ONE batched GPU call
# ---------------------------------------------------------------------------------------------------
def is_pure_cov(cov, hbar=2, rtol=1e-05, atol=1e-08):
    """Valid covariance with purity 1 / sqrt(det(2 cov / hbar)) = 1 (quantum/gaussian_checks.py:60-76)."""
    if not is_valid_cov(cov, hbar=hbar, rtol=rtol, atol=atol):
        return False
    purity = 1 / np.sqrt(np.linalg.det(2 * np.asarray(cov) / hbar))
    return bool(np.allclose(purity, 1.0, rtol=rtol, atol=atol))


def _pure_parts(mu, cov, hbar, tol):
    """(B*, gamma or None, alpha, det Q) of a pure state (fock_tensors.py:66-76): the amplitude of |i> is
    lhaf(B*, gamma, reps = i) with gamma = alpha - B* alpha*."""
    N = len(cov) // 2
    beta = complex_to_real_displacements(mu, hbar=hbar)
    Bc = np.ascontiguousarray(Amat(cov, hbar=hbar)[:N, :N].conj())
    alpha = beta[:N]
    gamma = None if np.linalg.norm(alpha) < tol else alpha - Bc @ alpha.conj()
    return Bc, gamma, alpha, np.linalg.det(Qmat(cov, hbar=hbar))


def _amplitudes(Bc, gamma, patterns, detQ, group, device):
    """lhaf(B*, gamma, i) / sqrt(prod i! sqrt(det Q)) for every row i of ``patterns`` (one GPU call)."""
    patterns = np.ascontiguousarray(patterns, dtype=np.int32)
    lh = lhaf_patterns(Bc, gamma, patterns, group=group, device=device)
    return lh * np.exp(-0.5 * _log_factorial_sums(patterns)) / np.sqrt(np.sqrt(detQ))


def pure_state_amplitude(mu, cov, i, include_prefactor=True, tol=1e-10, hbar=2, check_purity=True, *, device=None):
    """<i|psi> of a pure Gaussian state (fock_tensors.py:45-105)."""
    if check_purity and not is_pure_cov(cov, hbar=hbar, rtol=1e-05, atol=1e-08):
        raise ValueError("The covariance matrix does not correspond to a pure state")
    Bc, gamma, alpha, detQ = _pure_parts(mu, cov, hbar, tol)
    amp = complex(_amplitudes(Bc, gamma, np.array([list(i)]), detQ, None, device)[0])
    if include_prefactor:
        amp *= np.exp(-0.5 * (np.linalg.norm(alpha) ** 2 - alpha.conj() @ Bc @ alpha.conj()))
    return amp


def _insert_post_selected(free, N, post_select):
    """Full photon-number patterns [len(free), N] from the free-mode indices, post-selected modes filled in."""
    free = np.asarray(free, dtype=np.int32)
    full = np.zeros((len(free), N), dtype=np.int32)
    cols = [m for m in range(N) if m not in post_select]
    full[:, cols] = free
    for m, v in post_select.items():
        full[:, m] = v
    return full


def state_vector(mu, cov, post_select=None, normalize=False, cutoff=5, hbar=2, check_purity=True, *, group=None,
                 device=None, **kwargs):
    """State vector of a (PNR post-selected) pure Gaussian state, shape ``[cutoff] * M`` (fock_tensors.py:108-190).
    The reference takes the multidimensional-Hermite route without post-selection and one hafnian per element with
    it; here both are the same batched loop-hafnian call.  ``choi_r`` (the rescaling :func:`fock_tensor` asks for,
    :162-169) is honoured."""
    if check_purity and not is_pure_cov(cov, hbar=hbar, rtol=1e-05, atol=1e-08):
        raise ValueError("The covariance matrix does not correspond to a pure state")
    N = len(cov) // 2
    Bc, gamma, alpha, detQ = _pure_parts(mu, cov, hbar, 0.0)        # the reference never drops gamma here
    pref = np.exp(-0.5 * (np.linalg.norm(alpha) ** 2 - alpha.conj() @ Bc @ alpha.conj()))
    post_select = dict(post_select or {})
    choi_r = kwargs.get("choi_r", None)
    if choi_r is not None and not post_select:
        resc = np.concatenate([np.ones(N // 2), np.ones(N // 2) / np.tanh(choi_r)])
        Bc = np.ascontiguousarray(resc[:, None] * Bc * resc[None, :])
        gamma = resc * gamma
        detQ = np.linalg.det(Qmat(cov, hbar=hbar) / np.cosh(choi_r))
    M = N - len(post_select)
    free = np.array(list(product(range(cutoff), repeat=M)), dtype=np.int32).reshape(cutoff ** M, M)
    pats = _insert_post_selected(free, N, post_select)
    psi = (pref * _amplitudes(Bc, gamma, pats, detQ.real if not post_select else detQ, group, device)).reshape([cutoff] * M)
    if normalize:
        psi = psi / np.sqrt(np.sum(np.abs(psi) ** 2))
    return psi


def density_matrix(mu, cov, post_select=None, normalize=False, cutoff=5, hbar=2, *, group=None, device=None):
    """Density matrix of a (PNR post-selected) Gaussian state in the Strawberry Fields index order
    (ket_1, bra_1, ket_2, bra_2, ...), shape ``[cutoff] * 2M`` (fock_tensors.py:235-300).  Every element is
    ``density_matrix_element``; all ``cutoff^(2M)`` of them go through one batched GPU call."""
    N = len(mu) // 2
    post_select = dict(post_select or {})
    M = N - len(post_select)
    A, gamma = _state(mu, cov, hbar, 1e-10)
    free = np.array(list(product(range(cutoff), repeat=2 * M)), dtype=np.int32).reshape(cutoff ** (2 * M), 2 * M)
    el0 = _insert_post_selected(free[:, :M], N, post_select)
    el1 = _insert_post_selected(free[:, M:], N, post_select)
    rpt = np.concatenate([el0, el1], axis=1)
    vals = lhaf_patterns(A, gamma, rpt, group=group, device=device) * np.exp(-0.5 * _log_factorial_sums(rpt))
    vals = vals * _prefactor(mu, cov, hbar=hbar)
    # the reference stores element (i, j) at index (j_1, i_1, j_2, i_2, ...) (:283-285)
    rho = vals.reshape([cutoff] * (2 * M))                    # axes (i_1..i_M, j_1..j_M)
    order = [ax for m in range(M) for ax in (M + m, m)]
    rho = np.ascontiguousarray(rho.transpose(order))
    if normalize:
        back = np.arange(2 * M).reshape([M, 2]).T.flatten()
        rho = rho / np.trace(rho.transpose(back).reshape([cutoff**M, cutoff**M])).real
    return rho


def is_symplectic(S, rtol=1e-05, atol=1e-08):
    """S^T Omega S = Omega (thewalrus/symplectic.py is_symplectic)."""
    S = np.asarray(S)
    if S.ndim != 2 or S.shape[0] != S.shape[1] or S.shape[0] % 2:
        return False
    Om = sympmat(S.shape[0] // 2)
    return bool(np.allclose(S.T @ Om @ S, Om, rtol=rtol, atol=atol))


def fock_tensor(S, alpha, cutoff, choi_r=np.arcsinh(1.0), check_symplectic=True, sf_order=False, rtol=1e-05,
                atol=1e-08, *, group=None, device=None):
    """Fock representation of the Gaussian unitary (S, alpha) up to ``cutoff``, shape ``[cutoff] * 2l``
    (fock_tensors.py:303-389) by the Choi-Jamiolkowski trick: the unitary acts on one half of l two-mode squeezed
    vacua (squeezing ``choi_r``) and the tensor is that 2l-mode pure state's vector with the ancilla amplitudes
    rescaled — ``state_vector(..., choi_r=choi_r)``, i.e. ``cutoff^(2l)`` loop hafnians in one batched call.
    (The reference special-cases passive S through its Hermite-polynomial routine; the same path serves both here.)"""
    S = np.asarray(S)
    if check_symplectic and not is_symplectic(S, rtol=rtol, atol=atol):
        raise ValueError("The matrix S is not symplectic")
    l = S.shape[0] // 2
    if l != len(alpha):
        raise ValueError("The matrix S and the vector alpha do not have compatible dimensions")
    ch, sh, zh = np.cosh(choi_r) * np.identity(l), np.sinh(choi_r) * np.identity(l), np.zeros((l, l))
    S_choi = np.block([[ch, sh, zh, zh], [sh, ch, zh, zh], [zh, zh, ch, -sh], [zh, zh, -sh, ch]])
    S_big = np.identity(4 * l)                       # S acting on modes 0..l-1 of 2l modes
    w = np.arange(l)
    for (r0, c0), blk in (((0, 0), S[:l, :l]), ((0, 2 * l), S[:l, l:]), ((2 * l, 0), S[l:, :l]), ((2 * l, 2 * l), S[l:, l:])):
        S_big[np.ix_(w + r0, w + c0)] = blk
    S_exp = S_big @ S_choi
    alphat = np.concatenate([np.asarray(alpha, dtype=np.complex128), np.zeros(l)])
    mu = np.concatenate([2 * alphat.real, 2 * alphat.imag])
    tensor = state_vector(mu, S_exp @ S_exp.T, normalize=False, cutoff=cutoff, hbar=2, check_purity=False,
                          choi_r=choi_r, group=group, device=device)
    if sf_order:
        return tensor.transpose([ax for i in range(l) for ax in (i, i + l)])
    return tensor


def real_to_complex_displacements(beta, hbar=2):
    """xp means from (alpha, alpha^*), the inverse of complex_to_real_displacements (conversions.py:172-190)."""
    beta = np.asarray(beta)
    N = len(beta) // 2
    alpha = beta[:N]
    return np.sqrt(2 * hbar) * np.concatenate([alpha.real, alpha.imag])


def tvd_cutoff_bounds(mu, cov, cutoff, hbar=2, check_is_valid_cov=True, rtol=1e-05, atol=1e-08, *, device=None):
    """Upper bounds on the total variation distance between the exact GBS distribution and its truncations at local
    Fock dimensions 1..cutoff (fock_tensors.py:584-612): sum over modes of the single-mode tail probabilities
    ``1 - cumsum(p_k)``.  The single-mode distributions come from the batched probability call."""
    from .moments import reduced_gaussian

    mu, cov = np.asarray(mu), np.asarray(cov)
    if check_is_valid_cov and not is_valid_cov(cov, hbar=hbar, rtol=rtol, atol=atol):
        raise ValueError("The input covariance matrix violates the uncertainty relation.")
    bounds = np.zeros(cutoff)
    for k in range(len(cov) // 2):
        mu_k, cov_k = reduced_gaussian(mu, cov, [k])
        bounds += 1 - np.cumsum(probabilities(mu_k, cov_k, cutoff, hbar=hbar, device=device))
    return bounds


def n_body_marginals(mean, cov, cutoff, n, hbar=2, *, device=None):
    """Marginal photon-number distributions of all groups of up to ``n`` modes (fock_tensors.py:615-668): a list whose
    entry ``k - 1`` has shape ``[M] * k + [cutoff] * k``; entry ``[m_1..m_k]`` is the joint distribution of those modes
    (zero where indices repeat).  Each sorted group of distinct modes is ONE batched probability call on the reduced
    state; permuted index tuples are filled by transposition."""
    from .moments import reduced_gaussian

    mean, cov = np.asarray(mean), np.asarray(cov)
    if (len(mean), len(mean)) != cov.shape:
        raise ValueError("The covariance matrix and vector of means have incompatible dimensions")
    if len(mean) % 2 != 0:
        raise ValueError("The vector of means is not of even dimensions")
    M = len(mean) // 2
    if M < n:
        raise ValueError("The order of the correlations is higher than the number of modes")
    marginal = [np.zeros([M] * k + [cutoff] * k) for k in range(1, n + 1)]
    for ind in product(range(M), repeat=n):
        distinct_sorted = sorted(set(ind))
        k = len(distinct_sorted)
        if list(ind) == sorted(ind):
            sub_mean, sub_cov = reduced_gaussian(mean, cov, distinct_sorted)
            marginal[k - 1][tuple(distinct_sorted)] = probabilities(sub_mean, sub_cov, cutoff, hbar=hbar, device=device)
        else:
            first_seen = list(dict.fromkeys(ind))
            marginal[k - 1][tuple(first_seen)] = marginal[k - 1][tuple(distinct_sorted)].transpose(np.argsort(first_seen))
    return marginal


def find_classical_subsystem(cov, hbar=2, atol=1e-08):
    """Largest k such that modes 0..k-1 are in a classical state (fock_tensors.py:541-563)."""
    from .moments import reduced_gaussian

    cov = np.asarray(cov)
    N = len(cov) // 2
    if is_classical_cov(cov, hbar=hbar, atol=atol):
        return N
    zero = np.zeros(2 * N)
    k = 0
    while k < N and is_classical_cov(reduced_gaussian(zero, cov, list(range(k + 1)))[1], hbar=hbar, atol=atol):
        k += 1
    return k


def loss_mat(eta, cutoff):
    """Binomial loss matrix L[n, k] = C(n, k) eta^k (1 - eta)^(n - k) (fock_tensors.py:432-458)."""
    from scipy.special import comb

    if eta < 0.0 or eta > 1.0:
        raise ValueError("The transmission parameter eta should be a number between 0 and 1.")
    if eta == 1.0:
        return np.identity(cutoff)
    n, k = np.arange(cutoff)[:, None], np.arange(cutoff)[None, :]
    return np.where(k <= n, comb(n, k) * eta ** k * (1.0 - eta) ** np.maximum(n - k, 0), 0.0)


def update_probabilities_with_loss(etas, probs):
    """Photon-number distribution after per-mode loss (fock_tensors.py:461-486): contract every axis with its
    loss matrix."""
    probs = np.asarray(probs)
    if probs.ndim != len(etas):
        raise ValueError("The list of transmission etas and the tensor of probabilities probs have incompatible dimensions.")
    cutoff = probs.shape[0]
    for axis, eta in enumerate(etas):
        probs = np.moveaxis(np.tensordot(loss_mat(eta, cutoff), probs, axes=([0], [axis])), 0, axis)
    return probs


def update_probabilities_with_noise(probs_noise, probs):
    """Photon-number distribution after adding independent per-mode noise counts (fock_tensors.py:508-538): a
    truncated convolution along every axis."""
    probs = np.asarray(probs)
    if probs.ndim != len(probs_noise):
        raise ValueError("The list of probability distributions probs_noise and the tensor of probabilities probs have incompatible dimensions.")
    cutoff = probs.shape[0]
    for axis, noise in enumerate(probs_noise):
        noise = np.asarray(noise)
        moved = np.moveaxis(probs, axis, 0)
        out = np.zeros_like(moved)
        for i in range(cutoff):
            for j in range(min(i + 1, len(noise))):
                out[i] += moved[i - j] * noise[j]
        probs = np.moveaxis(out, 0, axis)
    return probs


# ---------------------------------------------------------------------------------------------------
# host-side state helpers used by the samplers (O(M^3) NumPy; no GPU work)
# ---------------------------------------------------------------------------------------------------
def Xmat(N):
    """[[0, I], [I, 0]] (conversions.py:46-66)."""
    Z, I = np.zeros((N, N)), np.identity(N)
    return np.block([[Z, I], [I, Z]])


def sympmat(N):
    """Symplectic form [[0, I], [-I, 0]] (thewalrus/symplectic.py sympmat)."""
    Z, I = np.zeros((N, N)), np.identity(N)
    return np.block([[Z, I], [-I, Z]])


def Covmat(Q, hbar=2):
    """xp Wigner covariance from the Husimi matrix, the inverse of Qmat (conversions.py:99-121)."""
    n = len(Q) // 2
    Nm = Q[:n, :n] - np.identity(n)     # <a_i^dagger a_j>
    Mm = Q[n:, :n]                      # <a_i a_j>
    xx = 2 * (Nm.real + Mm.real) + np.identity(n)
    pp = 2 * (Nm.real - Mm.real) + np.identity(n)
    xp = 2 * (Mm.imag + Nm.imag)
    return (hbar / 2) * np.block([[xx, xp], [xp.T, pp]])


def photon_number_mean_vector(mu, cov, hbar=2):
    """<n_j> = (mu_j^2 + mu_(j+N)^2 + cov_jj + cov_(j+N)(j+N) - hbar) / (2 hbar)
    (quantum/means_and_variances.py:31-66)."""
    mu, cov = np.asarray(mu), np.asarray(cov)
    N = len(mu) // 2
    d = np.diag(cov)
    return (mu[:N] ** 2 + mu[N:] ** 2 + d[:N] + d[N:] - hbar) / (2 * hbar)


def adj_scaling(A, n_mean):
    """Scale x such that the pure GBS state encoding x A has total mean photon number ``n_mean``:
    sum_i (x s_i)^2 / (1 - (x s_i)^2) = n_mean over the singular values s_i (quantum/adjacency_matrices.py:90-146)."""
    from scipy.optimize import brentq

    sv = np.linalg.svd(A, compute_uv=False)
    eps = 1e-10
    if 1000 * eps >= sv[0]:
        raise ValueError("The singular values of the matrix A are too small.")

    def excess(x):
        v2 = (x * sv) ** 2
        return np.sum(v2 / (1.0 - v2)) - n_mean

    return brentq(excess, 0.0, 1.0 / (eps + sv[0]), xtol=1e-15, rtol=1e-14)


def adj_to_qmat(A, n_mean):
    """Husimi matrix of the pure state encoding the graph A at mean photon number n_mean:
    Q = (I - X (xA (+) xA^*))^-1 (quantum/adjacency_matrices.py:149-170)."""
    A = np.asarray(A)
    if A.shape[0] != A.shape[1]:
        raise ValueError("Matrix must be square.")
    n = len(A)
    As = adj_scaling(A, n_mean) * A
    Z = np.zeros_like(As)
    return np.linalg.inv(np.identity(2 * n) - Xmat(n) @ np.block([[As, Z], [Z, As.conj()]]))


gen_Qmat_from_graph = adj_to_qmat


def is_valid_cov(cov, hbar=2, rtol=1e-05, atol=1e-08):
    """Square, symmetric, even-sized and cov + i hbar/2 Omega >= 0 (quantum/gaussian_checks.py:26-57)."""
    cov = np.asarray(cov)
    if cov.ndim != 2 or cov.shape[0] != cov.shape[1] or cov.shape[0] % 2:
        return False
    if not np.allclose(cov, cov.T, rtol=rtol, atol=atol):
        return False
    vals = np.linalg.eigvalsh(cov + 0.5j * hbar * sympmat(cov.shape[0] // 2))
    vals[np.abs(vals) < atol] = 0.0
    return bool(np.all(vals >= 0))


def is_classical_cov(cov, hbar=2, atol=1e-08):
    """Valid and cov - hbar/2 I >= 0, i.e. a positive P function (quantum/gaussian_checks.py:79-99)."""
    if not is_valid_cov(cov, hbar=hbar, atol=atol):
        return False
    vals = np.linalg.eigvalsh(np.asarray(cov) - 0.5 * hbar * np.identity(len(cov)))
    vals[np.abs(vals) < atol] = 0.0
    return bool(np.all(vals >= 0))


def williamson(V, rtol=1e-05, atol=1e-08):
    """Williamson form V = S D S^T with S symplectic and D = diag(nu, nu) (thewalrus/decompositions.py:50-102).

    Gamma = V^(-1/2) Omega V^(-1/2) is real antisymmetric; its real Schur form is block diagonal with blocks
    [[0, b_i], [-b_i, 0]], |b_i| = 1 / nu_i.  Orienting every block positively and regrouping the Schur vectors
    from (x1, p1, x2, p2, ...) to (x1.., p1..) gives the orthogonal O with S = V^(1/2) O diag(nu, nu)^(-1/2).
    """
    from scipy.linalg import schur, sqrtm

    V = np.asarray(V)
    if V.shape[0] != V.shape[1]:
        raise ValueError("The input matrix is not square")
    if not np.allclose(V, V.T, rtol=rtol, atol=atol):
        raise ValueError("The input matrix is not symmetric")
    if V.shape[0] % 2 != 0:
        raise ValueError("The input matrix must have an even number of rows/columns")
    n = V.shape[0] // 2
    if not np.all(np.linalg.eigvalsh(V) > 0):
        raise ValueError("Input matrix is not positive definite")
    # S is only fixed up to a rotation inside every mode's (x, p) plane; taking the root with sqrtm and its
    # inverse with inv, as the reference does, lands in the same gauge, so seeded sample streams agree.
    root = np.real_if_close(sqrtm(V))
    iroot = np.linalg.inv(root)
    blocks, O = schur(iroot @ sympmat(n) @ iroot)
    beta = np.diag(blocks, k=1)[::2]
    first = np.where(beta > 0, 2 * np.arange(n), 2 * np.arange(n) + 1)   # column playing "x" in each block
    second = np.where(beta > 0, 2 * np.arange(n) + 1, 2 * np.arange(n))
    O = O[:, np.concatenate([first, second])]
    nu = 1.0 / np.abs(beta)
    dd = np.concatenate([nu, nu])
    return np.diag(dd), (root @ O) / np.sqrt(dd)


# photon-number / click statistics live in thewalrus_b200.moments; the reference exposes them from
# thewalrus.quantum, so they are re-exported here (imported last: moments imports this module)
from .moments import (  # noqa: E402,F401
    click_cumulant,
    mean_clicks,
    normal_ordered_expectation,
    photon_number_covar,
    photon_number_covmat,
    photon_number_cumulant,
    photon_number_expectation,
    photon_number_mean,
    photon_number_moment,
    photon_number_squared_expectation,
    reduced_gaussian,
    s_ordered_expectation,
    variance_clicks,
)
