"""Host-side helpers and the closed-form / early-exit paths that never touch the GPU (run on the CPU-only box):
Gaussian-state conversions used by the samplers and the threshold-probability front end, the single-mode
torontonian / loop torontonian, input validation of the batched front ends."""
import numpy as np
import pytest

import thewalrus_b200 as wb
from oracle import walrus_oracle as wo
from thewalrus_b200 import quantum as q


def _random_cov(M, seed, mixed=True):
    rng = np.random.default_rng(seed)
    S = rng.standard_normal((2 * M, 2 * M))
    cov = S @ S.T / (2 * M) + (1.0 if mixed else 0.0) * np.identity(2 * M)
    return cov + np.identity(2 * M)      # comfortably above the vacuum: a valid (classical) covariance


def test_qmat_covmat_roundtrip_and_amat():
    cov = _random_cov(3, 1)
    Q = q.Qmat(cov)
    assert np.allclose(q.Covmat(Q), cov, atol=1e-12)
    assert np.allclose(Q, Q.conj().T)
    A = q.Amat(cov)
    assert np.allclose(A, A.T)                                     # hafnian matrix is symmetric
    assert np.allclose(A, q.Xmat(3) @ (np.identity(6) - np.linalg.inv(Q)).conj())
    assert np.allclose(q.Qmat(cov, hbar=0.7) , q.Qmat(cov * 2 / 0.7, hbar=2))


def test_williamson_is_symplectic_and_reconstructs():
    for M, seed in ((1, 2), (3, 3), (6, 4)):
        cov = _random_cov(M, seed)
        D, S = q.williamson(cov)
        Om = q.sympmat(M)
        assert np.allclose(S @ D @ S.T, cov, atol=1e-10)
        assert np.allclose(S @ Om @ S.T, Om, atol=1e-10)
        d = np.diag(D)
        assert np.allclose(d[:M], d[M:]) and np.all(d > 0)
    with pytest.raises(ValueError, match="not symmetric"):
        q.williamson(np.array([[1.0, 0.5], [0.1, 1.0]]))
    with pytest.raises(ValueError, match="even number"):
        q.williamson(np.identity(3))
    with pytest.raises(ValueError, match="positive definite"):
        q.williamson(np.diag([1.0, -1.0]))


def test_adjacency_scaling_hits_the_mean_photon_number():
    A = np.array([[0, 1, 1, 0], [1, 0, 1, 1], [1, 1, 0, 1], [0, 1, 1, 0]], dtype=float)
    for nbar in (0.3, 1.7):
        Q = q.adj_to_qmat(A, nbar)
        cov = q.Covmat(Q)
        assert abs(q.photon_number_mean_vector(np.zeros(8), cov).sum() - nbar) < 1e-9
        assert q.is_valid_cov(cov) and not q.is_classical_cov(cov)    # squeezed: valid but no positive P function
    assert q.is_classical_cov(_random_cov(2, 5))
    assert not q.is_valid_cov(np.identity(4) * 0.5)                    # below the vacuum noise
    with pytest.raises(ValueError, match="too small"):
        q.adj_scaling(np.zeros((3, 3)), 1.0)


def test_single_mode_torontonians_are_closed_forms():
    rng = np.random.default_rng(9)
    B = rng.standard_normal((2, 2)) + 1j * rng.standard_normal((2, 2))
    H = B @ B.conj().T
    O = 0.7 * H / np.linalg.norm(H, 2)
    g = 0.4 * (rng.standard_normal() + 1j * rng.standard_normal())
    gamma = np.array([g, np.conj(g)])
    assert abs(wb.tor(O) - wo.tor_direct(O)) < 1e-13
    assert abs(wb.ltor(O, gamma) - wo.ltor_direct(O, gamma)) < 1e-13
    assert wb.tor(np.zeros((0, 0))) == 1.0 and wb.ltor(np.zeros((0, 0), dtype=complex), np.zeros(0)) == 1.0
    # vacuum probability of a displaced thermal state against the density-matrix prefactor
    cov, mu = _random_cov(2, 6), np.array([0.3, -0.1, 0.2, 0.4])
    alpha = np.concatenate([mu[:2] + 1j * mu[2:], mu[:2] - 1j * mu[2:]]) / 2
    assert abs(wb.numba_vac_prob(alpha, q.Qmat(cov).conj()) - q._prefactor(mu, cov).real) < 1e-13
    # no detector clicks: the threshold probability is the vacuum probability, no GPU involved
    assert abs(wb.threshold_detection_prob(mu, cov, np.zeros(2, dtype=int)) - q._prefactor(mu, cov).real) < 1e-13
    assert abs(wb.threshold_detection_prob(0 * mu, cov, np.zeros(2, dtype=int)) - q._prefactor(0 * mu, cov).real) < 1e-13


def test_torontonian_input_checks():
    with pytest.raises(TypeError):
        wb.tor([[1.0]])
    with pytest.raises(ValueError, match="square"):
        wb.tor(np.zeros((2, 4)))
    with pytest.raises(ValueError, match="must be even"):
        wb.tor(np.zeros((3, 3)))
    with pytest.raises(ValueError, match="gamma must be a vector matching"):
        wb.ltor(np.zeros((4, 4)), np.zeros(3))


def test_hafnian_batch_host_paths():
    rng = np.random.default_rng(11)
    with pytest.raises(TypeError):
        wb.hafnian_batch([[[1.0]]])
    with pytest.raises(ValueError, match="symmetric"):
        wb.hafnian_batch(rng.standard_normal((2, 4, 4)))
    with pytest.raises(ValueError, match="square"):
        wb.hafnian_batch(np.zeros((2, 3, 4)))
    with pytest.raises(ValueError, match="NaNs"):
        wb.hafnian_batch(np.full((1, 2, 2), np.nan))
    assert wb.hafnian_batch(np.zeros((0, 4, 4))).shape == (0,)
    assert np.all(wb.hafnian_batch(np.zeros((3, 0, 0))) == 1)
    G = rng.standard_normal((4, 5, 5))
    assert np.all(wb.hafnian_batch(G + np.swapaxes(G, 1, 2)) == 0)     # odd size, no loops
    with pytest.raises(ValueError, match="needs A_index"):
        q.lhaf_patterns(np.zeros((2, 2, 2)), None, np.ones((2, 2), dtype=np.int32))
    with pytest.raises(ValueError, match="needs gamma_index"):
        q.lhaf_patterns(np.zeros((2, 2)), np.zeros((3, 2)), np.ones((2, 2), dtype=np.int32))


def test_log_factorial_table_matches_gammaln():
    from scipy.special import gammaln

    pats = np.random.default_rng(2).integers(0, 9, (500, 6))
    assert np.allclose(q._log_factorial_sums(pats), gammaln(pats + 1.0).sum(axis=1), rtol=1e-14, atol=1e-13)
    assert np.all(q._log_factorial_sums(np.zeros((3, 4), dtype=int)) == 0)
    assert q._log_factorial_sums(np.zeros((0, 4), dtype=int)).shape == (0,)


def test_permanent_repeated_host_logic(monkeypatch):
    """permanent_repeated builds [[0, A], [A^T, 0]] with (rpt, rpt) (thewalrus/_permanent.py:171-195).  The subset
    sum is replaced by the oracle here (no GPU on this box); the front-end bookkeeping is what runs."""
    from thewalrus_b200 import _hafnian

    def oracle_sum(Ax, Dx, edge_reps, oddloop, oddV, glynn, group, device):
        return wo.calc_hafnian(Ax, edge_reps, glynn) if Dx is None else wo.calc_loop_hafnian(Ax, Dx, edge_reps, oddloop, oddV, glynn)

    monkeypatch.setattr(_hafnian, "_subset_sum", oracle_sum)
    rng = np.random.default_rng(21)
    A = rng.standard_normal((3, 3)) + 1j * rng.standard_normal((3, 3))
    rpt = [2, 0, 1]
    rows = [0, 0, 2]
    want = wo.perm_bbfg(A[np.ix_(rows, rows)])
    got = wb.permanent_repeated(A, rpt)
    assert abs(got - want) < 1e-12 * max(1.0, abs(want))
    assert wb.version() == wb.__version__
