"""Fock-space callers of the hot path (pure_state_amplitude, state_vector, density_matrix) against outputs of the
reference itself (tests/golden/reference_quantum.json, made by tests/golden/make_golden_quantum.py).
On the CPU-only box the one GPU call (quantum.lhaf_patterns) is replaced by the C oracle, so these tests check the
host logic: prefactors, post-selection bookkeeping, the Strawberry Fields index order.  The GPU variants run the
same comparisons through the CUDA kernel."""
import json
import os

import numpy as np
import pytest
from conftest import ROOT, dec

from thewalrus_b200 import quantum as q


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(ROOT, "tests", "golden", "reference_quantum.json")) as fh:
        return json.load(fh)


@pytest.fixture
def cpu_kernel(monkeypatch):
    from oracle import c_oracle as co

    def oracle(A, gamma, rpt, glynn=True, *, gamma_index=None, A_index=None, group=None, device=None):
        assert gamma_index is None and A_index is None
        return co.lhaf_patterns(A, gamma, rpt, glynn)

    monkeypatch.setattr(q, "lhaf_patterns", oracle)


def _close(a, b, tol=1e-10):
    a, b = np.asarray(a), np.asarray(b)
    return np.max(np.abs(a - b)) <= tol * max(1.0, float(np.max(np.abs(b))))


def _check_pure(gold):
    for c in gold["pure"]:
        mu, cov = np.array(c["mu"]), np.array(c["cov"])
        for a in c["amps"]:
            assert _close(q.pure_state_amplitude(mu, cov, a["i"]), dec(a["with_pref"]))
            assert _close(q.pure_state_amplitude(mu, cov, a["i"], include_prefactor=False), dec(a["no_pref"]))
        assert _close(q.state_vector(mu, cov, cutoff=4), dec(c["state_vector"]))
        assert _close(q.state_vector(mu, cov, cutoff=4, normalize=True), dec(c["state_vector_norm"]))
        ps = {int(k): v for k, v in c["post_select"].items()}
        assert _close(q.state_vector(mu, cov, post_select=ps, cutoff=4), dec(c["state_vector_ps"]))
        assert _close(q.state_vector(mu, cov, post_select=ps, cutoff=4, normalize=True), dec(c["state_vector_ps_norm"]))


def _check_mixed(gold):
    for c in gold["mixed"]:
        mu, cov = np.array(c["mu"]), np.array(c["cov"])
        for e in c["elements"]:
            assert _close(q.density_matrix_element(mu, cov, e["i"], e["j"]), dec(e["value"]))
        rho = q.density_matrix(mu, cov, cutoff=3)
        assert rho.shape == tuple([3] * (2 * c["M"])) and _close(rho, dec(c["density_matrix"]))
        if "post_select" in c:
            ps = {int(k): v for k, v in c["post_select"].items()}
            assert _close(q.density_matrix(mu, cov, post_select=ps, cutoff=3), dec(c["density_matrix_ps"]))
            assert _close(q.density_matrix(mu, cov, post_select=ps, cutoff=3, normalize=True), dec(c["density_matrix_ps_norm"]))


def test_pure_state_callers_vs_reference(gold, cpu_kernel):
    _check_pure(gold)


def test_density_matrix_vs_reference(gold, cpu_kernel):
    _check_mixed(gold)


def test_purity_check():
    with pytest.raises(ValueError, match="pure state"):
        q.pure_state_amplitude(np.zeros(2), 3.0 * np.identity(2), [1])
    with pytest.raises(ValueError, match="pure state"):
        q.state_vector(np.zeros(2), 3.0 * np.identity(2))
    assert q.is_pure_cov(np.identity(4)) and not q.is_pure_cov(2 * np.identity(4))




# ------------------------------------------------------------------------------------------- fock_tensor, loss, noise
@pytest.fixture(scope="module")
def gold_ft():
    with open(os.path.join(ROOT, "tests", "golden", "reference_fock_tensor.json")) as fh:
        return json.load(fh)


def _check_fock_tensor(gold_ft):
    for c in gold_ft["fock_tensor"]:
        got = q.fock_tensor(np.array(c["S"]), dec(c["alpha"]), c["cutoff"], sf_order=c["sf_order"])
        want = dec(c["value"])
        assert got.shape == want.shape and _close(got, want, tol=1e-9)


def test_fock_tensor_vs_reference(gold_ft, cpu_kernel):
    _check_fock_tensor(gold_ft)
    with pytest.raises(ValueError, match="not symplectic"):
        q.fock_tensor(2 * np.identity(2), np.zeros(1), 3)
    with pytest.raises(ValueError, match="compatible dimensions"):
        q.fock_tensor(np.identity(2), np.zeros(2), 3)


def test_loss_and_noise_updates_vs_reference(gold_ft):
    for c in gold_ft["loss"]:
        p = np.array(c["probs"])
        assert np.allclose(q.update_probabilities_with_loss(c["etas"], p), c["lossy"], rtol=1e-12, atol=1e-15)
        assert np.allclose(q.update_probabilities_with_noise([np.array(n) for n in c["noise"]], p), c["noisy"], rtol=1e-12, atol=1e-15)
        assert np.allclose(q.loss_mat(0.3, 5), c["loss_mat"], rtol=1e-12, atol=1e-15)
    assert np.array_equal(q.loss_mat(1.0, 4), np.identity(4))
    with pytest.raises(ValueError, match="between 0 and 1"):
        q.loss_mat(1.5, 3)
    with pytest.raises(ValueError, match="incompatible dimensions"):
        q.update_probabilities_with_loss([0.5], np.ones((2, 2)))



# ------------------------------------------------------------------------------------------- marginals, tvd bounds
@pytest.fixture(scope="module")
def gold_marg():
    with open(os.path.join(ROOT, "tests", "golden", "reference_marginals.json")) as fh:
        return json.load(fh)


def _check_marginals(gold_marg):
    for c in gold_marg:
        mu, cov = np.array(c["mu"]), np.array(c["cov"])
        assert np.allclose(q.tvd_cutoff_bounds(mu, cov, 5), c["tvd"], rtol=1e-9, atol=1e-12)
        got = q.n_body_marginals(mu, cov, 3, 2)
        assert len(got) == 2
        for g, w in zip(got, c["marginals"]):
            assert g.shape == np.array(w).shape and np.allclose(g, w, rtol=1e-9, atol=1e-12)


def test_marginals_and_tvd_bounds_vs_reference(gold_marg, cpu_kernel):
    _check_marginals(gold_marg)
    for c in gold_marg:
        mu, cov = np.array(c["mu"]), np.array(c["cov"])
        assert q.find_classical_subsystem(cov) == c["classical"]
        assert np.allclose(q.real_to_complex_displacements(q.complex_to_real_displacements(mu)), c["r2c"])
    with pytest.raises(ValueError, match="higher than the number of modes"):
        q.n_body_marginals(np.zeros(2), np.identity(2), 3, 2)
    with pytest.raises(ValueError, match="violates the uncertainty relation"):
        q.tvd_cutoff_bounds(np.zeros(2), 0.1 * np.identity(2), 3)

