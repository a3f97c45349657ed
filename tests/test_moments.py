"""Photon-number / click statistics callers (thewalrus_b200.moments) against outputs of the reference itself
(tests/golden/reference_moments.json, made by tests/golden/make_golden_moments.py).  On the CPU-only box the GPU
calls (quantum.lhaf_patterns, tor, ltor) are replaced by the oracle; the GPU variants run the same comparisons
through the CUDA kernels."""
import json
import os

import numpy as np
import pytest
from conftest import ROOT

from thewalrus_b200 import _torontonian, moments
from thewalrus_b200 import quantum as q


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(ROOT, "tests", "golden", "reference_moments.json")) as fh:
        return json.load(fh)


@pytest.fixture
def cpu_kernels(monkeypatch):
    from oracle import c_oracle as co
    from oracle import walrus_oracle as wo

    def pats(A, gamma, rpt, glynn=True, *, gamma_index=None, A_index=None, group=None, device=None):
        return co.lhaf_patterns(A, gamma, rpt, glynn)

    monkeypatch.setattr(q, "lhaf_patterns", pats)
    monkeypatch.setattr(_torontonian, "tor", lambda A, recursive=True, **kw: np.complex128(wo.tor_direct(np.asarray(A, dtype=complex))))
    monkeypatch.setattr(_torontonian, "ltor", lambda A, g, recursive=True, **kw: np.complex128(co.ltor_direct(A, g)))


def _c(v):
    return complex(v[0], v[1])


def _ok(got, want, tol=1e-9):
    return abs(complex(got) - want) <= tol * max(1.0, abs(want))


def _check(gold):
    for rec in gold:
        mu, cov, M = np.array(rec["mu"]), np.array(rec["cov"]), rec["M"]
        assert np.allclose(moments.photon_number_covmat(mu, cov), rec["covmat"], rtol=1e-12, atol=1e-12)
        assert abs(moments.mean_clicks(cov) - rec["mean_clicks"]) < 1e-12
        assert abs(moments.variance_clicks(cov) - rec["variance_clicks"]) < 1e-12
        for c in rec["s_ordered"]:
            assert _ok(moments.s_ordered_expectation(mu, cov, c["rpt"], s=c["s"]), _c(c["value"])), c["rpt"]
        for c in rec["expectation"]:
            assert _ok(moments.photon_number_expectation(mu, cov, c["modes"]), _c(c["value"]))
        for c in rec["squared"]:
            assert _ok(moments.photon_number_squared_expectation(mu, cov, c["modes"]), _c(c["value"]))
        for c in rec["moment"]:
            ind = {int(k): v for k, v in c["indices"].items()}
            assert _ok(moments.photon_number_moment(mu, cov, ind), _c(c["value"]))
        for c in rec["cumulant"]:
            assert _ok(moments.photon_number_cumulant(mu, cov, c["modes"]), _c(c["value"]), tol=1e-8)
        for c in rec["click_cumulant"]:
            assert _ok(moments.click_cumulant(mu, cov, c["modes"]), _c(c["value"]), tol=1e-9)
        # second cumulant of two different modes is the covariance (closed form)
        if M >= 2:
            assert _ok(moments.photon_number_cumulant(mu, cov, [0, 1]), moments.photon_number_covar(mu, cov, 0, 1), tol=1e-9)


def test_moments_vs_reference(gold, cpu_kernels):
    _check(gold)


def test_host_only_pieces():
    assert moments.s_ordered_expectation(np.zeros(4), np.identity(4), [0, 0, 0, 0]) == 1.0
    assert [sorted(map(sorted, p)) for p in moments.partition([1, 2, 3])].__len__() == 5          # Bell number B_3
    mu, cov = moments.reduced_gaussian(np.arange(6.0), np.arange(36.0).reshape(6, 6), [2, 0])
    assert mu.tolist() == [2.0, 0.0, 5.0, 3.0] and cov[0, 1] == 12.0
    with pytest.raises(ValueError, match="larger than the number"):
        moments.reduced_gaussian(np.zeros(4), np.identity(4), [5])
    assert abs(moments._coeff_normal_ordered(3, 2) - 3.0) < 1e-12                                  # n^3 = :n: + 3 :n^2: + :n^3:

