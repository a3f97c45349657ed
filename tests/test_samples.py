"""Chain-rule samplers (thewalrus_b200.samples) against seeded sample streams of the reference itself
(tests/golden/reference_samples.json, made by tests/golden/make_golden_samples.py).

The non-GPU tests replace the one GPU call of a mode step (quantum.lhaf_patterns) by the CPU oracle, so they
check the host logic: state preparation (Williamson gauge included), heterodyne bookkeeping, the order in which
numpy.random is consumed, acceptance rules.  The GPU tests run the same comparisons through the CUDA kernel and
add a statistical check of the batched path.
"""
import json
import os

import numpy as np
import pytest

import thewalrus_b200 as wb
from oracle import walrus_oracle as wo
from thewalrus_b200 import quantum, samples

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(ROOT, "tests", "golden", "reference_samples.json")) as fh:
        return json.load(fh)


def _oracle_patterns(A, gamma, rpt, glynn=True, *, gamma_index=None, group=None, device=None):
    gamma = np.asarray(gamma)
    if gamma.ndim == 1:
        gamma = gamma[None, :]
    gi = np.zeros(len(rpt), dtype=int) if gamma_index is None else np.asarray(gamma_index)
    return np.array([wo.loop_hafnian(A, gamma[g], [int(v) for v in r]) for r, g in zip(rpt, gi)], dtype=np.complex128)


@pytest.fixture
def cpu_kernel(monkeypatch):
    monkeypatch.setattr(quantum, "lhaf_patterns", _oracle_patterns)


def _arr(x):
    return None if x is None else np.array(x)


def _check_hafnian(c):
    np.random.seed(c["seed"])
    got = samples.hafnian_sample_state(np.array(c["cov"]), len(c["samples"]), mean=_arr(c["mu"]), cutoff=c["cutoff"],
                                       max_photons=c["max_photons"], batch=1)
    assert got.tolist() == c["samples"]


def _check_torontonian(c):
    np.random.seed(c["seed"])
    got = samples.torontonian_sample_state(np.array(c["cov"]), len(c["samples"]), mu=_arr(c["mu"]), fanout=c["fanout"],
                                           cutoff=c["cutoff"], max_photons=c["max_photons"], batch=1)
    assert got.tolist() == c["samples"]


def test_decompose_cov_matches_reference(gold):
    for c in gold["decompose"]:
        T, sqrtW = samples.decompose_cov(np.array(c["cov"]))
        assert np.allclose(T, c["T"], atol=1e-11)
        assert np.allclose(sqrtW, c["sqrtW"], atol=1e-9)       # same Williamson gauge as the reference


def test_hafnian_sampler_reproduces_reference_stream(gold, cpu_kernel):
    for c in gold["hafnian"][:6]:
        _check_hafnian(c)


def test_torontonian_sampler_reproduces_reference_stream(gold, cpu_kernel):
    for c in gold["torontonian"][:6]:
        _check_torontonian(c)


def test_graph_samplers_reproduce_reference_stream(gold, cpu_kernel):
    for c in gold["graph"]:
        np.random.seed(c["seed"])
        A = np.array(c["A"])
        if c["kind"] == "hafnian":
            got = samples.hafnian_sample_graph(A, c["n_mean"], samples=len(c["samples"]), cutoff=c["cutoff"], batch=1)
        else:
            got = samples.torontonian_sample_graph(A, c["n_mean"], samples=len(c["samples"]), fanout=c["fanout"], batch=1)
        assert got.tolist() == c["samples"]


def test_batched_chains_equal_single_chains_given_the_same_draws(gold, cpu_kernel, monkeypatch):
    """Advancing S chains together is the same computation as S single chains: feed both the same normal and
    uniform draws (per chain) and compare."""
    c = gold["hafnian"][2]
    cov, mu = np.array(c["cov"]), _arr(c["mu"])
    ch = samples._Chain(cov, mu, 2)
    rng = np.random.default_rng(5)
    S, M = 4, ch.M
    normals = rng.standard_normal((2, S, 2 * M))
    uniforms = rng.random((M, S))

    def run(rows):
        calls = {"n": 0, "u": 0}

        def normal(size=None):
            out = normals[calls["n"]][rows]
            calls["n"] += 1
            return out

        def uniform(n):
            out = uniforms[calls["u"]][rows]
            calls["u"] += 1
            return out

        monkeypatch.setattr(np.random, "normal", normal)
        monkeypatch.setattr(np.random, "random_sample", uniform)
        return samples._hafnian_chains(ch, len(rows), 4, None)

    together = run(list(range(S)))
    for s in range(S):
        assert run([s])[0].tolist() == together[s].tolist()


def test_sampler_input_checks():
    with pytest.raises(TypeError, match="Covariance matrix must be a NumPy array."):
        samples.hafnian_sample_state([[1.0]], 1)
    with pytest.raises(ValueError, match="Covariance matrix must be square."):
        samples.hafnian_sample_state(np.ones((2, 3)), 1)
    with pytest.raises(ValueError, match="Covariance matrix must not contain NaNs."):
        samples.torontonian_sample_state(np.array([[0, 5], [0, np.nan]]), 1)
    with pytest.raises(ValueError, match="Not a classical covariance matrix"):
        samples.hafnian_sample_classical_state(np.diag([0.2, 5.0]), 3)


def test_classical_and_pmf_samplers():
    np.random.seed(3)
    cov = np.identity(4) * 1.5
    s = samples.hafnian_sample_classical_state(cov, 2000)
    assert s.shape == (2000, 2) and abs(s.mean() - 0.25) < 0.05          # thermal nbar = (1.5 - 1) / 2
    t = samples.torontonian_sample_classical_state(cov, 2000)
    assert set(np.unique(t)) <= {0, 1} and abs(t.mean() - 0.2) < 0.05     # 1 - 1/(1 + nbar)
    p = np.array([[0.1, 0.2], [0.3, 0.3]])
    draws = samples.photon_number_sampler(p, 4000)
    freq = np.zeros((2, 2))
    for d in draws:
        freq[d] += 1
    assert np.allclose(freq / 4000, p / p.sum(), atol=0.03)
    assert samples.photon_number_sampler(np.array([[0.0, 0.0], [0.0, 0.0]]), 3, out_of_bounds=-1) == [-1, -1, -1]
    r1 = samples.hafnian_sample_graph_rank_one(np.array([1.0, 2.0, 1.0]), 1.0, samples=500)
    assert r1.shape == (500, 3) and np.all(r1.sum(axis=1) % 2 == 0) and abs(r1.sum(axis=1).mean() - 1.0) < 0.35


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_hafnian_sampler_reproduces_reference_stream(gold):
    for c in gold["hafnian"]:
        _check_hafnian(c)


@pytest.mark.gpu
def test_gpu_torontonian_sampler_reproduces_reference_stream(gold):
    for c in gold["torontonian"]:
        _check_torontonian(c)


@pytest.mark.gpu
def test_gpu_multi_gamma_patterns_vs_oracle():
    rng = np.random.default_rng(8)
    n = 5
    G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    A = (G + G.T) / 3
    gam = rng.standard_normal((7, n)) + 1j * rng.standard_normal((7, n))
    rpt = rng.integers(0, 3, (40, n)).astype(np.int32)
    gi = rng.integers(0, 7, 40).astype(np.int32)
    got = wb.quantum.lhaf_patterns(A, gam, rpt, gamma_index=gi)
    want = _oracle_patterns(A, gam, rpt, gamma_index=gi)
    assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-12)) < 1e-10


@pytest.mark.gpu
def test_gpu_batched_hafnian_sampler_distribution(gold):
    """4000 chains in one batch against the exact photon-number distribution of the reference
    (thewalrus/tests/test_samples.py uses the same kind of frequency check)."""
    d = gold["distribution"]
    cov, mu, cutoff, probs = np.array(d["cov"]), np.array(d["mu"]), d["cutoff"], np.array(d["probs"])
    np.random.seed(2026)
    n = 4000
    s = samples.hafnian_sample_state(cov, n, mean=mu, cutoff=cutoff - 1, max_photons=2 * cutoff)
    assert s.shape == (n, 2) and s.max() < cutoff
    freq = np.zeros_like(probs)
    for a, b in s:
        freq[a, b] += 1
    # accepted samples follow probs restricted to the accepted set (last mode below the chain cutoff)
    ref = probs.copy()
    ref[:, cutoff - 1:] = 0
    ref[cutoff - 1:, :] = 0
    ref /= ref.sum()
    assert np.max(np.abs(freq / n - ref)) < 4.5 * np.sqrt(ref.max() / n)


@pytest.mark.gpu
def test_gpu_batched_torontonian_sampler_click_rates():
    """Click marginals of the batched threshold sampler against 1 - p(vacuum in mode j)."""
    np.random.seed(11)
    rng = np.random.default_rng(4)
    M = 3
    S = rng.standard_normal((2 * M, 2 * M))
    cov = 0.25 * S @ S.T / (2 * M) + np.identity(2 * M)
    mu = 0.3 * rng.standard_normal(2 * M)
    n = 3000
    s = samples.torontonian_sample_state(cov, n, mu=mu, fanout=8)
    assert s.shape == (n, M) and set(np.unique(s)) <= {0, 1}
    for j in range(M):
        det = np.zeros(M, dtype=int)
        det[j] = 1
        # marginal click probability of mode j: trace the others out = keep only mode j's block
        idx = [j, j + M]
        pj = wb.threshold_detection_prob(mu[idx], cov[np.ix_(idx, idx)], np.array([1]))
        assert abs(s[:, j].mean() - pj) < 4.5 * np.sqrt(pj * (1 - pj) / n) + 0.01
