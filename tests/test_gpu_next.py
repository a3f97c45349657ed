"""GPU parity tests for the SURVEY 8(f) components (run with -m gpu on a B200): the CUDA path through the public
API against outputs of the reference itself (tests/golden/reference_outputs_next.json), the CPU oracle on seeded
inputs, and the identities that tie each component to the core path.  Tolerance 1e-10 relative."""
import numpy as np
import pytest
from conftest import dec

import thewalrus_b200 as wb
from oracle import walrus_oracle as wo
from thewalrus_b200 import _engine

pytestmark = pytest.mark.gpu
TOL = 1e-10


def relv(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


# ------------------------------------------------------------------------------ loop_hafnian_batch_gamma
def test_batch_gamma_vs_reference_outputs(golden_next):
    for c in golden_next["batch_gamma"]:
        got = wb.loop_hafnian_batch_gamma(dec(c["A"]), dec(c["D"]), c["fixed"], c["cutoff"], glynn=c["glynn"])
        want = dec(c["value"])
        assert got.shape == want.shape and got.dtype == np.complex128
        assert relv(got, want) < TOL, (c["fixed"], c["cutoff"], c["glynn"])


def test_batch_gamma_rows_equal_batch():
    """gamma(A, D)[k] == loop_hafnian_batch(A, D[k]) on a larger case than the goldens hold."""
    rng = np.random.default_rng(41)
    n = 7
    G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    A = (G + G.T) / np.sqrt(n)
    D = rng.standard_normal((9, n)) + 1j * rng.standard_normal((9, n))
    for fixed, cutoff in (([1, 2, 0, 1, 1, 2], 6), ([1, 1, 1, 0, 2, 2], 5)):
        got = wb.loop_hafnian_batch_gamma(A, D, fixed, cutoff)
        for k in range(len(D)):
            assert relv(got[k], wb.loop_hafnian_batch(A, D[k], fixed, cutoff)) < TOL
        assert relv(got[3], wo.loop_hafnian_batch(A, D[3], fixed, cutoff)) < TOL


# ------------------------------------------------------------------------------ montrealer
def test_montrealer_vs_reference_outputs(golden_next):
    for c in golden_next["mtl"]:
        assert relv(wb.mtl(dec(c["A"])), dec(c["value"])) < TOL, (c["N"], c["kind"])
    for c in golden_next["lmtl"]:
        assert relv(wb.lmtl(dec(c["A"]), dec(c["zeta"])), dec(c["value"])) < TOL, (c["N"], c["kind"])


@pytest.mark.parametrize("n", [2, 5, 9, 11])
def test_montrealer_vs_oracle_and_ranges(n):
    rng = np.random.default_rng(100 + n)
    G = rng.standard_normal((2 * n, 2 * n)) + 1j * rng.standard_normal((2 * n, 2 * n))
    A = (G + G.T) / (2 * n)
    zeta = 0.5 * (rng.standard_normal(2 * n) + 1j * rng.standard_normal(2 * n))
    assert relv(wb.mtl(A), wo.mtl(A)) < TOL
    assert relv(wb.lmtl(A, zeta), wo.lmtl(A, zeta)) < TOL
    # contiguous label ranges add up (the C-ABI unit of multi-GPU sharding)
    cuts = [0, 1, (1 << n) // 3, (1 << n) - 5 if n > 2 else 2, 1 << n]
    parts = sum(_engine.mtl_range(A, zeta, a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a)
    full = _engine.mtl_range(A, zeta, 0, 1 << n)
    V = lambda t: complex(t[0] + t[1], t[2] + t[3])
    W = lambda t: complex(t[4] + t[5], t[6] + t[7])
    assert relv(V(parts), V(full)) < 1e-12 and relv(W(parts), W(full)) < 1e-12


def test_montrealer_input_checks():
    with pytest.raises(TypeError):
        wb.mtl([[0, 1], [1, 0]])
    with pytest.raises(ValueError):
        wb.mtl(np.zeros((3, 3)))
    with pytest.raises(ValueError):
        wb.lmtl(np.zeros((4, 4)), np.zeros(3))


# ------------------------------------------------------------------------------ Bristolian
def _brs_scale(A, E):
    return max(1.0, float(np.abs(wo.perm_bbfg(A.conj().T @ A + (0 if E is None else E))))) * 2 ** A.shape[0]


def test_bristolian_vs_reference_outputs(golden_next):
    for c in golden_next["brs"]:
        A, E = dec(c["A"]), dec(c["E"])
        assert abs(wb.brs(A, E) - dec(c["value"])) < TOL * _brs_scale(A, E), (c["m"], c["n"])
    for c in golden_next["ubrs"]:
        A = dec(c["A"])
        assert abs(wb.ubrs(A) - dec(c["value"])) < TOL * _brs_scale(A, None), (c["m"], c["n"])
    for c in golden_next["fock_threshold"]:
        U = dec(c["U"])
        assert abs(wb.fock_threshold_prob(c["n"], c["d"], U) - c["unitary"]) < 1e-12
        assert abs(wb.fock_threshold_prob(c["n"], c["d"], np.sqrt(0.8) * U) - c["lossy"]) < 1e-12


@pytest.mark.parametrize("m,n", [(3, 7), (6, 9), (9, 12), (4, 17), (12, 6)])
def test_bristolian_vs_oracle_and_ranges(m, n):
    rng = np.random.default_rng(1000 * m + n)
    A = (rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))) / np.sqrt(m)
    Eh = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    E = 0.05 * (Eh @ Eh.conj().T)
    if m + n <= 21:
        assert abs(wb.brs(A, E) - wo.brs(A, E)) < TOL * _brs_scale(A, E)
        assert abs(wb.ubrs(A) - wo.ubrs(A)) < TOL * _brs_scale(A, None)
    # brs with a single row subset label equals +-perm of the Gram matrix (ties the kernel to wb.perm)
    j = (1 << m) - 2                                    # all rows but the last
    Ay = A[: m - 1]
    one = _engine.brs_range(A, E, j, j + 1)
    got = complex(one[0] + one[1], one[2] + one[3]) / 2.0 ** (n - 1)
    want = -wb.perm(Ay.conj().T @ Ay + E)
    assert abs(got - want) < TOL * max(1.0, abs(want))
    # contiguous label ranges add up
    cuts = [0, 1, (1 << m) // 3, 1 << m]
    parts = sum(_engine.brs_range(A, E, a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a)
    full = _engine.brs_range(A, E, 0, 1 << m)
    c4 = lambda t: complex(t[0] + t[1], t[2] + t[3])
    assert abs(c4(parts) - c4(full)) < 1e-12 * _brs_scale(A, E) * 2.0 ** (n - 1)


def test_fock_prob_matches_permanent():
    rng = np.random.default_rng(8)
    Z = rng.standard_normal((6, 6)) + 1j * rng.standard_normal((6, 6))
    U = np.linalg.qr(Z)[0]
    n_in, n_out = [1, 0, 2, 1, 0, 1], [0, 2, 1, 0, 1, 1]
    p = wb.fock_prob(n_in, n_out, U)
    rows = [i for i, c in enumerate(n_out) for _ in range(c)]
    cols = [i for i, c in enumerate(n_in) for _ in range(c)]
    want = abs(wo.perm(U[np.ix_(rows, cols)])) ** 2 / (2 * 2)
    assert abs(p - want) < 1e-12
    with pytest.raises(ValueError):
        wb.fock_prob([1, 0], [1, 1], np.eye(2))


# ------------------------------------------------------------------------------ loop torontonian
def _rand_O_gamma(N, seed, scale=0.8):
    rng = np.random.default_rng(seed)
    B = rng.standard_normal((2 * N, 2 * N)) + 1j * rng.standard_normal((2 * N, 2 * N))
    H = B @ B.conj().T
    O = scale * H / np.linalg.norm(H, 2)
    g = 0.5 * (rng.standard_normal(N) + 1j * rng.standard_normal(N))
    return O, np.concatenate([g, g.conj()])


def test_ltor_vs_reference_outputs(golden_next):
    for c in golden_next["ltor"]:
        O, gamma = dec(c["O"]), dec(c["gamma"])
        got = wb.ltor(O, gamma)
        assert isinstance(got, np.complex128)
        assert relv(got, dec(c["direct"])) < TOL, len(gamma)
        # yardstick for the second variant: the gap between the reference's own two variants (rec vs direct)
        assert relv(got, dec(c["rec"])) < TOL + relv(dec(c["rec"]), dec(c["direct"])), len(gamma)


def test_ltor_vs_oracle_and_ranges():
    for N, seed in ((2, 1), (3, 2), (7, 3), (10, 4), (12, 5)):
        O, gamma = _rand_O_gamma(N, seed)
        want = wo.ltor_direct(O, gamma)
        assert relv(wb.ltor(O, gamma), want) < TOL, N
        total = _engine.tor_num_prefixes(N)
        if total >= 4:   # prefix ranges are additive (the multi-GPU shard unit)
            cut = [0, total // 4, total // 2 + 1, total]
            parts = [_engine.tor_range(O, cut[i], cut[i + 1], gamma=gamma) for i in range(3)]
            assert relv(sum(p[0] + p[1] for p in parts), want) < TOL, N


def test_ltor_without_displacement_is_tor():
    """thewalrus/tests/test_torontonian.py:260-273."""
    for N in (1, 2, 6, 11):
        O, _ = _rand_O_gamma(N, 20 + N)
        assert relv(wb.ltor(O, np.zeros(2 * N)), wb.tor(O)) < TOL


def test_ltor_input_checks():
    O, gamma = _rand_O_gamma(3, 9)
    with pytest.raises(ValueError, match="gamma must be a vector matching the dimension of A"):
        wb.ltor(O, gamma[:-1])
    with pytest.raises(TypeError):
        wb.ltor(O, list(gamma))


def test_threshold_detection_prob_vs_reference_outputs(golden_next):
    for c in golden_next["threshold"]:
        mu, cov = np.array(c["mu"]), np.array(c["cov"])
        assert abs(wb.threshold_detection_prob(mu, cov, c["det"]) - c["displaced"]) < 1e-12
        assert abs(wb.threshold_detection_prob(0 * mu, cov, c["det"]) - c["zero_mean"]) < 1e-12


def test_threshold_probabilities_sum_to_one():
    """All 2^M click patterns of a displaced 5-mode state (thewalrus/tests/test_torontonian.py:331-357 idea)."""
    from itertools import product

    rng = np.random.default_rng(77)
    M = 5
    S = rng.standard_normal((2 * M, 2 * M))
    cov = S @ S.T / (2 * M) + np.identity(2 * M)
    mu = 0.4 * rng.standard_normal(2 * M)
    tot = sum(wb.threshold_detection_prob(mu, cov, np.array(d)) for d in product([0, 1], repeat=M))
    assert abs(tot - 1.0) < TOL


# ------------------------------------------------------------------------------ batched-matrix front end
def test_hafnian_batch_vs_single_calls_and_oracle():
    rng = np.random.default_rng(61)
    for n, B in ((2, 5), (6, 40), (9, 12), (12, 30), (16, 6)):
        G = rng.standard_normal((B, n, n)) + 1j * rng.standard_normal((B, n, n))
        As = G + np.swapaxes(G, 1, 2)
        for loop in (False, True):
            got = wb.hafnian_batch(As, loop=loop)
            assert got.shape == (B,) and got.dtype == np.complex128
            if n % 2 == 1 and not loop:
                assert np.all(got == 0)
                continue
            want = np.array([wo.loop_hafnian(A, np.diag(A), [1] * n) if loop else wo.haf(A) for A in As])
            assert relv(got, want) < TOL, (n, loop)
            single = np.array([complex(wb.hafnian(np.ascontiguousarray(A), loop=loop)) for A in As[:3]])
            assert relv(got[:3], single) < TOL


def test_hafnian_batch_exact_matchings_and_checks():
    rng = np.random.default_rng(62)
    As = (rng.random((20, 10, 10)) < 0.5).astype(np.float64)
    As = np.triu(As, 1)
    As = As + np.swapaxes(As, 1, 2)
    got = wb.hafnian_batch(As)
    want = np.array([wo.hafnian_by_matchings(A) for A in As])
    assert np.all(np.rint(got.real) == want) and np.max(np.abs(got.imag)) < 1e-9   # integer perfect-matching counts
    with pytest.raises(ValueError, match="symmetric"):
        wb.hafnian_batch(rng.standard_normal((2, 4, 4)))
    with pytest.raises(ValueError, match="square"):
        wb.hafnian_batch(np.zeros((2, 3, 4)))
    assert wb.hafnian_batch(np.zeros((0, 4, 4))).shape == (0,)
    assert np.all(wb.hafnian_batch(np.zeros((3, 0, 0))) == 1)


# ------------------------------------------------------------------------------ remaining drop-in names
def test_permanent_repeated_and_driver_aliases():
    """permanent_repeated (thewalrus/_permanent.py:171-195) and the numba driver names the reference exports
    (thewalrus/__init__.py:126-134)."""
    rng = np.random.default_rng(71)
    A = rng.standard_normal((5, 5)) + 1j * rng.standard_normal((5, 5))
    rpt = [2, 1, 0, 3, 1]
    rows = [i for i, r in enumerate(rpt) for _ in range(r)]
    assert relv(wb.permanent_repeated(A, rpt), wo.perm_bbfg(A[np.ix_(rows, rows)])) < TOL
    O, gamma = _rand_O_gamma(5, 72)
    assert wb.numba_tor(O) == wb.tor(O) == wb.rec_torontonian(O)
    assert wb.numba_ltor(O, gamma) == wb.ltor(O, gamma) == wb.rec_ltorontonian(O, gamma)
