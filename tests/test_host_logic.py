"""Host-side logic of the drop-in API that needs no GPU: validation, early exits, closed forms, integer
bookkeeping.  Mirrors the wrapper tests of the reference (thewalrus/tests/test_hafnian.py:86-199,
test_permanent.py:39-101, test_torontonian.py:214-242, test_hafnian_repeated.py:38-84)."""
import numpy as np
import pytest

import thewalrus_b200 as wb
from oracle import walrus_oracle as wo
from thewalrus_b200._prep import dd_sum, glynn_steps, matched_reps, shard_range


def test_hafnian_validation_errors():
    with pytest.raises(TypeError, match="NumPy array"):
        wb.hafnian([[1, 2], [2, 1]])
    with pytest.raises(ValueError, match="square"):
        wb.hafnian(np.ones((2, 3)))
    with pytest.raises(ValueError, match="NaNs"):
        wb.hafnian(np.array([[np.nan, 1.0], [1.0, 1.0]]))
    with pytest.raises(ValueError, match="symmetric"):
        wb.hafnian(np.array([[1.0, 2.0], [3.0, 1.0]]))


def test_hafnian_early_exits_and_closed_forms():
    rng = np.random.default_rng(0)
    assert wb.hafnian(np.zeros((0, 0))) == 1
    assert wb.hafnian(np.ones((3, 3))) == 0.0
    assert wb.hafnian(np.diag([1.0, 2.0, 3.0, 4.0, 5.0, 6.0])) == 0
    assert wb.hafnian(np.diag([1.0, 2.0, 3.0, 4.0, 5.0, 6.0]), loop=True) == 720.0
    for n in (2, 4):
        G = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        A = G + G.T
        assert np.isclose(wb.hafnian(A), wo.haf(A))
        assert np.isclose(wb.hafnian(A, loop=True), wo.loop_hafnian(A))
    A3 = rng.standard_normal((3, 3))
    A3 = A3 + A3.T
    assert np.isclose(wb.hafnian(A3, loop=True), wo.loop_hafnian(A3))


def test_hafnian_approx_rejections():
    A = np.ones((6, 6)) * (1 + 1j)
    with pytest.raises(ValueError, match="must be real"):
        wb.hafnian(A, approx=True)
    B = -np.ones((6, 6))
    with pytest.raises(ValueError, match="negative"):
        wb.hafnian(B, approx=True)


def test_hafnian_repeated_validation_and_exits():
    A = np.ones((4, 4))
    with pytest.raises(ValueError, match="length len"):
        wb.hafnian_repeated(A, [1, 1])
    with pytest.raises(ValueError, match="non-negative integers"):
        wb.hafnian_repeated(A, [1, -1, 1, 1])
    with pytest.raises(ValueError, match="non-negative integers"):
        wb.hafnian_repeated(A, [1, 1.5, 1, 1])
    assert wb.hafnian_repeated(A, [0, 0, 0, 0]) == 1.0
    assert wb.hafnian_repeated(A, [1, 1, 1, 0]) == 0.0
    assert wb.hafnian_repeated(np.zeros((3, 3)), [1, 1, 2]) == 0
    mu = np.array([2.0, 3.0, 5.0])
    assert np.isclose(wb.hafnian_repeated(np.zeros((3, 3)), [1, 2, 1], mu=mu, loop=True), 2 * 9 * 5)
    with pytest.raises(ValueError, match="means vector"):
        wb.hafnian_repeated(A, [1, 1, 1, 1], mu=np.ones(3), loop=True)


def test_loop_hafnian_small_exits():
    A = np.arange(9.0).reshape(3, 3)
    A = A + A.T
    assert wb.loop_hafnian(A, reps=[0, 0, 0]) == 1.0
    assert wb.loop_hafnian(A, D=np.array([7.0, 8.0, 9.0]), reps=[0, 1, 0]) == 8.0
    assert wb._haf(A, reps=[0, 0, 0]) == 1.0
    assert wb._haf(A, reps=[1, 1, 1]) == 0.0


def test_perm_validation_and_closed_forms():
    with pytest.raises(TypeError):
        wb.perm([[1, 2], [3, 4]])
    with pytest.raises(ValueError, match="square"):
        wb.perm(np.ones((2, 3)))
    with pytest.raises(ValueError, match="NaNs"):
        wb.perm(np.array([[np.nan, 1.0], [1.0, 1.0]]))
    with pytest.raises(ValueError, match="method"):
        wb.perm(np.ones((5, 5)), method="bogus")
    rng = np.random.default_rng(1)
    assert wb.perm(np.zeros((0, 0))) == 1.0
    for n in (1, 2, 3):
        A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        assert np.isclose(wb.perm(A), wo.perm_ryser(A))


def test_tor_validation():
    with pytest.raises(TypeError):
        wb.tor([[1, 0], [0, 1]])
    with pytest.raises(ValueError, match="square"):
        wb.tor(np.ones((2, 4)))
    with pytest.raises(ValueError, match="even"):
        wb.tor(np.ones((3, 3)))
    assert wb.tor(np.zeros((0, 0))) == 1.0
    O = np.array([[0.3, 0.1], [0.1, 0.2]])
    assert np.isclose(wb.tor(O), wo.tor(O))  # single mode: closed form on the host


@pytest.mark.parametrize("seed", range(30))
def test_matched_reps_matches_oracle_and_is_consistent(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 9))
    reps = [int(r) for r in rng.integers(0, 5, n)]
    x, er, odd = matched_reps(reps)
    x2, er2, odd2 = wo.matched_reps(reps)
    assert list(x) == list(x2) and list(er) == list(er2) and odd == odd2
    # every vertex is used exactly reps[v] times
    used = np.zeros(n, dtype=int)
    ne = len(er)
    for i in range(ne):
        used[x[i]] += er[i]
        used[x[i + ne]] += er[i]
    if odd is not None:
        used[odd] += 1
    assert list(used) == reps


def test_matched_reps_all_ones_order():
    x, er, odd = matched_reps([1] * 8)
    assert list(x) == [7, 5, 3, 1, 6, 4, 2, 0] and list(er) == [1, 1, 1, 1] and odd is None


def test_steps_kept_edges_and_shards():
    assert glynn_steps([1] * 12) == 2**11
    assert glynn_steps([1] * 12, glynn=False) == 2**12
    assert glynn_steps([3, 2, 1], glynn=True) == 2 * 3 * 2
    assert glynn_steps([3, 2, 1], glynn=True, has_odd=True) == 4 * 3 * 2
    assert list(wb.find_kept_edges(5, np.array([1, 1, 1]))) == [1, 0, 1]
    assert list(wb.find_kept_edges(7, np.array([2, 1, 2]))) == list(wo.find_kept_edges(7, [2, 1, 2]))
    for total in (1, 7, 1 << 24, (1 << 39) + 5):
        for world in (1, 2, 3, 8):
            cuts = [shard_range(total, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_dd_sum_is_compensated():
    pairs = [(1e16, 0.0), (1.0, 0.0), (-1e16, 0.0), (1.0, 1e-17)]
    hi, lo = dd_sum(pairs)
    assert hi + lo == 2.0


def test_reduction():
    A = np.arange(9).reshape(3, 3)
    assert np.array_equal(wb.reduction(A, [2, 0, 1]), A[np.ix_([0, 0, 2], [0, 0, 2])])
    assert np.array_equal(wb.reduction(np.array([5, 6, 7]), [1, 2, 0]), [5, 6, 6])


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    A = np.ones((6, 6))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        wb.hafnian(A)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        wb.perm(np.ones((5, 5)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        wb.tor(np.eye(4) * 0.1)


def test_batch_edge_bookkeeping_matches_oracle():
    from thewalrus_b200.loop_hafnian_batch import add_batch_edges_even, add_batch_edges_odd

    assert add_batch_edges_even(np.array([], dtype=int)).tolist() == [0, 0]
    assert add_batch_edges_odd(np.array([], dtype=int), 0).tolist() == [1, 0, 1, 1]
    rng = np.random.default_rng(3)
    for _ in range(20):
        ne = 2 * int(rng.integers(1, 5))
        fe = rng.permutation(ne + 2)[:ne]
        assert add_batch_edges_even(fe).tolist() == wo._batch_edges_even(fe).tolist()
        odd = int(rng.integers(0, ne + 3))
        assert add_batch_edges_odd(fe, odd).tolist() == wo._batch_edges_odd(fe, odd).tolist()


def test_loop_hafnian_batch_assertions():
    A = np.ones((3, 3), dtype=complex)
    with pytest.raises(AssertionError):
        wb.loop_hafnian_batch(A, np.ones(2, dtype=complex), [1, 1], 2)
    with pytest.raises(AssertionError):
        wb.loop_hafnian_batch(A, np.ones(3, dtype=complex), [1], 2)


def test_gaussian_state_preparation_matches_reference(golden):
    """Qmat / Amat / displacement vector / prefactor against the reference's own outputs
    (thewalrus/quantum/conversions.py:70-169, fock_tensors.py:566-581)."""
    from conftest import dec
    from thewalrus_b200 import quantum as q

    d, c = golden["dme"], golden["conv"]
    mu, cov = np.array(d["mu"]), np.array(d["cov"])
    assert np.allclose(q.Qmat(cov), dec(c["Q"]), rtol=1e-13, atol=1e-14)
    assert np.allclose(q.Amat(cov), dec(c["A"]), rtol=1e-12, atol=1e-14)
    assert np.allclose(q.complex_to_real_displacements(mu), dec(c["beta"]), rtol=1e-14)
    assert np.isclose(q._prefactor(mu, cov), dec(c["prefactor"]), rtol=1e-12)
    with pytest.raises(ValueError, match="shape"):
        q.probabilities_batch(mu, cov, np.zeros((3, 5), dtype=int))


def test_gbs_state_generator_is_a_valid_covariance():
    import bench

    mu, cov, pats = bench.make_gbs_state(16, 500, seed=3016)
    assert cov.shape == (32, 32) and np.allclose(cov, cov.T)
    Om = np.block([[np.zeros((16, 16)), np.identity(16)], [-np.identity(16), np.zeros((16, 16))]])
    assert np.min(np.linalg.eigvalsh(cov + 1j * Om)) > -1e-12          # uncertainty relation (hbar = 2)
    assert pats.shape == (500, 16) and pats.sum(axis=1).max() <= 10 and pats.min() >= 0
