"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the public API and the
C ABI, against (a) committed outputs of the reference itself, (b) the CPU oracle on the same seeded
inputs, (c) exact integer answers and (d) size-independent properties at the headline sizes.

Tolerances: complex128 results within 1e-10 relative (BASELINE.json north_star); integer matching counts
exact; int64 permanents bit-exact."""
import ctypes
import math

import numpy as np
import pytest
from conftest import dec, random_symmetric, rel

import thewalrus_b200 as wb
from oracle import c_oracle as co
from oracle import walrus_oracle as wo
from thewalrus_b200 import _engine, _lib

pytestmark = pytest.mark.gpu
TOL = 1e-10


# ------------------------------------------------------------------------------ reference goldens
def test_hafnian_vs_reference_outputs(golden):
    for c in golden["hafnian"]:
        A = dec(c["A"])
        A = A.real if c["kind"] == "real" else A
        assert rel(wb.hafnian(A), dec(c["glynn"])) < TOL, c["n"]
        if c["n"] <= 16:
            assert rel(wb.hafnian(A, method="inclexcl"), dec(c["inclexcl"])) < TOL, c["n"]
        assert rel(wb.hafnian(A, method="recursive"), dec(c["glynn"])) < TOL


def test_loop_hafnian_vs_reference_outputs(golden):
    for c in golden["loop_hafnian"]:  # includes odd sizes (unpaired vertex -> general kernel)
        A = dec(c["A"])
        A = A.real if c["kind"] == "real" else A
        assert rel(wb.hafnian(A, loop=True), dec(c["value"])) < TOL, c["n"]


def test_repeated_vs_reference_outputs(golden):
    for c in golden["hafnian_repeated"]:
        assert rel(wb.hafnian_repeated(dec(c["A"]), c["rpt"], glynn=c["glynn"]), dec(c["value"])) < TOL, c["rpt"]
    for c in golden["loop_hafnian_reps"]:
        got = wb.hafnian_repeated(dec(c["A"]), c["rpt"], mu=dec(c["mu"]), loop=True, glynn=c["glynn"])
        assert rel(got, dec(c["value"])) < TOL, c["rpt"]


def test_perm_vs_reference_outputs(golden):
    for c in golden["perm"]:
        if c["kind"] == "int":
            A = np.array(c["A"], dtype=np.int64)
            assert wb.perm(A, "ryser") == c["ryser"] and isinstance(wb.perm(A, "ryser"), int)
            assert wb.perm(A, "bbfg") == c["bbfg"]
            continue
        A = dec(c["A"])
        A = A.real if c["kind"] == "real" else A
        exact = co.perm(A, "bbfg", long_double=True)
        for method in ("bbfg", "glynn", "ryser"):
            got = wb.perm(A, method)
            key = "ryser" if method == "ryser" else "bbfg"
            # the reference's own rounding error (vs 80-bit) is the yardstick where it exceeds 1e-10
            # (Ryser in plain FP64 is itself only ~1e-10 accurate at n = 16 on Haar blocks)
            ref_err = rel(dec(c[key]), exact)
            assert rel(got, dec(c[key])) < TOL + 3 * ref_err, (c["n"], method)
            assert rel(got, exact) < (TOL if method != "ryser" else TOL + 3 * ref_err), (c["n"], method)
        if c["kind"] == "real":
            assert isinstance(wb.perm(A), float)


def test_tor_vs_reference_outputs(golden):
    for c in golden["tor"]:
        O = dec(c["O"])
        O = O.real if c["kind"] == "real" else O
        tol = TOL
        if c["direct"] is not None:      # yardstick: the gap between the reference's own two variants (rec vs direct)
            tol += rel(dec(c["rec"]), dec(c["direct"]))
        got = wb.tor(O)
        assert rel(got, dec(c["rec"])) < tol, c["N"]
        assert rel(wb.tor(O, recursive=False), dec(c["rec"])) < tol
        assert isinstance(got, np.complex128 if c["kind"] == "complex" else np.float64)


# ------------------------------------------------------------------------------ exact integers
def test_perfect_matching_counts_are_exact(golden):
    for c in golden["int_hafnian"]:
        A = np.array(c["A"], dtype=np.int64)
        got = wb.hafnian(A)
        assert abs(got.imag) < 1e-3 and round(got.real) == c["value"], c["n"]


@pytest.mark.parametrize("n", [3, 5, 8, 11, 12])
def test_hafnian_of_ones_exact(n):
    want = math.factorial(2 * n) // (math.factorial(n) * 2**n)
    assert round(wb.hafnian(np.ones((2 * n, 2 * n))).real) == want
    T = [1, 1]
    for k in range(2, 2 * n + 1):
        T.append(T[-1] + (k - 1) * T[-2])
    assert rel(wb.hafnian(np.ones((2 * n, 2 * n)), loop=True).real, T[2 * n]) < 1e-12


def test_int64_perm_exact_and_wraps_like_int64():
    rng = np.random.default_rng(9)
    for n in (6, 10, 15):
        A = rng.integers(-3, 4, (n, n)).astype(np.int64)
        assert wb.perm(A, "ryser") == int(wo.perm_ryser(A.astype(object)))
    assert wb.perm(np.ones((12, 12), dtype=np.int64), "ryser") == math.factorial(12)
    assert wb.perm(np.ones((12, 12), dtype=np.int64)) == float(math.factorial(12))


# ------------------------------------------------------------------------------ oracle on seeded inputs
@pytest.mark.parametrize("n", [6, 8, 10, 18, 24, 26, 30, 34])
def test_hafnian_and_loop_vs_oracle(n):
    rng = np.random.default_rng(1000 + n)
    A = random_symmetric(rng, n)
    assert rel(wb.hafnian(A), co.hafnian(A)) < TOL
    assert rel(wb.hafnian(A, loop=True), co.hafnian(A, loop=True)) < TOL


def test_hafnian_haar_family_vs_long_double():
    rng = np.random.default_rng(77)
    n = 28
    Z = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    U = np.linalg.qr(Z)[0]
    A = 0.5 * (U @ U.T + (U @ U.T).T)
    assert rel(wb.hafnian(A), co.hafnian(A, long_double=True)) < TOL


def test_cabi_ranges_match_oracle_and_add_up():
    lib = _lib.load()
    rng = np.random.default_rng(5)
    n = 20
    A = random_symmetric(rng, n)
    x = co.matched_order(A)
    Ax = np.ascontiguousarray(A[np.ix_(x, x)])
    Dx = np.ascontiguousarray(np.diag(A)[x])
    steps = 1 << (n // 2 - 1)

    def run(j0, j1, D=None):
        out = np.zeros(4)
        rc = lib.wb200_hafnian_host(0, _lib.dptr(Ax.view(np.float64)), None if D is None else _lib.dptr(D.view(np.float64)),
                                    n, j0, j1, _lib.dptr(out), None)
        assert rc == 0
        return complex(out[0] + out[1], out[2] + out[3])

    for (a, b) in ((0, steps), (0, 1), (3, 4), (5, 130), (129, 130), (1, steps - 1), (steps - 3, steps)):
        assert rel(run(a, b), co.hafnian_range(Ax, a, b)) < TOL, (a, b)
        assert rel(run(a, b, Dx), co.hafnian_range(Ax, a, b, Dx)) < TOL, (a, b)
    assert run(7, 7) == 0
    cuts = [0, 1, 6, 100, 101, 400, steps]
    assert rel(sum(run(a, b) for a, b in zip(cuts[:-1], cuts[1:])), run(0, steps)) < 1e-12
    out = np.zeros(4)
    assert lib.wb200_hafnian_host(0, _lib.dptr(Ax.view(np.float64)), None, n, 0, steps + 1, _lib.dptr(out), None) == -1
    assert lib.wb200_hafnian_host(0, _lib.dptr(Ax.view(np.float64)), None, 7, 0, 1, _lib.dptr(out), None) == -1


def test_perm_cabi_ranges():
    lib = _lib.load()
    rng = np.random.default_rng(6)
    n = 18
    M = np.ascontiguousarray(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))

    def run(method, a, b):
        out = np.zeros(4)
        assert lib.wb200_perm_host(0, _lib.dptr(M.view(np.float64)), n, method, a, b, _lib.dptr(out), None) == 0
        return complex(out[0] + out[1], out[2] + out[3])

    for method, steps in ((0, 1 << (n - 1)), (1, 1 << n)):
        for (a, b) in ((0, steps), (0, 1), (1, 2), (63, 65), (1000, 77777), (steps - 1, steps)):
            assert rel(run(method, a, b), co.perm_range(M, method, a, b, long_double=True)) < TOL, (method, a, b)
        cuts = [0, 5, 64, 4097, steps // 2 + 3, steps]
        assert rel(sum(run(method, a, b) for a, b in zip(cuts[:-1], cuts[1:])), run(method, 0, steps)) < 1e-12


@pytest.mark.parametrize("n", [4, 7, 13, 20, 24])
def test_perm_vs_long_double_oracle(n):
    rng = np.random.default_rng(2000 + n)
    Z = rng.standard_normal((2 * n, 2 * n)) + 1j * rng.standard_normal((2 * n, 2 * n))
    U = np.linalg.qr(Z)[0][:n, :n]  # n x n block of a 2n-mode Haar unitary (BASELINE C2 recipe)
    exact = co.perm(U, "bbfg", long_double=True)
    assert rel(wb.perm(U), exact) < TOL
    if n <= 20:
        # Ryser in plain FP64 is itself not 1e-10 accurate on Haar blocks: the yardstick is the error of the
        # reference algorithm's own FP64 evaluation (the C port of perm_ryser) against the 80-bit value
        port_err = rel(co.perm(U, "ryser"), exact)
        err = rel(wb.perm(U, "ryser"), exact)
        print(f"\n[ryser n={n}] gpu rel_err {err:.2e}, reference-algorithm FP64 port {port_err:.2e}")
        assert err < TOL + port_err
    R = rng.standard_normal((n, n))
    assert rel(wb.perm(R), co.perm(R, "bbfg", long_double=True).real) < TOL


@pytest.mark.parametrize("N", [2, 3, 4, 7, 9, 10, 13, 16])
def test_tor_vs_oracle(N):
    rng = np.random.default_rng(3000 + N)
    B = rng.standard_normal((2 * N, 2 * N)) + 1j * rng.standard_normal((2 * N, 2 * N))
    H = B @ B.conj().T
    O = 0.9 * H / np.linalg.norm(H, 2)
    want = co.tor_recursive(O, long_double=True)
    assert rel(wb.tor(O), want) < TOL
    total = _engine.tor_num_prefixes(N)
    if total >= 4:  # prefix ranges add up
        cuts = [0, 1, total // 2 + 1, total]
        s = sum(sum(_engine.tor_range(O, a, b)) for a, b in zip(cuts[:-1], cuts[1:]))
        assert rel(s, want) < TOL


def test_general_kernel_matches_dmma_kernel_and_oracle():
    rng = np.random.default_rng(8)
    n = 14
    A = random_symmetric(rng, n)
    x, er, _ = wb.matched_reps([1] * n)
    Ax = A[np.ix_(x, x)].astype(np.complex128)
    Dx = np.diag(A)[x].astype(np.complex128)
    steps = 1 << (n // 2 - 1)
    fast = _engine.combine4([_engine.hafnian_range(Ax, Dx, 0, steps)])
    gen = _engine.combine4([_engine.lhaf_general_range(Ax, Dx, None, None, er, True, 0, steps)])
    assert rel(fast, gen) < 1e-12
    assert rel(gen, co.hafnian_range(Ax, 0, steps, Dx)) < 1e-12
    # sub-range of the mixed-radix index with repeated edges against the numpy oracle
    reps = np.array([2, 1, 3, 1, 2, 1, 1])
    g = _engine.combine4([_engine.lhaf_general_range(Ax, Dx, None, None, reps, True, 5, 77)])
    assert rel(g, wo.calc_loop_hafnian(Ax, Dx, reps, j0=5, j1=77, scale=False)) < 1e-11


def test_random_repeated_patterns_vs_oracle():
    rng = np.random.default_rng(10)
    for _ in range(12):
        n = int(rng.integers(2, 7))
        A = random_symmetric(rng, n)
        mu = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        rpt = [int(r) for r in rng.integers(0, 5, n)]
        if sum(rpt) < 2:
            rpt[0] += 2
        for glynn in (True, False):
            assert rel(wb.hafnian_repeated(A, rpt, mu=mu, loop=True, glynn=glynn), wo.loop_hafnian(A, mu, rpt, glynn)) < TOL
            if sum(rpt) % 2 == 0:
                assert rel(wb.hafnian_repeated(A, rpt, glynn=glynn), wo.haf(A, rpt, glynn)) < TOL


# ------------------------------------------------------------------------------ headline sizes: properties
def test_hafnian_n50_sampled_ranges_vs_extended_precision():
    """BASELINE metric size (n = 50, 2^24 subsets): sampled index ranges against the 80-bit oracle."""
    lib = _lib.load()
    n = 50
    rng = np.random.default_rng(1000 * 1 + n)
    A = random_symmetric(rng, n)
    x = co.matched_order(A)
    Ax = np.ascontiguousarray(A[np.ix_(x, x)])
    for j0 in (0, 123457, (1 << 24) - 40):
        out = np.zeros(4)
        assert lib.wb200_hafnian_host(0, _lib.dptr(Ax.view(np.float64)), None, n, j0, j0 + 37, _lib.dptr(out), None) == 0
        got = complex(out[0] + out[1], out[2] + out[3])
        assert rel(got, co.hafnian_range(Ax, j0, j0 + 37, long_double=True)) < TOL


def test_hafnian_n56_and_n64_sampled_ranges():
    lib = _lib.load()
    for n in (56, 64):
        rng = np.random.default_rng(1000 * 5 + n)
        A = random_symmetric(rng, n) / np.sqrt(n)
        x = co.matched_order(A)
        Ax = np.ascontiguousarray(A[np.ix_(x, x)])
        Dx = np.ascontiguousarray(np.diag(A)[x])
        j0 = (1 << (n // 2 - 2)) + 12345
        out = np.zeros(4)
        assert lib.wb200_hafnian_host(0, _lib.dptr(Ax.view(np.float64)), _lib.dptr(Dx.view(np.float64)), n, j0, j0 + 13,
                                      _lib.dptr(out), None) == 0
        got = complex(out[0] + out[1], out[2] + out[3])
        assert rel(got, co.hafnian_range(Ax, j0, j0 + 13, Dx, long_double=True)) < TOL


def test_full_size_properties_n40():
    """Homogeneity haf(cA) = c^(n/2) haf(A) and the block identity haf([[0,B],[B^T,0]]) = perm(B) on full runs."""
    rng = np.random.default_rng(40)
    n = 40
    A = random_symmetric(rng, n) / np.sqrt(n)
    h = wb.hafnian(A)
    c = 0.9 - 0.3j
    assert rel(wb.hafnian(c * A), c ** (n // 2) * h) < TOL
    k = 16
    B = (rng.standard_normal((k, k)) + 1j * rng.standard_normal((k, k))) / np.sqrt(k)
    Z = np.zeros((k, k))
    assert rel(wb.hafnian(np.block([[Z, B], [B.T, Z]])), co.perm(B, "bbfg", long_double=True)) < TOL
    assert rel(wb.perm(B), co.perm(B, "bbfg", long_double=True)) < TOL


def test_perm_properties_n30():
    rng = np.random.default_rng(30)
    n = 30
    P = np.eye(n)[rng.permutation(n)].astype(np.complex128)
    assert rel(wb.perm(P), 1.0) < 1e-12
    d = np.exp(1j * rng.uniform(0, 2 * np.pi, n))
    assert rel(wb.perm(d[:, None] * P), np.prod(d)) < 1e-10


def test_tor_tmsv_closed_form():
    r = 0.7
    c, s = np.cosh(r), np.sinh(r)
    Q = np.array([[c * c, 0, 0, c * s], [0, c * c, c * s, 0], [0, c * s, c * c, 0], [c * s, 0, 0, c * c]], dtype=complex)
    O = np.eye(4) - np.linalg.inv(Q)
    assert rel(wb.tor(O).real / np.sqrt(np.linalg.det(Q).real), np.tanh(r) ** 2) < 1e-12
    assert abs(wb.tor(np.zeros((8, 8)))) < 1e-12  # vacuum (test_torontonian.py:114-119)


def test_device_peak_probe():
    lib = _lib.load()
    t = ctypes.c_double(0)
    assert lib.wb200_fp64_peak(0, 1, ctypes.byref(t)) == 0
    assert 10 < t.value < 60


# ------------------------------------------------------------------------------ batched front ends (a8, a9)
def test_loop_hafnian_batch_vs_reference_outputs(golden):
    for c in golden["batch"]:
        got = wb.loop_hafnian_batch(dec(c["A"]), dec(c["D"]), c["fixed"], c["cutoff"], glynn=c["glynn"])
        want = dec(c["value"])
        assert got.shape == want.shape and got.dtype == np.complex128
        scale = max(np.max(np.abs(want)), 1e-300)
        assert np.max(np.abs(got - want)) / scale < TOL, (c["fixed"], c["cutoff"], c["glynn"])


def test_loop_hafnian_batch_identity_and_oracle():
    """batch[k] == loop_hafnian(A, D, reps = fixed + [k]) (the pin for the unpinned boundary)."""
    rng = np.random.default_rng(11)
    for trial in range(10):
        n = int(rng.integers(2, 7))
        A = random_symmetric(rng, n) / np.sqrt(n)
        D = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        fixed = [int(r) for r in rng.integers(0, 4, n - 1)]
        cutoff = int(rng.integers(1, 9))
        for glynn in (True, False):
            got = wb.loop_hafnian_batch(A, D, fixed, cutoff, glynn=glynn)
            want = np.array([wo.loop_hafnian(A, D, fixed + [k], glynn) for k in range(cutoff + 1)])
            assert np.max(np.abs(got - want)) / max(np.max(np.abs(want)), 1e-300) < TOL, (fixed, cutoff, glynn)
            orc = wo.loop_hafnian_batch(A, D, fixed, cutoff, glynn)
            assert np.max(np.abs(got - orc)) / max(np.max(np.abs(orc)), 1e-300) < TOL


def test_lhaf_patterns_vs_oracle_all_branches():
    from thewalrus_b200 import quantum as q

    rng = np.random.default_rng(12)
    nv = 8
    A = random_symmetric(rng, nv) / np.sqrt(nv)
    g = rng.standard_normal(nv) + 1j * rng.standard_normal(nv)
    rpt = rng.integers(0, 4, (200, nv)).astype(np.int32)
    rpt[0] = 0                       # N = 0 -> 1
    rpt[1] = 0; rpt[1, 3] = 1        # N = 1 -> gamma[3]
    rpt[2] = 0; rpt[2, 5] = 7        # one vertex, self-paired + odd leftover
    rpt[3] = 0; rpt[3, 2] = 6        # one vertex, self-paired
    got = q.lhaf_patterns(A, g, rpt)
    for b in range(len(rpt)):
        want = wo.loop_hafnian(A, g, [int(r) for r in rpt[b]])
        assert abs(got[b] - want) <= TOL * max(abs(want), 1e-12), (b, rpt[b])
    got0 = q.lhaf_patterns(A, None, rpt)      # no loops: odd totals vanish
    for b in range(len(rpt)):
        want = wo.haf(A, [int(r) for r in rpt[b]])
        assert abs(got0[b] - want) <= TOL * max(abs(want), 1e-12), (b, rpt[b])
    got_ie = q.lhaf_patterns(A, g, rpt[:40], glynn=False)
    for b in range(40):
        want = wo.loop_hafnian(A, g, [int(r) for r in rpt[b]], False)
        assert abs(got_ie[b] - want) <= TOL * max(abs(want), 1e-12), (b, rpt[b])
    assert q.lhaf_patterns(A, g, np.zeros((0, nv), dtype=np.int32)).shape == (0,)


def test_gbs_probabilities_vs_reference_outputs(golden):
    d = golden["dme"]
    mu, cov, pats = np.array(d["mu"]), np.array(d["cov"]), np.array(d["patterns"])
    p = wb.probabilities_batch(mu, cov, pats)
    assert np.max(np.abs(p - np.array(d["displaced"])) / np.maximum(np.array(d["displaced"]), 1e-30)) < TOL
    p0 = wb.probabilities_batch(0 * mu, cov, pats)
    want0 = np.array(d["zero_mean"])
    assert np.max(np.abs(p0 - want0)) < TOL * np.max(want0)
    one = wb.density_matrix_element(mu, cov, list(pats[3]), list(pats[3]))
    assert rel(one.real, d["displaced"][3]) < TOL


def test_gbs_probabilities_16_modes_sample_and_normalisation():
    """BASELINE config 3 state (16 modes, eta = 0.8, displaced): a pattern sample against the general kernel and
    the NumPy oracle, and sum_n p(n) -> 1 over a 2-mode marginal-free small state."""
    from thewalrus_b200 import quantum as q

    sys_path_bench = __import__("bench")
    mu, cov, pats = sys_path_bench.make_gbs_state(16, 2000, seed=3016)
    A, gamma = q._state(mu, cov, 2, 1e-10)
    rpt = np.concatenate([pats, pats], axis=1)
    got = q.lhaf_patterns(A, gamma, rpt)
    idx = np.argsort(pats.sum(axis=1))[[0, 10, 500, 1000, 1500, 1990, 1999]]
    for b in idx:
        r = [int(x) for x in rpt[b]]
        want = wb.hafnian_repeated(A, r, mu=gamma, loop=True)
        assert abs(got[b] - want) <= TOL * max(abs(want), 1e-300), (b, pats[b])
        if sum(r) <= 8:
            assert abs(got[b] - wo.loop_hafnian(A, gamma, r)) <= TOL * max(abs(want), 1e-300)
    # normalisation on a small lossy displaced state: probabilities over a generous cutoff sum to ~1
    mu2, cov2, _ = sys_path_bench.make_gbs_state(2, 1, seed=7, r=0.4)
    P = wb.probabilities(mu2, cov2, 14)
    assert abs(P.sum() - 1.0) < 1e-6 and P.min() >= 0.0     # truncation of the Fock space at 14 photons per mode, not rounding


def test_gbs_probabilities_config3_state_vs_reference_outputs(golden):
    """16-mode BASELINE config 3 state: the reference's own density_matrix_element on a pattern sample
    (0..10 photons), regenerated inputs checked against the stored ones."""
    import bench

    d = golden["dme16"]
    mu, cov, pats = bench.make_gbs_state(16, d["B"], seed=d["seed"])
    assert np.allclose(mu, np.array(d["mu"])) and np.allclose(cov, np.array(d["cov"]))
    sel = pats[d["index"]]
    assert sel.tolist() == d["patterns"]
    p = wb.probabilities_batch(mu, cov, pats)[d["index"]]
    want = np.array(d["displaced"])
    assert np.max(np.abs(p - want) / want) < TOL
    p0 = wb.probabilities_batch(0 * mu, cov, sel)
    want0 = np.array(d["zero_mean"])
    assert np.max(np.abs(p0 - want0)) < TOL * np.max(want0)
